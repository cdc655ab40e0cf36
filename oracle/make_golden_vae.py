"""TEST INFRASTRUCTURE ONLY — tests/golden/vae_decode.pt: AutoencoderKL.decode of the reference's diffusers run on seeded
latents with name-keyed synthetic weights (block_out_channels 32/64/64/64), via oracle/ref_shim.py.
Run in the build container (needs /root/reference):  python -m oracle.make_golden_vae"""
import os
import sys

import torch

from magicdrive_b200 import arch
from oracle import ref_shim
from oracle.make_golden import OUT


@torch.no_grad()
def main():
    R = ref_shim.load()
    cfg = arch.VaeConfig(block_out_channels=(32, 64, 64, 64))
    vae = R.AutoencoderKL(block_out_channels=list(cfg.block_out_channels), down_block_types=["DownEncoderBlock2D"] * 4,
                          up_block_types=["UpDecoderBlock2D"] * 4, latent_channels=4, layers_per_block=2)
    vae.load_state_dict(arch.synthetic_state_dict(arch.vae_decoder_param_shapes(cfg), 5), strict=False)
    z = torch.randn(2, 4, 6, 7, generator=torch.Generator().manual_seed(1))
    out = vae.decode(z).sample
    torch.save(dict(block_out_channels=cfg.block_out_channels, seed=5, z=z, sample=out), os.path.join(OUT, "vae_decode.pt"))
    print("vae_decode.pt", os.path.getsize(os.path.join(OUT, "vae_decode.pt")) // 1024, "KiB", tuple(out.shape))


if __name__ == "__main__":
    sys.exit(main())
