"""TEST INFRASTRUCTURE ONLY — BASELINE.json configs[0] fixture: the reference's stock `UNet2DConditionModel` (vendored
diffusers 0.17.1, through oracle/ref_shim.py), ONE view, text-only conditioning, fp32 on the CPU, tiny two-level config.

    python -m oracle.make_golden_plain_unet      # build container
"""
import os
import sys

import torch

from magicdrive_b200 import arch
from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "plain_unet.pt")


def tiny_plain_config():
    return arch.UNetConfig(block_out_channels=(64, 128), down_block_types=("CrossAttnDownBlock2D", "DownBlock2D"),
                           up_block_types=("UpBlock2D", "CrossAttnUpBlock2D"), layers_per_block=1, attention_head_dim=2,
                           neighboring_view_pair={})


@torch.no_grad()
def main():
    R = ref_shim.load()
    cfg = tiny_plain_config()
    net = R.UNet2DConditionModel(sample_size=cfg.sample_size, in_channels=4, out_channels=4,
                                 down_block_types=tuple(cfg.down_block_types), up_block_types=tuple(cfg.up_block_types),
                                 block_out_channels=tuple(cfg.block_out_channels), layers_per_block=cfg.layers_per_block,
                                 cross_attention_dim=cfg.cross_attention_dim, attention_head_dim=cfg.attention_head_dim,
                                 norm_num_groups=cfg.norm_num_groups).eval()
    sd = arch.synthetic_state_dict(arch.unet_param_shapes(cfg), 17)
    net.load_state_dict(sd, strict=True)  # same names and shapes as the reference's own state dict
    g = torch.Generator().manual_seed(4)
    sample = torch.randn(1, 4, 10, 13, generator=g)       # one view
    text = torch.randn(1, 77, 768, generator=g)           # text-only conditioning
    out = net(sample, 481, encoder_hidden_states=text).sample
    torch.save(dict(seed=17, sample=sample, text=text, t=481, eps=out), OUT)
    print(OUT, tuple(out.shape), float(out.norm()))


if __name__ == "__main__":
    sys.exit(main())
