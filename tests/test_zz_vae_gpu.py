"""GPU parity of the VAE decode of the generated views: the row-softmax kernel, and
magicdrive_b200.models.AutoencoderKL.decode / decode_latents against the oracle restatement of the reference's
AutoencoderKL.decode (oracle/torch_oracle.py: vae_decode, pinned to the reference class in tests/test_oracle_cpu.py), with
the same bf16 criterion as the denoising path: err(ours) <= 1.5 x err(reference arithmetic in bf16) + 2e-3."""
import os
import sys
from dataclasses import asdict

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from magicdrive_b200 import arch, ops  # noqa: E402
from magicdrive_b200.models import AutoencoderKL  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402  (checker only)
from tests.common import max_rel, rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def cuda_lib():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from magicdrive_b200 import _lib
    return _lib.lib()


@pytest.mark.parametrize("rows,cols,cols_out", [(130, 130, 192), (1400, 1400, 1408), (77, 5, 64)])
def test_softmax_rows(cuda_lib, rows, cols, cols_out):
    g = torch.Generator().manual_seed(rows)
    s = torch.randn(rows, cols_out + 8, generator=g) * 6
    out = ops.softmax_rows(s.to(DEV)[:, :cols_out], cols, cols_out).float().cpu()
    ref = torch.softmax(s[:, :cols], -1)
    assert out.shape == (rows, cols_out) and torch.all(out[:, cols:] == 0)
    torch.testing.assert_close(out[:, :cols], ref, atol=4e-3, rtol=8e-3)  # bf16 output


def _decode_case(cfg, n, h, w, seed):
    sd = arch.synthetic_state_dict(arch.vae_decoder_param_shapes(cfg), seed)
    vae = AutoencoderKL(**asdict(cfg))
    vae.load_state_dict(sd)
    vae = vae.to(DEV, torch.bfloat16)
    z = torch.randn(n, 4, h, w, generator=torch.Generator().manual_seed(seed + 1))
    truth = O.vae_decode({k: v.to(DEV) for k, v in sd.items()}, cfg, z.to(DEV))
    yard = O.vae_decode({k: v.to(DEV, torch.bfloat16) for k, v in sd.items()}, cfg, z.to(DEV, torch.bfloat16))
    return vae, sd, z, truth, yard


@torch.no_grad()
@pytest.mark.parametrize("name,cfg,n,h,w", [("small", arch.VaeConfig(block_out_channels=(64, 128, 128, 128)), 3, 10, 13),
                                            ("sd15", arch.VaeConfig(), 2, 28, 50)])
def test_vae_decode_vs_oracle(cuda_lib, name, cfg, n, h, w):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    vae, sd, z, truth, yard = _decode_case(cfg, n, h, w, 61)
    out = vae.decode(z.to(DEV, torch.bfloat16)).sample
    assert out.shape == truth.shape == (n, 3, 8 * h, 8 * w)
    e, ey = rel_l2(out, truth), rel_l2(yard, truth)
    print(f"[parity] vae decode {name}: rel-L2 ours {e:.3e} reference-bf16 {ey:.3e} max-rel ours {max_rel(out, truth):.3e}")
    assert e <= 1.5 * ey + 2e-3
    lat = (z * 0.18215)[None]
    imgs = vae.decode_latents(lat)
    ref = O.decode_latents({k: v.to(DEV) for k, v in sd.items()}, cfg, lat.to(DEV))
    assert imgs.shape == ref.shape and imgs.min() >= 0 and imgs.max() <= 1
    assert (imgs.float() - ref).abs().mean().item() < 2e-2
