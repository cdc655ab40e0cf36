"""Opt-in GPU tests of kernel variants that were written but not yet measured / validated on a GPU (so the driver's
`pytest -m gpu` does not depend on them):  MDB_TEST_EXPERIMENTAL=1 python -m pytest tests/test_zzz_experimental_gpu.py -m gpu
  * gn_cluster_kernel (MDB_GN_CLUSTER=1): pixel-major GroupNorm on a thread-block cluster per image (capi_norm.cu).
  * attention_tc3_kernel (MDB_ATTN_KERNEL=tc3): persistent form of attention_tc2 (attention_tc3.cuh).
A/B timing: tools/bench_norm.py, tools/bench_attn.py tc3 tc2 tc."""
import os

import pytest
import torch
import torch.nn.functional as F

from magicdrive_b200 import ops

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MDB_TEST_EXPERIMENTAL") != "1", reason="set MDB_TEST_EXPERIMENTAL=1")]


@pytest.fixture(scope="module")
def cuda_lib():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from magicdrive_b200 import _lib
    return _lib.lib()


@pytest.mark.parametrize("c0,c1,hw,n", [(320, 0, 1400, 3), (640, 320, 350, 2), (1280, 1280, 91, 5), (64, 0, 1400, 2),
                                        (1280, 0, 28, 12), (320, 0, 1400, 12), (640, 640, 350, 12), (128, 0, 5000, 2),
                                        (1920, 0, 350, 12), (320, 0, 5300, 2)])
@pytest.mark.parametrize("silu", [False, True])
def test_groupnorm_cluster_variant(cuda_lib, monkeypatch, c0, c1, hw, n, silu):
    monkeypatch.setenv("MDB_GN_CLUSTER", "1")
    g = torch.Generator(device="cuda").manual_seed(7)
    xa = (torch.randn(n * hw, c0, device="cuda", generator=g) * 2 + 0.5).bfloat16()
    xb = torch.randn(n * hw, c1, device="cuda", generator=g).bfloat16() if c1 else None
    c = c0 + c1
    gamma = torch.randn(c, device="cuda", generator=g)
    beta = torch.randn(c, device="cuda", generator=g)
    full = xa if xb is None else torch.cat([xa, xb], 1)
    ref = F.group_norm(full.float().reshape(n, hw, c).permute(0, 2, 1), 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(n * hw, c)
    out = ops.groupnorm(xa, c0, c0, n, hw, gamma, beta, 1e-5, silu, x1=xb, c1=c1, ld1=c1)
    monkeypatch.delenv("MDB_GN_CLUSTER")
    base = ops.groupnorm(xa, c0, c0, n, hw, gamma, beta, 1e-5, silu, x1=xb, c1=c1, ld1=c1)
    assert (out.float() - ref).abs().max().item() < 0.06
    assert ((out.float() - ref).norm() / ref.norm()).item() < 8e-3
    assert (out.float() - base.float()).abs().max().item() < 0.04  # both round the same fp32 values to bf16


@pytest.mark.parametrize("d,heads", [(40, 8), (32, 2), (64, 2)])
@pytest.mark.parametrize("lq,lk", [(1400, 1400), (350, 98), (91, 91), (28, 28), (70, 130), (130, 257), (200, 40), (129, 600)])
def test_attention_persistent_variant(cuda_lib, monkeypatch, d, heads, lq, lk):
    from tests.test_kernels_gpu import test_attention
    test_attention(cuda_lib, monkeypatch, d, heads, lq, lk, "tc3")


@pytest.mark.parametrize("d,heads,lq,lk", [(40, 8, 300, 700), (40, 8, 1400, 1400)])
def test_attention_persistent_variant_growing_scores(cuda_lib, monkeypatch, d, heads, lq, lk):
    from tests.test_kernels_gpu import test_attention_growing_scores
    test_attention_growing_scores(cuda_lib, monkeypatch, d, heads, lq, lk, "tc3")


@pytest.mark.parametrize("l,heads,d", [(1400, 8, 40), (130, 2, 32), (130, 2, 64), (35, 2, 64)])
def test_attention_persistent_variant_two_sets(cuda_lib, monkeypatch, l, heads, d):
    from tests.test_kernels_gpu import test_attention_two_sets_cross_view
    test_attention_two_sets_cross_view(cuda_lib, monkeypatch, "tc3", l, heads, d)
