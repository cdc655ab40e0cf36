"""GroupNorm(+SiLU) micro-benchmark over the shapes of the 6-view 224x400 CFG step: (image, group) kernel vs the pixel-major
cluster kernel (MDB_GN_ROWS), CUDA events over back-to-back launches queued behind a spin kernel (so the host launch rate
does not bound the short ones).  GB/s = one read + one write of the tensor."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magicdrive_b200 import ops  # noqa: E402

SHAPES = [(12, 1400, 320, 0), (12, 1400, 320, 320), (12, 1400, 640, 320), (12, 350, 320, 0), (12, 350, 640, 0), (12, 350, 640, 320),
          (12, 350, 640, 640), (12, 350, 1280, 640), (12, 91, 640, 0), (12, 91, 1280, 0), (12, 91, 1280, 640), (12, 91, 1280, 1280),
          (12, 28, 1280, 0), (12, 28, 1280, 1280)]


def time_mode(mode, n, hw, c0, c1, iters=40, cluster=None):
    os.environ["MDB_GN_ROWS"] = mode
    if cluster:
        os.environ["MDB_GN_ROWS_CLUSTER"] = str(cluster)
    else:
        os.environ.pop("MDB_GN_ROWS_CLUSTER", None)
    g = torch.Generator(device="cuda").manual_seed(1)
    xa = torch.randn(n * hw, c0, device="cuda", generator=g).bfloat16()
    xb = torch.randn(n * hw, c1, device="cuda", generator=g).bfloat16() if c1 else None
    gamma = torch.randn(c0 + c1, device="cuda", generator=g)
    beta = torch.randn(c0 + c1, device="cuda", generator=g)
    run = lambda: ops.groupnorm(xa, c0, c0, n, hw, gamma, beta, 1e-5, True, x1=xb, c1=c1, ld1=c1)  # noqa: E731
    for _ in range(3):
        out = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters, out


def main():
    print("# n hw c0+c1 : (image, group) kernel us GB/s | cluster rows kernel us GB/s | mismatching outputs")
    for n, hw, c0, c1 in SHAPES:
        byts = 2 * n * hw * (c0 + c1) * 2
        t0, o0 = time_mode("0", n, hw, c0, c1)
        t1, o1 = time_mode("1", n, hw, c0, c1)
        t12, _ = time_mode("1", n, hw, c0, c1, cluster=12)
        t16, _ = time_mode("1", n, hw, c0, c1, cluster=16)
        print(f"n={n} hw={hw} c={c0}+{c1}:  fused {t0:7.1f} us {byts / t0 / 1e3:7.0f} GB/s   rows {t1:7.1f} us {byts / t1 / 1e3:7.0f} GB/s   "
              f"rows/12 {t12:7.1f} us  rows/16 {t16:7.1f} us   diff {(o0 != o1).float().mean().item():.2e}")


if __name__ == "__main__":
    main()
