#!/usr/bin/env python
"""Device time of mdb_groupnorm on the shapes of one denoising step: production kernel vs MDB_GN_CLUSTER=1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magicdrive_b200 import ops  # noqa: E402

SHAPES = [(12, 1400, 320, 0), (12, 1400, 320, 320), (12, 1400, 640, 320), (12, 350, 640, 0), (12, 350, 640, 640),
          (12, 350, 1280, 640), (12, 91, 1280, 0), (12, 91, 1280, 1280), (12, 28, 1280, 0), (12, 28, 1280, 1280)]


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for n, hw, c0, c1 in SHAPES:
        xa = torch.randn(n * hw, c0, device="cuda", generator=g).bfloat16()
        xb = torch.randn(n * hw, c1, device="cuda", generator=g).bfloat16() if c1 else None
        gamma, beta = torch.ones(c0 + c1, device="cuda"), torch.zeros(c0 + c1, device="cuda")
        line = f"n={n} hw={hw} c={c0}+{c1}:"
        for name, env in (("fused", None), ("cluster", "1")):
            if env:
                os.environ["MDB_GN_CLUSTER"] = env
            else:
                os.environ.pop("MDB_GN_CLUSTER", None)
            run = lambda: ops.groupnorm(xa, c0, c0, n, hw, gamma, beta, 1e-5, True, x1=xb, c1=c1, ld1=c1)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 50
            line += f"  {name} {us:6.1f} us {4.0 * n * hw * (c0 + c1) / us / 1e3:6.0f} GB/s"
        print(line, flush=True)
    os.environ.pop("MDB_GN_CLUSTER", None)


if __name__ == "__main__":
    main()
