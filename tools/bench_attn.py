#!/usr/bin/env python
"""Device time of mdb_attention on the shapes of one denoising step, per kernel generation (MDB_ATTN_KERNEL)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magicdrive_b200 import ops  # noqa: E402

SHAPES = [  # (b, heads, lq, lk, d, sets)
    (12, 8, 1400, 1400, 40, 1), (12, 8, 1400, 1400, 40, 2), (12, 8, 1400, 78, 40, 1),
    (12, 8, 350, 350, 80, 1), (12, 8, 350, 350, 80, 2), (12, 8, 350, 78, 80, 1),
    (12, 8, 91, 91, 160, 2), (12, 8, 5300, 5300, 40, 1),
]


def trace_one(b, h, lq, lk, d, sets):
    """Phase stamps of the first CTA of attention_tc2 (mdb_attention_debug_trace): cycles per KV iteration."""
    from magicdrive_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(0)
    c = h * d
    q = torch.randn(b * lq, c, device="cuda", generator=g).bfloat16()
    k = torch.randn(b * lk, c, device="cuda", generator=g).bfloat16()
    v = torch.randn(b * lk, c, device="cuda", generator=g).bfloat16()
    run = lambda: ops.attention(q, k, v, b=b, heads=h, lq=lq, lk=lk, d=d, ldq=c, ldk=c, ldv=c, scale=d ** -0.5)
    for _ in range(3):
        run()
    tr = torch.zeros(3 * 16 * 8, dtype=torch.int64, device="cuda")
    _lib.lib().mdb_attention_debug_trace(tr.data_ptr())
    run()
    torch.cuda.synchronize()
    _lib.lib().mdb_attention_debug_trace(None)
    t = tr.view(3, 16, 8).cpu()
    t0 = int(t[0, 0, 0])
    print(f"B={b} H={h} Lq={lq} Lk={lk} D={d}: first CTA, cycles since the MMA warp's first wait")
    names = [["wait_p", "p_full", "pv_issued", "qk_issued"], ["wait_s", "s_full", "ldtm", "max", "exp", "sttm", "arrived"]]
    for it in range(min(16, (lk + 127) // 128 * sets)):
        mm = " ".join(f"{n}={int(t[0, it, i]) - t0}" for i, n in enumerate(names[0]))
        s0 = " ".join(f"{n}={int(t[1, it, i]) - t0}" for i, n in enumerate(names[1]))
        s1 = " ".join(f"{n}={int(t[2, it, i]) - t0}" for i, n in enumerate(names[1]))
        print(f"  it {it:2d} MMA: {mm}\n        sm0: {s0}\n        sm1: {s1}")


def main():
    if "--trace" in sys.argv:
        trace_one(12, 8, 1400, 1400, 40, 1)
        trace_one(12, 8, 350, 350, 80, 1)
        return
    kernels = sys.argv[1:] or ["tc2", "tc2d", "tc"]
    g = torch.Generator(device="cuda").manual_seed(0)
    nbr = {0: [5, 1], 1: [0, 2], 2: [1, 3], 3: [2, 4], 4: [3, 5], 5: [4, 0]}
    for (b, h, lq, lk, d, sets) in SHAPES:
        c = h * d
        q = torch.randn(b * lq, c, device="cuda", generator=g).bfloat16()
        k = torch.randn(b * lk, c, device="cuda", generator=g).bfloat16()
        v = torch.randn(b * lk, c, device="cuda", generator=g).bfloat16()
        idx = None
        if sets == 2:
            idx = torch.tensor([[s * 6 + nbr[i][0], s * 6 + nbr[i][1]] for s in range(b // 6) for i in range(6)],
                               dtype=torch.int32, device="cuda")
        flop = 4.0 * b * h * lq * lk * d * sets
        outs = {}
        line = f"B={b} H={h} Lq={lq} Lk={lk} D={d} sets={sets}:"
        for kern in kernels:
            os.environ["MDB_ATTN_KERNEL"] = kern
            run = lambda: ops.attention(q, k, v, b=b, heads=h, lq=lq, lk=lk, d=d, ldq=c, ldk=c, ldv=c, scale=d ** -0.5,
                                        kv_index=idx, n_sets=sets)
            for _ in range(3):
                outs[kern] = run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            line += f"  {kern} {us:7.1f} us {flop / us / 1e6:6.1f} TF/s"
        if len(kernels) > 1:
            a, bb = outs[kernels[0]].float(), outs[kernels[1]].float()
            line += f"  max|diff| {(a - bb).abs().max().item():.2e}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
