"""Micro-benchmark of mdb_gemm_conv on the shapes of one denoising step (CUDA events, L2 flushed between reps).

  python tools/bench_gemm.py [--variant 0|1] [--only SUBSTR] [--reps 20] [--profile]
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magicdrive_b200 import ops  # noqa: E402
from magicdrive_b200.params import pack_geglu  # noqa: E402

# (name, n_img, h, w, cin, cout, taps, residual, geglu)
SHAPES = [
    ("tok16800_320x320_res", 1, 1, 16800, 320, 320, 1, True, False),
    ("tok16800_320x320", 1, 1, 16800, 320, 320, 1, False, False),
    ("tok16800_320x960_qkv", 1, 1, 16800, 320, 960, 1, False, False),
    ("tok16800_320x2560_geglu", 1, 1, 16800, 320, 2560, 1, False, True),
    ("tok16800_1280x320_res", 1, 1, 16800, 1280, 320, 1, True, False),
    ("tok4200_640x640_res", 1, 1, 4200, 640, 640, 1, True, False),
    ("tok4200_640x1920_qkv", 1, 1, 4200, 640, 1920, 1, False, False),
    ("tok4200_640x5120_geglu", 1, 1, 4200, 640, 5120, 1, False, True),
    ("tok1092_1280x1280_res", 1, 1, 1092, 1280, 1280, 1, True, False),
    ("tok1092_1280x10240_geglu", 1, 1, 1092, 1280, 10240, 1, False, True),
    ("tok336_1280x1280", 1, 1, 336, 1280, 1280, 1, False, False),
    ("conv28x50_320x320_res", 12, 28, 50, 320, 320, 3, True, False),
    ("conv28x50_640x320", 12, 28, 50, 640, 320, 3, False, False),
    ("conv14x25_640x640_res", 12, 14, 25, 640, 640, 3, True, False),
    ("conv14x25_1280x1280", 12, 14, 25, 1280, 1280, 3, False, False),
    ("conv7x13_1280x1280_res", 12, 7, 13, 1280, 1280, 3, True, False),
    ("conv7x13_2560x1280", 12, 7, 13, 2560, 1280, 3, False, False),
    ("conv4x7_1280x1280_res", 12, 4, 7, 1280, 1280, 3, True, False),
    ("conv4x7_2560x1280", 12, 4, 7, 2560, 1280, 3, False, False),
]

ap = argparse.ArgumentParser()
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--debug-flags", type=int, default=0)
ap.add_argument("--bn", type=int, default=0)
ap.add_argument("--splits", type=int, default=0)
ap.add_argument("--trace", action="store_true", help="print per-CTA clock64 phase stamps of the persistent kernel")
ap.add_argument("--warm", action="store_true", help="no L2 flush between reps: operands L2-resident, like activations inside a step")
ap.add_argument("--inner", type=int, default=1, help="launches between one pair of events (the event clock ticks in ~2 us steps)")
ap.add_argument("--profile", action="store_true", help="one launch per shape between cudaProfilerStart/Stop (for ncu)")
args = ap.parse_args()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"# variant={args.variant}  name  us  TFLOP/s")
for name, n, h, w, ci, co, taps, res, geglu in SHAPES:
    if args.only and args.only not in name:
        continue
    pix = n * h * w
    x = torch.randn(pix, ci, device=dev, generator=g).bfloat16()
    wt = (torch.randn(co, taps * taps * ci, device=dev, generator=g) / math.sqrt(taps * taps * ci)).bfloat16()
    b = torch.randn(co, device=dev, generator=g)
    if geglu:
        wt, b = pack_geglu(wt, b)
    r = torch.randn(pix, co, device=dev, generator=g).bfloat16() if res else None
    kw = dict(n_img=n, h_in=h, w_in=w, c0=ci, lda0=ci, n_out=co, taps=taps, pad=taps // 2, bias=b, residual=r,
              ldr=co if res else 0, geglu=geglu, kernel_variant=args.variant, debug_flags=args.debug_flags,
              force_block_n=0 if geglu else args.bn, force_splits=0 if geglu else args.splits)
    for _ in range(3):
        ops.gemm_conv(x, wt, **kw)
    torch.cuda.synchronize()
    if args.trace and (args.variant or int(os.environ.get("MDB_GEMM_VARIANT", "0"))) in (3, 4):
        tr = torch.zeros(160 * 64, dtype=torch.int64, device=dev)
        if not args.warm:
            flush.zero_()
        ops.gemm_conv(x, wt, trace=tr, **kw)
        torch.cuda.synchronize()
        tr = tr.view(160, 64).cpu()
        live = [i for i in range(160) if tr[i, 0] != 0]
        t0 = min(int(tr[i, 0]) for i in live)
        names = ["entry", "setup", "tma_end", "mma_end", "aux_stores_issued", "aux_drained", "epi_end", "mma_tile0", "epi_tile0",
                 "pre_sync", "exit", "bar_init", "tmem_alloc", "cluster_sync", "first_full"]
        print(f"{name}: {len(live)} CTAs; ns since the first CTA's entry; kernel span {max(int(tr[i, 10]) for i in live) - t0} ns")
        dur = sorted(live, key=lambda i: int(tr[i, 10]))
        for cta in [live[0], live[1], dur[len(dur) // 2], dur[-2], dur[-1]]:
            print(f"   cta {cta:3d}: " + " ".join(f"{n}={int(tr[cta, i]) - t0 if tr[cta, i] else -1}" for i, n in enumerate(names)))
        # chunk-level stamps of the first epilogue warp (its first two tiles): accumulator ready, then per chunk it owns
        # (staging box ready, accumulator read, result stored, arrived)
        for cta in [live[0], dur[-1]]:
            for tile in range(2):
                row = [int(tr[cta, 16 + tile * 20 + i]) for i in range(17)]
                if row[0] == 0:
                    continue
                chunks = [row[1 + 4 * c: 5 + 4 * c] for c in range(4) if row[1 + 4 * c]]
                print(f"   cta {cta:3d} warp 3 tile {tile}: acc_full={row[0] - t0}  " +
                      "  ".join("[" + " ".join(str(v - t0) if v else "-" for v in ch) + "]" for ch in chunks))
        continue
    if args.trace:
        tr = torch.zeros(8 * 16, dtype=torch.int64, device=dev)
        flush.zero_()
        ops.gemm_conv(x, wt, trace=tr, **kw)
        torch.cuda.synchronize()
        tr = tr.view(8, 16).cpu()
        names = ["entry", "prologue_done", "tma_first", "tma_last", "mma_first_full", "mma_tile0_done", "mma_all_done",
                 "epi_tile0_ready", "epi_tile0_done", "epi_all_done", "exit"]
        print(f"{name}: per-CTA cycles since entry (CTA 0, 1, 7); tiles/CTA = {tr[0, 11].item()}")
        for cta in (0, 1, 7):
            print("   ", " ".join(f"{n}={int(tr[cta, i] - tr[cta, 0])}" for i, n in enumerate(names)),
                  "chunk_starts=" + ",".join(str(int(tr[cta, 12 + i] - tr[cta, 0])) for i in range(3)))
        continue
    if args.profile:
        torch.cuda.profiler.start()
        ops.gemm_conv(x, wt, **kw)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        continue
    tot = 0.0
    for _ in range(args.reps):
        if not args.warm:
            flush.zero_()
        else:
            torch.cuda._sleep(400000)  # ~0.2 ms spin: the launch below is enqueued before the GPU gets to it (no host-bound gap)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.inner):
            ops.gemm_conv(x, wt, **kw)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / args.reps / args.inner * 1e3
    fl = 2.0 * pix * co * taps * taps * ci
    print(f"{name:32s} {us:8.1f} {fl / us / 1e6:8.1f}")
