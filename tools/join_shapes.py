"""Join an ncu launch list (device time per launch, in order) with the ordered shape log of the same step
(tools/profile_step.py --shape-log) -> per-shape device time / TFLOP/s table."""
import collections
import csv
import re
import sys

launch_csv, shape_log = sys.argv[1], sys.argv[2]
lines = [l for l in open(launch_csv) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for row in r:
    v = float(row[vi].replace(",", ""))
    v = v / 1e3 if row[ui] == "ns" else (v * 1e3 if row[ui] == "ms" else v)
    launches.append((re.sub(r"\(.*", "", row[ki]), v))
shapes = [l.rstrip("\n").split("\t") for l in open(shape_log)]
tc = [(n, t) for n, t in launches if "gemm_tc" in n or "gemm_pair" in n or "attention" in n or "splitk_finalize" in n]
agg = collections.OrderedDict()
i = 0
for kind, fl, info in shapes:
    name, t = tc[i]
    i += 1
    if kind == "gemm_conv" and i < len(tc) and "splitk_finalize" in tc[i][0]:
        t += tc[i][1]
        i += 1
    a = agg.setdefault((kind, info), [0, 0.0, 0.0])
    a[0] += 1
    a[1] += float(fl)
    a[2] += t
assert i == len(tc), (i, len(tc))
tot = sum(a[2] for a in agg.values())
print(f"# device time per tensor-core launch (ncu, warm caches), joined with the call order; total {tot / 1e3:.3f} ms")
print("# time_us  share  launches  us/launch  TFLOP/s  kind  shape")
for (kind, info), (n, fl, t) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{t:9.1f} {100 * t / tot:5.1f}% {n:4d} {t / n:9.1f} {fl / t / 1e6:8.1f}  {kind}  {info}")
