"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in r:
    v = float(row[vi].replace(",", ""))
    v = v / 1e3 if row[ui] == "ns" else (v * 1e3 if row[ui] == "ms" else v)
    name = re.sub(r"\(.*", "", row[ki])
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.3f} ms summed device time (serialised launches; cache state as captured: see profiles/README.md)")
print("#   time_us  share  launches  avg_us  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:10.1f} {100 * t / tot:5.1f}% {n:5d} {t / n:8.1f}  {k[:110]}")
