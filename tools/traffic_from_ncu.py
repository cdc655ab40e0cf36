"""profiles/traffic_rN.json from an ncu capture (-k regex:gemm_, default --cache-control all so every replay
pass starts from a flushed L2) of one denoising step: average DRAM bytes (read + write) and duration per launch of the
dominant kernel.  usage: traffic_from_ncu.py report.ncu-rep out.json ["note"]"""
import csv
import json
import subprocess
import sys

rep, out_json = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    i = hdr.index(name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3, "%": 1.0}[units[i]]
    return [float(r[i].replace(",", "")) * scale for r in data]


rd, wr, dur = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
n = len(data)
res = {"kernel": "gemm_pair_kernel + gemm_tc2_kernel (tcgen05 GEMM / implicit-GEMM conv, all launches of one denoising step)", "launches_profiled": n,
       "dram_bytes_per_launch": (sum(rd) + sum(wr)) / n, "dram_read_bytes_profiled": sum(rd), "dram_write_bytes_profiled": sum(wr),
       "avg_us_per_launch_under_ncu": sum(dur) / n, "source": rep,
       "note": sys.argv[3] if len(sys.argv) > 3 else ""}
if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in hdr:
    tp = col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    res["tensor_pipe_active_pct_time_weighted"] = sum(t * d for t, d in zip(tp, dur)) / sum(dur)
json.dump(res, open(out_json, "w"), indent=1)
print(json.dumps(res))
