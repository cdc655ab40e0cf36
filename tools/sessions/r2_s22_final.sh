#!/bin/bash
# round-2 final measurement session: full GPU test suite, default bench line, launch list, per-shape event table, DRAM traffic of
# the GEMM launches (flushed L2), ncu --set full of the production GEMM and attention kernels
mkdir -p gpurun_out/s22
O=gpurun_out/s22
PT="-q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread"
rm -f profiles/parity_gpu_latest.txt gpurun_out/parity_gpu_latest.txt
timeout 1500 python -m pytest tests $PT 2>&1 | tail -15 > $O/pytest_gpu.log
cp profiles/parity_gpu_latest.txt $O/ 2>/dev/null
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python tools/profile_step.py --workload full --events $O/shape_times.txt > $O/events.log 2>&1
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_step.csv python tools/profile_step.py --workload full --shape-log $O/shape_log.txt > $O/ncu_launches.log 2>&1
python tools/summarize_launches.py $O/launches_step.csv > $O/launches_step.summary.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:gemm_ -o $O/traffic python tools/profile_step.py --workload full > $O/ncu_traffic.log 2>&1
python tools/traffic_from_ncu.py $O/traffic.ncu-rep $O/traffic_r2.json "one denoising step, full cond; default --cache-control all: every launch starts from a flushed L2" > $O/traffic.log 2>&1
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none \
   -k regex:gemm_pair --launch-skip 30 -c 24 -o $O/prof_pair python tools/profile_step.py --workload full > $O/ncu_pair.log 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none \
   -k regex:attention_tc2 -c 6 -o $O/prof_attn python tools/profile_step.py --workload full > $O/ncu_attn.log 2>&1
timeout 200 python tools/time_prepare.py > $O/time_prepare.txt 2>&1
tail -n 6 $O/pytest_gpu.log; tail -3 $O/bench_default.err; head -12 $O/shape_times.txt; head -22 $O/launches_step.summary.txt; cat $O/traffic.log | cut -c1-600
python -c "
import json
d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'], d.get('gpu_reference',{}).get('ms_per_step'), d.get('cpu_baseline',{}).get('value'), d.get('configs3_424x800'), d.get('vae_decode',{}).get('ms_per_scene'))"
