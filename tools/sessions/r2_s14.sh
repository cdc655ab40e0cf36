#!/bin/bash
# round-2 GPU session 14 (after the container was re-created): state of the tree -- full GPU test suite, default bench line,
# launch list of one step, per-shape event table, ncu --set full of the production GEMM and attention kernels
mkdir -p gpurun_out/s14
O=gpurun_out/s14
PT="-q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout 1500 python -m pytest tests $PT -x 2>&1 | tail -15 > $O/pytest_gpu.log
cp profiles/parity_gpu_latest.txt $O/ 2>/dev/null
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python tools/profile_step.py --workload full --events $O/shape_times.txt > $O/events.log 2>&1
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_step.csv python tools/profile_step.py --workload full --shape-log $O/shape_log.txt > $O/ncu_launches.log 2>&1
python tools/summarize_launches.py $O/launches_step.csv > $O/launches_step.summary.txt 2>&1
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none \
   -k regex:gemm_pair --launch-skip 30 -c 24 -o $O/prof_pair python tools/profile_step.py --workload full > $O/ncu_pair.log 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none \
   -k regex:attention_tc2 -c 6 -o $O/prof_attn python tools/profile_step.py --workload full > $O/ncu_attn.log 2>&1
tail -n 6 $O/pytest_gpu.log; tail -3 $O/bench_default.err; head -30 $O/shape_times.txt; head -20 $O/launches_step.summary.txt
python -c "
import json
d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'], d.get('gpu_reference',{}).get('ms_per_step'), d.get('vae_decode'))"
