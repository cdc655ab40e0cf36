#!/bin/bash
# round-2 GPU session 5: evidence captures on the final build — DRAM traffic of every GEMM launch of a step, ncu --set full of
# the production attention kernel and of the dominant GEMM instantiations, final launch list, final bench line
mkdir -p gpurun_out/s5
O=gpurun_out/s5
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > $O/pytest_gpu.log
cp gpurun_out/parity_gpu_latest.txt $O/parity_gpu.txt 2>/dev/null
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step.csv python tools/profile_step.py --workload full --shape-log $O/shapes.txt > $O/ncu_step.log 2>&1
timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_ -o $O/traffic_gemm python tools/profile_step.py --workload full > $O/ncu_traffic.log 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none -k regex:attention_tc2 -c 6 -o $O/prof_attention_tc2 python tools/profile_step.py --workload full > $O/ncu_attn.log 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none -k regex:gemm_pair --launch-skip 40 -c 40 -o $O/prof_gemm_pair python tools/profile_step.py --workload full > $O/ncu_gemm.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 > $O/bench_full.json 2> $O/bench_full.err
timeout 600 python bench.py --steps 30 --warmup 5 --workload cam --no-cpu-baseline --no-gpu-reference > $O/bench_cam.json 2> $O/bench_cam.err
timeout 600 python bench.py --steps 10 --warmup 3 --res 424x800 --no-cpu-baseline --no-gpu-reference --no-decode > $O/bench_424x800.json 2> $O/bench_424x800.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
tail -4 $O/pytest_gpu.log; for f in $O/bench_*.json; do echo $f; tail -c 400 $f; echo; done
