#!/bin/bash
# round-2 final measurement session (second part: the first one ran the whole GPU suite -- 396 passed -- but its files were larger
# than gpurun's 64 MiB return limit): default bench line, launch list, per-shape event table, DRAM traffic of the GEMM launches
# (flushed L2), ncu --set full of the production GEMM and attention kernels (summaries written here, reports kept small)
mkdir -p gpurun_out/s24
O=gpurun_out/s24
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python tools/profile_step.py --workload full --events $O/shape_times.txt > $O/events.log 2>&1
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $O/launches_step.csv python tools/profile_step.py --workload full --shape-log $O/shape_log.txt > $O/ncu_launches.log 2>&1
python tools/summarize_launches.py $O/launches_step.csv > $O/launches_step.summary.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   -k regex:gemm_ -o /tmp/traffic python tools/profile_step.py --workload full > $O/ncu_traffic.log 2>&1
python tools/traffic_from_ncu.py /tmp/traffic.ncu-rep $O/traffic_r2.json "one denoising step, full cond; default --cache-control all: every launch starts from a flushed L2" > $O/traffic.log 2>&1
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none \
   -k regex:gemm_pair --launch-skip 30 -c 16 -o $O/prof_pair python tools/profile_step.py --workload full > $O/ncu_pair.log 2>&1
python tools/ncu_summary.py $O/prof_pair.ncu-rep > $O/ncu_gemm_pair.summary.txt 2>&1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none \
   -k regex:attention_tc2 -c 3 -o $O/prof_attn python tools/profile_step.py --workload full > $O/ncu_attn.log 2>&1
python tools/ncu_summary.py $O/prof_attn.ncu-rep > $O/ncu_attention_tc2.summary.txt 2>&1
timeout 200 python tools/time_prepare.py > $O/time_prepare.txt 2>&1
du -sh $O; ls -la $O
if [ $(du -sm $O | cut -f1) -gt 55 ]; then rm -f $O/prof_pair.ncu-rep; fi
tail -3 $O/bench_default.err; head -40 $O/time_prepare.txt | cut -c1-200
