#!/bin/bash
# round-2 GPU session 1: (a) phase traces / timings of the round-1 single-CTA kernel (the per-launch fixed cost),
# (b) first run of gemm_pair_kernel: single-CTA form, then CTA pairs, (c) the never-run persistent attention (tc3)
mkdir -p gpurun_out/s1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1/smi.txt
export MDB_GEMM_VARIANT=2
timeout 200 python tools/bench_gemm.py --trace > gpurun_out/s1/gemm_trace.log 2>&1
timeout 200 python tools/bench_gemm.py > gpurun_out/s1/gemm_times_tc2.log 2>&1
timeout 200 python tools/bench_gemm.py --debug-flags 2 > gpurun_out/s1/gemm_times_tc2_noepiloads.log 2>&1
timeout 200 python tools/bench_gemm.py --debug-flags 3 > gpurun_out/s1/gemm_times_tc2_noepiloads_nostore.log 2>&1
unset MDB_GEMM_VARIANT
timeout 300 python -m pytest tests/test_gemm_pair_gpu.py -v -x -m gpu -k single 2>&1 | tail -40 > gpurun_out/s1/pair_tests_single.log
timeout 300 python -m pytest tests/test_gemm_pair_gpu.py -v -x -m gpu -k pair 2>&1 | tail -40 > gpurun_out/s1/pair_tests_pair.log
MDB_GEMM_VARIANT=4 timeout 200 python tools/bench_gemm.py > gpurun_out/s1/gemm_times_single3.log 2>&1
MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py > gpurun_out/s1/gemm_times_pair.log 2>&1
MDB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zzz_experimental_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -15 > gpurun_out/s1/experimental_attn.log
timeout 120 python tools/bench_attn.py tc3 tc2 > gpurun_out/s1/bench_attn_tc3.log 2>&1
MDB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zzz_experimental_gpu.py -q -m gpu -x -k "groupnorm" 2>&1 | tail -15 > gpurun_out/s1/experimental_gn.log
timeout 120 python tools/bench_norm.py > gpurun_out/s1/bench_norm.log 2>&1
tail -n 30 gpurun_out/s1/*.log
