#!/bin/bash
# round-2 GPU session 4: packed-math epilogue (G = 2 / 3 / 4 warp groups) correctness + timing, attention phase trace
mkdir -p gpurun_out/s4
O=gpurun_out/s4
V=magicdrive_b200/lib/variants
timeout 300 python -m pytest tests/test_gemm_pair_gpu.py -x -q -m gpu 2>&1 | tail -6 > $O/pair_tests_g4.log
for g in epi2 epi3; do
  MDB_LIB_PATH=$V/lib$g.so timeout 300 python -m pytest tests/test_gemm_pair_gpu.py -x -q -m gpu 2>&1 | tail -4 > $O/pair_tests_$g.log
done
MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm > $O/warm_pair_g4.log 2>&1
for g in epi2 epi3; do
  MDB_LIB_PATH=$V/lib$g.so MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm > $O/warm_pair_$g.log 2>&1
done
MDB_GEMM_VARIANT=2 timeout 200 python tools/bench_gemm.py --warm > $O/warm_tc2.log 2>&1
MDB_GEMM_VARIANT=4 timeout 200 python tools/bench_gemm.py --warm > $O/warm_single_g4.log 2>&1
MDB_GEMM_VARIANT=3 timeout 120 python tools/bench_gemm.py --trace --warm --only tok16800_320x > $O/trace_pair_warm.log 2>&1
timeout 120 python tools/bench_attn.py --trace > $O/attn_trace.log 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > $O/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-decode > $O/bench_full.json 2> $O/bench_full.err
for g in epi2 epi3; do
  MDB_LIB_PATH=$V/lib$g.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-decode > $O/bench_full_$g.json 2> $O/bench_full_$g.err
done
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step.csv python tools/profile_step.py --workload full --shape-log $O/shapes.txt > $O/ncu_step.log 2>&1
tail -n 8 $O/*.log; for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'])
except Exception as e: print('ERR', e)
"; done
# instruction-level profile of the pair kernel on the epilogue-bound token GEMMs (default build)
MDB_GEMM_VARIANT=3 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --cache-control none \
  -k regex:gemm_pair -o $O/prof_pair python tools/bench_gemm.py --profile --warm --only tok16800 > $O/ncu_pair.log 2>&1
tail -3 $O/ncu_pair.log
# cooperative pixel-major GroupNorm (opt-in): correctness, then the whole step with it
MDB_GN_GRID=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k groupnorm 2>&1 | tail -4 > $O/gn_grid_tests.log
MDB_GN_GRID=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-decode > $O/bench_full_gngrid.json 2> $O/bench_full_gngrid.err
MDB_GN_GRID=1 timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step_gngrid.csv python tools/profile_step.py --workload full > $O/ncu_step_gngrid.log 2>&1
tail -3 $O/gn_grid_tests.log; python -c "
import json
d=json.loads(open('$O/bench_full_gngrid.json').read().strip().splitlines()[-1]); print('gn grid step', d['ms_per_step'])"
# attention: pipelined speculative softmax (default build) vs the previous form (variant nopipe)
timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k attention 2>&1 | tail -4 > $O/attn_tests.log
timeout 120 python tools/bench_attn.py tc2 tc2d > $O/bench_attn_pipe.log 2>&1
MDB_LIB_PATH=$V/libnopipe.so timeout 120 python tools/bench_attn.py tc2 tc2d > $O/bench_attn_nopipe.log 2>&1
MDB_LIB_PATH=$V/libnopipe.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-decode > $O/bench_full_nopipe.json 2> $O/bench_full_nopipe.err
tail -3 $O/attn_tests.log; cat $O/bench_attn_pipe.log $O/bench_attn_nopipe.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_full_with_refs.json 2> $O/bench_full_with_refs.err
tail -2 $O/bench_full_with_refs.err; python -c "
import json
d=json.loads(open('$O/bench_full_with_refs.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['gpu_reference'], d['cpu_baseline'], d.get('vae_decode'), d['e2e_full_reencode'])"
