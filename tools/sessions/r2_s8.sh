#!/bin/bash
# round-2 GPU session 8: full GPU suite (per-test time-out, one process per file group), PDL A/B, epilogue groups 2 / 3,
# cluster-rows GroupNorm A/B (micro + step), epilogue trace after the instruction diet
mkdir -p gpurun_out/s8
O=gpurun_out/s8
V=magicdrive_b200/lib/variants
PT="-q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread"
B="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-decode"
timeout 400 python -m pytest tests/test_gemm_pair_gpu.py $PT 2>&1 | tail -12 > $O/pytest_pair.log
timeout 400 python -m pytest tests/test_kernels_gpu.py $PT 2>&1 | tail -12 > $O/pytest_kernels.log
timeout 900 python -m pytest tests/test_model_gpu.py $PT 2>&1 | tail -12 > $O/pytest_model.log
timeout 900 python -m pytest tests $PT --deselect tests/test_gemm_pair_gpu.py --deselect tests/test_kernels_gpu.py --deselect tests/test_model_gpu.py 2>&1 | tail -12 > $O/pytest_rest.log
tail -n 6 $O/pytest_*.log
timeout 300 python bench.py $B > $O/bench_full.json 2> $O/bench_full.err
MDB_PDL=1 timeout 300 python bench.py $B > $O/bench_full_pdl.json 2> $O/bench_full_pdl.err
MDB_LIB_PATH=$V/libepi3.so timeout 300 python bench.py $B > $O/bench_full_epi3.json 2> $O/bench_full_epi3.err
MDB_GN_ROWS=1 timeout 300 python bench.py $B > $O/bench_full_gnrows.json 2> $O/bench_full_gnrows.err
MDB_GN_ROWS=1 MDB_PDL=1 timeout 300 python bench.py $B > $O/bench_full_gnrows_pdl.json 2> $O/bench_full_gnrows_pdl.err
timeout 200 python tools/bench_norm.py > $O/bench_norm.log 2>&1
MDB_PDL=1 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_gemm_pair_gpu.py $PT 2>&1 | tail -8 > $O/pytest_pdl.log
MDB_GN_ROWS=1 timeout 600 python -m pytest tests/test_model_gpu.py $PT 2>&1 | tail -8 > $O/pytest_model_gnrows.log
MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm > $O/warm_pair.log 2>&1
MDB_GEMM_VARIANT=3 timeout 120 python tools/bench_gemm.py --trace --warm --only tok16800_320x > $O/trace_pair_warm.log 2>&1
MDB_GN_ROWS=1 timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step_gnrows.csv python tools/profile_step.py --workload full --shape-log $O/shapes.txt > $O/ncu_step.log 2>&1
tail -n 20 $O/bench_norm.log $O/pytest_pdl.log $O/pytest_model_gnrows.log; head -12 $O/trace_pair_warm.log; for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'], d['e2e'])
except Exception as e: print('ERR', e)
"; done
