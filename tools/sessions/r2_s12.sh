#!/bin/bash
# round-2 GPU session 12: column constants bulk-copied by the manager one tile ahead; relaxed cluster barrier in the tail
# copies and cluster sizes 8 / 12 / 16
mkdir -p gpurun_out/s12
O=gpurun_out/s12
V=magicdrive_b200/lib/variants
PT="-q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread"
B="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-decode"
timeout 300 python -m pytest tests/test_gemm_pair_gpu.py $PT 2>&1 | tail -8 > $O/pytest_pair.log
timeout 300 python -m pytest tests/test_kernels_gpu.py $PT -k "groupnorm" 2>&1 | tail -8 > $O/pytest_gn.log
MDB_GEMM_VARIANT=3 timeout 120 python tools/bench_gemm.py --trace --warm --only tok16800_320x > $O/trace_pair_warm.log 2>&1
timeout 300 python bench.py $B > $O/bench_full.json 2> $O/bench_full.err
MDB_LIB_PATH=$V/libepi3.so timeout 300 python bench.py $B > $O/bench_full_epi3.json 2> $O/bench_full_epi3.err
timeout 600 python -m pytest tests/test_model_gpu.py $PT 2>&1 | tail -8 > $O/pytest_model.log
MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm > $O/warm_pair.log 2>&1
timeout 300 python bench.py $B > $O/bench_full_b.json 2> $O/bench_full_b.err
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'])
except Exception as e: print('ERR', e)
"; done
