#!/bin/bash
# round-2 GPU session 13: non-blocking constant loads in the manager; PC-sampling profile of the pair kernel's epilogue
mkdir -p gpurun_out/s13
O=gpurun_out/s13
PT="-q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread"
B="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-decode"
timeout 300 python -m pytest tests/test_gemm_pair_gpu.py $PT 2>&1 | tail -8 > $O/pytest_pair.log
MDB_GEMM_VARIANT=3 timeout 120 python tools/bench_gemm.py --trace --warm --only tok16800_320x > $O/trace_pair_warm.log 2>&1
timeout 300 python bench.py $B > $O/bench_full.json 2> $O/bench_full.err
MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm > $O/warm_pair.log 2>&1
MDB_GEMM_VARIANT=3 timeout 600 ncu --profile-from-start off --set full --import-source on --sampling-interval 0 --clock-control none --cache-control none \
  -k regex:gemm_pair -o $O/prof_pair python tools/bench_gemm.py --profile --warm --only tok16800_320x > $O/ncu_pair.log 2>&1
timeout 600 python -m pytest tests/test_model_gpu.py $PT 2>&1 | tail -8 > $O/pytest_model.log
grep "warp 3\|^tok" $O/trace_pair_warm.log | cut -c1-200; tail -n 4 $O/pytest_*.log; tail -3 $O/ncu_pair.log; grep "tok" $O/warm_pair.log; for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'])
except Exception as e: print('ERR', e)
"; done
