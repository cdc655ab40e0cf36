#!/bin/bash
# round-2 GPU session 19: same-box A/B of epilogue group counts (2 = main, 3, 4), staging prefetch depth (kAhead 5), PDL on the
# decoder; GPU tests of the new config paths
mkdir -p gpurun_out/s19
O=gpurun_out/s19
V=magicdrive_b200/lib/variants
PT="-q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread"
B="--steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-decode --no-hires"
for v in main epi3 epi4 ahead5; do
  if [ $v = main ]; then L=""; else L="$V/lib$v.so"; fi
  MDB_LIB_PATH=$L timeout 200 python -m pytest tests/test_gemm_pair_gpu.py $PT 2>&1 | tail -2 > $O/pytest_pair_$v.log
  MDB_LIB_PATH=$L timeout 300 python bench.py $B > $O/bench_$v.json 2> $O/bench_$v.err
done
MDB_PDL_DECODER=1 timeout 300 python bench.py $B > $O/bench_pdl.json 2> $O/bench_pdl.err
timeout 300 python bench.py $B > $O/bench_main_b.json 2> $O/bench_main_b.err
timeout 600 python -m pytest tests/test_model_gpu.py $PT 2>&1 | tail -5 > $O/pytest_model.log
tail -n 3 $O/pytest_*.log
for f in $O/bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'])
except Exception as e: print('ERR', e)
"; done
