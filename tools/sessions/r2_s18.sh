#!/bin/bash
# round-2 GPU session 18: same-box A/B of the prologue (TMEM allocation beside the cluster barrier) and of the residual path
# (epilogue-warp register prefetch vs TMA boxes): s16 = old prologue + TMA residual, restma = new prologue + TMA residual,
# main = new prologue + register prefetch.  Plus the new config paths' GPU tests.
mkdir -p gpurun_out/s18
O=gpurun_out/s18
V=magicdrive_b200/lib/variants
PT="-q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread"
B="--steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-decode"
for rep in 1 2; do
for v in s16 restma main; do
  if [ $v = main ]; then L=""; else L="$V/lib$v.so"; fi
  MDB_LIB_PATH=$L MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm --only tok > $O/warm_${v}_$rep.log 2>&1
  MDB_LIB_PATH=$L timeout 300 python bench.py $B > $O/bench_${v}_$rep.json 2> $O/bench_${v}_$rep.err
done
done
timeout 600 python -m pytest tests/test_model_gpu.py $PT -k "attention_types or guess_mode" 2>&1 | tail -8 > $O/pytest_new.log
tail -n 5 $O/pytest_new.log
paste $O/warm_s16_1.log $O/warm_restma_1.log $O/warm_main_1.log | cut -c1-200
paste $O/warm_s16_2.log $O/warm_restma_2.log $O/warm_main_2.log | cut -c1-200
for f in $O/bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'])
except Exception as e: print('ERR', e)
"; done
