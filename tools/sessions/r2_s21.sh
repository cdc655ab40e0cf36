#!/bin/bash
# round-2 GPU session 21: multi-Q instantiation of the single-S attention kernel (one K/V tile, several query tiles per CTA)
mkdir -p gpurun_out/s21
O=gpurun_out/s21
PT="-q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread"
B="--steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-decode --no-hires"
timeout 300 python -m pytest tests/test_kernels_gpu.py $PT -x -k "multi_q" 2>&1 | tail -8 > $O/pytest_mq.log
cat $O/pytest_mq.log
if ! grep -q passed $O/pytest_mq.log || grep -q "failed\|Timeout\|error" $O/pytest_mq.log; then echo "multi-Q tests failed: stopping"; exit 1; fi
timeout 300 python -m pytest tests/test_kernels_gpu.py $PT -k "attention" 2>&1 | tail -4 > $O/pytest_attn.log
MDB_ATTN_MULTIQ=0 timeout 300 python bench.py $B > $O/bench_one.json 2> $O/bench_one.err
timeout 300 python bench.py $B > $O/bench_mq.json 2> $O/bench_mq.err
MDB_ATTN_MULTIQ=0 timeout 200 python tools/bench_attn.py > $O/attn_one.log 2>&1
timeout 200 python tools/bench_attn.py > $O/attn_mq.log 2>&1
timeout 600 python -m pytest tests/test_model_gpu.py $PT 2>&1 | tail -4 > $O/pytest_model.log
tail -n 3 $O/pytest_attn.log $O/pytest_model.log; paste $O/attn_one.log $O/attn_mq.log | cut -c1-260
for f in $O/bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'attn ms', d['roofline']['attention']['kernel_ms_per_step'], d['gpu_launches_per_step'])
except Exception as e: print('ERR', e)
"; done
