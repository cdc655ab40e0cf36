#!/bin/bash
# round-2 GPU session 7: full GPU suite on the current code; programmatic dependent launch (trigger after the prologue) A/B;
# epilogue groups 2 (default) / 3 / 4 after the relaxed remote arrive
mkdir -p gpurun_out/s7
O=gpurun_out/s7
V=magicdrive_b200/lib/variants
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > $O/pytest_gpu.log
B="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-decode"
timeout 300 python bench.py $B > $O/bench_full.json 2> $O/bench_full.err
MDB_PDL=1 timeout 300 python bench.py $B > $O/bench_full_pdl.json 2> $O/bench_full_pdl.err
MDB_PDL=1 timeout 900 python -m pytest tests/test_model_gpu.py tests/test_gemm_pair_gpu.py -q -m gpu 2>&1 | tail -8 > $O/pytest_pdl.log
for g in epi3 epi4; do
  MDB_LIB_PATH=$V/lib$g.so timeout 300 python bench.py $B > $O/bench_full_$g.json 2> $O/bench_full_$g.err
  MDB_PDL=1 MDB_LIB_PATH=$V/lib$g.so timeout 300 python bench.py $B > $O/bench_full_${g}_pdl.json 2> $O/bench_full_${g}_pdl.err
done
MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm > $O/warm_pair.log 2>&1
MDB_PDL=1 MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm > $O/warm_pair_pdl.log 2>&1
tail -n 12 $O/pytest_gpu.log $O/pytest_pdl.log; for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'], d['e2e'])
except Exception as e: print('ERR', e)
"; done
