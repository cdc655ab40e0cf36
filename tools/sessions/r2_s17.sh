#!/bin/bash
# round-2 GPU session 17: residual rows read by the epilogue warps (register prefetch one chunk ahead), TMEM allocation beside the
# cluster barrier; cfg_streams A/B
mkdir -p gpurun_out/s17
O=gpurun_out/s17
PT="-q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread"
B="--steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-decode"
timeout 300 python -m pytest tests/test_gemm_pair_gpu.py $PT -x 2>&1 | tail -8 > $O/pytest_pair.log
if ! grep -q passed $O/pytest_pair.log || grep -q failed $O/pytest_pair.log; then cat $O/pytest_pair.log; echo "pair tests failed: stopping"; exit 1; fi
timeout 300 python bench.py $B > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py $B --cfg-streams > $O/bench_cfgstreams.json 2> $O/bench_cfgstreams.err
MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm > $O/warm_pair.log 2>&1
for s in tok16800_320x320_res tok4200_640x640_res; do
  MDB_GEMM_VARIANT=3 timeout 120 python tools/bench_gemm.py --trace --warm --only $s >> $O/trace_pair_warm.log 2>&1
done
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py $PT 2>&1 | tail -8 > $O/pytest_model.log
tail -n 4 $O/pytest_*.log; cat $O/trace_pair_warm.log | cut -c1-330; cat $O/warm_pair.log
for f in $O/bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'], d['clocks'])
except Exception as e: print('ERR', e)
"; done
