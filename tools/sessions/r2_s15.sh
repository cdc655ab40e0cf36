#!/bin/bash
# round-2 GPU session 15: ControlNet residual additions fused into the zero convolutions (A/B + parity), phase traces of the
# small-K token GEMMs (where the per-launch fixed cost goes)
mkdir -p gpurun_out/s15
O=gpurun_out/s15
PT="-q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread"
B="--steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-decode"
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_zz_sampling_gpu.py $PT 2>&1 | tail -8 > $O/pytest_model.log
timeout 300 python bench.py $B > $O/bench_fused.json 2> $O/bench_fused.err
MDB_FUSE_RESIDUAL_ADDS=0 timeout 300 python bench.py $B > $O/bench_unfused.json 2> $O/bench_unfused.err
for s in tok16800_320x320_res tok16800_320x960 tok16800_320x2560 tok4200_640x640_res tok1092_1280x1280_res conv28x50_320x320_res; do
  MDB_GEMM_VARIANT=3 timeout 120 python tools/bench_gemm.py --trace --warm --only $s >> $O/trace_pair_warm.log 2>&1
done
MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm > $O/warm_pair.log 2>&1
MDB_GEMM_VARIANT=2 timeout 200 python tools/bench_gemm.py --warm > $O/warm_single.log 2>&1
tail -n 4 $O/pytest_model.log; cat $O/trace_pair_warm.log | cut -c1-260; paste $O/warm_pair.log $O/warm_single.log
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'])
except Exception as e: print('ERR', e)
"; done
