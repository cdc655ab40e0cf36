#!/bin/bash
# round-2 GPU session 9: chunk-level epilogue trace, L1 constant prefetch, PDL on the single-stream decoder, VAE fix
mkdir -p gpurun_out/s9
O=gpurun_out/s9
PT="-q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread"
B="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-decode"
MDB_GEMM_VARIANT=3 timeout 120 python tools/bench_gemm.py --trace --warm --only tok16800_320x > $O/trace_pair_warm.log 2>&1
timeout 300 python -m pytest tests/test_zz_vae_gpu.py tests/test_gemm_pair_gpu.py $PT 2>&1 | tail -8 > $O/pytest_vae_pair.log
timeout 300 python bench.py $B > $O/bench_full.json 2> $O/bench_full.err
MDB_PDL_DECODER=1 timeout 300 python bench.py $B > $O/bench_full_pdldec.json 2> $O/bench_full_pdldec.err
timeout 300 python bench.py $B > $O/bench_full_b.json 2> $O/bench_full_b.err
MDB_PDL_DECODER=1 timeout 300 python bench.py $B > $O/bench_full_pdldec_b.json 2> $O/bench_full_pdldec_b.err
MDB_GEMM_VARIANT=3 timeout 200 python tools/bench_gemm.py --warm > $O/warm_pair.log 2>&1
cat $O/trace_pair_warm.log | cut -c1-260; tail -n 5 $O/pytest_vae_pair.log; for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'gemm ms', d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step'])
except Exception as e: print('ERR', e)
"; done
