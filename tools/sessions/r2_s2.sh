#!/bin/bash
# round-2 GPU session 2: why are the no-residual K=320 GEMMs slow in CTA-pair mode (phase traces, mbarrier wait flavours),
# warm-L2 kernel comparison, the whole GPU test-suite on the new engine (pair kernel + folded LayerNorm), first bench line
mkdir -p gpurun_out/s2
O=gpurun_out/s2
V=magicdrive_b200/lib/variants
MDB_GEMM_VARIANT=3 timeout 120 python tools/bench_gemm.py --trace --only tok16800_320x3 > $O/trace_pair_cold.log 2>&1
MDB_GEMM_VARIANT=3 timeout 120 python tools/bench_gemm.py --trace --warm --only tok16800_320x > $O/trace_pair_warm.log 2>&1
MDB_GEMM_VARIANT=4 timeout 120 python tools/bench_gemm.py --trace --warm --only tok16800_320x3 > $O/trace_single_warm.log 2>&1
for w in wait1 wait2; do
  MDB_LIB_PATH=$V/lib$w.so MDB_GEMM_VARIANT=3 timeout 120 python tools/bench_gemm.py --only tok16800 > $O/pair_$w.log 2>&1
done
for v in 2 4 3; do
  MDB_GEMM_VARIANT=$v timeout 200 python tools/bench_gemm.py --warm > $O/warm_variant$v.log 2>&1
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > $O/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference > $O/bench_full.json 2> $O/bench_full.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --workload cam > $O/bench_cam.json 2> $O/bench_cam.err
MDB_GEMM_VARIANT=2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference > $O/bench_full_tc2.json 2> $O/bench_full_tc2.err
MDB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zzz_experimental_gpu.py -q -m gpu -x -k "groupnorm" 2>&1 | tail -15 > $O/experimental_gn.log
timeout 120 python tools/bench_norm.py > $O/bench_norm.log 2>&1
MDB_GN_CLUSTER=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference > $O/bench_full_gncluster.json 2> $O/bench_full_gncluster.err
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step.csv python tools/profile_step.py --workload full --shape-log $O/shapes.txt > $O/ncu_step.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_full_with_refs.json 2> $O/bench_full_with_refs.err
tail -n 12 $O/*.log; for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['gpu_launches_per_step'], d.get('gpu_reference'), d.get('cpu_baseline'))
except Exception as e: print('ERR', e)
"; done
