#!/bin/bash
# First GPU call of the next session (one B200, ~6 min): validates and measures everything that was written after the
# round-1 GPU budget ran out.  Outputs small text files under gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/first_gpu_session.sh'
mkdir -p gpurun_out
# 1. the additions that are ON by default or reachable through the public API (UniPC, given-view, VAE decode, speed ratio)
timeout 400 python -m pytest tests/test_zz_sampling_gpu.py tests/test_zz_vae_gpu.py tests/test_zz_speed_gpu.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/zz_tests.log
# 2. opt-in kernel candidates: correctness, then A/B timing
MDB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zzz_experimental_gpu.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/experimental_tests.log
timeout 120 python tools/bench_attn.py tc3 tc2 tc > gpurun_out/bench_attn_tc3.log 2>&1
timeout 120 python tools/bench_norm.py > gpurun_out/bench_norm.log 2>&1
# 3. whole step with each candidate switched on, and the new bench options
for v in "" "MDB_ATTN_KERNEL=tc3" "MDB_GN_CLUSTER=1" "MDB_ATTN_KERNEL=tc3 MDB_GN_CLUSTER=1"; do
  echo "== $v" >> gpurun_out/bench_variants.log
  env $v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'scene-steps/s', d['value'])" >> gpurun_out/bench_variants.log 2>&1
done
echo "== --cfg-streams" >> gpurun_out/bench_variants.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --cfg-streams 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'scene-steps/s', d['value'])" >> gpurun_out/bench_variants.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --scheduler unipc --decode 2>/dev/null | tail -1 > gpurun_out/bench_unipc_decode.json
tail -n +1 gpurun_out/zz_tests.log gpurun_out/experimental_tests.log gpurun_out/bench_attn_tc3.log gpurun_out/bench_norm.log gpurun_out/bench_variants.log
