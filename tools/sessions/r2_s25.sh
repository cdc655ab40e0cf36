#!/bin/bash
# round-2 GPU session 25: conv_direct with coalesced weights (BEV-map encoder, VAE conv_in): parity + per-call conditioning cost
mkdir -p gpurun_out/s25
O=gpurun_out/s25
PT="-q -m gpu -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_zz_vae_gpu.py tests/test_zz_sampling_gpu.py $PT 2>&1 | tail -6 > $O/pytest.log
timeout 200 python tools/time_prepare.py > $O/time_prepare.txt 2>&1
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-hires > $O/bench.json 2> $O/bench.err
cat $O/pytest.log; head -12 $O/time_prepare.txt | cut -c1-200
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'reencode', d['e2e_full_reencode']['ms_per_step'], 'vae', d['vae_decode']['ms_per_scene'], 'traffic', d['roofline']['traffic'])"
