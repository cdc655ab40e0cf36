#!/bin/bash
# round-2 multi-GPU session (gpurun --gpus N): NVLink peer-memory plumbing, sharded-mode parity, strong-scaling bench line
N=${1:-2}
mkdir -p gpurun_out/s6
O=gpurun_out/s6
nvidia-smi topo -m > $O/topo.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $RUN --master-port 29621 tools/check_peer.py > $O/check_peer_$N.log 2>&1
echo "check_peer rc=$?" >> $O/check_peer_$N.log
tail -8 $O/check_peer_$N.log
if grep -q "FAIL\|Error\|error" $O/check_peer_$N.log; then echo "peer plumbing failed: stopping"; exit 1; fi
MDB_CHECK_STEPS=3 timeout 300 $RUN --master-port 29631 tools/check_view_shard.py > $O/check_view_shard_$N.log 2>&1
echo "check_view_shard rc=$?" >> $O/check_view_shard_$N.log
tail -12 $O/check_view_shard_$N.log
if ! grep -q "check_view_shard rc=0" $O/check_view_shard_$N.log; then echo "sharded-mode check failed: skipping its bench"; SKIP_VIEWS=1; fi
[ -z "$SKIP_VIEWS" ] && timeout 300 $RUN --master-port 29641 bench.py --gpus $N --steps 20 --warmup 3 --shard views --no-decode --no-cpu-baseline --no-gpu-reference > $O/bench_views_$N.json 2> $O/bench_views_$N.err
tail -3 $O/bench_views_$N.err; tail -c 1500 $O/bench_views_$N.json
timeout 300 $RUN --master-port 29651 bench.py --gpus $N --steps 20 --warmup 3 --no-decode --no-cpu-baseline --no-gpu-reference > $O/bench_scenes_$N.json 2> $O/bench_scenes_$N.err
tail -c 600 $O/bench_scenes_$N.json
[ -z "$SKIP_VIEWS" ] && timeout 300 $RUN --master-port 29661 bench.py --gpus $N --steps 20 --warmup 3 --strong-scaling --no-decode --no-cpu-baseline --no-gpu-reference > $O/bench_scenes_strong_$N.json 2> $O/bench_scenes_strong_$N.err
tail -2 $O/bench_scenes_strong_$N.err; python -c "
import json
d=json.loads(open('$O/bench_scenes_strong_$N.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('strong_scaling'))"
