#!/bin/bash
# round-2 multi-GPU session 20 (gpurun --gpus N): NVLink peer-memory plumbing, sharded-mode parity vs the unsharded path,
# one bench line with the scene-replica value and the strong_scaling sub-record (one scene over all N GPUs)
N=${1:-2}
mkdir -p gpurun_out/s20
O=gpurun_out/s20
nvidia-smi topo -m > $O/topo_$N.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $RUN --master-port 29621 tools/check_peer.py > $O/check_peer_$N.log 2>&1
echo "check_peer rc=$?" >> $O/check_peer_$N.log
tail -6 $O/check_peer_$N.log
if grep -q "FAIL\|Error\|error" $O/check_peer_$N.log; then echo "peer plumbing failed: stopping"; exit 1; fi
MDB_CHECK_STEPS=3 timeout 300 $RUN --master-port 29631 tools/check_view_shard.py > $O/check_view_shard_$N.log 2>&1
echo "check_view_shard rc=$?" >> $O/check_view_shard_$N.log
grep "view-shard\|rc=" $O/check_view_shard_$N.log | tail -12
if ! grep -q "check_view_shard rc=0" $O/check_view_shard_$N.log; then echo "sharded-mode check failed: skipping its bench"; tail -20 $O/check_view_shard_$N.log; exit 1; fi
timeout 300 $RUN --master-port 29661 bench.py --gpus $N --steps 20 --warmup 3 --strong-scaling --no-decode --no-cpu-baseline --no-gpu-reference > $O/bench_strong_$N.json 2> $O/bench_strong_$N.err
tail -2 $O/bench_strong_$N.err; python -c "
import json
d=json.loads(open('$O/bench_strong_$N.json').read().strip().splitlines()[-1]); print('replicas', d['value'], d['ms_per_step'], 'strong', d.get('strong_scaling'))"
