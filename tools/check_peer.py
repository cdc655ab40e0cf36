#!/usr/bin/env python
"""Smallest possible check of the NVLink peer-memory plumbing (run before tools/check_view_shard.py):
symmetric allocation + rendezvous, the device-side barrier kernel, a peer write, and mdb_attention_multi reading K/V from the
neighbour GPU's buffer.   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
--master-port 29621 tools/check_peer.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magicdrive_b200 import ops  # noqa: E402
from magicdrive_b200.dist import PeerGroup, shutdown  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    grp = PeerGroup(list(range(world)), dist.new_group(list(range(world))), dev)
    print(f"[peer] rank {rank}: symmetric flags + rendezvous ok", flush=True)
    # 1. barrier + peer write visibility, several rounds (epochs)
    buf, hdl = grp.alloc((world, 4), torch.float32)
    buf.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    ok = True
    for it in range(5):
        for peer in range(world):
            grp.peer_view(hdl, peer, (world, 4), torch.float32)[rank].fill_(float(100 * it + rank))
        grp.barrier(0)
        torch.cuda.synchronize()
        grp.check()
        want = torch.tensor([[100.0 * it + r] * 4 for r in range(world)], device=dev)
        ok &= bool(torch.equal(buf, want))
        grp.barrier(1)
    print(f"[peer] rank {rank}: barrier + peer stores {'OK' if ok else 'FAIL'}", flush=True)
    # 2. attention with K/V read from the neighbour's buffer
    g = torch.Generator(device="cuda").manual_seed(7)  # same numbers on every rank
    b, h, L, d = 3, 8, 350, 40
    c = h * d
    q = torch.randn(b * L, c, device=dev, generator=g).bfloat16()
    kv_all = [torch.randn(b * L, 2 * c, device=dev, generator=g).bfloat16() for _ in range(world)]
    kv, khdl = grp.alloc((b * L, 2 * c), torch.bfloat16)
    kv.copy_(kv_all[rank])
    grp.barrier(0)
    peer = (rank + 1) % world
    pv = grp.peer_view(khdl, peer, (b * L, 2 * c), torch.bfloat16)
    idx = torch.tensor([[(1 << 24) | i, i] for i in range(b)], dtype=torch.int32, device=dev)  # set 0 from the peer, set 1 local
    out = ops.attention_multi(q, [(kv, kv[:, c:], 2 * c, b), (pv, pv[:, c:], 2 * c, b)], b=b, heads=h, lq=L, lk=L, d=d, ldq=c,
                              scale=d ** -0.5, kv_index=idx, n_sets=2)
    both = torch.cat([kv_all[rank], kv_all[peer]])
    idx2 = torch.tensor([[b + i, i] for i in range(b)], dtype=torch.int32, device=dev)
    ref = ops.attention(q, both, both[:, c:], b=b, b_kv=2 * b, heads=h, lq=L, lk=L, d=d, ldq=c, ldk=2 * c, ldv=2 * c,
                        scale=d ** -0.5, kv_index=idx2, n_sets=2)
    torch.cuda.synchronize()
    same = bool(torch.equal(out, ref))
    print(f"[peer] rank {rank}: attention over the neighbour's K/V {'OK' if same else 'FAIL'}", flush=True)
    grp.barrier(1)
    torch.cuda.synchronize()
    rc = 0 if (ok and same) else 1
    if rc:
        os._exit(rc)
    shutdown([])
    return rc


if __name__ == "__main__":
    sys.exit(main())
