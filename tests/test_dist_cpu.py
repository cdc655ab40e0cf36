"""world_size-2 gloo test of the scene-sharding host logic (the N>1 path has no data-path collective)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from magicdrive_b200.dist import gather_scenes, max_over_ranks, shard_range, shard_scene_inputs
from magicdrive_b200.synthetic import synthetic_inputs


def test_shard_range_is_a_partition():
    for n in (1, 2, 5, 8, 13):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_scenes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = synthetic_inputs(n_scenes, 6, 4, 6, n_box=3, map_hw=8, seed=5)
    mine = shard_scene_inputs(full, rank, world)
    b, e = shard_range(n_scenes, rank, world)
    assert mine["latents"].shape[0] == e - b and mine["bboxes_3d_data"]["masks"].shape[0] == e - b
    # stand-in for the per-rank denoising: a deterministic per-scene function of the local inputs
    local = mine["latents"][:, None].expand(-1, 6, -1, -1, -1) * 2.0 + mine["camera_param"][:, :, 0, 6, None, None, None]
    out = gather_scenes(local.contiguous(), n_scenes)
    ref = full["latents"][:, None].expand(-1, 6, -1, -1, -1) * 2.0 + full["camera_param"][:, :, 0, 6, None, None, None]
    ok = torch.equal(out, ref)
    t = max_over_ranks(float(rank + 1), "cpu")
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, t))


def test_scene_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]
    assert all(r[2] == 2.0 for r in res)
