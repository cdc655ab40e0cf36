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


# ---------------------------------------------------------------- guidance-half x view sharding of ONE scene
_PAIRS = {0: [5, 1], 1: [0, 2], 2: [1, 3], 3: [2, 4], 4: [3, 5], 5: [4, 0]}


def test_shard_plan_covers_every_sample_once_and_addresses_the_right_neighbours():
    from magicdrive_b200.dist import ShardPlan
    n_cam, S, rows = 6, 2, 3
    pairs = [_PAIRS[i] for i in range(n_cam)]
    for cfg in (True, False):
        for world in (1, 2, 3, 4, 6, 8, 12):
            try:
                plans = [ShardPlan(r, world, n_cam, cfg, pairs) for r in range(world)]
            except ValueError:
                assert (world // 2 if (cfg and world % 2 == 0) else world) > n_cam
                continue
            halves = 2 if plans[0].split_cfg else 1
            # every (half, view) is owned by exactly one rank; partners own the same views in the other half
            owned = sorted((p.half, v) for p in plans for v in range(*p.views))
            assert owned == [(h, v) for h in range(halves) for v in range(n_cam)]
            for p in plans:
                assert plans[p.partner].views == p.views and (plans[p.partner].half != p.half) == p.split_cfg
                assert max(q.n_local for q in plans) - min(q.n_local for q in plans) <= 1
            # K/V stand-in per rank of a half: batch (s, local view j) holds 100*s + global view
            n_samples = S * (1 if plans[0].split_cfg else (2 if cfg else 1))
            for p in plans:
                group = [q for q in plans if q.half == p.half]
                bufs = {q.vg: torch.tensor([[100.0 * s + q.views[0] + j] * rows for s in range(n_samples) for j in range(q.n_local)])
                        for q in group}
                srcs, idx = p.kv_sources(n_samples)
                assert srcs[0] == p.vg and len(srcs) <= 3 and len(idx) == 2 * n_samples * p.n_local
                for s in range(n_samples):
                    for j in range(p.n_local):
                        for side in range(2):
                            e = idx[(s * p.n_local + j) * 2 + side]
                            got = bufs[srcs[e >> 24]][e & 0xffffff][0].item()
                            assert got == 100.0 * s + _PAIRS[p.views[0] + j][side]


def _gather_worker(rank, world, port, q):
    from magicdrive_b200.dist import ShardContext
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = ShardContext(6, False, [_PAIRS[i] for i in range(6)], "cpu", peer_memory=False)  # 4 ranks, views 2+2+1+1
    pl = ctx.plan
    full = torch.arange(2 * 6 * 5, dtype=torch.float32).view(2, 6, 5)
    mine = pl.slice_views(dict(latents=full.view(2, 6, 5, 1, 1).expand(-1, -1, -1, 1, 1).contiguous(), camera_param=full))
    ok = torch.equal(mine["camera_param"], full[:, pl.views[0]:pl.views[1]])
    out = ctx.gather_views(mine["camera_param"].contiguous())
    ok &= torch.equal(out, full)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok), pl.views))


def test_uneven_view_gather_world4_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True] * 4
    assert [r[2] for r in res] == [(0, 2), (2, 4), (4, 5), (5, 6)]
