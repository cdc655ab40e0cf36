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


# ---------------------------------------------------------------- view sharding (cameras of one scene across ranks)
_PAIRS = {0: [5, 1], 1: [0, 2], 2: [1, 3], 3: [2, 4], 4: [3, 5], 5: [4, 0]}


def test_view_shard_kv_index_addresses_the_gathered_neighbours():
    from magicdrive_b200.dist import ViewShard
    n_cam, n_samples, rows = 6, 4, 3  # samples = cfg halves x scenes
    # global K/V stand-in: batch (s, g) holds the value 100*s + g in every row
    glob = torch.tensor([[100.0 * s + g] * rows for s in range(n_samples) for g in range(n_cam)])
    for world in (1, 2, 3, 6):
        shards = [ViewShard(r, world, n_cam) for r in range(world)]
        nl = n_cam // world
        local = [glob.view(n_samples, n_cam, rows)[:, r * nl:(r + 1) * nl].reshape(-1, rows) for r in range(world)]
        gathered = torch.cat(local)  # what all_gather_rows returns (rank-major)
        for r, sh in enumerate(shards):
            idx = sh.kv_index(n_samples * nl, [_PAIRS[i] for i in range(n_cam)])
            assert len(idx) == n_samples * nl
            for s in range(n_samples):
                for j in range(nl):
                    g = r * nl + j
                    for side in range(2):
                        assert gathered[idx[s * nl + j][side]][0].item() == 100.0 * s + _PAIRS[g][side]


def _view_worker(rank, world, port, q):
    from magicdrive_b200.dist import ViewShard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = synthetic_inputs(2, 6, 4, 6, n_box=3, map_hw=8, seed=7)
    sh = ViewShard(rank, world, 6)
    lat5 = torch.stack([full["latents"]] * 6, dim=1) + torch.arange(6.0).view(1, 6, 1, 1, 1)
    mine = sh.slice_views(dict(camera_param=full["camera_param"], bboxes_3d_data=full["bboxes_3d_data"], latents=lat5,
                               prompt_embeds=full["prompt_embeds"]))
    b, e = sh.views
    ok = mine["camera_param"].shape[1] == e - b and mine["bboxes_3d_data"]["masks"].shape[1] == e - b
    ok &= mine["prompt_embeds"].shape == full["prompt_embeds"].shape and torch.equal(mine["latents"], lat5[:, b:e])
    # "cross-view attention" stand-in: out(view) = x(view) + mean of the two neighbours' K rows, K = 3 * x
    S, nl = 2, e - b
    x_loc = mine["latents"].reshape(S * nl, -1)
    kv = sh.all_gather_rows((3.0 * x_loc).contiguous())
    idx = torch.tensor(sh.kv_index(S * nl, [_PAIRS[i] for i in range(6)]))
    out_loc = x_loc + 0.5 * (kv[idx[:, 0]] + kv[idx[:, 1]])
    out = sh.gather_views(out_loc.view(S, nl, -1))
    xg = lat5.reshape(S, 6, -1)
    nb = torch.tensor([_PAIRS[i] for i in range(6)])
    ref = xg + 0.5 * (3.0 * xg[:, nb[:, 0]] + 3.0 * xg[:, nb[:, 1]])
    ok &= torch.equal(out, ref)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_view_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_view_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]
