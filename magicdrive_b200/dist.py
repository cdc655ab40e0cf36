"""Multi-GPU plumbing: one process per GPU, scenes sharded across ranks.

The path shards over (scene, view, cfg-half) samples; the only coupling is cross-view attention inside one scene
and cfg-half (magicdrive/networks/blocks.py:113-121).  Scene-sharding therefore needs NO data-path collective
(BASELINE.json configs[4]): every rank denoises its own scenes and the finished latents are
gathered once.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for that gather, the barrier
and the max-over-ranks timing."""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of `n_items` scenes for `rank` (first n_items % world ranks get one more)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_scene_inputs(inputs: dict, rank: int, world: int) -> dict:
    """Slice every per-scene tensor (leading dim = scenes) of a pipeline input dict."""
    n = inputs["camera_param"].shape[0]
    b, e = shard_range(n, rank, world)

    def cut(v):
        if isinstance(v, dict):
            return {k: cut(x) for k, x in v.items()}
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
            return v[b:e]
        return v
    return {k: cut(v) for k, v in inputs.items()}


def gather_scenes(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather per-rank latents (scenes_local, ...) into (n_total, ...) in scene order (ragged shards allowed)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(e - b for b, e in counts)
    pad = torch.zeros((mx, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([bufs[r][: e - b] for r, (b, e) in enumerate(counts)], dim=0)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
