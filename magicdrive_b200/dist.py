"""Multi-GPU plumbing: one process per GPU, scenes sharded across ranks.

The path shards over (scene, view, cfg-half) samples; the only coupling is cross-view attention inside one scene
and cfg-half (magicdrive/networks/blocks.py:113-121).  Scene-sharding therefore needs NO data-path collective
(BASELINE.json configs[4]): every rank denoises its own scenes and the finished latents are
gathered once.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for that gather, the barrier
and the max-over-ranks timing.

`ViewShard` is the second mode (BASELINE.json north_star: "NCCL all-gather of cross-view KV"): the cameras of ONE
scene are split across ranks, for latency rather than throughput.  Everything on the path is per view except the
neighbour-view attention, so the only exchange is one all-gather of that block's K/V projections per multiview
transformer (16 per step); queries, softmax and the output stay local and address the gathered K/V through the
kernel's kv_index (include/magicdrive_b200.h: mdb_attention, b_kv > b)."""
import gc
import os
import sys
import threading
from typing import List, Optional, Sequence, Tuple

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


def shutdown(denoisers=(), timeout_s: float = 20.0, exit_code: int = 0, hard_exit_on_timeout: bool = True) -> None:
    """Tear the process group down at the end of a run.  A CUDA graph that captured NCCL collectives (view-sharded mode)
    keeps the communicator busy: release the graphs first; if the communicator still does not come down within
    `timeout_s` (observed on 2 x B200, NCCL 2.28.9: destroy_process_group never returned with a live graph), warn, flush
    and leave the process with `exit_code` (the status the caller would have returned) without running the remaining
    teardown; with hard_exit_on_timeout=False the caller gets control back instead."""
    for d in denoisers:
        d.release_graph()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if not dist.is_available() or not dist.is_initialized():
        return
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout_s)
    if t.is_alive():
        sys.stderr.write(f"[magicdrive_b200.dist] destroy_process_group still running after {timeout_s:.0f} s; "
                         + ("leaving the process without it\n" if hard_exit_on_timeout else "returning without it\n"))
        sys.stdout.flush()
        sys.stderr.flush()
        if hard_exit_on_timeout:
            os._exit(exit_code)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


class ViewShard:
    """Contiguous split of the n_cam cameras of every scene over the ranks of `group` (n_cam % world == 0).

    Local sample order on every rank is (cfg-half, scene, local view), i.e. the unsharded order with the view axis cut;
    `all_gather_rows` stacks the ranks' K/V row blocks in rank order, so the K/V batch of (sample s, global view g) sits
    at  (g // n_local) * n_samples * n_local + s * n_local + g % n_local  — what `kv_index` returns for the two ring
    neighbours of each local view (neighboring_view_pair, configs/dataset/Nuscenes.yaml:27-33)."""

    def __init__(self, rank: int, world: int, n_cam: int, group: Optional["dist.ProcessGroup"] = None):
        if n_cam % world:
            raise ValueError(f"view sharding needs n_cam ({n_cam}) divisible by the group size ({world})")
        self.rank, self.world, self.n_cam, self.group = rank, world, n_cam, group
        self.n_local = n_cam // world

    @property
    def views(self) -> Tuple[int, int]:
        return self.rank * self.n_local, (self.rank + 1) * self.n_local

    def gathered_batch(self, sample: int, view: int, n_samples: int) -> int:
        r, j = divmod(view, self.n_local)
        return r * n_samples * self.n_local + sample * self.n_local + j

    def kv_index(self, n_local_views: int, pairs: Sequence[Sequence[int]]) -> List[List[int]]:
        assert n_local_views % self.n_local == 0 and len(pairs) == self.n_cam
        n_samples = n_local_views // self.n_local
        b, _ = self.views
        return [[self.gathered_batch(s, pairs[b + j][0], n_samples), self.gathered_batch(s, pairs[b + j][1], n_samples)]
                for s in range(n_samples) for j in range(self.n_local)]

    def slice_views(self, inputs: dict) -> dict:
        """Cut the view axis (dim 1) of camera_param / bboxes_3d_data / 5-D latents; per-scene tensors pass through."""
        b, e = self.views
        n_cam = self.n_cam

        def cut(k, v):
            if isinstance(v, dict):
                return {kk: cut(kk, x) for kk, x in v.items()}
            if torch.is_tensor(v) and v.dim() >= 2 and v.shape[1] == n_cam and (k != "latents" or v.dim() == 5) \
                    and k not in ("prompt_embeds", "negative_prompt_embeds", "image", "bev_map"):
                return v[:, b:e]
            return v
        return {k: cut(k, v) for k, v in inputs.items()}

    def all_gather_rows(self, t: torch.Tensor) -> torch.Tensor:
        """[rows, cols] contiguous -> [world * rows, cols], rank-major, on the current stream."""
        out = torch.empty((self.world * t.shape[0], t.shape[1]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out

    def gather_views(self, local: torch.Tensor) -> torch.Tensor:
        """(S, n_local, ...) per rank -> (S, n_cam, ...) on every rank."""
        bufs = [torch.empty_like(local) for _ in range(self.world)]
        dist.all_gather(bufs, local.contiguous(), group=self.group)
        return torch.cat(bufs, dim=1)


class ShardPlan:
    """How ONE scene's 12 guidance x view samples are spread over `world` ranks for low latency (BASELINE.json north_star,
    SURVEY.md section 8e).  Two independent cuts:
      * guidance halves (uncond | cond): no coupling inside the networks at all; the halves meet only in the guidance
        combine of the scheduler step (pipeline_bev_controlnet.py:426-428), where each rank reads its partner's predicted
        noise.  Used whenever classifier-free guidance is on and `world` is even.
      * camera views: contiguous ranges of the ring FL, F, FR, BR, B, BL over the `groups` ranks of a half (uneven allowed,
        e.g. 2+2+1+1 on 4 ranks).  The only coupling is the neighbour-view attention (blocks.py:113-121): a rank needs the K/V
        of the view left of its first view and right of its last view, i.e. one view of each ring neighbour rank — read in
        place through NVLink peer memory (no gather, no copy).
    world 2 -> halves only (zero exchange inside the UNet); 4 -> halves x (3+3 views); 8 -> halves x (2+2+1+1 views);
    without guidance or with an odd world the ranks only split the views."""

    def __init__(self, rank: int, world: int, n_cam: int, cfg: bool, pairs: Sequence[Sequence[int]]):
        if world < 1 or not 0 <= rank < world:
            raise ValueError(f"bad rank/world {rank}/{world}")
        self.rank, self.world, self.n_cam, self.cfg = rank, world, n_cam, cfg
        self.split_cfg = bool(cfg and world % 2 == 0)
        self.groups = world // 2 if self.split_cfg else world        # ranks sharing the views of one half
        if self.groups > n_cam:
            raise ValueError(f"{world} ranks cannot share {n_cam} views" + (" (two guidance halves)" if self.split_cfg else ""))
        self.half = rank // self.groups if self.split_cfg else 0      # 0 = unconditional, 1 = conditional
        self.vg = rank % self.groups                                  # position on the view ring of this half
        self.views = shard_range(n_cam, self.vg, self.groups)         # [begin, end) global view indices
        self.n_local = self.views[1] - self.views[0]
        self.pairs = [list(p) for p in pairs]
        self.partner = (rank + self.groups) % world if self.split_cfg else rank   # same views, other guidance half
        self.half_ranks = [self.half * self.groups + g for g in range(self.groups)]

    def owner(self, view: int) -> Tuple[int, int]:
        """(rank inside the half group, local view index on that rank) of a global view."""
        for g in range(self.groups):
            b, e = shard_range(self.n_cam, g, self.groups)
            if b <= view < e:
                return g, view - b
        raise ValueError(view)

    def local_views_of(self, g: int) -> int:
        b, e = shard_range(self.n_cam, g, self.groups)
        return e - b

    def kv_sources(self, n_samples: int):
        """For local batch (sample s, local view j) and each of its two ring neighbours: which K/V buffer holds it and at
        which batch index.  Returns (sources, index): `sources` = ordered list of half-group ranks whose buffers are read
        (this rank first), `index[(s * n_local + j) * 2 + side]` = (position in `sources`) << 24 | batch index there."""
        sources = [self.vg]
        idx = []
        b0 = self.views[0]
        for s in range(n_samples):
            for j in range(self.n_local):
                for side in range(2):
                    g, lj = self.owner(self.pairs[b0 + j][side])
                    if g not in sources:
                        sources.append(g)
                    idx.append((sources.index(g) << 24) | (s * self.local_views_of(g) + lj))
        return sources, idx

    def slice_views(self, inputs: dict) -> dict:
        """Cut the view axis (dim 1) of camera_param / bboxes_3d_data / 5-D latents to this rank's views."""
        b, e = self.views
        n_cam = self.n_cam

        def cut(k, v):
            if isinstance(v, dict):
                return {kk: cut(kk, x) for kk, x in v.items()}
            if torch.is_tensor(v) and v.dim() >= 2 and v.shape[1] == n_cam and (k != "latents" or v.dim() == 5) \
                    and k not in ("prompt_embeds", "negative_prompt_embeds", "image", "bev_map"):
                return v[:, b:e]
            return v
        return {k: cut(k, v) for k, v in inputs.items()}
