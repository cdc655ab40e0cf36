#!/usr/bin/env python
"""Parity of the sharded low-latency mode (one scene's guidance halves x camera views spread over the GPUs, neighbour K/V and
the partner half's noise exchanged through NVLink peer memory) against the unsharded path on the same inputs.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/check_view_shard.py
Prints one line per rank; exits non-zero when the sharded result drifts from the single-GPU one."""
import os
import sys
from dataclasses import asdict

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from magicdrive_b200 import arch  # noqa: E402
from magicdrive_b200.dist import ShardContext, shutdown  # noqa: E402
from magicdrive_b200.models import BEVControlNetModel, UNet2DConditionModelMultiview  # noqa: E402
from magicdrive_b200.pipeline import BEVControlNetDenoiser  # noqa: E402
from magicdrive_b200.synthetic import synthetic_inputs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    steps = int(os.environ.get("MDB_CHECK_STEPS", "3"))
    un = UNet2DConditionModelMultiview(**asdict(arch.UNetConfig())).reset_parameters_synthetic(11).to(dev, torch.bfloat16)
    cn = BEVControlNetModel(**asdict(arch.ControlNetConfig())).reset_parameters_synthetic(12).to(dev, torch.bfloat16)
    inp = synthetic_inputs(2, 6, 28, 50, n_box=20, map_hw=200, seed=3)
    kw = dict(image=inp["bev_map"], camera_param=inp["camera_param"], prompt_embeds=inp["prompt_embeds"],
              negative_prompt_embeds=inp["negative_prompt_embeds"], latents=inp["latents"], num_inference_steps=steps,
              guidance_scale=2.0, bev_controlnet_kwargs={"bboxes_3d_data": inp["bboxes_3d_data"]})
    ref = BEVControlNetDenoiser(un, cn, use_cuda_graph=False)(**kw)
    rc = 0
    dens = []
    ctx = ShardContext(6, True, [arch.DEFAULT_NEIGHBORS[i] for i in range(6)], dev)
    print(f"[view-shard] rank {rank}: half {ctx.plan.half} views {ctx.plan.views} partner {ctx.plan.partner}", flush=True)
    for graph in (False, True):
        den = BEVControlNetDenoiser(un, cn, use_cuda_graph=graph, view_shard=ctx)
        dens.append(den)
        out = den(**kw)
        torch.cuda.synchronize()
        den.check_peers()
        err = ((out - ref).norm() / ref.norm()).item()
        ok = out.shape == ref.shape and err < 2e-2
        print(f"[view-shard] rank {rank}/{world} graph={graph} rel-L2 vs unsharded {err:.3e} {'OK' if ok else 'FAIL'}", flush=True)
        rc |= 0 if ok else 1
    un.set_view_shard(None)
    sys.stdout.flush()
    if rc:
        os._exit(rc)
    shutdown(dens)
    return rc


if __name__ == "__main__":
    sys.exit(main())
