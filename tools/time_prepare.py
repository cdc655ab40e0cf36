#!/usr/bin/env python
"""Where the per-call conditioning cost goes (bench.py's e2e_full_reencode = prepare() + one step): wall clock and device time of
BEVControlNetDenoiser.prepare() on host inputs, and the kernels / host ops it spends them in (torch.profiler)."""
import os
import sys
import time
from dataclasses import asdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magicdrive_b200 import arch  # noqa: E402
from magicdrive_b200.models import BEVControlNetModel, UNet2DConditionModelMultiview  # noqa: E402
from magicdrive_b200.pipeline import BEVControlNetDenoiser  # noqa: E402
from magicdrive_b200.synthetic import synthetic_inputs  # noqa: E402

dev = torch.device("cuda", 0)
un = UNet2DConditionModelMultiview(**asdict(arch.UNetConfig())).reset_parameters_synthetic(11).to(dev, torch.bfloat16)
cn = BEVControlNetModel(**asdict(arch.ControlNetConfig())).reset_parameters_synthetic(12).to(dev, torch.bfloat16)
pipe = BEVControlNetDenoiser(un, cn)
inp = synthetic_inputs(1, 6, 28, 50, n_box=20, map_hw=200, seed=0)
host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in inp.items()}
host["bboxes_3d_data"] = {k: v.pin_memory() for k, v in inp["bboxes_3d_data"].items()}


def prep():
    return pipe.prepare(host["latents"], host["prompt_embeds"], host["negative_prompt_embeds"], host["camera_param"],
                        host["bboxes_3d_data"], host["bev_map"], guidance_scale=2.0)


for _ in range(3):
    st = prep()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
n = 10
for _ in range(n):
    st = prep()
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"prepare(): host enqueue {1e3 * (t1 - t0) / n:.2f} ms/call, device span {e0.elapsed_time(e1) / n:.2f} ms/call, "
      f"wall incl. sync {1e3 * (t2 - t0) / n:.2f} ms/call")
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        st = prep()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=70))
