"""One steady-state denoising step (eager launches) bracketed by cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel \
      --launch-skip 60 -c 3 -o gpurun_out/prof_gemm python tools/profile_step.py
"""
import argparse
import os
import sys
from dataclasses import asdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magicdrive_b200 import arch  # noqa: E402
from magicdrive_b200.models import BEVControlNetModel, UNet2DConditionModelMultiview  # noqa: E402
from magicdrive_b200.pipeline import BEVControlNetDenoiser  # noqa: E402
from magicdrive_b200.synthetic import synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cam")
ap.add_argument("--scenes", type=int, default=1)
ap.add_argument("--shape-log", default=None, help="write the ordered list of tensor-core launches (kind, shape) of the profiled step")
ap.add_argument("--events", default=None, help="write a per-launch CUDA-event table of the tensor-core kernels here")
args = ap.parse_args()
dev = torch.device("cuda", 0)
un = UNet2DConditionModelMultiview(**asdict(arch.UNetConfig())).reset_parameters_synthetic(11).to(dev, torch.bfloat16)
cn = BEVControlNetModel(**asdict(arch.ControlNetConfig())).reset_parameters_synthetic(12).to(dev, torch.bfloat16)
pipe = BEVControlNetDenoiser(un, cn, use_cuda_graph=False, overlap_controlnet=False)
inp = synthetic_inputs(args.scenes, 6, 28, 50, n_box=20 if args.workload == "full" else 0, map_hw=200, seed=0)
if args.workload == "cam":
    inp["bev_map"] = torch.zeros_like(inp["bev_map"])
st = pipe.prepare(inp["latents"], inp["prompt_embeds"], inp["negative_prompt_embeds"], inp["camera_param"],
                  inp["bboxes_3d_data"], inp["bev_map"], guidance_scale=2.0)
pipe.set_schedule(st, 50)
pipe.run_steps(st, 0, 2)
torch.cuda.synchronize()
if args.events:
    from magicdrive_b200 import ops
    import collections
    torch.cuda.synchronize()
    torch.cuda._sleep(int(50e6))  # enqueue the whole step behind a spin so events see no host gaps
    ops.start_profile()
    pipe.run_steps(st, 2, 3)
    rec = ops.stop_profile(with_info=True)
    agg = collections.OrderedDict()
    for kind, fl, sec, info in rec:
        a = agg.setdefault((kind, info), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += fl
        a[2] += sec
    tot = sum(a[2] for a in agg.values())
    with open(args.events, "w") as f:
        f.write(f"# per-launch CUDA events (eager), one denoising step; total tensor-kernel time {tot * 1e3:.3f} ms\n")
        f.write("# time_us  share  launches  us/launch  TFLOP/s  kind  shape\n")
        for (kind, info), (n, fl, sec) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
            f.write(f"{sec * 1e6:9.1f} {100 * sec / tot:5.1f}% {n:4d} {sec * 1e6 / n:9.1f} {fl / sec / 1e12:8.1f}  {kind}  {info}\n")
    print("wrote", args.events)
else:
    from magicdrive_b200 import ops
    if args.shape_log:
        ops.start_profile()
    torch.cuda.profiler.start()
    pipe.run_steps(st, 2, 3)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    if args.shape_log:
        rec = ops.stop_profile(with_info=True)
        with open(args.shape_log, "w") as f:
            for kind, fl, sec, info in rec:
                f.write(f"{kind}\t{fl:.0f}\t{info}\n")
    print("profiled one step")
