#!/usr/bin/env python
"""Algorithmic bytes / FLOPs of every operator launch of one denoising step (bench.py's default workload), computed on the
host by running the real engines on the `meta` device with shape-only operator stand-ins, and joined with the measured
per-kernel device times of an ncu launch list: achieved GB/s of the memory-bound kernels against the measured HBM peak.

  python tools/algorithmic_bytes.py profiles/launches_step_r1_final.summary.txt > profiles/membound_kernels_r1.txt
"""
import collections
import json
import os
import re
import sys
from dataclasses import asdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from magicdrive_b200 import arch, models, ops  # noqa: E402
from magicdrive_b200.pipeline import BEVControlNetDenoiser  # noqa: E402
from magicdrive_b200.synthetic import synthetic_inputs  # noqa: E402

M = "meta"
REC = collections.OrderedDict()  # kernel class -> [launches, bytes, flops]
ON = [False]


def rec(kind, nbytes, flops=0.0):
    if ON[0]:
        r = REC.setdefault(kind, [0, 0.0, 0.0])
        r[0] += 1
        r[1] += nbytes
        r[2] += flops


def e(shape, dt=torch.bfloat16):
    return torch.empty(shape, dtype=dt, device=M)


def gemm_conv(a0, w, *, n_img, h_in, w_in, c0, lda0, n_out, taps=1, stride=1, pad=0, h_out=None, w_out=None, a1=None, c1=0,
              lda1=0, bias=None, rowbias=None, residual=None, ldr=0, out=None, ldo=None, out_f32=False, out_scale=1.0,
              geglu=False, **_):
    h_out = h_out or (h_in + 2 * pad - taps) // stride + 1
    w_out = w_out or (w_in + 2 * pad - taps) // stride + 1
    pix, width = n_img * h_out * w_out, n_out // 2 if geglu else n_out
    k = taps * taps * (c0 + c1)
    by = n_img * h_in * w_in * (c0 + c1) * 2 + n_out * k * 2 + pix * width * (4 if out_f32 else 2) + (pix * width * 2 if residual is not None else 0)
    rec("gemm_tc2_kernel", by, 2.0 * pix * n_out * k)
    return out if out is not None else e((pix, width), torch.float32 if out_f32 else torch.bfloat16)


def linear(x, w, bias=None, residual=None, out=None, ldo=None, geglu=False, out_f32=False, out_scale=1.0, **kw):
    m, k = x.shape
    return gemm_conv(x, w, n_img=1, h_in=1, w_in=m, c0=k, lda0=k, n_out=w.shape[0], bias=bias, residual=residual, out=out,
                     geglu=geglu, out_f32=out_f32)


def groupnorm(x0, c0, ld0, n_img, hw, gamma, beta, eps, silu, x1=None, c1=0, ld1=0, groups=32):
    rec("gn_fused_kernel", 2 * n_img * hw * (c0 + c1) * 2)
    return e((n_img * hw, c0 + c1))


def layernorm(x, gamma, beta, eps=1e-5):
    rec("layernorm_kernel", 2 * x.shape[0] * x.shape[1] * 2)
    return e(tuple(x.shape))


def attention(q, k, v, *, b, heads, lq, lk, d, ldq, ldk, ldv, scale, kv_index=None, n_sets=1, out=None, b_kv=None):
    c = heads * d
    rec("attention_tc_kernel", (2 * b * lq * c + 2 * b * n_sets * lk * c) * 2, 4.0 * b * heads * lq * lk * d * n_sets)
    return e((b * lq, c))


def add(a, b):
    rec("add_kernel", 3 * a.numel() * 2)
    return e(tuple(a.shape))


def upsample_nearest(x, n, h, w, c, ho, wo):
    rec("upsample_nearest_kernel", (n * h * w * c + n * ho * wo * c) * 2)
    return e((n * ho * wo, c))


def pack_latents(x, cpad=64, repeat=1):
    rec("pack_latents_kernel", x.numel() * 4 + repeat * x.shape[0] * cpad * 2)
    return e((repeat * x.shape[0], cpad))


def cfg_ddim_step(eps, latents, coef, cfg, guidance, c=4):
    rec("cfg_ddim_kernel", eps.shape[0] * c * 4 + 2 * latents.numel() * 4)
    return latents


def _f32(shape):
    return e(shape, torch.float32)


STANDINS = dict(
    gemm_conv=gemm_conv, linear=linear, groupnorm=groupnorm, layernorm=layernorm, attention=attention, add=add,
    upsample_nearest=upsample_nearest, pack_latents=pack_latents, cfg_ddim_step=cfg_ddim_step,
    linear_small=lambda x, w, bias=None, pre_silu=False, post_silu=False: _f32((x.shape[0], w.shape[0])),
    timestep_embedding=lambda t, dim, *a, **k: _f32((t.numel(), dim)),
    fourier_embed=lambda x, nf: _f32((x.shape[0], x.shape[1] * (1 + 2 * nf))),
    f32_to_bf16=lambda x: e(tuple(x.shape)),
    nchw_to_nhwc=lambda x: e((x.shape[0] * x.shape[2] * x.shape[3], x.shape[1])),
    conv_direct=lambda x, wgt, bias, *, n, h, w, cin, cout, k, stride=(1, 1), pad=(1, 1), silu=False, residual=None, out_f32=False:
        e((n, (h + 2 * pad[0] - k) // stride[0] + 1, (w + 2 * pad[1] - k) // stride[1] + 1, cout), torch.float32 if out_f32 else torch.bfloat16),
)


def main():
    for name, fn in STANDINS.items():
        setattr(ops, name, fn)
    models._B200Module._get_engine = lambda self, cls_: self.__dict__.setdefault("_eng", cls_(self.arch_cfg, dict(self.state_dict()), torch.device(M)))
    models._B200Module.device = property(lambda self: torch.device(M))
    with torch.device(M):
        un = models.UNet2DConditionModelMultiview(**asdict(arch.UNetConfig()))
        cn = models.BEVControlNetModel(**asdict(arch.ControlNetConfig()))
    inp = synthetic_inputs(1, 6, 28, 50, n_box=0, map_hw=200, seed=0)
    inp = {k: (v.to(M) if torch.is_tensor(v) else v) for k, v in inp.items()}
    pipe = BEVControlNetDenoiser(un, cn, use_cuda_graph=False, overlap_controlnet=False)
    st = pipe.prepare(inp["latents"], inp["prompt_embeds"], inp["negative_prompt_embeds"], inp["camera_param"], None,
                      inp["bev_map"], guidance_scale=2.0)
    pipe.scheduler.set_timesteps(50)
    st["u_temb"], st["c_temb"] = _f32((1, 32 * 1280)), _f32((1, 16 * 1280))  # stride-0 time shifts (shape-only)
    ON[0] = True
    pipe._step(st)
    ON[0] = False
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    times = {}
    if len(sys.argv) > 1:
        for line in open(sys.argv[1]):
            m = re.match(r"\s*([\d.]+)\s+[\d.]+%\s+(\d+)\s+[\d.]+\s+(?:void )?(?:<unnamed>::|mdb::)?(\w+)", line)
            if m:
                t = times.setdefault(m.group(3), [0.0, 0])
                t[0] += float(m.group(1))
                t[1] += int(m.group(2))
    print(f"# one denoising step, bench.py default workload (configs[1], CFG, V = 12): algorithmic bytes per kernel class from the")
    print(f"# real engines run shape-only on the meta device; device time from {sys.argv[1] if len(sys.argv) > 1 else '-'} (ncu, warm L2)")
    print(f"# measured HBM copy peak {peaks['hbm_gbs']:.0f} GB/s, sustained bf16 {peaks['bf16_tflops_sustained']:.0f} TFLOP/s (MEASURED_PEAKS.json)")
    print("# kernel class            launches(model/ncu)  algorithmic MB  TFLOP   time us   GB/s   of HBM peak   TFLOP/s")
    for kind, (n, by, fl) in REC.items():
        t, nn = times.get(kind, (0.0, 0))
        gbs = by / t / 1e3 if t else 0.0
        print(f"{kind:24s} {n:6d} / {nn:<6d} {by / 1e6:14.1f} {fl / 1e12:7.3f} {t:9.1f} {gbs:7.0f} {gbs / peaks['hbm_gbs']:10.2f} {fl / t / 1e6 if t else 0:12.1f}")


if __name__ == "__main__":
    main()
