#!/usr/bin/env python
"""Numerical study (CPU) for the planned LayerNorm -> GEMM fusion: y = LN(x) W^T + b computed as
rstd_m * (x . (W*gamma)^T) - rstd_m * mu_m * s_n + c_n  with  s_n = sum_k (W*gamma)_nk,  c_n = sum_k beta_k W_nk + b_n,
row statistics from fp32 sums of the bf16 activations (E[x^2] - mu^2).  Activations are the 85 LayerNorm inputs of one
denoising step of the four-level test model run through tests/ops_emulator.py with bf16 activation rounding.
Result (printed): the folded form is closer to fp32 than LN-then-GEMM in bf16 (1.8e-3 vs 2.5e-3 rel-L2; |mean|/std of the
residual stream is 0.1 .. 0.17, so the variance formula does not cancel)."""
import sys; sys.path.insert(0,'/root/repo')
import torch
from dataclasses import asdict
from _pytest.monkeypatch import MonkeyPatch
from tests import ops_emulator, test_engine_host_cpu as T
from magicdrive_b200 import ops, arch, models
from magicdrive_b200.synthetic import synthetic_inputs
mp=MonkeyPatch(); ops_emulator.install(mp); ops_emulator.ROUND_ACTIVATIONS=True
captured=[]
orig_ln=ops.layernorm
def ln_hook(x, gamma, beta, eps=1e-5):
    captured.append((x.clone(), gamma.clone(), beta.clone()))
    return orig_ln(x, gamma, beta, eps)
mp.setattr(ops, "layernorm", ln_hook)
with torch.no_grad():
    ucfg, ccfg = T._four_level_configs()
    un, cn, usd, csd = T._modules(ucfg, ccfg, 31)
    inp = synthetic_inputs(1, 6, 28, 50, n_box=4, map_hw=200, seed=8)
    lat5 = torch.stack([inp["latents"]]*6,1); t=torch.tensor([481])
    down, mid, ctx = cn(lat5, t, inp["camera_param"], inp["bboxes_3d_data"], inp["prompt_embeds"], inp["bev_map"], return_dict=False)
    eps = un(lat5.reshape(-1,4,28,50), t[0], encoder_hidden_states=ctx, down_block_additional_residuals=down, mid_block_additional_residual=mid).sample
print("captured", len(captured))
bf=lambda t: t.to(torch.bfloat16).float()
g=torch.Generator().manual_seed(0)
worst=(0,0)
rows=[]
for x,gamma,beta in captured:
    C=x.shape[1]; N=2*C
    W=torch.randn(N,C,generator=g)/C**0.5; b=0.05*torch.randn(N,generator=g)
    mu=x.mean(1,keepdim=True); var=x.var(1,unbiased=False,keepdim=True); rstd=(var+1e-5).rsqrt()
    ln=(x-mu)*rstd*gamma+beta
    ref=ln@W.t()+b
    unf=bf(ln)@bf(W).t()+b
    # fused: stats from sums of x (bf16 values) in fp32: E[x^2]-mu^2
    s1=x.sum(1,keepdim=True); s2=(x*x).sum(1,keepdim=True)
    mu_f=s1/C; var_f=(s2/C-mu_f*mu_f).clamp_min(0); rstd_f=(var_f+1e-5).rsqrt()
    Wg=bf(W*gamma[None]); s_n=Wg.sum(1); c_n=(W*beta[None]).sum(1)+b
    fused=rstd_f*(x@Wg.t()) - rstd_f*mu_f*s_n[None] + c_n[None]
    e_u=((unf-ref).norm()/ref.norm()).item(); e_f=((fused-ref).norm()/ref.norm()).item()
    rows.append((C, x.shape[0], (mu.abs().mean()/var.sqrt().mean()).item(), e_u, e_f))
import collections
for C in sorted(set(r[0] for r in rows)):
    rr=[r for r in rows if r[0]==C]
    print(f"C={C}: sites {len(rr)}  |mean|/std {max(r[2] for r in rr):.3f}  unfused bf16 err {max(r[3] for r in rr):.2e}  fused err {max(r[4] for r in rr):.2e}")
mp.undo()
