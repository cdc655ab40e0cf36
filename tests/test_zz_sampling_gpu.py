"""GPU parity of the sampling-side additions, against the oracle and fixtures produced by the reference pipelines
themselves: the fused guidance + UniPC step (mdb_cfg_unipc_step; the reference's default sampler;
tests/golden/tiny_pipeline_unipc.pt, oracle/make_golden_unipc.py) and the given-view pipeline (mdb_pin_views;
tests/golden/tiny_given_view.pt, oracle/make_golden_given_view.py)."""
import os
import sys
from dataclasses import asdict

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from magicdrive_b200 import ops  # noqa: E402
from magicdrive_b200.models import BEVControlNetModel, UNet2DConditionModelMultiview  # noqa: E402
from magicdrive_b200.pipeline import BEVControlNetDenoiser, UniPCSchedule  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402  (checker only)
from tests.common import golden, max_rel, rel_l2, tiny_configs, tiny_state_dicts  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def cuda_lib():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from magicdrive_b200 import _lib
    return _lib.lib()


@pytest.mark.parametrize("cfg", [False, True])
@pytest.mark.parametrize("n_steps", [20, 3, 1])
def test_cfg_unipc_step_follows_the_oracle_scheduler(cuda_lib, cfg, n_steps):
    g = torch.Generator().manual_seed(5 + n_steps)
    pix, c, ld = 6 * 10 * 13, 4, 8
    sch, orc = UniPCSchedule(), O.UniPC()
    ts = sch.set_timesteps(n_steps)
    orc.set_timesteps(n_steps)
    x = torch.randn(pix, c, generator=g)
    lat = x.to(DEV)
    hist = [torch.zeros_like(lat) for _ in range(3)]
    xo = x.clone()
    for i, t in enumerate(ts):
        eps = torch.randn((2 if cfg else 1) * pix, ld, generator=g)
        e = eps[:, :c]
        e = e[:pix] + 2.5 * (e[pix:] - e[:pix]) if cfg else e
        xo = orc.step(e, t, xo)
        coef = torch.tensor(sch.coefs[i], dtype=torch.float32, device=DEV)
        ops.cfg_unipc_step(eps.to(DEV), lat, hist[0], hist[1], hist[2], coef, cfg, 2.5, c=c)
        assert ((lat.cpu() - xo).abs().max() / xo.abs().max()).item() < 2e-5, (i, t)


@torch.no_grad()
@pytest.mark.parametrize("graph", [False, True])
def test_tiny_pipeline_unipc_vs_reference_fixture(cuda_lib, graph):
    p = golden("tiny_pipeline_unipc.pt")
    inp = golden(p["inputs_from"])["inputs"]
    ucfg, ccfg = tiny_configs()
    usd, csd = tiny_state_dicts(p["seed"])
    un = UNet2DConditionModelMultiview(**asdict(ucfg))
    cn = BEVControlNetModel(**asdict(ccfg))
    un.load_state_dict(usd)
    cn.load_state_dict(csd)
    pipe = BEVControlNetDenoiser(un.to(DEV), cn.to(DEV), use_cuda_graph=graph, scheduler="unipc")
    kw = dict(image=inp["bev_map"], camera_param=inp["camera_param"], prompt_embeds=inp["prompt_embeds"],
              negative_prompt_embeds=inp["negative_prompt_embeds"], latents=inp["latents"], num_inference_steps=p["steps"],
              guidance_scale=p["guidance"], bev_controlnet_kwargs={"bboxes_3d_data": inp["bboxes_3d_data"]})
    out = pipe(**kw)
    e = rel_l2(out, p["latents_out"])
    print(f"[parity] unipc pipeline(graph={graph}): rel-L2 {e:.3e} max-rel {max_rel(out, p['latents_out']):.3e}")
    assert out.shape == p["latents_out"].shape and e < 2e-2
    # a second call on the same denoiser must restart the multistep history (and reuse the captured graph)
    out2 = pipe(**kw)
    assert rel_l2(out2, out) < 1e-6


def test_pin_views_kernel(cuda_lib):
    g = torch.Generator().manual_seed(3)
    n_views, rows, c, ld = 12, 130, 4, 8
    dst = torch.randn(n_views * rows, ld, generator=g)
    a, b = torch.randn(n_views * rows, c, generator=g), torch.randn(n_views * rows, c, generator=g)
    mask = torch.tensor([1, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0], dtype=torch.int32)
    coef = torch.tensor([0.6, -1.7])
    for with_a in (True, False):
        ref = dst.clone()
        sel = mask.bool().repeat_interleave(rows)
        ref[sel, :c] = (coef[0] * a[sel] if with_a else 0) + coef[1] * b[sel]
        out = dst.clone().to(DEV)
        ops.pin_views(out, a.to(DEV) if with_a else None, b.to(DEV), coef.to(DEV), mask.to(DEV), rows, c=c)
        torch.testing.assert_close(out.cpu(), ref, rtol=1e-6, atol=1e-6)


@torch.no_grad()
@pytest.mark.parametrize("case,scheduler,change", [("ddim_change", "ddim", True), ("ddim_once", "ddim", False),
                                                   ("unipc_change", "unipc", True)])
def test_tiny_given_view_pipeline_vs_reference_fixture(cuda_lib, case, scheduler, change):
    from oracle.make_golden_given_view import pinned_latents
    p = golden("tiny_given_view.pt")
    inp = golden(p["inputs_from"])["inputs"]
    ucfg, ccfg = tiny_configs()
    usd, csd = tiny_state_dicts(p["seed"])
    un = UNet2DConditionModelMultiview(**asdict(ucfg))
    cn = BEVControlNetModel(**asdict(ccfg))
    un.load_state_dict(usd)
    cn.load_state_dict(csd)
    pipe = BEVControlNetDenoiser(un.to(DEV), cn.to(DEV), use_cuda_graph=True, scheduler=scheduler)
    out = pipe(image=inp["bev_map"], camera_param=inp["camera_param"], prompt_embeds=inp["prompt_embeds"],
               negative_prompt_embeds=inp["negative_prompt_embeds"], latents=inp["latents"], num_inference_steps=p["steps"],
               guidance_scale=p["guidance"], bev_controlnet_kwargs={"bboxes_3d_data": inp["bboxes_3d_data"]},
               conditional_latents=pinned_latents(p["pinned_seed"]), conditional_latents_change_every_input=change)
    ref = p["outputs"][case]
    e = rel_l2(out, ref)
    print(f"[parity] given-view {case}: rel-L2 {e:.3e} max-rel {max_rel(out, ref):.3e}")
    assert out.shape == ref.shape and e < 2e-2
