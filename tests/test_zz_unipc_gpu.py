"""GPU parity of the fused guidance + UniPC step (mdb_cfg_unipc_step) and of the denoiser running the reference's
default sampler, against the oracle and the fixture produced by the reference pipeline itself
(tests/golden/tiny_pipeline_unipc.pt, oracle/make_golden_unipc.py)."""
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
