"""The reference's 272x736 experiment (configs/exp/272x736.yaml) at full SD-1.5 size on the GPU: 34 x 92 latents (odd sizes down
the pyramid: 17 x 46, 9 x 23, 5 x 12; 3128-token level-0 attention) and the BEVControlNetConditioningEmbeddingPlus map encoder pooling a
200 x 200 BEV map to the latent grid.  Same criterion as tests/test_model_gpu.py: no worse than the reference arithmetic in bf16
against the fp32 oracle, at every tap."""
from dataclasses import asdict

import pytest
import torch

pytestmark = pytest.mark.gpu

from magicdrive_b200 import arch  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402  (checker only)
from tests.common import to_dev  # noqa: E402
from tests.test_model_gpu import DEV, _bf16_yardstick, _check, _models  # noqa: E402


@torch.no_grad()
def test_sd15_size_272x736_forward_vs_fp32_oracle(cuda_lib):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from magicdrive_b200.synthetic import synthetic_inputs
    h, w, map_hw = 34, 92, 200
    ucfg = arch.UNetConfig()
    ccfg = arch.ControlNetConfig(map_size=(8, map_hw, map_hw), map_embedding_size=(h, w))
    usd = arch.synthetic_state_dict(arch.unet_param_shapes(ucfg), 11)
    csd = arch.synthetic_state_dict(arch.controlnet_param_shapes(ccfg), 12)
    un, cn = _models(ucfg, ccfg, usd, csd, torch.bfloat16)
    inp = to_dev(synthetic_inputs(1, 6, h, w, n_box=20, map_hw=map_hw, seed=5), DEV)
    lat5 = torch.stack([inp["latents"]] * 6, 1)
    t = torch.tensor([601], device=DEV)
    down, mid, ctx = cn(lat5.bfloat16(), t, inp["camera_param"], inp["bboxes_3d_data"], inp["prompt_embeds"], inp["bev_map"],
                        return_dict=False)
    eps = un(lat5.reshape(-1, 4, h, w).bfloat16(), t[0], encoder_hidden_states=ctx, down_block_additional_residuals=down,
             mid_block_additional_residual=mid).sample
    uf = {k: v.to(DEV) for k, v in usd.items()}
    cf = {k: v.to(DEV) for k, v in csd.items()}
    d32, m32, c32 = O.controlnet_forward(cf, ccfg, lat5, t, inp["camera_param"], inp["bboxes_3d_data"], inp["prompt_embeds"],
                                         inp["bev_map"])
    e32 = O.unet_forward(uf, ucfg, lat5.reshape(-1, 4, h, w), t[0], c32, d32, m32)

    def yard(ub, cb, dt):
        l5 = lat5.to(dt)
        d, m, c = O.controlnet_forward(cb, ccfg, l5, t, inp["camera_param"].to(dt), to_dev(inp["bboxes_3d_data"], DEV, dt),
                                       inp["prompt_embeds"].to(dt), inp["bev_map"].to(dt))
        return d, m, c, O.unet_forward(ub, ucfg, l5.reshape(-1, 4, h, w), t[0], c, d, m)
    yd, ym, yc, ye = _bf16_yardstick(yard, usd, csd)
    assert eps.shape == e32.shape == (6, 4, h, w)
    _check("272x736 ctx", ctx, c32, yc)
    for i in (0, 3, 6, 9, 11):
        assert down[i].shape == d32[i].shape
        _check(f"272x736 down[{i}]", down[i], d32[i], yd[i])
    _check("272x736 mid", mid, m32, ym)
    _check("272x736 eps", eps, e32, ye)
