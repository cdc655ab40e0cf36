"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, checkpoint-name
compatibility with the reference, host helpers mirroring the reference's pipeline-facing methods."""
import ctypes
import json
import os
import re
from dataclasses import asdict

import pytest
import torch

from magicdrive_b200 import _lib, arch
from magicdrive_b200.models import BEVControlNetModel, UNet2DConditionModelMultiview
from tests.common import GOLDEN, tiny_configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "magicdrive_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 18
    L = _lib.lib()  # raises if the .so is missing
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/magicdrive_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    assert L.mdb_version() >= 100
    assert L.mdb_device_ok() in (0, 1)


def test_gemm_desc_layout_matches_header():
    # field order / count of the ctypes mirror vs the C struct
    src = open(os.path.join(ROOT, "include", "magicdrive_b200.h")).read()
    body = src[src.index("typedef struct {"):src.index("} mdb_gemm_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct {", "").strip()
        if not decl:
            continue
        for part in decl.split(","):
            c_fields.append(part.replace("*", " ").split()[-1])
    assert c_fields == [f[0] for f in _lib.GemmDesc._fields_]


def test_errors_are_reported_not_swallowed():
    L = _lib.lib()
    d = _lib.GemmDesc()
    rc = L.mdb_gemm_conv(ctypes.byref(d), None)
    assert rc != 0 and b"null pointer" in L.mdb_last_error()
    assert L.mdb_attention(None, 8, None, 8, None, 8, None, 8, 1, 1, 1, 1, 1, 40, None, 1, 1.0, None) != 0


@pytest.mark.parametrize("name", ["sd15", "tiny"])
def test_parameter_names_match_reference_checkpoints(name):
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))[name]
    u, c = (arch.UNetConfig(), arch.ControlNetConfig()) if name == "sd15" else tiny_configs()
    assert {k: list(v) for k, v in arch.unet_param_shapes(u).items()} == ref["unet"]
    assert {k: list(v) for k, v in arch.controlnet_param_shapes(c).items()} == ref["controlnet"]


def test_modules_state_dict_roundtrip_and_no_cpu_fallback():
    u, c = tiny_configs()
    un = UNet2DConditionModelMultiview(**asdict(u))
    cn = BEVControlNetModel(**asdict(c))
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))["tiny"]
    assert {k: list(v.shape) for k, v in un.state_dict().items()} == ref["unet"]
    assert {k: list(v.shape) for k, v in cn.state_dict().items()} == ref["controlnet"]
    sd = arch.synthetic_state_dict(arch.unet_param_shapes(u), 3)
    un.load_state_dict(sd)
    assert all(torch.equal(un.state_dict()[k], v) for k, v in sd.items())
    assert un.config.in_channels == 4 and un.dtype == torch.float32
    with pytest.raises(_lib.MdbError):
        un(torch.zeros(6, 4, 10, 13), 5, torch.zeros(6, 10, 768))  # CPU tensors: must fail loudly, never fall back
    # zero-initialised modules of the reference must NOT be zero in the synthetic weights
    for k in ("down_blocks.0.attentions.0.transformer_blocks.0.connector.weight",):
        assert sd[k].abs().sum() > 0


def test_uncond_helpers_match_reference_fixture():
    e = torch.load(os.path.join(GOLDEN, "tiny_encoders.pt"), weights_only=False)
    _, c = tiny_configs()
    cn = BEVControlNetModel(**asdict(c))
    cn.load_state_dict(arch.synthetic_state_dict(arch.controlnet_param_shapes(c), e["seed"] + 1))
    assert torch.equal(cn.uncond_cam_param([2, 6]), e["uncond_cam"])
    cam = e["camera_param"]
    boxes = {k: v.reshape(1, 6, *v.shape[1:]) for k, v in e["boxes"].items()}
    kw = cn.add_uncond_to_kwargs(camera_param=cam, bboxes_3d_data=boxes, image=torch.zeros(1, 8, 4, 4), max_len=9)
    assert kw["camera_param"].shape == (2, 6, 3, 7) and torch.equal(kw["camera_param"][1], cam[0])
    assert kw["bboxes_3d_data"]["bboxes"].shape == (2, 6, 9, 8, 3)
    assert not kw["bboxes_3d_data"]["masks"][0].any() and kw["bboxes_3d_data"]["masks"].dtype == torch.bool
    assert torch.equal(kw["bboxes_3d_data"]["classes"][1, :, :5], boxes["classes"][0])


def test_unipc_schedule_coefficients_reproduce_the_oracle():
    """pipeline.UniPCSchedule reduces corrector + history shift + predictor to per-step scalar coefficients; the
    recurrence mdb_cfg_unipc_step evaluates (include/magicdrive_b200.h), emulated here in torch, must follow the oracle's
    (= the reference scheduler's) trajectory for every order / warm-up / final-step case."""
    from magicdrive_b200.pipeline import UniPCSchedule
    from oracle import torch_oracle as O
    for n in (20, 50, 5, 3, 2, 1):
        sch, orc = UniPCSchedule(), O.UniPC()
        ts = sch.set_timesteps(n)
        assert ts == orc.set_timesteps(n).tolist() and len(sch.coefs) == len(ts)
        g = torch.Generator().manual_seed(100 + n)
        x = torch.randn(2, 4, 10, 13, generator=g)
        xo, last, m0, m1 = x.clone(), torch.zeros_like(x), torch.zeros_like(x), torch.zeros_like(x)
        for i, t in enumerate(ts):
            eps = torch.randn(2, 4, 10, 13, generator=g)
            c = torch.tensor(sch.coefs[i], dtype=torch.float32)
            assert len(c) == UniPCSchedule.ROW and (c[9] != 0) == (i > 0)
            x0 = c[0] * x + c[1] * eps
            xc = c[2] * last + c[3] * m0 + c[4] * m1 + c[5] * x0 if c[9] != 0 else x
            x, last, m1, m0 = c[6] * xc + c[7] * x0 + c[8] * m0, xc, m0, x0
            xo = orc.step(eps, t, xo)
            assert ((x - xo).abs().max() / xo.abs().max()).item() < 1e-5


def test_prepare_initialises_class_tokens_like_the_reference():
    """BEVControlNetModel.prepare (unet_addon_rawbox.py:704-705 -> bbox_embedder.py:117-136) with stub tokenizer / encoder."""
    from types import SimpleNamespace

    from magicdrive_b200 import models
    names = ["car", "truck", "bus"]

    class Tok:
        def __call__(self, texts, padding, return_tensors):
            assert padding == "do_not_pad" and return_tensors == "pt"
            return SimpleNamespace(input_ids=torch.tensor([[49406, 7 + len(texts[0]), 49407]]))

    class Enc:
        def __call__(self, ids):
            g = torch.Generator().manual_seed(int(ids[0, 1]))
            return SimpleNamespace(pooler_output=torch.randn(1, 768, generator=g))

    cfg = SimpleNamespace(dataset=SimpleNamespace(object_classes=names))
    _, ccfg = tiny_configs()
    cn = models.BEVControlNetModel(**asdict(ccfg), bbox_embedder_param=dict(n_classes=10, class_token_dim=768,
                                                                            use_text_encoder_init=True))
    cn.reset_parameters_synthetic(1)
    before = cn.bbox_embedder._class_tokens.clone()
    cn.prepare(cfg, tokenizer=Tok(), text_encoder=Enc())
    after = cn.bbox_embedder._class_tokens
    for i, n in enumerate(names):
        assert torch.equal(after[i], Enc()(Tok()([n], "do_not_pad", "pt").input_ids).pooler_output[0])
    assert torch.equal(after[len(names):], before[len(names):])
    cn2 = models.BEVControlNetModel(**asdict(ccfg))  # use_text_encoder_init absent -> untouched
    cn2.reset_parameters_synthetic(1)
    cn2.prepare(cfg, tokenizer=None, text_encoder=None)
    assert torch.equal(cn2.bbox_embedder._class_tokens, before)


def test_unsupported_configurations_raise_like_the_reference_would():
    """Bad / unbuilt configurations are Python exceptions at construction or call time, never a silent different result
    (errors = ValueError for bad config, unet_2d_condition_multiview.py:413-416)."""
    from magicdrive_b200 import arch, models
    from magicdrive_b200.dist import ShardPlan
    from magicdrive_b200.pipeline import BEVControlNetDenoiser, UniPCSchedule
    ucfg, ccfg = tiny_configs()
    with pytest.raises(ValueError):
        models.BEVControlNetModel(**asdict(ccfg), bbox_embedder_param=dict(mode="owhr"))
    with pytest.raises(ValueError):
        models.BEVControlNetModel(**asdict(ccfg), map_embedder_cls="my.module.Cls")
    with pytest.raises(ValueError):
        models.AutoencoderKL(act_fn="gelu")
    with pytest.raises(ValueError):
        ShardPlan(0, 7, 6, False, [arch.DEFAULT_NEIGHBORS[i] for i in range(6)])  # 6 cameras do not spread over 7 ranks
    with pytest.raises(ValueError):
        UniPCSchedule(solver_order=3)
    un, cn = models.UNet2DConditionModelMultiview(**asdict(ucfg)), models.BEVControlNetModel(**asdict(ccfg))
    with pytest.raises(ValueError):
        BEVControlNetDenoiser(un, cn, scheduler="pndm")
    # CPU tensors: no fallback, a clear error instead
    with pytest.raises(Exception) as ei:
        un(torch.zeros(6, 4, 10, 13), 10, encoder_hidden_states=torch.zeros(6, 78, 768))
    assert "CUDA" in str(ei.value)


@pytest.mark.parametrize("fmt", ["safetensors", "bin"])
def test_from_pretrained_reads_diffusers_save_pretrained_directories(tmp_path, fmt):
    """config.json + diffusion_pytorch_model.{safetensors,bin} as written by ModelMixin.save_pretrained
    (multiview_runner.py:233-242): constructor kwargs incl. diffusers' private '_class_name' keys, dtype cast."""
    import json

    from magicdrive_b200 import arch, models
    ucfg, ccfg = tiny_configs()
    vcfg = arch.VaeConfig(block_out_channels=(64, 64, 64, 64))
    cases = [("unet", models.UNet2DConditionModelMultiview, asdict(ucfg), arch.unet_param_shapes(ucfg)),
             ("controlnet", models.BEVControlNetModel, asdict(ccfg), arch.controlnet_param_shapes(ccfg)),
             ("vae", models.AutoencoderKL, asdict(vcfg), arch.vae_decoder_param_shapes(vcfg))]
    for sub, cls, cfg, shapes in cases:
        d = tmp_path / sub
        d.mkdir()
        sd = arch.synthetic_state_dict(shapes, 3)
        (d / "config.json").write_text(json.dumps({"_class_name": cls.__name__, "_diffusers_version": "0.17.1",
                                                   **{k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}}))
        if fmt == "safetensors":
            from safetensors.torch import save_file
            save_file({k: v.contiguous() for k, v in sd.items()}, str(d / "diffusion_pytorch_model.safetensors"))
        else:
            torch.save(sd, d / "diffusion_pytorch_model.bin")
        m = cls.from_pretrained(str(tmp_path), subfolder=sub, torch_dtype=torch.bfloat16)
        got = m.state_dict()
        assert set(got) == set(sd) and m.dtype == torch.bfloat16
        for k in sd:
            if sd[k].is_floating_point():
                assert torch.equal(got[k], sd[k].to(torch.bfloat16)), k
