"""TEST INFRASTRUCTURE ONLY — import the *reference itself* (read-only tree at /root/reference) on a 2026 stack.

In the build container it reads /root/reference in place; on the GPU box (no /root/reference) it reads the byte-for-byte
snapshot of the same .py files that `oracle/make_ref_snapshot.py` put under oracle/_ref (git-ignored).  It is used by
`oracle/make_golden*.py` to generate the committed fixtures under tests/golden/, by `oracle/ref_runner.py`
(`bench.py --impl reference` and the gpu_reference measurement) and by tests that are skipped when neither tree exists.  The reference's vendored diffusers 0.17.1 does not import against transformers 5.x / huggingface_hub 1.x,
so the package is registered as a namespace module and a handful of removed symbols are stubbed (SURVEY.md §8c,
Appendix C).  Nothing in the product imports this file.
"""
import importlib
import os
import sys
import types

_SNAPSHOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # oracle/make_ref_snapshot.py
REF = os.environ.get("MAGICDRIVE_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/magicdrive") else _SNAPSHOT)
SRC = os.path.join(REF, "third_party/diffusers/src/diffusers")
_loaded = None


def available() -> bool:
    return os.path.isdir(SRC)


def load():
    """Returns a namespace with the reference classes (UNet2DConditionModel, UNet2DConditionModelMultiview,
    BEVControlNetModel, StableDiffusionBEVControlNetPipeline, DDIMScheduler, AutoencoderKL)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    import huggingface_hub
    import huggingface_hub.constants as HC
    HC.hf_cache_home = os.path.expanduser("~/.cache/huggingface")  # utils/constants.py:16
    for n in ("HfFolder", "cached_download"):  # removed from hub >= 1.0
        if not hasattr(huggingface_hub, n):
            setattr(huggingface_hub, n, lambda *a, **k: None)
    import transformers.utils as TU
    if not hasattr(TU, "FLAX_WEIGHTS_NAME"):
        TU.FLAX_WEIGHTS_NAME = "flax_model.msgpack"  # pipelines/pipeline_utils.py:66

    def ns(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    pkg = ns("diffusers", SRC)
    pkg.__version__ = "0.17.1"
    importlib.import_module("diffusers.models.unet_2d_condition")  # first: breaks the loaders<->models cycle
    M = importlib.import_module("diffusers.models")
    S = importlib.import_module("diffusers.schedulers")
    ns("diffusers.pipelines", SRC + "/pipelines")
    sd = ns("diffusers.pipelines.stable_diffusion", SRC + "/pipelines/stable_diffusion")
    ns("diffusers.pipelines.controlnet", SRC + "/pipelines/controlnet")
    sd.StableDiffusionSafetyChecker = importlib.import_module(
        "diffusers.pipelines.stable_diffusion.safety_checker").StableDiffusionSafetyChecker
    from dataclasses import dataclass

    from diffusers.utils import BaseOutput

    @dataclass
    class StableDiffusionPipelineOutput(BaseOutput):
        images: object
        nsfw_content_detected: object

    sd.StableDiffusionPipelineOutput = StableDiffusionPipelineOutput
    pc = importlib.import_module("diffusers.pipelines.controlnet.pipeline_controlnet")
    pkg.StableDiffusionControlNetPipeline = pc.StableDiffusionControlNetPipeline
    pkg.UNet2DConditionModel, pkg.ModelMixin, pkg.AutoencoderKL = M.UNet2DConditionModel, M.ModelMixin, M.AutoencoderKL
    importlib.import_module("diffusers.image_processor")
    acc = types.ModuleType("accelerate")
    acc.state = types.ModuleType("accelerate.state")
    acc.utils = types.ModuleType("accelerate.utils")  # magicdrive/misc/common.py:6-8
    acc.state.AcceleratorState = object
    acc.state.is_initialized = lambda: False
    acc.utils.recursively_apply = lambda f, t, **k: f(t)
    sys.modules.update({"accelerate": acc, "accelerate.state": acc.state, "accelerate.utils": acc.utils})
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from magicdrive.networks.unet_2d_condition_multiview import UNet2DConditionModelMultiview
    from magicdrive.networks.unet_addon_rawbox import BEVControlNetModel
    from magicdrive.pipeline.pipeline_bev_controlnet import StableDiffusionBEVControlNetPipeline

    out = types.SimpleNamespace(
        models=M, schedulers=S, UNet2DConditionModel=M.UNet2DConditionModel, AutoencoderKL=M.AutoencoderKL,
        UNet2DConditionModelMultiview=UNet2DConditionModelMultiview, BEVControlNetModel=BEVControlNetModel,
        StableDiffusionBEVControlNetPipeline=StableDiffusionBEVControlNetPipeline, DDIMScheduler=S.DDIMScheduler)
    _loaded = out
    return out


def build_reference_models(ucfg, ccfg, img_size=(224, 400), **controlnet_kwargs):
    """Instantiate the reference UNet2DConditionModelMultiview + BEVControlNetModel for our config dataclasses."""
    R = load()
    base = R.UNet2DConditionModel(
        sample_size=ucfg.sample_size, in_channels=ucfg.in_channels, out_channels=ucfg.out_channels,
        down_block_types=tuple(ucfg.down_block_types), up_block_types=tuple(ucfg.up_block_types),
        block_out_channels=tuple(ucfg.block_out_channels), layers_per_block=ucfg.layers_per_block,
        cross_attention_dim=ucfg.cross_attention_dim, attention_head_dim=ucfg.attention_head_dim,
        norm_num_groups=ucfg.norm_num_groups)
    mv = R.UNet2DConditionModelMultiview.from_unet_2d_condition(
        base, neighboring_view_pair=dict(ucfg.neighboring_view_pair), img_size=list(img_size),
        neighboring_attn_type=ucfg.neighboring_attn_type)
    cn = R.BEVControlNetModel.from_unet(
        base, map_size=list(ccfg.map_size),
        conditioning_embedding_out_channels=list(ccfg.conditioning_embedding_out_channels),
        cam_embedder_param=dict(input_dims=3, num_freqs=ccfg.cam_num_freqs, include_input=True, log_sampling=True),
        bbox_embedder_cls="magicdrive.networks.bbox_embedder.ContinuousBBoxWithTextEmbedding",
        bbox_embedder_param=dict(n_classes=ccfg.bbox_n_classes, class_token_dim=ccfg.bbox_class_token_dim,
                                 trainable_class_token=False, use_text_encoder_init=False,
                                 embedder_num_freq=ccfg.bbox_num_freqs, proj_dims=list(ccfg.bbox_proj_dims),
                                 mode="all-xyz", minmax_normalize=False), **controlnet_kwargs)
    return mv.eval(), cn.eval()
