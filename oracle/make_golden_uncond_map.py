"""TEST INFRASTRUCTURE ONLY — tests/golden/tiny_uncond_map.pt: the reference pipeline (2 DDIM steps, CFG 2.0, tiny models)
with (a) use_zero_map_as_unconditional=True and (b) a ControlNet built with use_uncond_map="negative1" (its `uncond_map`
buffer replaces the unconditional half's BEV map), (c) bbox_max_length=9 (5 boxes padded to 9 masked slots).  Run in the build container:  python -m oracle.make_golden_uncond_map"""
import os
import sys

import torch

from magicdrive_b200 import arch
from oracle import ref_shim
from oracle.make_golden import OUT, synthetic_inputs, tiny_configs


@torch.no_grad()
def main():
    R = ref_shim.load()
    ucfg, ccfg = tiny_configs()
    inp = synthetic_inputs(1, 6, 10, 13, n_box=5, map_hw=52, seed=3)
    usd = arch.synthetic_state_dict(arch.unet_param_shapes(ucfg), 7)
    csd = arch.synthetic_state_dict(arch.controlnet_param_shapes(ccfg), 8)

    class Pipe(R.StableDiffusionBEVControlNetPipeline):
        def prepare_extra_step_kwargs(self, generator, eta):
            return {"eta": eta}

    class TextStub(torch.nn.Module):
        dtype = torch.float32

        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        @property
        def device(self):
            return self.p.device

    vae = R.AutoencoderKL(block_out_channels=[32, 64, 64, 64], down_block_types=["DownEncoderBlock2D"] * 4,
                          up_block_types=["UpDecoderBlock2D"] * 4, latent_channels=4)
    out = {}
    for name, cn_kw, call_kw in (("zero_map", {}, dict(use_zero_map_as_unconditional=True)),
                                 ("negative1", dict(use_uncond_map="negative1", drop_cond_ratio=0.25), {}),
                                 ("max_len9", {}, dict(bbox_max_length=9))):
        mv, cn = ref_shim.build_reference_models(ucfg, ccfg, **cn_kw)
        mv.load_state_dict(usd, strict=True)
        missing = cn.load_state_dict(csd, strict=False)
        assert set(missing.missing_keys) <= {"uncond_map"} and not missing.unexpected_keys, missing
        sched = R.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                                set_alpha_to_one=False, steps_offset=1)
        pipe = Pipe(vae=vae, text_encoder=TextStub(), unet=mv, controlnet=cn, scheduler=sched, tokenizer=None)
        pipe.set_progress_bar_config(disable=True)
        res = pipe(prompt=None, image=inp["bev_map"], camera_param=inp["camera_param"], height=80, width=104,
                   num_inference_steps=2, guidance_scale=2.0, latents=inp["latents"].clone(),
                   prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
                   output_type="latent", bev_controlnet_kwargs={"bboxes_3d_data": inp["bboxes_3d_data"]}, **call_kw)
        out[name] = res.images.clone()
        print(name, float(res.images.abs().mean()))
    torch.save(dict(inputs_from="tiny_pipeline.pt", steps=2, guidance=2.0, seed=7, outputs=out), os.path.join(OUT, "tiny_uncond_map.pt"))


if __name__ == "__main__":
    sys.exit(main())
