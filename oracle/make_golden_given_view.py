"""TEST INFRASTRUCTURE ONLY — fixtures of the given-view pipeline produced by the REFERENCE ITSELF
(StableDiffusionBEVControlNetGivenViewPipeline.__call__, magicdrive/pipeline/pipeline_bev_controlnet_given_view.py, via
oracle/ref_shim.py): tiny models, views 0 and 3 of the scene pinned to seeded clean latents, 3 steps, CFG 2.0;
both `conditional_latents_change_every_input` modes with DDIM (eta 0) and the default mode with UniPC.
Run in the build container (needs /root/reference):  python -m oracle.make_golden_given_view"""
import importlib
import os
import sys

import torch

from oracle import ref_shim
from oracle.make_golden import OUT, load_ref, synthetic_inputs, tiny_configs

PINNED = (0, 3)


def pinned_latents(seed=77, h=10, w=13, n_cam=6):
    g = torch.Generator().manual_seed(seed)
    return [[torch.randn(4, h, w, generator=g) * 0.8 if j in PINNED else None for j in range(n_cam)]]


@torch.no_grad()
def main():
    R = ref_shim.load()
    GV = importlib.import_module("magicdrive.pipeline.pipeline_bev_controlnet_given_view").StableDiffusionBEVControlNetGivenViewPipeline
    UniPC = importlib.import_module("diffusers.schedulers.scheduling_unipc_multistep").UniPCMultistepScheduler
    ucfg, ccfg = tiny_configs()
    mv, cn, _, _ = load_ref(ucfg, ccfg, seed=7)
    inp = synthetic_inputs(1, 6, 10, 13, n_box=5, map_hw=52, seed=3)

    class PipeDDIM(GV):
        def prepare_extra_step_kwargs(self, generator, eta):  # DDIM eta = 0 is deterministic
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
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    out = {}
    for name, cls, sched, change in (
            ("ddim_change", PipeDDIM, R.DDIMScheduler(clip_sample=False, set_alpha_to_one=False, steps_offset=1, **kw), True),
            ("ddim_once", PipeDDIM, R.DDIMScheduler(clip_sample=False, set_alpha_to_one=False, steps_offset=1, **kw), False),
            ("unipc_change", GV, UniPC(**kw), True)):
        pipe = cls(vae=vae, text_encoder=TextStub(), unet=mv, controlnet=cn, scheduler=sched, tokenizer=None)
        pipe.set_progress_bar_config(disable=True)
        res = pipe(prompt=None, image=inp["bev_map"], camera_param=inp["camera_param"], height=80, width=104,
                   conditional_latents=pinned_latents(), conditional_latents_change_every_input=change,
                   num_inference_steps=3, guidance_scale=2.0, latents=inp["latents"].clone(),
                   prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
                   output_type="latent", bev_controlnet_kwargs={"bboxes_3d_data": inp["bboxes_3d_data"]})
        out[name] = res.images.clone()
        print(name, tuple(res.images.shape), float(res.images.abs().mean()))
    torch.save(dict(inputs_from="tiny_pipeline.pt", steps=3, guidance=2.0, seed=7, pinned_seed=77, outputs=out),
               os.path.join(OUT, "tiny_given_view.pt"))
    print("tiny_given_view.pt", os.path.getsize(os.path.join(OUT, "tiny_given_view.pt")) // 1024, "KiB")


if __name__ == "__main__":
    sys.exit(main())
