"""TEST INFRASTRUCTURE ONLY — UniPC fixtures produced by the REFERENCE ITSELF (via oracle/ref_shim.py):
 (1) tests/golden/unipc_scheduler.pt: UniPCMultistepScheduler (SD-1.5 betas, defaults) stepped over seeded tensors;
 (2) tests/golden/tiny_pipeline_unipc.pt: the unmodified StableDiffusionBEVControlNetPipeline.__call__ with its default
     sampler (UniPC, magicdrive/misc/test_utils.py:129), 4 steps, CFG 2.0, tiny models.
Run in the build container (needs /root/reference):  python -m oracle.make_golden_unipc"""
import importlib
import os
import sys

import torch

from oracle import ref_shim
from oracle.make_golden import OUT, load_ref, synthetic_inputs, tiny_configs


@torch.no_grad()
def main():
    R = ref_shim.load()
    UniPC = importlib.import_module("diffusers.schedulers.scheduling_unipc_multistep").UniPCMultistepScheduler
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000)
    # ---- (1) scheduler alone
    cases = {}
    for n in (20, 4, 2):
        sch = UniPC(**kw)
        sch.set_timesteps(n)
        g = torch.Generator().manual_seed(40 + n)
        x = torch.randn(2, 4, 6, 7, generator=g)
        eps = [torch.randn(2, 4, 6, 7, generator=g) for _ in range(n)]
        x0, traj = x.clone(), []
        for e, t in zip(eps, sch.timesteps.tolist()):
            x = sch.step(e, t, x).prev_sample
            traj.append(x.clone())
        cases[n] = dict(timesteps=sch.timesteps.clone(), x=x0, eps=torch.stack(eps), traj=torch.stack(traj))
    torch.save(cases, os.path.join(OUT, "unipc_scheduler.pt"))
    # ---- (2) the reference pipeline with its default sampler
    ucfg, ccfg = tiny_configs()
    mv, cn, _, _ = load_ref(ucfg, ccfg, seed=7)
    inp = synthetic_inputs(1, 6, 10, 13, n_box=5, map_hw=52, seed=3)

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
    pipe = R.StableDiffusionBEVControlNetPipeline(vae=vae, text_encoder=TextStub(), unet=mv, controlnet=cn,
                                                  scheduler=UniPC(**kw), tokenizer=None)
    pipe.set_progress_bar_config(disable=True)
    out = pipe(prompt=None, image=inp["bev_map"], camera_param=inp["camera_param"], height=80, width=104,
               num_inference_steps=4, guidance_scale=2.0, latents=inp["latents"].clone(),
               prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
               output_type="latent", bev_controlnet_kwargs={"bboxes_3d_data": inp["bboxes_3d_data"]})
    same = torch.load(os.path.join(OUT, "tiny_pipeline.pt"), weights_only=False)["inputs"]
    assert all(torch.equal(same[k], inp[k]) for k in ("latents", "camera_param", "bev_map", "prompt_embeds"))
    torch.save(dict(inputs_from="tiny_pipeline.pt", steps=4, guidance=2.0, latents_out=out.images, seed=7),
               os.path.join(OUT, "tiny_pipeline_unipc.pt"))
    for f in ("unipc_scheduler.pt", "tiny_pipeline_unipc.pt"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    sys.exit(main())
