"""The denoising loop of StableDiffusionBEVControlNetPipeline.__call__ (magicdrive/pipeline/pipeline_bev_controlnet.py:
303-451) on top of the B200 engines: classifier-free-guidance batching ([uncond ; cond]), ControlNet -> UNet ->
guidance -> DDIM update per step, with everything step-invariant hoisted and one whole step captured in a CUDA graph.

Differences from the reference that do not change results: latents stay fp32 and NHWC-resident between steps
(the reference re-stacks / rearranges 5-D tensors every step); the DDIM update (eta = 0) is fused with the guidance
combine; `timestep` and the two DDIM coefficients are read from device memory so the captured graph is replayed
unchanged for every step.  The reference refuses DDIM only because `scheduler.step` accepts a `generator`
(:93-97); eta = 0 is deterministic, so no generator is needed.
"""
from typing import Dict, Optional

import torch

from . import ops
from .models import BEVControlNetModel, UNet2DConditionModelMultiview

F32, BF16 = torch.float32, torch.bfloat16


class DDIMSchedule:
    """DDIMScheduler(beta 0.00085-0.012 scaled_linear, clip_sample False, set_alpha_to_one False, steps_offset 1,
    'leading' spacing): scheduling_ddim.py:120-160, 287-323, 325-445 with eta = 0 reduced to x' = c0 x + c1 eps."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0).double()  # fp32 table like the reference, fp64 math after
        self.T = num_train_timesteps
        self.steps_offset = steps_offset

    def set_timesteps(self, n: int):
        ratio = self.T // n
        self.timesteps = [int(round(i * ratio)) + self.steps_offset for i in range(n)][::-1]
        coefs = []
        for t in self.timesteps:
            prev = t - ratio
            a_t = self.alphas_cumprod[t]
            a_p = self.alphas_cumprod[prev] if prev >= 0 else self.alphas_cumprod[0]
            c0 = (a_p / a_t) ** 0.5
            c1 = (1 - a_p) ** 0.5 - (a_p * (1 - a_t) / a_t) ** 0.5
            coefs.append([float(c0), float(c1)])
        self.coefs = coefs
        return self.timesteps


class BEVControlNetDenoiser:
    """Call-compatible core of StableDiffusionBEVControlNetPipeline for `output_type="latent"` with precomputed
    prompt embeddings (the CLIP text encoder and the VAE sit outside the hot path: SURVEY.md §2.1)."""

    def __init__(self, unet: UNet2DConditionModelMultiview, controlnet: BEVControlNetModel, use_cuda_graph: bool = True,
                 overlap_controlnet: bool = True, view_shard=None):
        """view_shard: a dist.ViewShard to split the cameras of each scene across the ranks of its group (inputs are
        still passed with all n_cam views on every rank; the result is gathered back to (S, n_cam, ...))."""
        self.unet, self.controlnet = unet, controlnet
        self.overlap_controlnet = overlap_controlnet
        self.view_shard = view_shard
        unet.engine().set_view_shard(view_shard)
        self._side = {}
        self.scheduler = DDIMSchedule()
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._graph_key = None
        self._static = None

    def release_graph(self):
        """Drop the captured CUDA graph (it is re-captured on the next run_steps)."""
        self._graph = None
        self._graph_key = None
        self._graph_state = None

    def _side_stream(self, device):
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(device=device)
        return self._side[key]

    # ------------------------------------------------------------------ one step on resident buffers
    def _step(self, st):
        ue, ce = st["ue"], st["ce"]
        V, h, w = st["V"], st["h"], st["w"]
        lat = st["latents"]  # fp32 [S*ncam*h*w, 4] NHWC, S scenes (no CFG duplication)
        # bf16, channel-padded to one K block; CFG: [uncond ; cond] share the latents (:352-354) -> repeat = 2
        x = ops.pack_latents(lat, ue.CIN_PAD, repeat=2 if st["cfg"] else 1)
        if self.overlap_controlnet and st.get("u_temb") is not None:
            # The ControlNet and the UNet's down/mid path only meet at the skip additions: run them on two streams so
            # that each one's small-grid kernels and per-kernel tails are filled by the other (captured as two branches
            # of the same CUDA graph).
            main = torch.cuda.current_stream()
            side = self._side_stream(lat.device)
            side.wait_stream(main)
            with torch.cuda.stream(side), ops.workspace_slot(1):
                down, mid, _, _ = ce.forward(x, V, h, w, st["t_dev"], st["c_kv"], st["lc"], st["map"], st["cond_scale"],
                                             temb_all=st["c_temb"])
            xe, skips = ue.forward_encoder(x, V, h, w, st["u_temb"], st["u_kv"], st["lc"])
            main.wait_stream(side)
            eps = ue.forward_decoder(xe, skips, st["u_temb"], st["u_kv"], st["lc"], down, mid)
        else:
            down, mid, _, _ = ce.forward(x, V, h, w, st["t_dev"], st["c_kv"], st["lc"], st["map"], st["cond_scale"],
                                         temb_all=st.get("c_temb"))
            eps = ue.forward(x, V, h, w, st["t_dev"], st["u_kv"], st["lc"], down, mid, temb_all=st.get("u_temb"))
        ops.cfg_ddim_step(eps, lat, st["coef_dev"], st["cfg"], st["guidance"], c=lat.shape[1])

    @torch.no_grad()
    def prepare(self, latents, prompt_embeds, negative_prompt_embeds, camera_param, bboxes_3d_data, image,
                guidance_scale=2.0, controlnet_conditioning_scale=1.0):
        """Host -> device staging + all step-invariant work.  latents: (S, 4, h, w) initial noise shared by the views
        (:326) or (S, n_cam, 4, h, w)."""
        dev = self.unet.device
        cn, un = self.controlnet, self.unet
        cfg = guidance_scale > 1.0
        if self.view_shard is not None:
            if latents.dim() == 4:
                latents = torch.stack([latents] * camera_param.shape[1], dim=1)
            cut = self.view_shard.slice_views(dict(camera_param=camera_param, bboxes_3d_data=bboxes_3d_data, latents=latents))
            camera_param, bboxes_3d_data, latents = cut["camera_param"], cut["bboxes_3d_data"], cut["latents"]
        camera_param = camera_param.to(dev, F32)
        S, n_cam = camera_param.shape[:2]
        prompt_embeds = prompt_embeds.to(dev, F32)
        image = image.to(dev, F32)
        boxes = None if bboxes_3d_data is None else {k: v.to(dev) for k, v in bboxes_3d_data.items()}
        if cfg:
            negative_prompt_embeds = negative_prompt_embeds.to(dev, F32)
            kw = cn.add_uncond_to_kwargs(camera_param=camera_param, bboxes_3d_data=boxes, image=image)
            camera_param, boxes = kw["camera_param"], kw["bboxes_3d_data"]
            text = torch.cat([negative_prompt_embeds, prompt_embeds])
            image = torch.cat([image, image])
        else:
            text = prompt_embeds
        cond = cn.prepare_conditions(camera_param, boxes, text, image)
        u_kv, lc = un.prepare_context(cond["ctx"])
        lat = latents.to(dev, F32)
        if lat.dim() == 4:
            lat = torch.stack([lat] * n_cam, dim=1)
        S_, _, c, h, w = lat.shape
        lat_nhwc = lat.reshape(S * n_cam, c, h, w).permute(0, 2, 3, 1).contiguous().view(-1, c)
        V = S * n_cam * (2 if cfg else 1)
        sig = (V, h, w, cfg, lc, S, n_cam)
        st = self._static
        if st is not None and st["sig"] == sig:
            # same shapes as the resident state: refresh its buffers in place so a captured graph stays valid
            st["latents"].copy_(lat_nhwc)
            st["map"].copy_(cond["map"])
            for k, v in cond["kv"].items():
                st["c_kv"][k].copy_(v)
            for k, v in u_kv.items():
                st["u_kv"][k].copy_(v)
            st["guidance"], st["cond_scale"] = float(guidance_scale), float(controlnet_conditioning_scale)
            return st
        st = dict(ue=un.engine(), ce=cn.engine(), V=V, h=h, w=w, S=S, n_cam=n_cam, cfg=cfg, sig=sig,
                  guidance=float(guidance_scale), cond_scale=float(controlnet_conditioning_scale), latents=lat_nhwc,
                  c_kv={k: v.clone() for k, v in cond["kv"].items()}, u_kv={k: v.clone() for k, v in u_kv.items()},
                  lc=lc, map=cond["map"].clone(),
                  t_dev=torch.zeros(V, dtype=F32, device=dev), coef_dev=torch.zeros(2, dtype=F32, device=dev))
        self._static, self._graph = st, None
        return st

    def _set_step(self, st, i):
        st["t_dev"].copy_(st["t_table"][i], non_blocking=True)
        st["coef_dev"].copy_(st["coef_table"][i], non_blocking=True)
        st["u_temb"].copy_(st["u_temb_table"][i:i + 1], non_blocking=True)
        st["c_temb"].copy_(st["c_temb_table"][i:i + 1], non_blocking=True)

    def set_schedule(self, st, num_inference_steps):
        ts = self.scheduler.set_timesteps(num_inference_steps)
        dev = st["latents"].device
        st["t_table"] = torch.tensor(ts, dtype=F32, device=dev)[:, None].expand(-1, st["V"]).contiguous()
        st["coef_table"] = torch.tensor(self.scheduler.coefs, dtype=F32, device=dev)
        # the time-embedding MLP + all time_emb_proj layers depend only on t: one table for the whole schedule
        # (every view-sample of a step shares t, so one row serves all images: rowbias stride 0)
        tt = torch.tensor(ts, dtype=F32, device=dev)
        st["u_temb_table"] = st["ue"].time_embed(tt)
        st["c_temb_table"] = st["ce"].time_embed(tt)
        if "u_temb" not in st or st["u_temb"].shape[1] != st["u_temb_table"].shape[1]:
            st["u_temb"] = torch.zeros_like(st["u_temb_table"][:1])
            st["c_temb"] = torch.zeros_like(st["c_temb_table"][:1])
        return ts

    def run_steps(self, st, first: int, last: int):
        """Run denoising steps [first, last) on the resident state (eager on the first use, then graph replay)."""
        key = (st["sig"], id(st["latents"]), st["guidance"], st["cond_scale"])
        for i in range(first, last):
            self._set_step(st, i)
            if not self.use_cuda_graph:
                self._step(st)
                continue
            if self._graph is None or self._graph_key != key:
                # one eager step sizes workspaces / sets kernel attributes, then capture the same step
                saved = st["latents"].clone()
                self._step(st)
                st["latents"].copy_(saved)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step(st)
                st["latents"].copy_(saved)
                self._graph, self._graph_key, self._graph_state = g, key, st
            self._graph.replay()

    @torch.no_grad()
    def __call__(self, image, camera_param, prompt_embeds, negative_prompt_embeds=None, latents=None,
                 num_inference_steps: int = 50, guidance_scale: float = 2.0, bev_controlnet_kwargs: Optional[Dict] = None,
                 controlnet_conditioning_scale: float = 1.0, output_type: str = "latent"):
        """Same argument meaning as the reference pipeline call (:114-160).  Returns latents (S, n_cam, 4, h, w) fp32."""
        if output_type != "latent":
            raise NotImplementedError("VAE decode is outside the hot path; use output_type='latent'")
        boxes = (bev_controlnet_kwargs or {}).get("bboxes_3d_data")
        st = self.prepare(latents, prompt_embeds, negative_prompt_embeds, camera_param, boxes, image, guidance_scale,
                          controlnet_conditioning_scale)
        self.set_schedule(st, num_inference_steps)
        self.run_steps(st, 0, num_inference_steps)
        return self.latents_out(st)

    def latents_out(self, st):
        S, n_cam, h, w = st["S"], st["n_cam"], st["h"], st["w"]
        out = st["latents"].view(S, n_cam, h, w, -1).permute(0, 1, 4, 2, 3).contiguous()
        return out if self.view_shard is None else self.view_shard.gather_views(out)
