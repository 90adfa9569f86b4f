"""Pipelines — drop-ins for the reference's three denoising drivers, all on one fused engine:

* `B200Stage2InpaintPipeline`   <- `Stage2_InpaintDiffusionPipeline`
  (/root/reference/src/pipelines/stage2_inpaint_pipeline.py:71-541): same constructor modules (vae, unet, scheduler),
  same `__call__` keyword surface (:391-417) and `.images` output (:541), same conditioning layout (:427-466, CFG batch
  = [uncond ; cond], only tokens + class embedding zeroed for the unconditional half).
* `B200PCDMsPipeline`           <- `PCDMsPipeline`, the demo driver
  (/root/reference/src/pipelines/PCDMs_pipeline.py:889-1180; used by pcdms_demo.ipynb): no class embedding, the
  conditioning tokens are `prompt_embeds` / `negative_prompt_embeds` (the projected DINOv2 tokens and the projection of
  zeros), 9-channel input `[latents, mask, simg_mask_latents]` (:1115), `cond_pose` added after conv_in.
* `B200Stage3RefinedPipeline`   <- `Stage3_RefinedPipeline`
  (/root/reference/src/pipelines/stage3_refined_pipeline.py:443-578): stock 8-channel UNet, input
  `[latents, vae(gen_t_image)]` with the image latents zeroed in the unconditional half (:484-491, :538).

The denoising loop (stage2 :496-525) is executed as ONE CUDA graph replayed `num_inference_steps` times: the graph
holds the whole UNet forward (~400 launches of the sm_100a kernels) plus the fused CFG + scheduler-step + input-rebuild
kernel (DDIM or UniPC); step-dependent scalars (scheduler coefficients, timestep) are read from device tables indexed
by a device-side step counter, so the captured graph is step-invariant.  Step-invariant work (cross-attention K/V, pose
/ mask / masked-latent layout) is done once per call.  When handed a foreign unet / scheduler the generic protocol loop
is used instead.

`vae` is a `B200AutoencoderKL` (pcdms_b200/vae.py) or any object with diffusers' AutoencoderKL encode / decode surface,
or None — then latents must be passed in and `output_type="latent"` is the only output.
"""
from __future__ import annotations

import inspect
import json
import os
import warnings
from types import SimpleNamespace
from typing import Optional

import torch

from . import ops
from .scheduler import B200DDIMScheduler, B200UniPCMultistepScheduler
from .unet import B200AttnProcessor, B200UNet2DConditionModel


class Stage2PipelineOutput(SimpleNamespace):
    pass


class _B200DenoisingPipeline:
    """What the three reference pipelines share: module registry, the diffusers conveniences their drivers call, the
    fused (one CUDA graph per step) denoising engine and the generic protocol loop."""

    def __init__(self, vae=None, unet: Optional[B200UNet2DConditionModel] = None, scheduler=None, **ignored):
        if unet is None or scheduler is None:
            raise ValueError("unet and scheduler are required")
        if hasattr(scheduler, "config") and getattr(scheduler.config, "steps_offset", 1) != 1:
            # reference :86-98: an outdated steps_offset is patched to 1 with a deprecation warning, never an error
            # (UniPC's default 0 with linspace spacing does not use the offset at all)
            warnings.warn(f"The configuration file of this scheduler: {scheduler} is outdated. `steps_offset` should be "
                          f"set to 1 instead of {scheduler.config.steps_offset}.", FutureWarning)
            try:
                scheduler.config["steps_offset"] = 1
            except TypeError:
                pass   # immutable foreign config: the reference swaps _internal_dict; nothing here depends on it
        self.vae, self.unet, self.scheduler = vae, unet, scheduler
        self.vae_scale_factor = (2 ** (len(vae.config.block_out_channels) - 1)) if vae is not None else 8
        self._graphs = {}
        self.use_cuda_graph = True

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, torch_dtype=torch.float16, **kw):
        """Mirror of `<Pipeline>.from_pretrained(path, torch_dtype=..., unet=..., scheduler=...)`
        (stage2_batchtest_inpaint_model.py:123, pcdms_demo.ipynb): components passed by keyword win; the others are
        loaded from the `unet/` and `vae/` subfolders."""
        unet = kw.get("unet") or B200UNet2DConditionModel.from_pretrained(
            pretrained_model_name_or_path, subfolder="unet", torch_dtype=torch_dtype)
        vae = kw.get("vae")
        if vae is None and os.path.isdir(os.path.join(str(pretrained_model_name_or_path), "vae")):
            from .vae import B200AutoencoderKL
            vae = B200AutoencoderKL.from_pretrained(pretrained_model_name_or_path, subfolder="vae",
                                                    torch_dtype=torch_dtype)
            if not vae._loaded:
                vae = None
        scheduler = kw.get("scheduler")
        if scheduler is None:   # <path>/scheduler/scheduler_config.json when present (diffusers layout), else DDIM defaults
            sc = os.path.join(str(pretrained_model_name_or_path), "scheduler", "scheduler_config.json")
            cfg = {}
            if os.path.exists(sc):
                with open(sc) as f:
                    cfg = json.load(f)
            known = set(inspect.signature(B200DDIMScheduler.__init__).parameters) - {"self"}
            scheduler = B200DDIMScheduler(**{k: v for k, v in cfg.items() if k in known})
        return cls(vae=vae, unet=unet, scheduler=scheduler)

    # -- diffusers pipeline conveniences the reference drivers call ------------------------------------------------
    def to(self, *a, **k):
        return self

    @property
    def device(self):
        return self.unet.device

    @property
    def _execution_device(self):
        return self.unet.device

    def enable_xformers_memory_efficient_attention(self, *a, **k):  # stage2_batchtest_inpaint_model.py:133
        return None

    def enable_vae_slicing(self):
        return None

    def enable_vae_tiling(self):
        return None

    def set_progress_bar_config(self, **k):
        return None

    def check_inputs(self, height, width, callback_steps):
        """reference :324-369"""
        if height is None or width is None or height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if (callback_steps is None) or (not isinstance(callback_steps, int) or callback_steps <= 0):
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type"
                             f" {type(callback_steps)}.")

    def prepare_extra_step_kwargs(self, generator, eta):
        """reference :307-322"""
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        extra = {}
        if "eta" in params:
            extra["eta"] = eta
        if "generator" in params:
            extra["generator"] = generator
        return extra

    def _common_checks(self, cross_attention_kwargs, guidance_rescale, guidance_scale, height, width, callback_steps):
        if cross_attention_kwargs is not None:
            raise NotImplementedError("cross_attention_kwargs are not supported")
        self.check_inputs(height, width, callback_steps)
        if not guidance_scale > 1.0:
            raise NotImplementedError("guidance_scale <= 1: the reference's non-CFG branch is inconsistent "
                                      "(stage2_inpaint_pipeline.py:449,465) and is never used by its drivers")

    def _initial_latents(self, latents, n, h, w, generator):
        """reference prepare_latents (:371-386): randn * init_noise_sigma; the scheduler state is kept in fp32."""
        dev = self.unet.device
        if latents is None:
            latents = torch.randn((n, 4, h, w), generator=generator,
                                  device=generator.device if generator is not None else dev, dtype=torch.float32)
        return latents.to(device=dev, dtype=torch.float32) * self.scheduler.init_noise_sigma

    def _denoise(self, latents, extra, pose_cond, feature_f, class_labels, guidance_scale, num_inference_steps, eta,
                 generator, callback, callback_steps, guidance_rescale=0.0):
        """latents [n,4,h,w] fp32; extra [2n,Ce,h,w] (the non-latent input channels of both CFG halves); pose_cond
        [2n,320,h,w] or None; feature_f [2n,S,D]; class_labels [2n,1,D] or None."""
        fast = (isinstance(self.unet, B200UNet2DConditionModel)
                and isinstance(self.scheduler, (B200DDIMScheduler, B200UniPCMultistepScheduler))
                and eta == 0.0 and callback is None
                and all(isinstance(a.processor, B200AttnProcessor) for a in self.unet._attn.values()))
        if fast:
            st = self.prepare_fused(latents.contiguous().clone(), extra, pose_cond, feature_f, class_labels,
                                    float(guidance_scale), num_inference_steps, float(guidance_rescale or 0.0))
            self.replay_fused(st)
            return st.latents.clone()
        return self._denoise_generic(latents, extra, pose_cond, feature_f, class_labels, guidance_scale,
                                     self.scheduler.timesteps, eta, generator, callback, callback_steps,
                                     float(guidance_rescale or 0.0))

    def _finish(self, latents, output_type, return_dict, generator=None):
        dt = self.unet.dtype
        if output_type == "latent" or self.vae is None:
            if output_type != "latent" and self.vae is None:
                raise ValueError("no VAE: use output_type='latent'")
            image = latents
        else:
            image = self.vae.decode((latents / self.vae.config.scaling_factor).to(dt), return_dict=False)[0]  # :528
            image = self._postprocess(image, output_type)
        if not return_dict:
            return (image, None)
        return Stage2PipelineOutput(images=image, nsfw_content_detected=None)

    @staticmethod
    def _postprocess(image, output_type):
        """diffusers VaeImageProcessor.postprocess (denormalize -> numpy -> PIL), as the reference's :530-534."""
        if output_type == "pt":
            return image
        image = (image / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).float().numpy()
        if output_type == "np":
            return image
        from PIL import Image
        return [Image.fromarray((im * 255).round().astype("uint8")) for im in image]

    # ------------------------------------------------------------------------------------------------------------
    # fused engine
    # ------------------------------------------------------------------------------------------------------------
    def prepare_fused(self, latents, extra, pose_cond, feature_f, class_labels, guidance, steps, rescale=0.0):
        """Per-call set-up of the fused loop: lays the conditioning out in the graph's static buffers (NHWC, 16-bit),
        projects the cross-attention K/V once, uploads the step tables and (re)captures the one-step CUDA graph when
        the shapes / guidance / step count changed.  Returns the state object `replay_fused` consumes."""
        unet, sch = self.unet, self.scheduler
        dev, dt = unet.device, unet.dtype
        n, _, h, w = latents.shape
        B = 2 * n
        unipc = isinstance(sch, B200UniPCMultistepScheduler)
        has_pose, has_cls = pose_cond is not None, class_labels is not None
        key = (n, h, w, feature_f.shape[1], dt, unipc, has_pose, has_cls, extra.shape[1])
        st = self._graphs.get(key)
        if st is None:
            # UniPC keeps {sample, last_sample, model_outputs[-1], model_outputs[-2]} as four fp32 planes
            state = torch.zeros((4 if unipc else 1, n, 4, h, w), device=dev, dtype=torch.float32)
            st = SimpleNamespace(
                x9=torch.zeros((B, h, w, 64), device=dev, dtype=dt),
                x9_init=torch.zeros((B, h, w, 64), device=dev, dtype=dt),
                state=state, unipc=unipc,
                latents=state[0],
                latents_init=torch.empty((n, 4, h, w), device=dev, dtype=torch.float32),
                pose=torch.empty((B, h, w, pose_cond.shape[1]), device=dev, dtype=dt) if has_pose else None,
                cls=torch.empty((B, class_labels.shape[-1]), device=dev, dtype=dt) if has_cls else None,
                t_cur=torch.zeros(1, device=dev, dtype=torch.float32),
                counter=torch.zeros(2, device=dev, dtype=torch.int32),
                guidance=None, graph=None, kv=None, coef=None, t_table=None, steps=None, wver=None, rescale=0.0,
                ratio=torch.ones(n, device=dev, dtype=torch.float32),
                launches_per_step=0)
            self._graphs[key] = st
        x_nchw = torch.cat([torch.cat([latents, latents], dim=0), extra.to(dev, torch.float32)], dim=1).contiguous()
        ops.nchw_to_nhwc_pad(x_nchw, 64, dt, out=st.x9_init)                                    # ref :499-501
        st.latents_init.copy_(latents)
        if has_pose:
            ops.nchw_to_nhwc_pad(pose_cond.contiguous(), pose_cond.shape[1], dt, out=st.pose)
        if has_cls:
            st.cls.copy_(class_labels.reshape(B, -1))
        first = st.kv is None
        st.kv = unet.context_kv(feature_f, out=st.kv)   # K/V GEMMs write straight into the graph's static buffers
        coef = sch.coefficient_table(dev)
        t_table = torch.cat([sch.timesteps.to(dev, torch.float32), torch.zeros(1, device=dev)]).contiguous()
        wver = getattr(unet, "_weights_version", 0)
        rebuild = (st.graph is None or first or st.guidance != guidance or st.steps != steps or st.wver != wver
                   or st.rescale != rescale)
        st.guidance, st.steps, st.wver, st.rescale = guidance, steps, wver, rescale
        if st.coef is None or st.coef.shape != coef.shape:
            st.coef, st.t_table = coef, t_table
            rebuild = True
        else:
            st.coef.copy_(coef)
            st.t_table.copy_(t_table)
        if self.use_cuda_graph and rebuild:
            # warm-up on a side stream (allocator + lazy kernel attribute set-up), then capture
            self._reset_state(st)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._one_step(st)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            from . import lib as _l
            before = _l.launch_count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._one_step(st)
            st.launches_per_step = _l.launch_count - before
            st.graph = g
        return st

    def _one_step(self, st):
        eps_rows = self.unet.forward_nhwc(st.x9, st.t_cur, st.kv, st.cls, st.pose)
        ratio = None
        if st.rescale > 0.0:   # reference rescale_noise_cfg (:52-63, :514-516): per-sample std ratio, one small launch
            ratio = ops.cfg_rescale_ratio(eps_rows, st.guidance, out=st.ratio, nhwc_channels=4)
        if st.unipc:
            ops.cfg_unipc_step(eps_rows, st.state, st.x9, st.coef, st.counter, st.guidance, st.t_table, st.t_cur,
                               ratio=ratio, guidance_rescale=st.rescale)
        else:
            ops.cfg_ddim_step(eps_rows, st.latents, st.x9, st.coef, st.counter, st.guidance, st.t_table, st.t_cur,
                              ratio=ratio, guidance_rescale=st.rescale)

    def _reset_state(self, st):
        st.x9.copy_(st.x9_init)
        if st.unipc:
            st.state[1:].zero_()
        st.latents.copy_(st.latents_init)
        st.counter.zero_()
        st.t_cur.copy_(st.t_table[:1])

    def replay_fused(self, st):
        """The denoising loop proper (reference :496-525): `steps` replays of the captured one-step graph."""
        self._reset_state(st)
        if self.use_cuda_graph:
            for _ in range(st.steps):
                st.graph.replay()
        else:
            for _ in range(st.steps):
                self._one_step(st)

    def _denoise_generic(self, latents, extra, pose_cond, feature_f, class_labels, guidance_scale, timesteps, eta,
                         generator, callback, callback_steps, guidance_rescale=0.0):
        """Protocol loop (reference :496-525) for foreign unet / scheduler objects, callbacks and custom attention
        processors.  The CFG combine (+ rescale) is the pcdm_cfg_combine kernel — no eager-torch arithmetic here."""
        dt = self.unet.dtype
        step_kw = self.prepare_extra_step_kwargs(generator, eta)
        extra = extra.to(latents.device, dt)
        latents = latents.to(dt)
        kw = {}
        if class_labels is not None:
            kw["class_labels"] = class_labels
        if pose_cond is not None:
            kw["my_pose_cond"] = pose_cond
        for i, t in enumerate(timesteps):
            x = torch.cat([latents] * 2)
            x = self.scheduler.scale_model_input(x, t)
            xin = torch.cat([x, extra], dim=1).to(dt)
            eps = self.unet(xin, t, encoder_hidden_states=feature_f, return_dict=False, **kw)[0]
            eps = ops.cfg_combine(eps.contiguous(), float(guidance_scale), guidance_rescale)       # :510-516
            latents = self.scheduler.step(eps, t, latents, **step_kw, return_dict=False)[0]
            if callback is not None and i % callback_steps == 0:
                callback(i, t, latents)
        return latents.float()


class B200Stage2InpaintPipeline(_B200DenoisingPipeline):
    @torch.no_grad()
    def __call__(self, prompt=None, height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, negative_prompt=None,
                 num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None, latents=None,
                 prompt_embeds=None, negative_prompt_embeds=None, output_type: Optional[str] = "pil",
                 return_dict: bool = True, callback=None, callback_steps: int = 1, cross_attention_kwargs=None,
                 guidance_rescale: float = 0.0,
                 vae_image=None, mask=None, s_img_proj_f=None, st_pose_f=None, pred_t_img_embed=None,
                 masked_latents=None):
        self._common_checks(cross_attention_kwargs, guidance_rescale, guidance_scale, height, width, callback_steps)
        dev, dt = self.unet.device, self.unet.dtype
        bs, num, _ = s_img_proj_f.shape
        if bs != 1:
            raise NotImplementedError("the reference drivers run one source/target pair per call (bs == 1)")
        n = bs * num_images_per_prompt
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        B = 2 * n

        # --- conditioning (reference :430-462), built once per call -------------------------------------------
        pose_cond = torch.cat([st_pose_f.to(dev)] * B).to(dt)                                   # :430-431
        if mask is None:                                                                         # :434-437
            m1 = torch.ones((bs, 1, h, w // 2), dtype=torch.float32, device=dev)
            m0 = torch.zeros((bs, 1, h, w // 2), dtype=torch.float32, device=dev)
            mask = torch.cat([m1, m0], dim=3)
        mask = torch.cat([mask.to(dev, torch.float32)] * B)
        if masked_latents is None:                                                               # :443-444
            if self.vae is None:
                raise ValueError("pass masked_latents= when the pipeline has no VAE")
            masked_latents = self.vae.encode(vae_image.to(device=dev, dtype=dt)).latent_dist.sample(generator=generator)
            masked_latents = masked_latents * self.vae.config.scaling_factor
        masked_latents = torch.cat([masked_latents.to(dev, torch.float32)] * B)                  # :445
        feature_f = torch.cat([s_img_proj_f, pred_t_img_embed], dim=1).to(dev)                   # :448
        feature_f = feature_f.repeat(n, 1, 1).to(dt)
        prior_embed = pred_t_img_embed.to(dev).repeat(n, 1, 1).to(dt)                            # :452
        feature_f = torch.cat([torch.zeros_like(feature_f), feature_f], dim=0)                   # :455-458
        prior_embed = torch.cat([torch.zeros_like(prior_embed), prior_embed], dim=0)             # :461-462

        # --- timesteps & latents (reference :472-487) -----------------------------------------------------------
        self.scheduler.set_timesteps(num_inference_steps, device=dev)
        latents = self._initial_latents(latents, n, h, w, generator)
        extra = torch.cat([mask, masked_latents], dim=1)
        latents = self._denoise(latents, extra, pose_cond, feature_f, prior_embed, guidance_scale,
                                num_inference_steps, eta, generator, callback, callback_steps, guidance_rescale)
        return self._finish(latents, output_type, return_dict)


class B200SimpleStage2InpaintPipeline(_B200DenoisingPipeline):
    """`Simple_Stage2_InpaintDiffusionPipeline` (/root/reference/src/pipelines/stage2_inpaint_pipeline.py:544-877): the
    stage-2 call surface without the predicted-embedding path — tokens are `s_img_proj_f` alone, the UNet has no class
    embedding; `pred_t_img_embed` is accepted and ignored, as in the reference (:783, never read)."""

    @torch.no_grad()
    def __call__(self, prompt=None, height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, negative_prompt=None,
                 num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None, latents=None,
                 prompt_embeds=None, negative_prompt_embeds=None, output_type: Optional[str] = "pil",
                 return_dict: bool = True, callback=None, callback_steps: int = 1, cross_attention_kwargs=None,
                 guidance_rescale: float = 0.0,
                 vae_image=None, mask=None, s_img_proj_f=None, st_pose_f=None, pred_t_img_embed=None,
                 masked_latents=None):
        self._common_checks(cross_attention_kwargs, guidance_rescale, guidance_scale, height, width, callback_steps)
        dev, dt = self.unet.device, self.unet.dtype
        bs, num, _ = s_img_proj_f.shape
        if bs != 1:
            raise NotImplementedError("the reference drivers run one source/target pair per call (bs == 1)")
        n = bs * num_images_per_prompt
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        B = 2 * n
        pose_cond = torch.cat([st_pose_f.to(dev)] * B).to(dt)                                    # :793-794
        if mask is None:                                                                          # :797-800
            m1 = torch.ones((bs, 1, h, w // 2), dtype=torch.float32, device=dev)
            m0 = torch.zeros((bs, 1, h, w // 2), dtype=torch.float32, device=dev)
            mask = torch.cat([m1, m0], dim=3)
        mask = torch.cat([mask.to(dev, torch.float32)] * B)                                       # :801-803
        if masked_latents is None:                                                                # :806-807
            if self.vae is None:
                raise ValueError("pass masked_latents= when the pipeline has no VAE")
            masked_latents = self.vae.encode(vae_image.to(device=dev, dtype=dt)).latent_dist.sample(generator=generator)
            masked_latents = masked_latents * self.vae.config.scaling_factor
        masked_latents = torch.cat([masked_latents.to(dev, torch.float32)] * B)                   # :808
        feature_f = s_img_proj_f.to(dev).repeat(n, 1, 1).to(dt)                                   # :811
        feature_f = torch.cat([torch.zeros_like(feature_f), feature_f], dim=0)                    # :814-816
        self.scheduler.set_timesteps(num_inference_steps, device=dev)
        latents = self._initial_latents(latents, n, h, w, generator)
        extra = torch.cat([mask, masked_latents], dim=1)
        latents = self._denoise(latents, extra, pose_cond, feature_f, None, guidance_scale, num_inference_steps, eta,
                                generator, callback, callback_steps, guidance_rescale)
        return self._finish(latents, output_type, return_dict)


class B200PCDMsPipeline(_B200DenoisingPipeline):
    """Demo driver surface (PCDMs_pipeline.py:889-926 signature; pcdms_demo.ipynb call): image-token conditioning comes
    in as `prompt_embeds` / `negative_prompt_embeds`, there is no class embedding."""

    @torch.no_grad()
    def __call__(self, simg_mask_latents=None, mask=None, cond_pose=None, prompt=None, height: Optional[int] = None,
                 width: Optional[int] = None, num_inference_steps: int = 50, timesteps=None,
                 guidance_scale: float = 7.5, negative_prompt=None, num_images_per_prompt: Optional[int] = 1,
                 eta: float = 0.0, generator=None, latents=None, prompt_embeds=None, negative_prompt_embeds=None,
                 ip_adapter_image=None, output_type: Optional[str] = "pil", return_dict: bool = True,
                 cross_attention_kwargs=None, guidance_rescale: float = 0.0, clip_skip=None,
                 callback_on_step_end=None, callback_on_step_end_tensor_inputs=("latents",), **kwargs):
        callback, callback_steps = kwargs.pop("callback", None), kwargs.pop("callback_steps", None) or 1
        if prompt is not None or negative_prompt is not None or ip_adapter_image is not None or timesteps is not None \
                or callback_on_step_end is not None:
            raise NotImplementedError("B200PCDMsPipeline: text prompts, IP-adapter images, custom timesteps and "
                                      "step-end callbacks are not part of the PCDMs demo path (pcdms_demo.ipynb)")
        if prompt_embeds is None or negative_prompt_embeds is None:
            raise ValueError("prompt_embeds and negative_prompt_embeds (projected image tokens) are required")
        if self.unet.config.time_cond_proj_dim is not None:
            raise NotImplementedError("time_cond_proj_dim")
        self._common_checks(cross_attention_kwargs, guidance_rescale, guidance_scale, height, width, callback_steps)
        dev, dt = self.unet.device, self.unet.dtype
        n = prompt_embeds.shape[0] * num_images_per_prompt     # encode_prompt repeats per image (:458-462)
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        pe = prompt_embeds.to(dev, dt).repeat_interleave(num_images_per_prompt, dim=0)
        ne = negative_prompt_embeds.to(dev, dt).repeat_interleave(num_images_per_prompt, dim=0)
        feature_f = torch.cat([ne, pe])                                                          # :1062-1063

        def rows(t, what):   # [latents, mask, simg_mask_latents] is concatenated BEFORE the CFG duplication (:1115-1117)
            t = t.to(dev, torch.float32)
            if t.shape[0] not in (1, n):
                raise ValueError(f"{what}: batch {t.shape[0]} does not match {n} images")
            return t.expand(n, *t.shape[1:]) if t.shape[0] == 1 else t

        extra = torch.cat([rows(mask, "mask"), rows(simg_mask_latents, "simg_mask_latents")], dim=1)
        extra = torch.cat([extra, extra], dim=0)
        cp = cond_pose.to(dev, dt)
        pose_cond = cp.expand(2 * n, *cp.shape[1:]) if cp.shape[0] == 1 else torch.cat([rows(cp, "cond_pose")] * 2)
        self.scheduler.set_timesteps(num_inference_steps, device=dev)
        latents = self._initial_latents(latents, n, h, w, generator)
        latents = self._denoise(latents, extra, pose_cond.contiguous(), feature_f, None, guidance_scale,
                                num_inference_steps, eta, generator, callback, callback_steps, guidance_rescale)
        return self._finish(latents, output_type, return_dict, generator)


class B200Stage3RefinedPipeline(_B200DenoisingPipeline):
    """`Stage3_RefinedPipeline.__call__` (stage3_refined_pipeline.py:443-578).  The reference builds a CFG batch of 2
    whatever `num_images_per_prompt` is (:484-491) and therefore only runs with bs * num_images_per_prompt == 1; here
    the conditioning is repeated per image, which is the same computation at 1 and well-defined above."""

    @torch.no_grad()
    def __call__(self, prompt=None, height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, negative_prompt=None,
                 num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None, latents=None,
                 prompt_embeds=None, negative_prompt_embeds=None, output_type: Optional[str] = "pil",
                 return_dict: bool = True, callback=None, callback_steps: int = 1, cross_attention_kwargs=None,
                 guidance_rescale: float = 0.0, vae_gen_t_image=None, s_img_proj_f=None, gen_t_img_latents=None):
        self._common_checks(cross_attention_kwargs, guidance_rescale, guidance_scale, height, width, callback_steps)
        dev, dt = self.unet.device, self.unet.dtype
        bs, num, _ = s_img_proj_f.shape
        n = bs * num_images_per_prompt
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        if gen_t_img_latents is None:                                                            # :479-480
            if self.vae is None:
                raise ValueError("pass gen_t_img_latents= when the pipeline has no VAE")
            g = self.vae.encode(vae_gen_t_image.to(device=dev, dtype=dt)).latent_dist.sample(generator=generator)
            gen_t_img_latents = g * self.vae.config.scaling_factor
        g = gen_t_img_latents.to(dev, torch.float32).repeat_interleave(num_images_per_prompt, dim=0)
        f = s_img_proj_f.to(dev, dt).repeat_interleave(num_images_per_prompt, dim=0)
        feature_f = torch.cat([torch.zeros_like(f), f], dim=0)                                   # :487-488
        extra = torch.cat([torch.zeros_like(g), g], dim=0)                                       # :490-491
        self.scheduler.set_timesteps(num_inference_steps, device=dev)
        latents = self._initial_latents(latents, n, h, w, generator)
        latents = self._denoise(latents, extra, None, feature_f, None, guidance_scale, num_inference_steps, eta,
                                generator, callback, callback_steps, guidance_rescale)
        return self._finish(latents, output_type, return_dict)
