"""ORACLE (test infrastructure only) — CPU restatement of the stage-2 pipeline's conditioning set-up and
denoising loop, /root/reference/src/pipelines/stage2_inpaint_pipeline.py:420-525.

Out of scope here exactly as in SURVEY.md §8: the VAE encode (:443) and decode (:528) — callers hand in
`masked_latents` (already multiplied by the VAE scaling factor) and receive the final latents.
"""
from __future__ import annotations

import torch


def prepare_conditioning(*, s_img_proj_f, pred_t_img_embed, st_pose_f, masked_latents, height, width,
                         num_images_per_prompt=1, guidance_scale=2.0, mask=None, dtype=torch.float32):
    """ref :427-466.  Returns dict(pose_cond, mask, masked_latents, feature_f, prior_embed) with the CFG batch layout
    [uncond ; cond].  Quirk kept: only the tokens and the class embedding are zeroed for the unconditional half;
    pose, mask and masked latents are NOT dropped (:455-462)."""
    bs = s_img_proj_f.shape[0]
    cfg = guidance_scale > 1.0
    rep = 2 * num_images_per_prompt if cfg else 1
    pose_cond = torch.cat([st_pose_f] * rep).to(dtype)                                   # :430-431
    if mask is None:                                                                      # :434-437
        mask1 = torch.ones((bs, 1, int(height / 8), int(width / 16)), dtype=torch.float32)
        mask0 = torch.zeros((bs, 1, int(height / 8), int(width / 16)), dtype=torch.float32)
        mask = torch.cat([mask1, mask0], dim=3)
    mask = torch.cat([mask] * rep).to(dtype)                                              # :438-440
    masked_latents = torch.cat([masked_latents] * rep).to(dtype)                          # :445
    feature_f = torch.cat([s_img_proj_f, pred_t_img_embed], dim=1)                        # :448
    feature_f = feature_f.repeat(bs * num_images_per_prompt, 1, 1).to(dtype)              # :449
    prior_embed = pred_t_img_embed.repeat(bs * num_images_per_prompt, 1, 1).to(dtype)     # :452
    if cfg:                                                                               # :455-462
        feature_f = torch.cat([torch.zeros_like(feature_f), feature_f], dim=0)
        prior_embed = torch.cat([torch.zeros_like(prior_embed), prior_embed], dim=0)
    else:
        raise NotImplementedError("reference's non-CFG branch repeats feature_f twice (:465, a latent bug); "
                                  "the drivers always run guidance_scale=2")
    return dict(pose_cond=pose_cond, mask=mask, masked_latents=masked_latents, feature_f=feature_f,
                prior_embed=prior_embed)


def cfg_combine(noise_pred, guidance_scale):
    """ref :510-512"""
    u, c = noise_pred.chunk(2)
    return u + guidance_scale * (c - u)


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """ref :52-63 (disabled by the drivers: guidance_rescale = 0.0)"""
    std_text = noise_pred_text.std(dim=list(range(1, noise_pred_text.ndim)), keepdim=True)
    std_cfg = noise_cfg.std(dim=list(range(1, noise_cfg.ndim)), keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


@torch.no_grad()
def denoise_loop(unet, scheduler, *, latents, cond, num_inference_steps, guidance_scale=2.0, dtype=torch.float32,
                 return_trajectory=False, guidance_rescale=0.0):
    """ref :472-520.  `latents`: [n, 4, h, w] initial noise (already times init_noise_sigma); `cond` from
    prepare_conditioning.  Returns the final latents (and, optionally, the per-step (eps, latents) list)."""
    scheduler.set_timesteps(num_inference_steps)
    latents = latents.to(dtype)
    traj = []
    for t in scheduler.timesteps:
        x = torch.cat([latents] * 2)                                                      # :499
        x = scheduler.scale_model_input(x, t)                                             # :500
        x9 = torch.cat([x, cond["mask"], cond["masked_latents"]], dim=1).to(dtype)       # :501
        eps = unet(x9, t, class_labels=cond["prior_embed"], encoder_hidden_states=cond["feature_f"],
                   my_pose_cond=cond["pose_cond"], return_dict=False)[0]                  # :504-506
        eps_text = eps.chunk(2)[1]
        eps = cfg_combine(eps, guidance_scale)                                            # :510-512
        if guidance_rescale > 0.0:                                                        # :514-516
            eps = rescale_noise_cfg(eps, eps_text, guidance_rescale=guidance_rescale)
        latents = scheduler.step(eps, t, latents, return_dict=False)[0]                   # :519
        if return_trajectory:
            traj.append((eps.clone(), latents.clone()))
    return (latents, traj) if return_trajectory else latents


@torch.no_grad()
def denoise_loop_simple(unet, scheduler, *, latents, s_img_proj_f, st_pose_f, masked_latents, height, width,
                        num_inference_steps, num_images_per_prompt=1, guidance_scale=2.0, mask=None,
                        dtype=torch.float32):
    """`Simple_Stage2_InpaintDiffusionPipeline.__call__`, /root/reference/src/pipelines/stage2_inpaint_pipeline.py:
    757-877: the stage-2 loop WITHOUT the predicted-embedding inputs — tokens are `s_img_proj_f` alone (:815), zeros for
    the unconditional half (:818-820), no class embedding in the UNet call (:861-863); pose / mask / masked latents
    are repeated for both CFG halves (:793-812)."""
    bs = s_img_proj_f.shape[0]
    rep = 2 * num_images_per_prompt
    pose_cond = torch.cat([st_pose_f] * rep).to(dtype)                                    # :793-794
    if mask is None:                                                                      # :797-800
        mask = torch.cat([torch.ones((bs, 1, int(height / 8), int(width / 16))),
                          torch.zeros((bs, 1, int(height / 8), int(width / 16)))], dim=3)
    mask = torch.cat([mask.float()] * rep).to(dtype)                                      # :801-803
    masked = torch.cat([masked_latents] * rep)                                            # :808
    feature_f = s_img_proj_f.repeat(bs * num_images_per_prompt, 1, 1).to(dtype)           # :811
    feature_f = torch.cat([torch.zeros_like(feature_f), feature_f], dim=0)                # :814-816
    scheduler.set_timesteps(num_inference_steps)
    latents = latents.to(dtype)
    for t in scheduler.timesteps:
        x = torch.cat([latents] * 2)                                                      # :853
        x = scheduler.scale_model_input(x, t)
        x9 = torch.cat([x, mask, masked], dim=1).to(dtype)                                # :856
        eps = unet(x9, t, encoder_hidden_states=feature_f, my_pose_cond=pose_cond, return_dict=False)[0]
        eps = cfg_combine(eps, guidance_scale)                                            # :866-868
        latents = scheduler.step(eps, t, latents, return_dict=False)[0]                   # :872
    return latents


@torch.no_grad()
def denoise_loop_pcdms(unet, scheduler, *, latents, mask, simg_mask_latents, cond_pose, prompt_embeds,
                       negative_prompt_embeds, num_inference_steps, guidance_scale=2.0, dtype=torch.float32):
    """Demo driver loop, /root/reference/src/pipelines/PCDMs_pipeline.py:1062-1063 (CFG token batch = [negative ;
    positive]) and :1107-1150: the 9-channel input is assembled BEFORE the CFG duplication (:1115-1117), there is no
    class embedding, `cond_pose` ([1, C0, h, w]) broadcasts over the batch of 2 inside UNet.forward (:742)."""
    scheduler.set_timesteps(num_inference_steps)
    latents = latents.to(dtype)
    feature_f = torch.cat([negative_prompt_embeds, prompt_embeds]).to(dtype)
    for t in scheduler.timesteps:
        x = torch.cat([latents, mask, simg_mask_latents], dim=1).to(dtype)                # :1115
        x = torch.cat([x] * 2)                                                            # :1117
        x = scheduler.scale_model_input(x, t)                                             # :1118
        eps = unet(x, t, encoder_hidden_states=feature_f, my_pose_cond=cond_pose.to(dtype), return_dict=False)[0]
        eps = cfg_combine(eps, guidance_scale)                                            # :1133-1135
        latents = scheduler.step(eps, t, latents, return_dict=False)[0]                   # :1142
    return latents


@torch.no_grad()
def denoise_loop_stage3(unet, scheduler, *, latents, gen_t_img_latents, s_img_proj_f, num_inference_steps,
                        guidance_scale=2.0, dtype=torch.float32, unet_dtype=torch.float32):
    """/root/reference/src/pipelines/stage3_refined_pipeline.py:484-491 (CFG halves: zero tokens AND zero image
    latents for the unconditional one) and :530-556 (loop; 8-channel input `cat([latents x2, gen_t_img_f])`, :538).
    Only defined for bs * num_images_per_prompt == 1, like the reference.  `dtype` is the type of the conditioning and
    of the initial latents (fp16 in the reference, :487-491,:519); the UNet input and tokens are cast to `unet_dtype`
    (fp32 in the reference, :538,:542), so after the first scheduler step the latents live in the promoted type."""
    assert latents.shape[0] == 1 and s_img_proj_f.shape[0] == 1
    scheduler.set_timesteps(num_inference_steps)
    # :487-488 — the tokens keep the caller's dtype (the fp16 zeros are promoted by the cat)
    feature_f = torch.cat([torch.zeros(s_img_proj_f.shape, dtype=dtype), s_img_proj_f], dim=0)
    g = gen_t_img_latents.to(dtype)                                                       # vae.encode(x.half()), :479
    g = torch.cat([torch.zeros(g.shape, dtype=dtype), g], dim=0)                          # :490-491
    latents = latents.to(dtype)
    for t in scheduler.timesteps:
        x = torch.cat([latents] * 2)
        x = scheduler.scale_model_input(x, t)
        x8 = torch.cat([x, g], dim=1).to(unet_dtype)                                      # :538
        eps = unet(x8, t, encoder_hidden_states=feature_f.to(unet_dtype), return_dict=False)[0]   # :541-543
        eps = cfg_combine(eps, guidance_scale)
        latents = scheduler.step(eps, t, latents, return_dict=False)[0]                   # :556
    return latents
