"""ORACLE (test infrastructure only) — CPU restatement of the diffusers-0.24.0 schedulers used on the path.

Call sites in the reference: scheduler.set_timesteps / scale_model_input / step at
/root/reference/src/pipelines/stage2_inpaint_pipeline.py:472,500,519; DDIM configuration from
/root/reference/pcdms_demo.ipynb:106-114 (scaled_linear 0.00085 -> 0.012, clip_sample=False, set_alpha_to_one=False,
steps_offset=1, 1000 train steps); DDPMScheduler.add_noise at /root/reference/stage2_train_inpaint_model.py:361.
The arithmetic itself is diffusers' (un-vendored, pinned 0.24.0, README.md:37) and is restated from the published
DDIM algorithm (Song et al. 2020, eq. 12; eta = 0 on the reference drivers' path, eta > 0 as the stochastic variant) as diffusers implements it.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


def scaled_linear_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class OracleDDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 clip_sample=False, set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon",
                 timestep_spacing="leading"):
        if beta_schedule != "scaled_linear" or prediction_type != "epsilon" or timestep_spacing != "leading":
            raise NotImplementedError
        if clip_sample:
            raise NotImplementedError("reference config has clip_sample=False")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, clip_sample=clip_sample,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                      prediction_type=prediction_type, timestep_spacing=timestep_spacing)
        self.alphas_cumprod = scaled_linear_alphas_cumprod(num_train_timesteps, beta_start, beta_end)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output=False, generator=None,
             variance_noise=None, return_dict: bool = True):
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta_t = 1 - a_t
        pred_x0 = (sample - beta_t ** 0.5 * model_output) / a_t ** 0.5
        # diffusers 0.24.0 scheduling_ddim.py: _get_variance and the eta > 0 branch of step()
        variance = (1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)
        std_dev_t = eta * variance ** 0.5
        pred_dir = (1 - a_prev - std_dev_t ** 2) ** 0.5 * model_output
        prev_sample = a_prev ** 0.5 * pred_x0 + pred_dir
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise.")
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev_sample = prev_sample + std_dev_t * variance_noise
        if not return_dict:
            return (prev_sample,)
        return SimpleNamespace(prev_sample=prev_sample, pred_original_sample=pred_x0)


def ddpm_add_noise(original_samples, noise, timesteps, alphas_cumprod=None):
    """DDPMScheduler.add_noise: sqrt(abar_t) x0 + sqrt(1 - abar_t) noise, abar cast to the sample dtype."""
    if alphas_cumprod is None:
        alphas_cumprod = scaled_linear_alphas_cumprod()
    ac = alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
    timesteps = timesteps.to(original_samples.device)
    sa = ac[timesteps] ** 0.5
    sb = (1 - ac[timesteps]) ** 0.5
    while sa.dim() < original_samples.dim():
        sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
    return sa * original_samples + sb * noise
