"""Scheduler boundary: `B200DDIMScheduler` / `B200DDPMScheduler` mirror the diffusers scheduler protocol the
reference's pipelines use — `.config` (steps_offset, clip_sample), `set_timesteps(n, device=)`, `.timesteps`,
`.init_noise_sigma`, `.order`, `scale_model_input(x, t)`, `step(model_output, t, sample, eta=0.0, generator=None,
return_dict=False)[0]` (/root/reference/src/pipelines/stage2_inpaint_pipeline.py:307-322,386,472-473,494,500,519;
`step`'s signature is introspected for `eta` / `generator` at :313-321) and `DDPMScheduler.add_noise`
(/root/reference/stage2_train_inpaint_model.py:361).  DDIM defaults are the reference demo's
(/root/reference/pcdms_demo.ipynb:106-114).

Host side only holds the schedule tables (a 1000-entry fp32 cumprod — not hot-path arithmetic); the per-element math
of `step` / `add_noise` runs in the CUDA kernels (pcdm_cfg_ddim_step / pcdm_add_noise).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

from . import ops


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule):
    if beta_schedule == "scaled_linear":
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    elif beta_schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    else:
        raise NotImplementedError(beta_schedule)
    return torch.cumprod(1.0 - betas, dim=0)


class B200DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 clip_sample=False, set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon",
                 timestep_spacing="leading"):
        if prediction_type != "epsilon" or timestep_spacing != "leading":
            raise NotImplementedError("B200DDIMScheduler: epsilon prediction with leading spacing only")
        if clip_sample:
            raise NotImplementedError("B200DDIMScheduler: clip_sample=True is not on the reference path")
        self.config = _AttrDict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                beta_schedule=beta_schedule, clip_sample=clip_sample,
                                set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                prediction_type=prediction_type, timestep_spacing=timestep_spacing)
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **kw):
        keys = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "clip_sample", "set_alpha_to_one",
                "steps_offset", "prediction_type", "timestep_spacing")
        d = {k: config[k] for k in keys if k in config}
        d.update(kw)
        return cls(**d)

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError("num_inference_steps larger than num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step_coefficients(self, timestep):
        """(1/sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)) as python floats computed from the fp32 table."""
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
        return (1.0 / math.sqrt(a_t), math.sqrt(1.0 - a_t), math.sqrt(a_prev), math.sqrt(1.0 - a_prev))

    def coefficient_table(self, device):
        """[num_inference_steps, 4] fp32 device table for the fused per-step kernel."""
        rows = [self.step_coefficients(t) for t in self.timesteps.tolist()]
        return torch.tensor(rows, dtype=torch.float32, device=device).contiguous()

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        """x_{t-1} = sqrt(a_prev) (x_t - sqrt(1-a_t) eps)/sqrt(a_t) + sqrt(1-a_prev) eps  (eta = 0): one launch of
        pcdm_ddim_step.  (The B200 pipeline itself uses the fused pcdm_cfg_ddim_step instead.)"""
        if eta != 0.0:
            raise NotImplementedError("B200DDIMScheduler: eta must be 0 (as in the reference drivers)")
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps() first")
        if not sample.is_cuda:
            raise RuntimeError("B200DDIMScheduler.step runs on CUDA tensors only (no CPU fallback)")
        c = self.step_coefficients(timestep)
        prev = ops.ddim_step(model_output.contiguous(), sample.contiguous(), c)
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev)

    def add_noise(self, original_samples, noise, timesteps):
        return _add_noise(self.alphas_cumprod, original_samples, noise, timesteps)

    def __len__(self):
        return self.config.num_train_timesteps


class B200DDPMScheduler:
    """Only what the training caller needs (stage2_train_inpaint_model.py:361): add_noise on the B200."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear"):
        self.config = _AttrDict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                beta_schedule=beta_schedule)
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule)

    def add_noise(self, original_samples, noise, timesteps):
        return _add_noise(self.alphas_cumprod, original_samples, noise, timesteps)


_ac_cache = {}


def _add_noise(alphas_cumprod, x0, noise, timesteps):
    if not x0.is_cuda:
        raise RuntimeError("pcdm_b200 add_noise runs on CUDA tensors only (no CPU fallback)")
    key = (x0.device, alphas_cumprod.data_ptr())
    ac = _ac_cache.get(key)
    if ac is None:
        ac = alphas_cumprod.to(device=x0.device, dtype=torch.float32).contiguous()
        _ac_cache[key] = ac
    return ops.add_noise(x0.contiguous(), noise.contiguous(), ac, timesteps.to(device=x0.device, dtype=torch.int64))
