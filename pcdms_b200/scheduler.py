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


def _config_dict(config):
    """A scheduler config as a plain dict: diffusers hands a FrozenDict (a mapping), some callers an attribute bag."""
    if hasattr(config, "keys"):
        return {k: config[k] for k in config.keys()}
    return dict(vars(config))


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
    elif beta_schedule == "squaredcos_cap_v2":   # diffusers betas_for_alpha_bar (stage-1 training: stage1_train_prior_model.py:155)
        def abar(x):
            return math.cos((x + 0.008) / 1.008 * math.pi / 2) ** 2
        n = num_train_timesteps
        betas = torch.tensor([min(1 - abar((i + 1) / n) / abar(i / n), 0.999) for i in range(n)], dtype=torch.float32)
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
        config = _config_dict(config)
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
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps() first")
        ops.require_cuda(sample, "B200DDIMScheduler.step")
        c = self.step_coefficients(timestep)
        if eta != 0.0:
            # the stochastic step of diffusers' DDIMScheduler (eta > 0; the reference pipelines forward `eta` and
            # `generator` to step(), stage2_inpaint_pipeline.py:307-322,519; its drivers leave eta = 0):
            # sigma = eta sqrt((1-a_prev)/(1-a_t)) sqrt(1 - a_t/a_prev), direction sqrt(1 - a_prev - sigma^2) eps, noise drawn
            # like diffusers' randn_tensor (on the generator's device, in the model output's dtype)
            if use_clipped_model_output:
                raise NotImplementedError("B200DDIMScheduler: use_clipped_model_output (clip_sample is off on this path)")
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either `generator` or"
                                 " `variance_noise` stays `None`.")
            a_t, a_prev = 1.0 - c[1] ** 2, c[2] ** 2
            sigma = eta * math.sqrt((1.0 - a_prev) / (1.0 - a_t)) * math.sqrt(max(1.0 - a_t / a_prev, 0.0))
            if variance_noise is None:
                gdev = generator.device if generator is not None else sample.device
                variance_noise = torch.randn(model_output.shape, generator=generator, device=gdev,
                                             dtype=model_output.dtype).to(sample.device)
            prev = ops.ddim_step_eta(model_output.contiguous(), sample.contiguous(),
                                     variance_noise.to(sample.dtype).contiguous(), c,
                                     math.sqrt(max(1.0 - a_prev - sigma * sigma, 0.0)), sigma)
        else:
            prev = ops.ddim_step(model_output.contiguous(), sample.contiguous(), c)
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev)

    def add_noise(self, original_samples, noise, timesteps):
        return _add_noise(self.alphas_cumprod, original_samples, noise, timesteps)

    def __len__(self):
        return self.config.num_train_timesteps


class B200DDPMScheduler:
    """Only what the training callers need: add_noise on the B200 (stage2_train_inpaint_model.py:361 with the SD
    "scaled_linear" schedule; stage1_train_prior_model.py:155,287 with `beta_schedule="squaredcos_cap_v2",
    prediction_type="sample"`)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 prediction_type="epsilon"):
        self.config = _AttrDict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                beta_schedule=beta_schedule, prediction_type=prediction_type)
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule)

    def add_noise(self, original_samples, noise, timesteps):
        return _add_noise(self.alphas_cumprod, original_samples, noise, timesteps)


class B200UniPCMultistepScheduler:
    """diffusers `UniPCMultistepScheduler` surface (the scheduler the reference's batch-test drivers install:
    /root/reference/stage2_batchtest_inpaint_model.py:132; protocol use at
    /root/reference/src/pipelines/stage2_inpaint_pipeline.py:472,500,519): predict_x0, bh1/bh2, solver_order <= 2,
    epsilon prediction, corrector on every step but the first.

    The host only derives the per-step scalars of the multistep update from the noise schedule (fp32 torch scalar
    arithmetic in the published order: sigma -> (alpha_t, sigma_t) -> lambda -> h, rk, h_phi_1, B_h, rho);
    `coefficient_table` lays them out as 16 floats per step for `pcdm_cfg_unipc_step` (fused, graph-replayable) and
    `step` feeds one row to `pcdm_unipc_step`.  Sample history (last_sample, two converted model outputs) lives in
    fp32 device buffers."""
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, solver_order=2, prediction_type="epsilon", thresholding=False,
                 dynamic_thresholding_ratio=0.995, sample_max_value=1.0, predict_x0=True, solver_type="bh2",
                 lower_order_final=True, disable_corrector=(), solver_p=None, use_karras_sigmas=False,
                 timestep_spacing="linspace", steps_offset=0):
        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"B200UniPCMultistepScheduler: {what}")
        need(trained_betas is None and solver_p is None and not use_karras_sigmas and not thresholding,
             "trained_betas / solver_p / karras sigmas / thresholding are not on the reference path")
        need(prediction_type == "epsilon" and predict_x0, "epsilon prediction with predict_x0 only")
        need(solver_type in ("bh1", "bh2") and solver_order in (1, 2), "solver bh1/bh2 with solver_order <= 2")
        need(timestep_spacing in ("linspace", "leading", "trailing"), f"timestep_spacing {timestep_spacing}")
        self.config = _AttrDict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                beta_schedule=beta_schedule, solver_order=solver_order,
                                prediction_type=prediction_type, thresholding=thresholding, predict_x0=predict_x0,
                                solver_type=solver_type, lower_order_final=lower_order_final,
                                disable_corrector=list(disable_corrector), use_karras_sigmas=use_karras_sigmas,
                                timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps,
                                                      dtype=np.float32)[::-1].copy())
        self.sigmas = None
        self._step_index = None
        self._hist = None
        self._rows = None

    @classmethod
    def from_config(cls, config, **kw):
        """`UniPCMultistepScheduler.from_config(pipe.scheduler.config)`: keys of another scheduler's config that this
        one does not know (skip_prk_steps, set_alpha_to_one, clip_sample, ...) are ignored, as diffusers does."""
        import inspect
        known = set(inspect.signature(cls.__init__).parameters) - {"self"}
        d = {k: v for k, v in _config_dict(config).items() if k in known}
        d.update(kw)
        return cls(**d)

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        T = c.num_train_timesteps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            ratio = T // (num_inference_steps + 1)
            ts = (np.arange(0, num_inference_steps + 1) * ratio).round()[::-1][:-1].copy().astype(np.int64)
            ts += c.steps_offset
        else:
            ts = np.arange(T, 0, -T / num_inference_steps).round().copy().astype(np.int64) - 1
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        sig_last = ((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5
        self.sigmas = torch.from_numpy(np.concatenate([sig, [sig_last]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(ts)
        self._step_index = None
        self._hist = None
        self._rows = None

    def scale_model_input(self, sample, *a, **k):
        return sample

    # -- schedule-only scalars ------------------------------------------------------------------------------------
    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def _lambda(self, i):
        a, s = self._alpha_sigma(self.sigmas[i])
        return torch.log(a) - torch.log(s)

    def _update_scalars(self, i_t, i_s0, i_s1, order, corrector):
        """(sigma_t/sigma_s0, alpha_t h_phi_1, alpha_t B_h, rk, rho_0, rho_last) of one bh update from sigma index
        i_s0 to i_t; i_s1 is the index of the older model output (order 2)."""
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i_t])
        _, sigma_s0 = self._alpha_sigma(self.sigmas[i_s0])
        lam_s0 = self._lambda(i_s0)
        h = self._lambda(i_t) - lam_s0
        rks = []
        if order == 2:
            rks.append((self._lambda(i_s1) - lam_s0) / h)
        rks.append(1.0)
        rks_t = torch.tensor(rks)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        fact = 1
        B_h = hh if self.config.solver_type == "bh1" else torch.expm1(hh)
        R, b = [], []
        for k in range(1, order + 1):
            R.append(torch.pow(rks_t, k - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= k + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        if corrector:
            rhos = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(torch.stack(R), torch.tensor(b))
        else:
            rhos = torch.tensor([0.5])   # order-2 predictor: the simplified rho = 1/2; unused at order 1
        rk = float(rks[0]) if order == 2 else 1.0
        rho0 = float(rhos[0]) if (order == 2) else 0.0
        return (float(sigma_t / sigma_s0), float(alpha_t * h_phi_1), float(alpha_t * B_h), rk, rho0, float(rhos[-1]))

    def _orders(self):
        N, so = self.num_inference_steps, self.config.solver_order
        out = []
        for i in range(N):
            o = min(so, N - i) if self.config.lower_order_final else so
            out.append(min(o, min(i, so) + 1))
        return out

    def coefficient_rows(self):
        """[steps][16] python floats (layout: include/pcdm_b200.h, pcdm_cfg_unipc_step)."""
        if self._rows is not None:
            return self._rows
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps() first")
        orders = self._orders()
        rows = []
        for i in range(self.num_inference_steps):
            alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i])
            row = [float(sigma_t), float(alpha_t)]
            use_c = i > 0 and (i - 1) not in self.config.disable_corrector
            if use_c:
                oc = orders[i - 1]
                a, bq, cq, rk, rho0, rho_last = self._update_scalars(i, i - 1, i - 2, oc, corrector=True)
                row += [1.0, a, bq, cq, rk, rho0, rho_last, float(oc)]
            else:
                row += [0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 1.0]
            op = orders[i]
            a, bq, cq, rk, _, rho = self._update_scalars(i + 1, i, i - 1, op, corrector=False)
            row += [a, bq, cq, rk, rho, float(op)]
            rows.append(row)
        self._rows = rows
        return rows

    def coefficient_table(self, device):
        return torch.tensor(self.coefficient_rows(), dtype=torch.float32, device=device).contiguous()

    # -- scheduler protocol ---------------------------------------------------------------------------------------
    def _init_step_index(self, timestep):
        t = int(timestep)
        idx = (self.timesteps.cpu() == t).nonzero()
        if len(idx) == 0:
            self._step_index = len(self.timesteps) - 1
        elif len(idx) > 1:
            self._step_index = int(idx[1])
        else:
            self._step_index = int(idx[0])

    def step(self, model_output, timestep, sample, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        ops.require_cuda(sample, "B200UniPCMultistepScheduler.step")
        if self._step_index is None:
            self._init_step_index(timestep)
        if self._hist is None or self._hist.shape[1] != sample.numel() or self._hist.device != sample.device:
            self._hist = torch.zeros((3, sample.numel()), dtype=torch.float32, device=sample.device)
        row = self.coefficient_rows()[self._step_index]
        prev = ops.unipc_step(model_output.contiguous(), sample.contiguous(), self._hist[0], self._hist[1],
                              self._hist[2], row)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev)

    def add_noise(self, original_samples, noise, timesteps):
        return _add_noise(self.alphas_cumprod, original_samples, noise, timesteps)

    def __len__(self):
        return self.config.num_train_timesteps


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


class B200UnCLIPScheduler:
    """diffusers `UnCLIPScheduler` surface — the scheduler `Stage1_PriorPipeline` samples the stage-1 prior with
    (/root/reference/src/pipelines/stage1_prior_pipeline.py:23,445-446,478-483; kandinsky-2-2-prior's
    scheduler_config.json: squaredcos_cap_v2 betas, prediction_type "sample", variance_type "fixed_small_log",
    clip_sample True with range 10).  Timesteps are spread evenly over [0, T-1] including both ends;
    `step(model_output, timestep, sample, prev_timestep=None, generator=None, return_dict=True)` is the DDPM
    posterior mean on the clipped x0 prediction plus "fixed_small_log" noise for t > 0.

    The host derives the schedule-only scalars of each step with the published fp32 tensor arithmetic
    (`step_coefficients`); the per-element update runs in `pcdm_unclip_step` / `pcdm_cfg_unclip_step`."""
    order = 1
    ROW = 8

    def __init__(self, num_train_timesteps=1000, variance_type="fixed_small_log", clip_sample=True,
                 clip_sample_range=1.0, prediction_type="epsilon", beta_schedule="squaredcos_cap_v2"):
        if beta_schedule != "squaredcos_cap_v2":
            raise ValueError("UnCLIPScheduler only supports `beta_schedule`: 'squaredcos_cap_v2'")
        if variance_type != "fixed_small_log":
            raise NotImplementedError("B200UnCLIPScheduler: variance_type 'fixed_small_log' only (no learned variance "
                                      "on the reference path)")
        if prediction_type not in ("sample", "epsilon"):
            raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon` or `sample`")
        self.config = _AttrDict(num_train_timesteps=num_train_timesteps, variance_type=variance_type,
                                clip_sample=clip_sample, clip_sample_range=clip_sample_range,
                                prediction_type=prediction_type, beta_schedule=beta_schedule)

        def abar(x):
            return math.cos((x + 0.008) / 1.008 * math.pi / 2) ** 2
        n = num_train_timesteps
        self.betas = torch.tensor([min(1 - abar((i + 1) / n) / abar(i / n), 0.999) for i in range(n)],
                                  dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, n)[::-1].copy())

    @classmethod
    def from_config(cls, config, **kw):
        keys = ("num_train_timesteps", "variance_type", "clip_sample", "clip_sample_range", "prediction_type",
                "beta_schedule")
        config = _config_dict(config)
        d = {k: config[k] for k in keys if k in config}
        d.update(kw)
        return cls(**d)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = (self.config.num_train_timesteps - 1) / (num_inference_steps - 1)
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def step_coefficients(self, timestep, prev_timestep=None):
        """The 8 floats of one `pcdm_unclip_step` row: {c_x0, c_xt, std, clip, sqrt(abar_t), sqrt(1 - abar_t),
        prediction is epsilon, 0}."""
        t = int(timestep)
        prev_t = t - 1 if prev_timestep is None else int(prev_timestep)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        if prev_t == t - 1:
            beta, alpha = self.betas[t], self.alphas[t]
        else:
            beta = 1 - a_t / a_prev
            alpha = 1 - beta
        c_x0 = (a_prev ** 0.5 * beta) / b_t
        c_xt = alpha ** 0.5 * b_prev / b_t
        std = 0.0
        if t > 0:
            std = float(torch.exp(0.5 * torch.log(torch.clamp(b_prev / b_t * beta, min=1e-20))))
        clip = float(self.config.clip_sample_range) if self.config.clip_sample else float("inf")
        return (float(c_x0), float(c_xt), std, clip, float(a_t ** 0.5), float(b_t ** 0.5),
                1.0 if self.config.prediction_type == "epsilon" else 0.0, 0.0)

    def coefficient_table(self, device):
        """[num_inference_steps, 8] fp32 device table for the fused per-step kernel (prev_timestep = the next entry of
        `timesteps`, None on the last step: stage1_prior_pipeline.py:473-476)."""
        ts = self.timesteps.tolist()
        rows = [self.step_coefficients(t, ts[i + 1] if i + 1 < len(ts) else None) for i, t in enumerate(ts)]
        return torch.tensor(rows, dtype=torch.float32, device=device).contiguous()

    def step(self, model_output, timestep, sample, prev_timestep=None, generator=None, return_dict: bool = True):
        if not sample.is_cuda:
            raise RuntimeError("B200UnCLIPScheduler.step runs on CUDA tensors only (no CPU fallback)")
        row = self.step_coefficients(timestep, prev_timestep)
        noise = None
        if row[2] != 0.0:
            noise = torch.randn(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
        prev = ops.unclip_step(model_output.contiguous(), sample.contiguous(), noise, row)
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev)

    def __len__(self):
        return self.config.num_train_timesteps
