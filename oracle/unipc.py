"""TEST INFRASTRUCTURE ONLY — CPU restatement of diffusers 0.24.0 `UniPCMultistepScheduler`, the scheduler the
reference's batch-test drivers install (/root/reference/stage2_batchtest_inpaint_model.py:132,
stage3_batchtest_refined_model.py: `UniPCMultistepScheduler.from_config(pipe.scheduler.config)`; the pipeline calls it
at /root/reference/src/pipelines/stage2_inpaint_pipeline.py:472,500,519).

diffusers is not vendored in /root/reference and cannot be installed here, so this file restates the published
algorithm (Zhao et al., "UniPC", 2023; diffusers `schedulers/scheduling_unipc_multistep.py` at v0.24.0, sigma-based
formulation) operation by operation in torch fp32.  PARITY UNPINNED against diffusers itself; what pins it
(tests/test_unipc.py): (1) the order-1 predictor step is algebraically the DDIM step, checked against
oracle.schedulers.DDIMScheduler; (2) on an analytic Gaussian diffusion (exact epsilon known in closed form) the
sampler converges to the exact probability-flow solution with order >= 2 as the step count grows; (3) timestep /
sigma tables against closed forms.

`from_config` takes the SD-2.1-base scheduler config the reference passes (PNDM config: scaled_linear betas
0.00085..0.012, steps_offset 1, timestep_spacing "leading") — see SURVEY.md App. A.6.
"""
from __future__ import annotations

import numpy as np
import torch


class UniPCMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 solver_order=2, prediction_type="epsilon", thresholding=False, predict_x0=True, solver_type="bh2",
                 lower_order_final=True, disable_corrector=(), timestep_spacing="linspace", steps_offset=0):
        if beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        assert prediction_type == "epsilon" and not thresholding and solver_type in ("bh1", "bh2")
        self.config = dict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                           beta_schedule=beta_schedule, solver_order=solver_order, prediction_type=prediction_type,
                           predict_x0=predict_x0, solver_type=solver_type, lower_order_final=lower_order_final,
                           disable_corrector=list(disable_corrector), timestep_spacing=timestep_spacing,
                           steps_offset=steps_offset)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alpha_t = torch.sqrt(self.alphas_cumprod)
        self.sigma_t = torch.sqrt(1 - self.alphas_cumprod)
        self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
        self.init_noise_sigma = 1.0
        self.predict_x0 = predict_x0
        self.num_inference_steps = None
        ts = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=np.float32)[::-1].copy()
        self.timesteps = torch.from_numpy(ts)
        self.model_outputs = [None] * solver_order
        self.timestep_list = [None] * solver_order
        self.lower_order_nums = 0
        self.disable_corrector = list(disable_corrector)
        self.last_sample = None
        self._step_index = None

    @classmethod
    def from_config(cls, config, **kw):
        keys = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "solver_order", "prediction_type",
                "predict_x0", "solver_type", "lower_order_final", "timestep_spacing", "steps_offset")
        config = {k: config[k] for k in config.keys()} if hasattr(config, "keys") else dict(vars(config))
        d = {k: config[k] for k in keys if k in config}
        d.update(kw)
        return cls(**d)

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps, device=None):
        c = self.config
        if c["timestep_spacing"] == "linspace":
            timesteps = (np.linspace(0, c["num_train_timesteps"] - 1, num_inference_steps + 1)
                         .round()[::-1][:-1].copy().astype(np.int64))
        elif c["timestep_spacing"] == "leading":
            step_ratio = c["num_train_timesteps"] // (num_inference_steps + 1)
            timesteps = (np.arange(0, num_inference_steps + 1) * step_ratio).round()[::-1][:-1].copy().astype(np.int64)
            timesteps += c["steps_offset"]
        elif c["timestep_spacing"] == "trailing":
            step_ratio = c["num_train_timesteps"] / num_inference_steps
            timesteps = np.arange(c["num_train_timesteps"], 0, -step_ratio).round().copy().astype(np.int64)
            timesteps -= 1
        else:
            raise ValueError(c["timestep_spacing"])
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        sigma_last = ((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5
        sigmas = np.concatenate([sigmas, [sigma_last]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None] * c["solver_order"]
        self.lower_order_nums = 0
        self.last_sample = None
        self._step_index = None

    def scale_model_input(self, sample, *a, **k):
        return sample

    @staticmethod
    def _sigma_to_alpha_sigma_t(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        sigma_t = sigma * alpha_t
        return alpha_t, sigma_t

    def convert_model_output(self, model_output, sample):
        sigma = self.sigmas[self.step_index]
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(sigma)
        if self.predict_x0:
            return (sample - sigma_t * model_output) / alpha_t
        return model_output

    def _bh_terms(self, h, rks, order):
        hh = -h if self.predict_x0 else h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        factorial_i = 1
        B_h = hh if self.config["solver_type"] == "bh1" else torch.expm1(hh)
        R, b = [], []
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * factorial_i / B_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        return h_phi_1, B_h, torch.stack(R), torch.tensor(b)

    def multistep_uni_p_bh_update(self, model_output, sample, order):
        m0 = self.model_outputs[-1]
        x = sample
        sigma_t, sigma_s0 = self.sigmas[self.step_index + 1], self.sigmas[self.step_index]
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(sigma_t)
        alpha_s0, sigma_s0 = self._sigma_to_alpha_sigma_t(sigma_s0)
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        rks, D1s = [], []
        for i in range(1, order):
            si = self.step_index - i
            mi = self.model_outputs[-(i + 1)]
            alpha_si, sigma_si = self._sigma_to_alpha_sigma_t(self.sigmas[si])
            lambda_si = torch.log(alpha_si) - torch.log(sigma_si)
            rk = (lambda_si - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks)
        h_phi_1, B_h, R, b = self._bh_terms(h, rks, order)
        if len(D1s) > 0:
            D1s = torch.stack(D1s, dim=1)
            if order == 2:
                rhos_p = torch.tensor([0.5], dtype=x.dtype)
            else:
                rhos_p = torch.linalg.solve(R[:-1, :-1], b[:-1])
        else:
            D1s = None
        if self.predict_x0:
            x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
            pred_res = torch.einsum("k,bkc...->bc...", rhos_p, D1s) if D1s is not None else 0
            x_t = x_t_ - alpha_t * B_h * pred_res
        else:
            x_t_ = alpha_t / alpha_s0 * x - sigma_t * h_phi_1 * m0
            pred_res = torch.einsum("k,bkc...->bc...", rhos_p, D1s) if D1s is not None else 0
            x_t = x_t_ - sigma_t * B_h * pred_res
        return x_t.to(x.dtype)

    def multistep_uni_c_bh_update(self, this_model_output, last_sample, this_sample, order):
        m0 = self.model_outputs[-1]
        x = last_sample
        model_t = this_model_output
        sigma_t, sigma_s0 = self.sigmas[self.step_index], self.sigmas[self.step_index - 1]
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(sigma_t)
        alpha_s0, sigma_s0 = self._sigma_to_alpha_sigma_t(sigma_s0)
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        rks, D1s = [], []
        for i in range(1, order):
            si = self.step_index - (i + 1)
            mi = self.model_outputs[-(i + 1)]
            alpha_si, sigma_si = self._sigma_to_alpha_sigma_t(self.sigmas[si])
            lambda_si = torch.log(alpha_si) - torch.log(sigma_si)
            rk = (lambda_si - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks)
        h_phi_1, B_h, R, b = self._bh_terms(h, rks, order)
        D1s = torch.stack(D1s, dim=1) if len(D1s) > 0 else None
        if order == 1:
            rhos_c = torch.tensor([0.5], dtype=x.dtype)
        else:
            rhos_c = torch.linalg.solve(R, b)
        if self.predict_x0:
            x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
            corr_res = torch.einsum("k,bkc...->bc...", rhos_c[:-1], D1s) if D1s is not None else 0
            D1_t = model_t - m0
            x_t = x_t_ - alpha_t * B_h * (corr_res + rhos_c[-1] * D1_t)
        else:
            x_t_ = alpha_t / alpha_s0 * x - sigma_t * h_phi_1 * m0
            corr_res = torch.einsum("k,bkc...->bc...", rhos_c[:-1], D1s) if D1s is not None else 0
            D1_t = model_t - m0
            x_t = x_t_ - sigma_t * B_h * (corr_res + rhos_c[-1] * D1_t)
        return x_t.to(x.dtype)

    def _init_step_index(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(self.timesteps.device)
        idx = (self.timesteps == timestep).nonzero()
        if len(idx) == 0:
            step_index = len(self.timesteps) - 1
        elif len(idx) > 1:
            step_index = idx[1].item()
        else:
            step_index = idx[0].item()
        self._step_index = step_index

    def step(self, model_output, timestep, sample, return_dict=True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        if self.step_index is None:
            self._init_step_index(timestep)
        c = self.config
        use_corrector = (self.step_index > 0 and self.step_index - 1 not in self.disable_corrector
                         and self.last_sample is not None)
        model_output_convert = self.convert_model_output(model_output, sample=sample)
        if use_corrector:
            sample = self.multistep_uni_c_bh_update(this_model_output=model_output_convert,
                                                    last_sample=self.last_sample, this_sample=sample,
                                                    order=self.this_order)
        for i in range(c["solver_order"] - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
            self.timestep_list[i] = self.timestep_list[i + 1]
        self.model_outputs[-1] = model_output_convert
        self.timestep_list[-1] = timestep
        if c["lower_order_final"]:
            this_order = min(c["solver_order"], len(self.timesteps) - self.step_index)
        else:
            this_order = c["solver_order"]
        self.this_order = min(this_order, self.lower_order_nums + 1)
        assert self.this_order > 0
        self.last_sample = sample
        prev_sample = self.multistep_uni_p_bh_update(model_output=model_output, sample=sample, order=self.this_order)
        if self.lower_order_nums < c["solver_order"]:
            self.lower_order_nums += 1
        self._step_index += 1
        if not return_dict:
            return (prev_sample,)
        from types import SimpleNamespace
        return SimpleNamespace(prev_sample=prev_sample)
