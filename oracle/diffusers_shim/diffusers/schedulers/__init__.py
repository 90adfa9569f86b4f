"""Shim of diffusers.schedulers: DDIM from the oracle; the other names exist only so the imports resolve."""
from oracle.schedulers import OracleDDIMScheduler as DDIMScheduler  # noqa: F401
from oracle.unipc import UniPCMultistepScheduler  # noqa: F401


class KarrasDiffusionSchedulers:
    pass


class _NotOnPath:
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: scheduler not restated")


class DPMSolverMultistepScheduler(_NotOnPath):
    pass


class EulerAncestralDiscreteScheduler(_NotOnPath):
    pass


class EulerDiscreteScheduler(_NotOnPath):
    pass


class LMSDiscreteScheduler(_NotOnPath):
    pass


class PNDMScheduler(_NotOnPath):
    pass


class UnCLIPScheduler:
    """diffusers' UnCLIPScheduler call surface (stage1_prior_pipeline.py:445-446,478-483) over the oracle restatement."""

    def __init__(self, **kw):
        from oracle.prior import UnCLIPScheduler as _Impl
        self._impl = _Impl(**kw)
        self.config = self._impl.config
        self.init_noise_sigma = self._impl.init_noise_sigma

    @property
    def timesteps(self):
        return self._impl.timesteps

    def set_timesteps(self, num_inference_steps, device=None):
        self._impl.set_timesteps(num_inference_steps, device=device)

    def step(self, model_output, timestep, sample, prev_timestep=None, generator=None, return_dict=True):
        from types import SimpleNamespace
        prev = self._impl.step(model_output, timestep, sample, prev_timestep=prev_timestep, generator=generator)
        return SimpleNamespace(prev_sample=prev) if return_dict else (prev,)
