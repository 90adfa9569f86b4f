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
