"""Shim of diffusers.pipelines.stable_diffusion.safety_checker (type annotation only)."""


class StableDiffusionSafetyChecker:
    pass
