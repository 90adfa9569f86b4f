"""Shim of diffusers.models."""


class AutoencoderKL:  # type annotation only on the reference path
    pass


class UNet2DConditionModel:  # type annotation only (stage3_refined_pipeline.py:10,69)
    pass
