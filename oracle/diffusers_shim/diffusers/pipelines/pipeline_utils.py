"""Shim of diffusers.pipelines.pipeline_utils."""
from ..pipeline_utils import DiffusionPipeline  # noqa: F401
