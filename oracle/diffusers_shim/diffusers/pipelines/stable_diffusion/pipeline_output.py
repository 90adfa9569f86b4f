"""Shim of diffusers.pipelines.stable_diffusion.pipeline_output."""
from dataclasses import dataclass
from typing import Any

from ...utils import BaseOutput


@dataclass
class StableDiffusionPipelineOutput(BaseOutput):
    images: Any = None
    nsfw_content_detected: Any = None
