"""Shim of diffusers.models.attention (stage1_prior_transformer.py:10)."""
from oracle.blocks import BasicTransformerBlock  # noqa: F401
