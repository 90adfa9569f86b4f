from oracle.blocks import Attention, AttentionProcessor, AttnProcessor  # noqa: F401
