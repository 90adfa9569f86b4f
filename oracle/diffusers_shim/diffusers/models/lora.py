"""Shim of diffusers.models.lora (LoRA is not on the path)."""


def adjust_lora_scale_text_encoder(*args, **kwargs):
    return None
