"""Shim of diffusers.utils.torch_utils.randn_tensor (CPU generator path)."""
import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    if isinstance(generator, list):
        raise NotImplementedError
    return torch.randn(shape, generator=generator, device="cpu", dtype=dtype).to(device or "cpu")
