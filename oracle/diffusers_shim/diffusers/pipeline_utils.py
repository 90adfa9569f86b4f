"""Shim of diffusers.DiffusionPipeline: module registry, device, progress bar."""
import contextlib

import torch


class _Bar:
    def update(self, n=1):
        return None


class DiffusionPipeline:
    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def device(self):
        unet = getattr(self, "unet", None)
        return unet.device if unet is not None else torch.device("cpu")

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()
