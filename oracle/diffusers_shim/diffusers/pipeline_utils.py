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

    def register_to_config(self, **kwargs):
        from types import SimpleNamespace
        cfg = dict(getattr(self, "_config", {}))
        cfg.update(kwargs)
        self._config = cfg
        self.config = SimpleNamespace(**cfg)

    @property
    def _execution_device(self):
        return self.device

    def maybe_free_model_hooks(self):
        return None

    @property
    def device(self):
        unet = getattr(self, "unet", None)
        return unet.device if unet is not None else torch.device("cpu")

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()
