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
        for name in ("unet", "prior"):
            m = getattr(self, name, None)
            if m is not None:
                return m.device
        return torch.device("cpu")

    def progress_bar(self, iterable=None, total=None):
        if iterable is not None:   # `for t in self.progress_bar(timesteps)` (stage1_prior_pipeline.py:456)
            return iterable
        return contextlib.nullcontext(_Bar())
