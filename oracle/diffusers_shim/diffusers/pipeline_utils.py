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

    def enable_xformers_memory_efficient_attention(self, attention_op=None):
        """diffusers 0.24.0 DiffusionPipeline.set_use_memory_efficient_attention_xformers: every registered component
        that is a torch.nn.Module is walked recursively (children()) and each module offering
        `set_use_memory_efficient_attention_xformers` is called; other component objects are left alone."""
        def walk(m):
            if hasattr(m, "set_use_memory_efficient_attention_xformers"):
                m.set_use_memory_efficient_attention_xformers(True, attention_op)
            for c in m.children():
                walk(c)
        for name in ("vae", "text_encoder", "tokenizer", "unet", "scheduler", "safety_checker", "feature_extractor"):
            m = getattr(self, name, None)
            if isinstance(m, torch.nn.Module):
                walk(m)

    def progress_bar(self, iterable=None, total=None):
        if iterable is not None:   # `for t in self.progress_bar(timesteps)` (stage1_prior_pipeline.py:456)
            return iterable
        return contextlib.nullcontext(_Bar())
