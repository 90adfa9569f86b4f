"""Shim of diffusers.utils."""
import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields, is_dataclass


class BaseOutput(OrderedDict):
    """dataclass-style output that also supports attribute access (like diffusers.utils.BaseOutput)."""

    def __post_init__(self):
        if is_dataclass(self):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    self[f.name] = v

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class _Logging:
    @staticmethod
    def get_logger(name=None):
        return _pylogging.getLogger(name or "diffusers")


logging = _Logging()


def deprecate(*args, **kwargs):
    return None


def is_accelerate_available():
    return False


def is_accelerate_version(*args, **kwargs):
    return False


USE_PEFT_BACKEND = False


def replace_example_docstring(doc):
    return lambda fn: fn


def scale_lora_layers(*args, **kwargs):
    return None


def unscale_lora_layers(*args, **kwargs):
    return None
