"""Shim of diffusers.configuration_utils: ConfigMixin / register_to_config / FrozenDict (attribute-style config)."""
import functools
import inspect


class FrozenDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in self.items():
            object.__setattr__(self, k, v)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e


class ConfigMixin:
    config_name = None

    def register_to_config(self, **kwargs):
        cur = dict(getattr(self, "_internal_dict", {}))
        cur.update(kwargs)
        self._internal_dict = FrozenDict(cur)

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = [p for n, p in sig.parameters.items() if n != "self"]
        cfg = {p.name: p.default for p in params if p.default is not inspect.Parameter.empty}
        for p, a in zip(params, args):
            cfg[p.name] = a
        cfg.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        cfg.setdefault("_diffusers_version", "0.24.0")
        init(self, *args, **{k: v for k, v in kwargs.items() if not k.startswith("_")})
        merged = dict(cfg)
        merged.update(getattr(self, "_internal_dict", {}))  # values re-registered inside __init__ win
        self._internal_dict = FrozenDict(merged)

    return inner
