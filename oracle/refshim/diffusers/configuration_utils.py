import functools
import inspect


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        if not hasattr(self, "_internal_dict"):
            object.__setattr__(self, "_internal_dict", _AttrDict())
        self._internal_dict.update(kwargs)

    @property
    def config(self):
        if not hasattr(self, "_internal_dict"):
            object.__setattr__(self, "_internal_dict", _AttrDict())
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        config = dict(config)
        config.update(kwargs)
        sig = inspect.signature(cls.__init__).parameters
        init_kwargs = {k: v for k, v in config.items() if k in sig and not k.startswith("_")}
        return cls(**init_kwargs)


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k not in ("self", "kwargs")}
        init(self, *args, **kwargs)
        ConfigMixin.register_to_config(self, **cfg)

    return inner
