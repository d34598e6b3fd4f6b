import logging as _pylogging
from collections import OrderedDict


class BaseOutput(OrderedDict):
    def __post_init__(self):
        import dataclasses
        for f in dataclasses.fields(self):
            self[f.name] = getattr(self, f.name)

    def __getitem__(self, k):
        if isinstance(k, int):
            return list(self.values())[k]
        return super().__getitem__(k)

    def to_tuple(self):
        return tuple(self.values())


def is_torch_version(op, version):
    return True


def is_scipy_available():
    return True


def deprecate(*args, **kwargs):
    pass


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


logging = _Logging()
