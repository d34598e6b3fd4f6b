import contextlib
from dataclasses import dataclass

import torch


@dataclass
class ImagePipelineOutput:
    images: object


class _Bar:
    def update(self, n=1):
        pass


class DiffusionPipeline:
    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def _execution_device(self):
        tr = getattr(self, "transformer", None)
        if tr is not None:
            return tr.device
        return torch.device("cpu")

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()

    def set_progress_bar_config(self, **kwargs):
        pass
