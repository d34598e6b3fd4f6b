import torch
from torch import nn


class ModelMixin(nn.Module):
    _supports_gradient_checkpointing = False

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    @property
    def dtype(self):
        try:
            return next(self.parameters()).dtype
        except StopIteration:
            return torch.float32

    def __getattr__(self, name):
        # diffusers ModelMixin falls back to config keys (deprecated direct access)
        try:
            return super().__getattr__(name)
        except AttributeError:
            d = self.__dict__.get("_internal_dict")
            if d is not None and name in d:
                return d[name]
            raise
