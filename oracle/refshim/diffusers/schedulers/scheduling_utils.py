from dataclasses import dataclass
from enum import Enum

import torch

from ..utils import BaseOutput


class KarrasDiffusionSchedulers(Enum):
    DDIMScheduler = 1


class SchedulerMixin:
    pass


@dataclass
class SchedulerOutput(BaseOutput):
    prev_sample: torch.Tensor
