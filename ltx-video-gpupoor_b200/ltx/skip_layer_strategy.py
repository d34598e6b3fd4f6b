"""ltx_video/utils/skip_layer_strategy.py:4 — same names and order."""
from enum import Enum, auto


class SkipLayerStrategy(Enum):
    AttentionSkip = auto()
    AttentionValues = auto()
    Residual = auto()
    TransformerBlock = auto()
