"""Which part of a skipped transformer block the spatio-temporal-guidance pass bypasses (API of ltx_video/utils/skip_layer_strategy.py:4:
same member names, order and values 1..4, so the reference's `stg_mode` mapping in ltxv.py:399-407 keeps working).

  AttentionSkip     the self-attention output of the perturbed rows is replaced by the attention INPUT   (transformer3d.py, stg_blend)
  AttentionValues   ... by the projected VALUES (the released presets' "attention_values")
  Residual          no effect in the reference for this model (needs attn.residual_connection, never set) and none here
  TransformerBlock  the whole block is bypassed for the perturbed rows
"""
from enum import Enum

SkipLayerStrategy = Enum("SkipLayerStrategy", ["AttentionSkip", "AttentionValues", "Residual", "TransformerBlock"], module=__name__)
