"""Minimal stand-in for the `diffusers` package (TEST INFRASTRUCTURE ONLY).

The reference (/root/reference) imports diffusers>=0.31 for a handful of small
modules (AdaLayerNormSingle, RMSNorm, GELU, PixArtAlpha* embeddings, mixins).
diffusers is not installed in this image and there is no network, so this shim
restates the published diffusers v0.31 semantics of exactly those pieces so that
the UNMODIFIED reference files can be imported on CPU to generate golden
vectors (oracle/gen_golden.py).  Nothing in the product path imports this.
State-dict key names follow diffusers (emb.timestep_embedder.linear_1, ...).
"""
from .configuration_utils import ConfigMixin, register_to_config  # noqa
from .models.modeling_utils import ModelMixin  # noqa
from .models import AutoencoderKL  # noqa
