"""`torch.ops.ltxb200.*` — the hot-path kernels registered as torch custom ops (torch.library), each a thin call into the C ABI.

The drop-in classes call `ops.py` directly (same C entry points, no dispatcher round trip on the 389-launch step); this module
exposes the same kernels to code that wants dispatcher-visible ops: `torch.compile` graphs of a surrounding application (the ops
are opaque to the tracer and carry shape-only fake implementations), `torch.library.opcheck`, profilers that group by op name.
Functional forms only: every op returns a new tensor; the in-place / strided-output variants stay in `ops.py`.

    import ltx_video_gpupoor_b200.custom_ops          # registers the ops
    y = torch.ops.ltxb200.gemm(a, w, bias, act)        # act: 0 none, 1 gelu-tanh, 2 silu, 3 gelu-erf
    o = torch.ops.ltxb200.attention(q, k, v, key_bias)
"""
from typing import Optional

import torch
from torch import Tensor

from . import ops

BF16 = torch.bfloat16
_lib = torch.library


@_lib.custom_op("ltxb200::gemm", mutates_args=())
def gemm(a: Tensor, w: Tensor, bias: Optional[Tensor] = None, act: int = 0) -> Tensor:
    """(a @ w.T + bias) -> act; a [M, K], w [N, K] bf16 -> [M, N] bf16 (ltxb200_gemm_bf16)"""
    return ops.gemm(a, w, bias, act=act)


@gemm.register_fake
def _(a, w, bias=None, act=0):
    return a.new_empty(a.shape[0], w.shape[0])


@_lib.custom_op("ltxb200::gemm_gate_residual", mutates_args=())
def gemm_gate_residual(a: Tensor, w: Tensor, bias: Optional[Tensor], residual: Tensor, gate: Optional[Tensor], rows_per_gate: int) -> Tensor:
    """residual + gate[m // rows_per_gate] * (a @ w.T + bias) — the AdaLN-gated projection epilogue (ltxb200_gemm_bf16)"""
    return ops.gemm(a, w, bias, residual=residual, gate=gate, rows_per_gate=rows_per_gate)


@gemm_gate_residual.register_fake
def _(a, w, bias, residual, gate, rows_per_gate):
    return torch.empty_like(residual)


@_lib.custom_op("ltxb200::attention", mutates_args=())
def attention(q: Tensor, k: Tensor, v: Tensor, key_bias: Optional[Tensor] = None, scale: float = 0.0) -> Tensor:
    """softmax(scale * q k^T + key_bias) v; q [B, Lq, H, d], k / v [B, Lk, H, d] bf16, d in {64, 128} (ltxb200_attention_bf16)"""
    return ops.attention(q, k, v, key_bias=key_bias, scale=scale)


@attention.register_fake
def _(q, k, v, key_bias=None, scale=0.0):
    return q.new_empty(q.shape)


@_lib.custom_op("ltxb200::norm_mod", mutates_args=())
def norm_mod(x: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], rows_per_group: int, eps: float, layer_norm: bool) -> Tensor:
    """RMSNorm / LayerNorm without affine, then (1 + scale[g]) * y + shift[g] with g = row // rows_per_group (ltxb200_norm_mod_bf16)"""
    return ops.norm_mod(x, scale, shift, rows_per_group=rows_per_group, eps=eps, layer_norm=layer_norm)


@norm_mod.register_fake
def _(x, scale, shift, rows_per_group, eps, layer_norm):
    return torch.empty_like(x)


@_lib.custom_op("ltxb200::qk_norm_rope", mutates_args=())
def qk_norm_rope(qk: Tensor, wq: Tensor, wk: Tensor, cos: Tensor, sin: Tensor, tokens_per_batch: int, eps: float) -> Tensor:
    """qk [M, 2D] = q | k columns: RMSNorm over D with weights, then the interleaved LTX RoPE with bf16 tables [tokens, D]
    (ltxb200_qk_norm_rope_bf16 on a copy; ops.qk_norm_rope is the in-place form used by the model)"""
    out = qk.clone()
    D = qk.shape[1] // 2
    ops.qk_norm_rope(out[:, :D], out[:, D:], wq, wk, cos, sin, tokens_per_batch=tokens_per_batch, eps=eps)
    return out


@qk_norm_rope.register_fake
def _(qk, wq, wk, cos, sin, tokens_per_batch, eps):
    return torch.empty_like(qk)


@_lib.custom_op("ltxb200::conv3d", mutates_args=())
def conv3d(x: Tensor, w: Tensor, bias: Optional[Tensor], causal: bool) -> Tensor:
    """3x3x3 convolution, replicate padding in time / zero padding in space, NDHWC bf16; w [Cout, 27*Cin] tap-major (ltxb200_conv3d_bf16)"""
    return ops.conv3d(x, w, bias, causal=causal)


@conv3d.register_fake
def _(x, w, bias, causal):
    return x.new_empty(*x.shape[:4], w.shape[0])


OPS = ("gemm", "gemm_gate_residual", "attention", "norm_mod", "qk_norm_rope", "conv3d")
