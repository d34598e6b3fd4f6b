"""Attention entry point — drop-in for utils/attention.py (== wan/modules/attention.py) of the reference.

`pay_attention(qkv_list, ...)` keeps the reference's calling convention (utils/attention.py:161-398):
the caller hands over a Python list [q, k, v] that is cleared here (ownership transfer, :185-186);
tensors are [B, tokens, heads, head_dim]; the result is [B, Lq, H, d] in q's dtype.  There is exactly
one backend: the sm_100a tcgen05/TMEM flash kernel in libltx_b200.so — no sdpa/flash/sage/xformers
dispatch, no CPU fallback.  `offload.shared_state["_attention"]` is ignored on purpose.
"""
from typing import List, Optional

import torch

from . import ops

__all__ = ["pay_attention", "get_attention_modes", "get_supported_attention_modes"]


def get_attention_modes() -> List[str]:
    """utils/attention.py:119-130 — the reference lists installed backends; here there is one."""
    return ["b200"]


def get_supported_attention_modes() -> List[str]:
    """utils/attention.py:132-142"""
    return ["b200"]


def _key_bias_from_mask(attention_mask: torch.Tensor, B: int, H: int, Lk: int) -> torch.Tensor:
    """The reference passes an additive mask shaped [B, 1, H, Lk] (transposed to [B, H, 1, Lk] inside
    sdpa_wrapper, utils/attention.py:110-111).  The kernel takes a per-key bias [B, Lk]; masks that vary
    over heads or queries are not produced anywhere on the reference path and are rejected."""
    m = attention_mask
    if m.dim() == 4:
        if m.shape[1] != 1 and m.shape[2] == 1:
            m = m.transpose(1, 2)
        if m.shape[1] != 1:
            raise ValueError("pay_attention: per-query attention masks are not supported")
        if m.shape[2] > 1 and not bool((m[:, :, :1] == m).all()):
            raise ValueError("pay_attention: per-head attention masks are not supported")
        m = m[:, 0, 0, :]
    elif m.dim() == 3:
        m = m[:, 0, :]
    if m.shape != (B, Lk):
        m = m.expand(B, Lk)
    return m.to(torch.float32).contiguous()


@torch.compiler.disable()
def pay_attention(qkv_list, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1),
                  deterministic=False, version=None, force_attention=None, attention_mask=None,
                  cross_attn=False, q_lens=None, k_lens=None) -> torch.Tensor:
    if causal or dropout_p != 0.0 or window_size != (-1, -1):
        raise NotImplementedError("pay_attention: only dense non-causal attention exists on this path")
    q, k, v = qkv_list
    qkv_list.clear()
    out_dtype = q.dtype
    b, lq, lk = q.size(0), q.size(1), k.size(1)
    if q.dtype != torch.bfloat16:
        q = q.to(torch.bfloat16)
    k = k.to(q.dtype)
    v = v.to(q.dtype)

    # variable key lengths (utils/attention.py:197-237): batch > 1 -> runs of equal k_len are attended separately
    if b > 1 and k_lens is not None:
        assert attention_mask is None and q_lens is None
        outs = []
        i = 0
        lens = [int(x) for x in k_lens]
        while i < b:
            j = i
            while j < b and lens[j] == lens[i]:
                j += 1
            outs.append(pay_attention([q[i:j], k[i:j, :lens[i]], v[i:j, :lens[i]]]))
            i = j
        return torch.cat(outs, dim=0).to(out_dtype)
    final_padding = 0
    if q_lens is not None or k_lens is not None:
        assert b == 1
        szq = int(q_lens[0]) if q_lens is not None else lq
        szk = int(k_lens[0]) if k_lens is not None else lk
        final_padding = lq - szq
        q, k, v = q[:, :szq], k[:, :szk], v[:, :szk]

    H, d = q.shape[2], q.shape[3]
    fix = lambda t: t if (t.stride(3) == 1 and t.stride(2) == d) else t.contiguous()
    q, k, v = fix(q), fix(k), fix(v)
    bias = None if attention_mask is None else _key_bias_from_mask(attention_mask, b, H, k.shape[1])
    # like the sdpa path of the reference (sdpa_wrapper never receives softmax_scale, utils/attention.py:99-116, 283-288), the scale
    # is head_dim ** -0.5 whatever `softmax_scale` says — pinned by oracle/gen_golden_attention.py case "dtypes"
    out = ops.attention(q, k, v, key_bias=bias, scale=0.0)
    if final_padding > 0:   # utils/attention.py:395-396: tail is uninitialised padding
        out = torch.cat([out, torch.empty(b, final_padding, H, d, device=out.device, dtype=out.dtype)], dim=1)
    return out.to(out_dtype)
