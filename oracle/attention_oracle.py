"""CPU oracle for the attention ENTRY POINT (TEST INFRASTRUCTURE ONLY — see ltx_oracle.py header).

Restates `pay_attention` of utils/attention.py:161-398 (== wan/modules/attention.py) on its `sdpa` path, i.e. what the reference
computes when `offload.shared_state["_attention"] == "sdpa"` or a mask forces it (:179-180), with the calling conventions a
drop-in has to keep:
  * the caller's list [q, k, v] is emptied (:185-186), tensors are [B, tokens, heads, head_dim];
  * q and k are cast to v's dtype (:195-196), the result comes back in q's ORIGINAL dtype (:187, 394);
  * b > 1 with k_lens: runs of equal key length are attended separately on keys [:k_len] and concatenated (:197-226);
  * b == 1 with q_lens / k_lens: queries [:q_len] against keys [:k_len]; the output is padded back to lq rows whose content
    is uninitialised (:227-234, 395-396);
  * attention_mask is additive, [B, 1, H|1, Lk] on entry and transposed to [B, H|1, 1, Lk] (sdpa_wrapper :110-111);
  * softmax_scale is NOT honoured on this path (sdpa_wrapper never receives it): the scale is head_dim ** -0.5.
Pinned by oracle/gen_golden_attention.py against the unmodified reference function.
"""
from __future__ import annotations

from typing import List, Optional

import torch

Tensor = torch.Tensor


def _sdpa(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor]) -> Tensor:
    """[B, L, H, d] in / out; fp32 softmax, scale d^-0.5, additive mask broadcast to [B, H, Lq, Lk]."""
    d = q.shape[-1]
    qh, kh, vh = (t.transpose(1, 2).float() for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) * (d ** -0.5)
    if mask is not None:
        s = s + mask.transpose(1, 2).float()
    return (torch.softmax(s, dim=-1) @ vh).transpose(1, 2).to(v.dtype)


def pay_attention(qkv_list: List[Tensor], attention_mask: Optional[Tensor] = None, q_lens=None, k_lens=None) -> Tensor:
    q, k, v = qkv_list
    qkv_list.clear()
    out_dtype = q.dtype
    b, lq, lk = q.size(0), q.size(1), k.size(1)
    q, k = q.to(v.dtype), k.to(v.dtype)
    final_padding = 0
    if b > 1 and k_lens is not None:
        assert attention_mask is None and q_lens is None
        lens = [int(x) for x in k_lens]
        outs, i = [], 0
        while i < b:
            j = i
            while j < b and lens[j] == lens[i]:
                j += 1
            outs.append(_sdpa(q[i:j], k[i:j, : lens[i]], v[i:j, : lens[i]], None))
            i = j
        return torch.cat(outs, dim=0).to(out_dtype)
    if q_lens is not None or k_lens is not None:
        assert b == 1
        szq = int(q_lens[0]) if q_lens is not None else lq
        szk = int(k_lens[0]) if k_lens is not None else lk
        final_padding = lq - szq
        q, k, v = q[:, :szq], k[:, :szk], v[:, :szk]
    x = _sdpa(q, k, v, attention_mask).to(out_dtype)
    if final_padding > 0:
        x = torch.cat([x, torch.zeros(x.shape[0], final_padding, *x.shape[-2:], dtype=x.dtype)], 1)   # reference: torch.empty
    return x
