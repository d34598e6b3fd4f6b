"""Ulysses (DeepSpeed-style) head<->sequence exchange for sequence-parallel self-attention — the behaviour of
`xFuserLongContextAttention` as called at wan/distributed/xdit_context_parallel.py:179-184 (xfuser itself is
not vendored by the reference; semantics restated from its published design).

Layout choice: the send buffer is [P, n_loc, B, 3, H/P, d] (peer-major, then token-major).  After
`all_to_all_single` the receive buffer read as [N, B, 3, H/P, d] is already in GLOBAL token order, so q/k/v
are strided views the attention kernel consumes directly (token stride B*3*Hp*d) — no gather copy — and the
attention output written token-major [N, B, Hp, d] is already peer-major for the way back.
"""
from typing import Callable

import torch


def pack_qkv(qkv: torch.Tensor, B: int, n_loc: int, P: int, H: int, d: int) -> torch.Tensor:
    """qkv [B*n_loc, 3*H*d] (columns q|k|v, heads packed) -> send buffer [P, n_loc, B, 3, H/P, d]."""
    Hp = H // P
    return qkv.view(B, n_loc, 3, P, Hp, d).permute(3, 1, 0, 2, 4, 5).contiguous()


def qkv_views(recv: torch.Tensor, B: int, N: int, Hp: int, d: int):
    """recv [P, n_loc, B, 3, Hp, d] -> q, k, v as [B, N, Hp, d] strided views in global token order."""
    full = recv.view(N, B, 3, Hp, d)
    return tuple(full[:, :, i].permute(1, 0, 2, 3) for i in range(3))


def unpack_out(back: torch.Tensor, B: int, n_loc: int, P: int, Hp: int, d: int) -> torch.Tensor:
    """back [P(head group), n_loc, B, Hp, d] -> rows [B*n_loc, H*d] with heads in global order."""
    return back.view(P, n_loc, B, Hp, d).permute(2, 1, 0, 3, 4).reshape(B * n_loc, P * Hp * d)


def ulysses_self_attention(qkv: torch.Tensor, B: int, n_loc: int, H: int, d: int, group,
                           attn_fn: Callable[[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor], None]) -> torch.Tensor:
    """qkv: local tokens, all heads (q/k already normed + RoPE'd with global positions).  attn_fn(q, k, v, out)
    writes attention into `out` ([B, N, Hp, d] view).  Returns rows [B*n_loc, H*d] for the local tokens."""
    import torch.distributed as dist
    P = dist.get_world_size(group)
    Hp = H // P
    N = P * n_loc
    send = pack_qkv(qkv, B, n_loc, P, H, d)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    q, k, v = qkv_views(recv, B, N, Hp, d)
    o_tok = torch.empty(N, B, Hp, d, device=qkv.device, dtype=qkv.dtype)
    attn_fn(q, k, v, o_tok.permute(1, 0, 2, 3))
    back = torch.empty_like(o_tok)
    dist.all_to_all_single(back, o_tok, group=group)
    return unpack_out(back, B, n_loc, P, Hp, d)
