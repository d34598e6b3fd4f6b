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


class PeerExchange:
    """The same exchange without a collective call: q/k/v head groups and attention outputs are stored straight into
    the destination rank's buffers by the kernels that produce them (NVLink P2P stores through cudaIpc mappings, flag
    per (destination, source) pair, see csrc/comm.cuh).  One instance per (group, B, n_loc, H, d); buffers are
    double-buffered by call parity.  torch.distributed is used once, at construction, to swap the 64-byte IPC handles.

      recv[par]  [N, B, 3, H/P, d]   filled by every rank's qk_norm_rope_wan_scatter kernel
      back[par]  [B*n_loc, H*d]      filled by every rank's attention epilogue
      flags      [2 exchanges][2 parities][P] uint32 epochs, counters [2] (local)
    """

    def __init__(self, group, B: int, n_loc: int, H: int, d: int, device):
        import ctypes
        import torch.distributed as dist
        from ... import _lib
        self.lib = _lib.lib()
        self._check = _lib.check
        self.group, self.P, self.rank = group, dist.get_world_size(group), dist.get_rank(group)
        P = self.P
        assert H % P == 0 and P <= 8
        self.B, self.n_loc, self.H, self.d, self.Hp, self.N = B, n_loc, H, d, H // P, P * n_loc
        self.device = device
        recv_b = self.N * B * 3 * self.Hp * d * 2
        back_b = B * n_loc * H * d * 2
        self.sizes = dict(recv0=recv_b, recv1=recv_b, back0=back_b, back1=back_b, ctl=4096)
        self.local, handles = {}, {}
        for name, nbytes in self.sizes.items():
            ptr = ctypes.c_void_p()
            h = ctypes.create_string_buffer(64)
            self._check(self.lib.ltxb200_comm_alloc(nbytes, ctypes.byref(ptr), h), "comm_alloc")
            self.local[name], handles[name] = ptr.value, h.raw
        gathered = [None] * P
        dist.all_gather_object(gathered, handles, group=group)
        self.peer = {name: [] for name in self.sizes}          # name -> P mapped pointers
        self._opened = []
        for r in range(P):
            for name in self.sizes:
                if r == self.rank:
                    self.peer[name].append(self.local[name])
                else:
                    ptr = ctypes.c_void_p()
                    self._check(self.lib.ltxb200_comm_open(gathered[r][name], ctypes.byref(ptr)), "comm_open")
                    self.peer[name].append(ptr.value)
                    self._opened.append(ptr.value)
        dist.barrier(group=group)
        VP = ctypes.c_void_p * P
        self._arr = lambda vals: VP(*vals)
        # flag words: exchange e (0 = qkv scatter, 1 = attention return), parity par -> ctl + (e*2+par)*64 bytes; counters after
        self.flag_off = lambda e, par: (e * 2 + par) * 64
        self.counter_off = lambda e: 1024 + e * 64
        self.calls = 0

    def _view(self, name, shape):
        """torch view of a LOCAL comm buffer (no copy): wraps the raw pointer through the CUDA array interface."""
        n = 1
        for s in shape:
            n *= s

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (n,), "typestr": "<u2", "data": (self.local[name], False), "version": 2}
        return torch.as_tensor(raw, device=self.device).view(torch.bfloat16).view(*shape)

    def self_attention(self, qkv, wq, wk, cos, sin, eps, stream_ptr):
        """qkv [B*n_loc, 3*H*d] raw projection rows -> attention output rows [B*n_loc, H*d] (view of back[par])."""
        par = self.calls & 1
        epoch = self.calls // 2 + 1
        self.calls += 1
        B, n_loc, H, d, Hp, N, P = self.B, self.n_loc, self.H, self.d, self.Hp, self.N, self.P
        D = H * d
        lib, ctl = self.lib, "ctl"
        f0 = self._arr([p + self.flag_off(0, par) for p in self.peer[ctl]])
        f1 = self._arr([p + self.flag_off(1, par) for p in self.peer[ctl]])
        recv, back = f"recv{par}", f"back{par}"
        from ... import ops
        with ops._Prof("qk_norm_rope_wan_scatter_bf16", "byte", 2.0 * 2 * B * n_loc * 3 * D):
          self._check(lib.ltxb200_qk_norm_rope_wan_scatter_bf16(
            qkv.data_ptr(), qkv.stride(0), B * n_loc, D, wq.data_ptr(), wk.data_ptr(), cos.data_ptr(), sin.data_ptr(), d,
            n_loc, self.rank * n_loc, float(eps), B, P, self.rank, self._arr(self.peer[recv]), f0, epoch,
            self.local[ctl] + self.counter_off(0), stream_ptr), "qk_norm_rope_wan_scatter")
        with ops._Prof("comm_wait", "byte", 0.0):
            self._check(lib.ltxb200_comm_wait(self.local[ctl] + self.flag_off(0, par), P, epoch, stream_ptr), "comm_wait")
        base = self.local[recv]
        tok = B * 3 * Hp * d                    # elements per global token in recv
        q, k, v = base, base + Hp * d * 2, base + 2 * Hp * d * 2
        with ops._Prof("attention_bf16", "flop", 4.0 * B * Hp * N * N * d):
          self._check(lib.ltxb200_attention_scatter_bf16(
            q, tok, 3 * Hp * d, k, tok, 3 * Hp * d, v, tok, 3 * Hp * d, D, B, Hp, N, N, d, 0.0, None, P, self.rank,
            self._arr(self.peer[back]), f1, epoch, self.local[ctl] + self.counter_off(1), n_loc, self.rank * Hp, stream_ptr),
            "attention_scatter")
        with ops._Prof("comm_wait", "byte", 0.0):
            self._check(lib.ltxb200_comm_wait(self.local[ctl] + self.flag_off(1, par), P, epoch, stream_ptr), "comm_wait")
        return self._view(back, (B * n_loc, D))

    def close(self):
        for p in self._opened:
            self.lib.ltxb200_comm_close(p)
        for p in self.local.values():
            self.lib.ltxb200_comm_free(p)
        self._opened, self.local = [], {}
