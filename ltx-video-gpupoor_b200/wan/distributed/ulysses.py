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
      gath       [B, P, n_loc, out]  fp32 head output of every rank (the final all-gather, also peer stores)
      flags      [3 exchanges][2 parities][P] uint32 epochs, counters [3] (local), status word (device) + one pinned host word

    The way in CAN be issued in `chunks` token chunks (LTXB200_SP_CHUNKS > 1): chunk c's QKV projection on the caller's stream and
    chunk c-1's norm + RoPE + scatter on a side stream.  Measured and NOT the default: on one GPU the QKV GEMM (179 us) and the scatter
    (65 us) take 229-233 us on two streams against 244 us back to back (profiles/scripts/overlap_probe.py) — the GEMM runs at the
    L2 -> SM bandwidth cap, the scatter is pure memory traffic, so the two serialise on the memory system — and the Wan-1.3B step at
    2 / 4 GPUs is 0.4-1.3 ms SLOWER with 4 chunks (smaller GEMMs, more launches): profiles/r02_wan_sp_chunk_ab.md.  What did pay
    was making the scatter itself cheaper (bounded grid-stride grid, one RoPE table read per row: 11.8 -> 8.3 ms per step at SP2).

    Failure handling: every wait is bounded (csrc/comm.cuh); a peer that never publishes makes the waiting kernel record
    (source rank, epoch) in a host-visible word instead of spinning forever, and the next call here raises.
    """

    def __init__(self, group, B: int, n_loc: int, H: int, d: int, device, gather_width: int = 0):
        import ctypes
        import os
        import torch.distributed as dist
        from ... import _lib
        self.lib = _lib.lib()
        self._check = _lib.check
        self._err = _lib.LtxB200Error
        self.group, self.P, self.rank = group, dist.get_world_size(group), dist.get_rank(group)
        P = self.P
        assert H % P == 0 and P <= 8
        self.B, self.n_loc, self.H, self.d, self.Hp, self.N = B, n_loc, H, d, H // P, P * n_loc
        self.device = device
        self.gather_width = gather_width
        recv_b = self.N * B * 3 * self.Hp * d * 2
        back_b = B * n_loc * H * d * 2
        self.sizes = dict(recv0=recv_b, recv1=recv_b, back0=back_b, back1=back_b, ctl=4096)
        if gather_width:
            self.sizes["gath0"] = self.sizes["gath1"] = B * P * n_loc * gather_width * 4
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
        # flag words: exchange e (0 = qkv scatter, 1 = attention return, 2 = head gather), parity par -> ctl + (e*2+par)*64 bytes;
        # counters and the device status word after them
        self.flag_off = lambda e, par: (e * 2 + par) * 64
        self.counter_off = lambda e: 1024 + e * 64
        self.status_off = 2048
        self.status_host = torch.zeros(16, dtype=torch.int32).pin_memory()     # UVA: the kernel stores into it directly
        self.calls = 0
        self.gather_calls = 0
        self.chunks = max(1, int(os.environ.get("LTXB200_SP_CHUNKS", "1")))
        self.min_chunk_rows = 1024                 # below this a chunk's GEMM no longer fills the machine (tests lower it)
        self.v_from_gemm = os.environ.get("LTXB200_SP_V_FROM_GEMM", "0") == "1"
        self.side = torch.cuda.Stream(device=device)
        self._closed = False

    def _view(self, name, shape, dtype=torch.bfloat16):
        """torch view of a LOCAL comm buffer (no copy): wraps the raw pointer through the CUDA array interface."""
        n = 1
        for s in shape:
            n *= s
        item = torch.empty(0, dtype=dtype).element_size()

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (n,), "typestr": "<u2" if item == 2 else "<u4", "data": (self.local[name], False),
                                        "version": 2}
        return torch.as_tensor(raw, device=self.device).view(dtype).view(*shape)

    def check(self):
        """Raise if a bounded wait of an earlier call gave up (polls the pinned host word; no synchronisation)."""
        code = int(self.status_host[0])
        if code:
            raise self._err(f"sequence-parallel exchange timed out on rank {self.rank}: peer {(code & 0xff) - 1} never published epoch "
                            f"{code >> 8} (a rank crashed or issued a different kernel sequence); results since then are invalid")

    def _wait(self, e, par, epoch, stream_ptr):
        from ... import ops
        with ops._Prof("comm_wait", "byte", 0.0):
            self._check(self.lib.ltxb200_comm_wait_status(self.local["ctl"] + self.flag_off(e, par), self.P, epoch,
                                                          self.local["ctl"] + self.status_off, self.status_host.data_ptr(), 0,
                                                          stream_ptr), "comm_wait")

    def self_attention(self, x_mod, w_qkv, b_qkv, wq, wk, cos, sin, eps):
        """x_mod [B*n_loc, D] modulated rows -> attention output rows [B*n_loc, H*d] (view of back[par]).  Runs the fused QKV
        projection (chunked), q/k norm + RoPE + head scatter, attention with the return scatter in its epilogue, and the waits."""
        from ... import ops
        self.check()
        par = self.calls & 1
        epoch = self.calls // 2 + 1
        self.calls += 1
        B, n_loc, H, d, Hp, N, P = self.B, self.n_loc, self.H, self.d, self.Hp, self.N, self.P
        D = H * d
        M = B * n_loc
        lib, ctl = self.lib, "ctl"
        f0 = self._arr([p + self.flag_off(0, par) for p in self.peer[ctl]])
        f1 = self._arr([p + self.flag_off(1, par) for p in self.peer[ctl]])
        recv, back = f"recv{par}", f"back{par}"
        recv_ptrs = self._arr(self.peer[recv])
        main = torch.cuda.current_stream()
        # chunk boundaries: multiples of 256 rows (the GEMM's CTA-pair tile) so that no chunk ends in a partial tile but the last
        nch = min(self.chunks, max(1, M // self.min_chunk_rows))
        align = 256 if M // nch >= 512 else 8
        step = ((M + nch - 1) // nch + align - 1) // align * align
        bounds = [(r0, min(M, r0 + step)) for r0 in range(0, M, step)]
        # LTXB200_SP_V_FROM_GEMM=1: V (a third of the exchange, no normalisation needed) leaves from the QKV GEMM's own epilogue and the
        # scatter kernel handles q and k only.  Parity-tested, measured neutral at 4 GPUs (scatter -1.0 ms, GEMM +1.1 ms per step: the
        # exchange is NVLink-bound wherever its stores are issued), so off by default
        vgemm = self.v_from_gemm
        nsel = 2 if vgemm else 3
        total_ctas = sum(int(lib.ltxb200_scatter_signal_ctas(r1 - r0)) * nsel // 3 for r0, r1 in bounds)
        qkv = torch.empty(M, 3 * D, device=x_mod.device, dtype=x_mod.dtype)
        overlap = len(bounds) > 1 and ops.PROFILER is None       # the per-launch event timing of the bench probe needs one stream
        def scatter(r0, r1, st):
            with ops._Prof("qk_norm_rope_wan_scatter_bf16", "byte", 2.0 * 2 * (r1 - r0) * 3 * D):
                self._check(lib.ltxb200_qk_norm_rope_wan_scatter_rows_bf16(
                    qkv.data_ptr(), qkv.stride(0), M, r0, r1 - r0, D, wq.data_ptr(), wk.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                    d, n_loc, self.rank * n_loc, float(eps), B, P, self.rank, recv_ptrs, f0, epoch,
                    self.local[ctl] + self.counter_off(0), total_ctas, nsel, st.cuda_stream), "qk_norm_rope_wan_scatter")

        def qkv_gemm(r0, r1):
            if not vgemm:
                return ops.gemm(x_mod[r0:r1], w_qkv, b_qkv, out=qkv[r0:r1])
            a, o = x_mod[r0:r1], qkv[r0:r1]
            with ops._Prof('gemm_bf16', 'flop', 2.0 * (r1 - r0) * 3 * D * a.shape[1]):
                self._check(lib.ltxb200_gemm_qkv_vscatter_bf16(
                    a.data_ptr(), a.stride(0), w_qkv.data_ptr(), w_qkv.stride(0), r1 - r0, a.shape[1], D, o.data_ptr(), o.stride(0),
                    b_qkv.data_ptr(), d, n_loc, self.rank * n_loc, r0, B, P, self.rank, recv_ptrs, main.cuda_stream), "gemm_qkv_vscatter")

        # chunk c's GEMM is enqueued BEFORE chunk c-1's scatter: both become runnable when GEMM c-1 retires and the block scheduler
        # serves launches in order, so the GEMM's CTAs (one per SM, ~200 KB of shared memory) take their slots first and the
        # scatter's small CTAs fill in beside them
        prev = None
        for r0, r1 in bounds:
            qkv_gemm(r0, r1)
            if not overlap:
                scatter(r0, r1, main)
                continue
            ev = torch.cuda.Event()
            ev.record(main)
            if prev is not None:
                self.side.wait_event(prev[2])
                scatter(prev[0], prev[1], self.side)
            prev = (r0, r1, ev)
        if overlap:
            self.side.wait_event(prev[2])
            scatter(prev[0], prev[1], self.side)
        if overlap:
            main.wait_stream(self.side)          # also orders every later reuse of `qkv`'s memory after the side stream's reads
        sp = main.cuda_stream
        self._wait(0, par, epoch, sp)
        base = self.local[recv]
        tok = B * 3 * Hp * d                    # elements per global token in recv
        q, k, v = base, base + Hp * d * 2, base + 2 * Hp * d * 2
        with ops._Prof("attention_bf16", "flop", 4.0 * B * Hp * N * N * d):
            self._check(lib.ltxb200_attention_scatter_bf16(
                q, tok, 3 * Hp * d, k, tok, 3 * Hp * d, v, tok, 3 * Hp * d, D, B, Hp, N, N, d, 0.0, None, P, self.rank,
                self._arr(self.peer[back]), f1, epoch, self.local[ctl] + self.counter_off(1), n_loc, self.rank * Hp, sp),
                "attention_scatter")
        self._wait(1, par, epoch, sp)
        return self._view(back, (B * n_loc, D))

    def all_gather_rows(self, out: torch.Tensor) -> torch.Tensor:
        """out [B, n_loc, W] fp32 (this rank's token shard of the head output) -> [B, N, W] with every rank's shard, by peer stores
        (xdit_context_parallel.py:142).  Returns a view of a double-buffered comm buffer: valid until the call after the next one."""
        from ... import ops
        assert self.gather_width and out.dtype == torch.float32 and out.is_contiguous()
        B, n_loc, W, P = self.B, self.n_loc, self.gather_width, self.P
        assert tuple(out.shape) == (B, n_loc, W)
        par = self.gather_calls & 1
        epoch = self.gather_calls // 2 + 1
        self.gather_calls += 1
        name = f"gath{par}"
        sp = torch.cuda.current_stream().cuda_stream
        f2 = self._arr([p + self.flag_off(2, par) for p in self.peer["ctl"]])
        with ops._Prof("peer_allgather", "byte", 4.0 * out.numel() * (P + 1)):
            self._check(self.lib.ltxb200_peer_allgather(out.data_ptr(), n_loc * W * 4, B, P, self.rank, self._arr(self.peer[name]), f2,
                                                        epoch, self.local["ctl"] + self.counter_off(2), sp), "peer_allgather")
        self._wait(2, par, epoch, sp)
        return self._view(name, (B, P * n_loc, W), torch.float32)

    def close(self):
        """Collective over the group: unmap the peers' buffers, then (after a barrier, so that nobody frees memory a peer still
        has mapped) free the local ones.  Idempotent."""
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        for p in self._opened:
            self.lib.ltxb200_comm_close(p)
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier(group=self.group)
        except Exception:
            pass
        for p in self.local.values():
            self.lib.ltxb200_comm_free(p)
        self._opened, self.local = [], {}
