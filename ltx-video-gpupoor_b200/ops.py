"""Torch-facing wrappers over the C ABI: torch owns device memory and streams, every arithmetic op
below is a hand-written sm_100a kernel in libltx_b200.so.  All tensors must be CUDA tensors."""
from typing import Optional

import torch

from . import _lib

BF16 = torch.bfloat16


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Optional per-launch instrumentation (bench.py's roofline probe, never on during timed steps): when PROFILER is a
# list, every wrapper appends (kernel_name, "flop"|"byte", algorithmic_amount, start_event, end_event), the events
# recorded on the launching (current) stream around the single kernel launch.
PROFILER = None


class _Prof:
    __slots__ = ("name", "kind", "amount", "e0")

    def __init__(self, name, kind, amount):
        self.name, self.kind, self.amount, self.e0 = name, kind, amount, None

    def __enter__(self):
        if PROFILER is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILER.append((self.name, self.kind, float(self.amount), self.e0, e1))
        return False


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype=BF16, name="tensor"):
    if not t.is_cuda:
        raise _lib.LtxB200Error(f"{name}: expected a CUDA tensor (this package has no CPU path)")
    if t.dtype != dtype:
        raise _lib.LtxB200Error(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise _lib.LtxB200Error(f"{name}: last dimension must be contiguous")
    return t


ACT_NONE, ACT_GELU_TANH, ACT_SILU, ACT_GELU_ERF = 0, 1, 2, 3


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         residual: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None, rows_per_gate: int = 1,
         out: Optional[torch.Tensor] = None, out_f32: bool = False) -> torch.Tensor:
    """out[M,N] = ((a @ w.T + bias) -> act) * gate[m // rows_per_gate] + residual.  a [M,K], w [N,K] bf16."""
    _req(a, name="a"); _req(w, name="w")
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if out_f32 else BF16)
    _req(out, torch.float32 if out_f32 else BF16, "out")
    if bias is not None:
        _req(bias, name="bias"); assert bias.numel() == N
    if residual is not None:
        _req(residual, name="residual"); assert residual.shape == (M, N)
    if gate is not None:
        _req(gate, name="gate"); assert gate.dim() == 2 and gate.shape[1] == N
    with _Prof('gemm_bf16', 'flop', 2.0 * M * N * K):
        rc = _lib.lib().ltxb200_gemm_bf16(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, out.data_ptr(), out.stride(0),
        1 if out_f32 else 0, _p(bias), act, _p(residual), residual.stride(0) if residual is not None else 0,
        _p(gate), gate.stride(0) if gate is not None else 0, rows_per_gate, _stream())
    _lib.check(rc, "gemm_bf16")
    return out


def gemm_f32res(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
                residual: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None, rows_per_gate: int = 1,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`mixed` precision form of gemm: out fp32 [M,N] = act(a @ w.T + bias) * gate + residual with fp32 gate / residual (out may be residual)."""
    _req(a, name="a"); _req(w, name="w")
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    _req(out, torch.float32, "out")
    if bias is not None:
        _req(bias, name="bias"); assert bias.numel() == N
    if residual is not None:
        _req(residual, torch.float32, "residual"); assert residual.shape == (M, N)
    if gate is not None:
        _req(gate, torch.float32, "gate"); assert gate.dim() == 2 and gate.shape[1] == N
    with _Prof('gemm_bf16', 'flop', 2.0 * M * N * K):
        rc = _lib.lib().ltxb200_gemm_bf16_f32res(
            a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K, out.data_ptr(), out.stride(0), _p(bias), act,
            _p(residual), residual.stride(0) if residual is not None else 0, _p(gate), gate.stride(0) if gate is not None else 0,
            rows_per_gate, _stream())
    _lib.check(rc, "gemm_bf16_f32res")
    return out


def norm_mod_f32in(x: torch.Tensor, scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None, rows_per_group: int = 0,
                   eps: float = 1e-6, layer_norm: bool = False) -> torch.Tensor:
    """x [M,D] fp32, scale / shift [G,D] fp32 row views -> bf16 (the `mixed` precision norm + AdaLN modulate)."""
    _req(x, torch.float32, "x"); assert x.dim() == 2
    M, D = x.shape
    out = torch.empty(M, D, device=x.device, dtype=BF16)
    mod_ld = 0
    if scale is not None:
        _req(scale, torch.float32, "scale"); _req(shift, torch.float32, "shift")
        assert scale.dim() == 2 and shift.dim() == 2 and scale.stride(0) == shift.stride(0)
        mod_ld = scale.stride(0)
    with _Prof('norm_mod_bf16', 'byte', 6.0 * M * D):
        rc = _lib.lib().ltxb200_norm_mod_f32in(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), M, D, _p(scale), _p(shift), mod_ld,
                                               rows_per_group, float(eps), 1 if layer_norm else 0, _stream())
    _lib.check(rc, "norm_mod_f32in")
    return out


def ada_add_f32(table: torch.Tensor, temb: torch.Tensor) -> torch.Tensor:
    """table [L, J, D], temb [G, J*D] bf16 -> [L, G, J, D] fp32 (sum formed in fp32)"""
    _req(table, name="table"); _req(temb, name="temb")
    L, J, D = table.shape
    G = temb.shape[0]
    assert table.is_contiguous() and temb.is_contiguous() and temb.shape[1] == J * D
    out = torch.empty(L, G, J, D, device=table.device, dtype=torch.float32)
    _lib.check(_lib.lib().ltxb200_ada_add_f32(table.data_ptr(), temb.data_ptr(), out.data_ptr(), L, G, J * D, _stream()), "ada_add_f32")
    return out


CONV_NDHWC, CONV_D2S, CONV_UNPATCH = 0, 1, 2


def conv3d(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], causal: bool = False,
           store: int = CONV_NDHWC, residual: Optional[torch.Tensor] = None, out_f32: bool = False) -> torch.Tensor:
    """x [B,T,H,W,Cin] bf16 (NDHWC, contiguous); w [Cout, 27*Cin] bf16 (tap-major K)."""
    _req(x, name="x"); _req(w, name="w")
    assert x.is_contiguous() and w.is_contiguous() and x.dim() == 5
    B, T, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == 27 * Cin
    if store == CONV_NDHWC:
        out = torch.empty(B, T, H, W, Cout, device=x.device, dtype=BF16)
    elif store == CONV_D2S:
        out = torch.empty(B, 2 * T - 1, 2 * H, 2 * W, Cout // 8, device=x.device, dtype=BF16)
    else:
        out = torch.empty(B, Cout // 16, T, 4 * H, 4 * W, device=x.device, dtype=torch.float32 if out_f32 else BF16)
    if residual is not None:
        _req(residual, name="residual"); assert residual.is_contiguous() and residual.shape == out.shape
    with _Prof('conv3d_bf16', 'flop', 2.0 * B * T * H * W * Cout * 27 * Cin):
        rc = _lib.lib().ltxb200_conv3d_bf16(x.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), B, T, H, W, Cin, Cout,
                                        1 if causal else 0, store, 1 if out_f32 else 0, _p(residual), _stream())
    _lib.check(rc, "conv3d_bf16")
    return out


def conv3d_norm(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], causal: bool = False,
                residual: Optional[torch.Tensor] = None, keep_raw: bool = True, eps: float = 1e-8):
    """conv3d (NDHWC) whose epilogue also writes silu(pixelnorm(y)) of its own output row (Cout <= 256): returns (y | None, y_norm).
    keep_raw=False skips the raw output (it is only needed as the next block's residual)."""
    _req(x, name="x"); _req(w, name="w")
    assert x.is_contiguous() and w.is_contiguous() and x.dim() == 5
    B, T, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == 27 * Cin and Cout <= 256
    out = torch.empty(B, T, H, W, Cout, device=x.device, dtype=BF16) if keep_raw else None
    out2 = torch.empty(B, T, H, W, Cout, device=x.device, dtype=BF16)
    if residual is not None:
        _req(residual, name="residual"); assert residual.is_contiguous() and tuple(residual.shape) == tuple(out2.shape)
    with _Prof('conv3d_bf16', 'flop', 2.0 * B * T * H * W * Cout * 27 * Cin):
        rc = _lib.lib().ltxb200_conv3d_norm_bf16(x.data_ptr(), w.data_ptr(), _p(bias), _p(out), out2.data_ptr(), B, T, H, W, Cin, Cout,
                                                 1 if causal else 0, _p(residual), 1 if keep_raw else 2, float(eps), _stream())
    _lib.check(rc, "conv3d_norm_bf16")
    return out, out2


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key_bias: Optional[torch.Tensor] = None,
              scale: float = 0.0, out: Optional[torch.Tensor] = None, accumulate: bool = False,
              key_lens: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B,Lq,H,d], k/v [B,Lk,H,d] bf16 (strided views allowed, head stride must be d) -> [B,Lq,H,d].
    accumulate: out += attention (out must be given).  key_lens int32 [B] (device): batch element b attends to its first key_lens[b]
    keys only (a right-padded prompt mask without the bias pass / the padded key blocks); exclusive with key_bias."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, name=n)
        assert t.dim() == 4 and t.stride(2) == t.shape[3], f"{n}: heads must be packed (stride(2) == d)"
    B, Lq, H, d = q.shape
    Lk = k.shape[1]
    if accumulate and out is None:
        raise ValueError("accumulate=True needs out=")
    if out is None:
        out = torch.empty(B, Lq, H, d, device=q.device, dtype=BF16)
    _req(out, name="out")
    if key_bias is not None:
        _req(key_bias, torch.float32, "key_bias"); assert key_bias.shape == (B, Lk) and key_bias.is_contiguous()
    args = (q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0),
            v.data_ptr(), v.stride(1), v.stride(0), out.data_ptr(), out.stride(1), out.stride(0), B, H, Lq, Lk, d, float(scale))
    with _Prof('attention_bf16', 'flop', 4.0 * B * H * Lq * Lk * d):
        if key_lens is not None:
            _req(key_lens, torch.int32, "key_lens")
            assert key_bias is None and not accumulate and key_lens.shape == (B,) and key_lens.is_contiguous()
            rc = _lib.lib().ltxb200_attention_klens_bf16(*args, key_lens.data_ptr(), _stream())
        else:
            fn = _lib.lib().ltxb200_attention_acc_bf16 if accumulate else _lib.lib().ltxb200_attention_bf16
            rc = fn(*args, _p(key_bias), _stream())
    _lib.check(rc, "attention_bf16")
    return out


def norm_mod(x: torch.Tensor, scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
             rows_per_group: int = 0, weight: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
             eps: float = 1e-6, layer_norm: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [M,D]; scale/shift [G, D] row views with a common row stride."""
    _req(x, name="x"); assert x.dim() == 2
    M, D = x.shape
    if out is None:
        out = torch.empty(M, D, device=x.device, dtype=BF16)
    mod_ld = 0
    if scale is not None:
        _req(scale, name="scale"); _req(shift, name="shift")
        assert scale.dim() == 2 and shift.dim() == 2 and scale.stride(0) == shift.stride(0)
        mod_ld = scale.stride(0)
    with _Prof('norm_mod_bf16', 'byte', 4.0 * M * D):
        rc = _lib.lib().ltxb200_norm_mod_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), M, D, _p(scale),
                                          _p(shift), mod_ld, rows_per_group, _p(weight), _p(bias), float(eps),
                                          1 if layer_norm else 0, _stream())
    _lib.check(rc, "norm_mod_bf16")
    return out


def qk_norm_rope(q: Optional[torch.Tensor], k: Optional[torch.Tensor], wq, wk, cos=None, sin=None,
                 tokens_per_batch: int = 0, eps: float = 1e-5):
    """In-place on 2-D row views q [Mq,D], k [Mk,D]."""
    D = (q if q is not None else k).shape[1]
    rows = (q.shape[0] if q is not None else 0) + (k.shape[0] if k is not None else 0)
    # algorithmic bytes: q/k rows read + written once; the cos/sin tables once per launch (they are shared by the batch)
    with _Prof('qk_norm_rope_bf16', 'byte', 4.0 * D * rows + (0.0 if cos is None else 4.0 * D * tokens_per_batch)):
        rc = _lib.lib().ltxb200_qk_norm_rope_bf16(
        _p(q), q.stride(0) if q is not None else 0, q.shape[0] if q is not None else 0,
        _p(k), k.stride(0) if k is not None else 0, k.shape[0] if k is not None else 0, D,
        _p(wq), _p(wk), _p(cos), _p(sin), tokens_per_batch, float(eps), _stream())
    _lib.check(rc, "qk_norm_rope_bf16")


def ada_add(table: torch.Tensor, temb: torch.Tensor) -> torch.Tensor:
    """table [L, J, D], temb [G, J*D] -> [L, G, J, D]"""
    _req(table, name="table"); _req(temb, name="temb")
    L, J, D = table.shape
    G = temb.shape[0]
    assert table.is_contiguous() and temb.is_contiguous() and temb.shape[1] == J * D
    out = torch.empty(L, G, J, D, device=table.device, dtype=BF16)
    _lib.check(_lib.lib().ltxb200_ada_add_bf16(table.data_ptr(), temb.data_ptr(), out.data_ptr(), L, G, J * D, _stream()),
               "ada_add_bf16")
    return out


def act(x: torch.Tensor, mode: int) -> torch.Tensor:
    _req(x, name="x"); assert x.is_contiguous()
    y = torch.empty_like(x)
    _lib.check(_lib.lib().ltxb200_act_bf16(x.data_ptr(), y.data_ptr(), x.numel(), mode, _stream()), "act_bf16")
    return y


def stg_blend(a: torch.Tensor, v: torch.Tensor, mask: torch.Tensor):
    """a [B, rows, D] contiguous (in place); v [B*rows, D] row view; mask [B] fp32"""
    B, rows, D = a.shape
    _req(a, name="a"); _req(v, name="v"); _req(mask, torch.float32, "mask")
    assert a.is_contiguous()
    _lib.check(_lib.lib().ltxb200_stg_blend_bf16(a.data_ptr(), v.data_ptr(), v.stride(0), mask.data_ptr(), B, rows, D,
                                                 _stream()), "stg_blend_bf16")


def axpby(x: torch.Tensor, y: torch.Tensor, a: float = 1.0, b: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = bf16(a*x + b*y) (out may be x or y)."""
    _req(x, name="x"); _req(y, name="y")
    assert x.is_contiguous() and y.is_contiguous() and x.shape == y.shape
    if out is None:
        out = torch.empty_like(x)
    _req(out, name="out"); assert out.is_contiguous() and out.shape == x.shape
    _lib.check(_lib.lib().ltxb200_axpby_bf16(x.data_ptr(), y.data_ptr(), out.data_ptr(), x.numel(), float(a), float(b),
                                             _stream()), "axpby_bf16")
    return out


def rel_l1(a: torch.Tensor, b: torch.Tensor) -> float:
    """`((a - b).abs().mean() / b.abs().mean()).cpu().item()` with ATen's bf16 roundings (model.py:1039); syncs like the
    reference's .item()."""
    _req(a, name="a"); _req(b, name="b")
    assert a.is_contiguous() and b.is_contiguous() and a.shape == b.shape
    out = torch.empty(2, device=a.device, dtype=torch.float32)
    _lib.check(_lib.lib().ltxb200_rel_l1_bf16(a.data_ptr(), b.data_ptr(), a.numel(), out.data_ptr(), _stream()), "rel_l1_bf16")
    s = (out.cpu() / a.numel()).to(BF16)           # the two means, rounded to the tensors' dtype
    return float((s[0] / s[1]).item())             # bf16 division


def timestep_embed(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    _req(t, torch.float32, "t"); assert t.dim() == 1 and t.is_contiguous()
    out = torch.empty(t.shape[0], dim, device=t.device, dtype=BF16)
    _lib.check(_lib.lib().ltxb200_timestep_embed(t.data_ptr(), out.data_ptr(), t.shape[0], dim, _stream()), "timestep_embed")
    return out


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, torch.float32, "x"); assert x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=BF16)
    _lib.check(_lib.lib().ltxb200_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "cast")
    return out


def guidance_step(pred: torch.Tensor, latents: torch.Tensor, timesteps: torch.Tensor, t: float, *, num_conds: int,
                  has_cfg: bool, has_stg: bool, do_rescale: bool, guidance_scale: float, stg_scale: float,
                  rescale: float, channels: int, cond_mask: Optional[torch.Tensor], scratch: Optional[torch.Tensor],
                  latents_bf16: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                  cond_stride: Optional[int] = None):
    """One sample: latents [n] fp32 in place; its prediction for cond c starts c * cond_stride elements after `pred` (bf16;
    default cond_stride = n, i.e. pred [num_conds, n]; a batch of b samples laid out cond-major passes the sample's slice and
    cond_stride = b * n); `noise` [n] fp32 selects the stochastic update (rf.py:370-373)."""
    _req(pred, name="pred"); _req(latents, torch.float32, "latents"); _req(timesteps, torch.float32, "timesteps")
    assert pred.is_contiguous() and latents.is_contiguous()
    n = latents.numel()
    cond_stride = n if cond_stride is None else int(cond_stride)
    assert cond_stride >= n and pred.numel() >= (num_conds - 1) * cond_stride + n
    args = (pred.data_ptr(), cond_stride, n, channels, int(has_cfg), int(has_stg), int(do_rescale), float(guidance_scale), float(stg_scale),
            float(rescale), latents.data_ptr(), _p(latents_bf16), timesteps.data_ptr(), timesteps.numel(), float(t), _p(cond_mask),
            _p(scratch))
    with _Prof('guidance_step', 'byte', n * (2.0 * num_conds + 10.0)):
        if noise is None:
            rc = _lib.lib().ltxb200_guidance_step(*args, _stream())
        else:
            _req(noise, torch.float32, "noise"); assert noise.is_contiguous() and noise.numel() == n
            rc = _lib.lib().ltxb200_guidance_step_stochastic(*args, noise.data_ptr(), _stream())
    _lib.check(rc, "guidance_step")


def pixelnorm_silu(x: torch.Tensor, silu: bool = True, eps: float = 1e-8, scale: Optional[torch.Tensor] = None,
                   shift: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PixelNorm [-> x*(1+scale)+shift, scale/shift bf16 [C]] [-> SiLU] on [..., C] bf16."""
    _req(x, name="x"); assert x.is_contiguous()
    C = x.shape[-1]
    y = torch.empty_like(x)
    if scale is not None:
        _req(scale, name="scale"); _req(shift, name="shift")
        assert scale.is_contiguous() and shift.is_contiguous() and scale.numel() == C and shift.numel() == C
        with _Prof('pixelnorm_silu_bf16', 'byte', 4.0 * x.numel()):
            rc = _lib.lib().ltxb200_pixelnorm_mod_silu_bf16(x.data_ptr(), y.data_ptr(), x.numel() // C, C, float(eps),
                                                            scale.data_ptr(), shift.data_ptr(), int(silu), _stream())
        _lib.check(rc, "pixelnorm_mod_silu")
        return y
    with _Prof('pixelnorm_silu_bf16', 'byte', 4.0 * x.numel()):
        rc = _lib.lib().ltxb200_pixelnorm_silu_bf16(x.data_ptr(), y.data_ptr(), x.numel() // C, C, float(eps), int(silu), _stream())
    _lib.check(rc, "pixelnorm_silu")
    return y


def latent_to_ndhwc(z: torch.Tensor, stdv: Optional[torch.Tensor], meanv: Optional[torch.Tensor]) -> torch.Tensor:
    assert z.is_cuda and z.is_contiguous() and z.dim() == 5 and z.dtype in (torch.float32, BF16)
    B, C, F_, H, W = z.shape
    out = torch.empty(B, F_, H, W, C, device=z.device, dtype=BF16)
    _lib.check(_lib.lib().ltxb200_latent_to_ndhwc(z.data_ptr(), int(z.dtype == torch.float32), out.data_ptr(), B, C,
                                                  F_ * H * W, _p(stdv), _p(meanv), _stream()), "latent_to_ndhwc")
    return out


def qk_norm_rope_wan(q: Optional[torch.Tensor], k: Optional[torch.Tensor], wq, wk, cos=None, sin=None, head_dim: int = 128,
                     tokens_per_batch: int = 0, token_offset: int = 0, eps: float = 1e-6):
    """In-place on 2-D row views q [Mq,D], k [Mk,D]; cos/sin fp32 [tokens, head_dim]."""
    D = (q if q is not None else k).shape[1]
    if cos is not None:
        _req(cos, torch.float32, "cos"); _req(sin, torch.float32, "sin")
        assert cos.is_contiguous() and sin.is_contiguous() and cos.shape[1] == head_dim
    with _Prof("qk_norm_rope_wan_bf16", "byte", (4.0 * D) * ((q.shape[0] if q is not None else 0) + (k.shape[0] if k is not None else 0))):
        rc = _lib.lib().ltxb200_qk_norm_rope_wan_bf16(
            _p(q), q.stride(0) if q is not None else 0, q.shape[0] if q is not None else 0,
            _p(k), k.stride(0) if k is not None else 0, k.shape[0] if k is not None else 0, D,
            _p(wq), _p(wk), _p(cos), _p(sin), head_dim, tokens_per_batch, token_offset, float(eps), _stream())
    _lib.check(rc, "qk_norm_rope_wan_bf16")


def lincomb(out: torch.Tensor, terms) -> torch.Tensor:
    """out = sum(c * x for c, x in terms); fp32 contiguous tensors of identical numel (out may alias an x)."""
    import ctypes
    n = out.numel()
    k = len(terms)
    _req(out, torch.float32, "out"); assert out.is_contiguous()
    ptrs = (ctypes.c_void_p * k)()
    cs = (ctypes.c_float * k)()
    for j, (c, x) in enumerate(terms):
        _req(x, torch.float32, "x"); assert x.is_contiguous() and x.numel() == n
        ptrs[j] = x.data_ptr(); cs[j] = float(c)
    with _Prof("lincomb_f32", "byte", 4.0 * n * (k + 1)):
        rc = _lib.lib().ltxb200_lincomb_f32(out.data_ptr(), n, k, ptrs, cs, _stream())
    _lib.check(rc, "lincomb_f32")
    return out


def rf_step_tokens(x: torch.Tensor, v: torch.Tensor, tok_timesteps: torch.Tensor, schedule: torch.Tensor,
                   noise: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rf.py:361-375 with per-token timesteps: x, v (and noise) [tokens, C] fp32, tok_timesteps [tokens] fp32, schedule [S] fp32 descending."""
    _req(x, torch.float32, "x"); _req(v, torch.float32, "v"); _req(tok_timesteps, torch.float32, "tok_timesteps"); _req(schedule, torch.float32, "schedule")
    assert x.dim() == 2 and x.shape == v.shape and x.is_contiguous() and v.is_contiguous()
    tokens, C = x.shape
    assert tok_timesteps.is_contiguous() and tok_timesteps.numel() == tokens and schedule.is_contiguous() and C % 4 == 0
    if noise is not None:
        _req(noise, torch.float32, "noise"); assert noise.is_contiguous() and noise.shape == x.shape
    out = torch.empty_like(x) if out is None else out
    _req(out, torch.float32, "out"); assert out.is_contiguous() and out.shape == x.shape
    with _Prof("rf_step_tokens_f32", "byte", 4.0 * x.numel() * (3 + (noise is not None))):
        rc = _lib.lib().ltxb200_rf_step_tokens_f32(out.data_ptr(), x.data_ptr(), v.data_ptr(), _p(noise), tok_timesteps.data_ptr(), tokens, C,
                                                   schedule.data_ptr(), schedule.numel(), _stream())
    _lib.check(rc, "rf_step_tokens_f32")
    return out


def cfg_combine(cond: torch.Tensor, uncond: torch.Tensor, guide_scale: float, use_alpha: bool,
                scratch: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(cond, torch.float32, "cond"); _req(uncond, torch.float32, "uncond")
    assert cond.is_contiguous() and uncond.is_contiguous() and cond.shape == uncond.shape
    if out is None:
        out = torch.empty_like(cond)
    if use_alpha and scratch is None:
        scratch = torch.empty(2 * 148, device=cond.device, dtype=torch.float32)
    with _Prof("cfg_combine_f32", "byte", 4.0 * cond.numel() * (5 if use_alpha else 3)):
        rc = _lib.lib().ltxb200_cfg_combine_f32(cond.data_ptr(), uncond.data_ptr(), out.data_ptr(), cond.numel(),
                                                float(guide_scale), int(use_alpha), _p(scratch), _stream())
    _lib.check(rc, "cfg_combine_f32")
    return out


# ------------------------------------------------------------------------------------------------------------------
# Wan VAE decode pieces
# ------------------------------------------------------------------------------------------------------------------
def conv_taps(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], taps_t: int, taps_hw: int,
              zero_pad_t: bool = True, residual: Optional[torch.Tensor] = None, centered: bool = False) -> torch.Tensor:
    """x [B,T,H,W,Cin] bf16 NDHWC; w [Cout, taps_t*taps_hw^2*Cin] bf16 (tap-major K); causal in time, or `centered`
    (taps t-1,t,t+1, zero padding: nn.Conv3d(kernel 3, padding 1))."""
    _req(x, name="x"); _req(w, name="w")
    assert x.is_contiguous() and w.is_contiguous() and x.dim() == 5
    B, T, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == taps_t * taps_hw * taps_hw * Cin
    out = torch.empty(B, T, H, W, Cout, device=x.device, dtype=BF16)
    if residual is not None:
        _req(residual, name="residual"); assert residual.is_contiguous() and residual.shape == out.shape
    with _Prof('conv_taps_bf16', 'flop', 2.0 * B * T * H * W * Cout * w.shape[1]):
        rc = _lib.lib().ltxb200_conv_taps_bf16(x.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), B, T, H, W, Cin, Cout,
                                               taps_t, taps_hw, 2 if centered else (1 if zero_pad_t else 0), _p(residual),
                                               _stream())
    _lib.check(rc, "conv_taps_bf16")
    return out


def conv_taps_strided(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], taps_t: int, taps_hw: int,
                      stride_t: int = 1, stride_hw: int = 1, off_hw: int = 0) -> torch.Tensor:
    """Zero-padded causal conv with output strides (Wan encoder Resample, wan/modules/vae.py:90-97,150-165): x [B,T,H,W,Cin]
    bf16 NDHWC, w [Cout, taps*Cin] tap-major; off_hw = 1 -> spatial taps (h, h+1, h+2) = ZeroPad2d((0,1,0,1)) + Conv2d."""
    _req(x, name="x"); _req(w, name="w")
    assert x.is_contiguous() and w.is_contiguous() and x.dim() == 5
    B, T, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == taps_t * taps_hw * taps_hw * Cin
    To, Ho, Wo = (T - 1) // stride_t + 1, (H - 1 - off_hw) // stride_hw + 1, (W - 1 - off_hw) // stride_hw + 1
    out = torch.empty(B, To, Ho, Wo, Cout, device=x.device, dtype=BF16)
    with _Prof('conv_taps_bf16', 'flop', 2.0 * B * To * Ho * Wo * Cout * w.shape[1]):
        rc = _lib.lib().ltxb200_conv_taps_strided_bf16(x.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), B, T, H, W, Cin,
                                                       Cout, taps_t, taps_hw, stride_t, stride_hw, off_hw, _stream())
    _lib.check(rc, "conv_taps_strided_bf16")
    return out


def l2norm_silu(x: torch.Tensor, gamma: torch.Tensor, c_real: int, silu: bool = True) -> torch.Tensor:
    """x [..., C] bf16 contiguous (C = stored, 64-padded channels) -> RMS_norm(+SiLU)."""
    _req(x, name="x"); _req(gamma, name="gamma")
    assert x.is_contiguous()
    C = x.shape[-1]
    y = torch.empty_like(x)
    with _Prof('l2norm_silu_bf16', 'byte', 4.0 * x.numel()):
        rc = _lib.lib().ltxb200_l2norm_silu_bf16(x.data_ptr(), y.data_ptr(), x.numel() // C, C, c_real, gamma.data_ptr(),
                                                 1 if silu else 0, _stream())
    _lib.check(rc, "l2norm_silu_bf16")
    return y


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    """x [F, H, W, C] bf16 -> [F, 2H, 2W, C] (nearest)."""
    _req(x, name="x"); assert x.is_contiguous() and x.dim() == 4
    Fr, H, W, C = x.shape
    y = torch.empty(Fr, 2 * H, 2 * W, C, device=x.device, dtype=BF16)
    with _Prof('upsample2x_bf16', 'byte', 2.5 * y.numel()):
        rc = _lib.lib().ltxb200_upsample2x_nhwc_bf16(x.data_ptr(), y.data_ptr(), Fr, H, W, C, _stream())
    _lib.check(rc, "upsample2x_nhwc_bf16")
    return y


def softmax_rows(s: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """s [rows, cols] fp32 (row-strided view allowed) -> softmax(scale*s) bf16."""
    _req(s, torch.float32, "s"); assert s.dim() == 2 and s.stride(1) == 1
    p = out if out is not None else torch.empty(s.shape, device=s.device, dtype=BF16)
    _req(p, name="out"); assert p.shape == s.shape and p.stride(1) == 1
    _lib.check(_lib.lib().ltxb200_softmax_rows_f32_bf16(s.data_ptr(), s.stride(0), p.data_ptr(), p.stride(0), s.shape[0], s.shape[1],
                                                        float(scale), _stream()), "softmax_rows_f32_bf16")
    return p


def conv3d_strided(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], stride_t: int = 2, stride_hw: int = 2) -> torch.Tensor:
    """Causal 3x3x3 conv with output strides: x [B,T,H,W,Cin] bf16 NDHWC, w [Cout, 27*Cin] tap-major."""
    _req(x, name="x"); _req(w, name="w")
    assert x.is_contiguous() and w.is_contiguous() and x.dim() == 5
    B, T, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == 27 * Cin
    To, Ho, Wo = (T - 1) // stride_t + 1, (H - 1) // stride_hw + 1, (W - 1) // stride_hw + 1
    out = torch.empty(B, To, Ho, Wo, Cout, device=x.device, dtype=BF16)
    with _Prof('conv3d_bf16', 'flop', 2.0 * B * To * Ho * Wo * Cout * 27 * Cin):
        rc = _lib.lib().ltxb200_conv3d_strided_bf16(x.data_ptr(), w.data_ptr(), _p(bias), out.data_ptr(), B, T, H, W, Cin, Cout,
                                                    stride_t, stride_hw, _stream())
    _lib.check(rc, "conv3d_strided_bf16")
    return out


# ------------------------------------------------------------------------------------------------------------------
# LTX multi-scale flow pieces (latent upsampler, AdaIN, resize)
# ------------------------------------------------------------------------------------------------------------------
def groupnorm_silu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, residual: Optional[torch.Tensor] = None,
                   silu: bool = True, eps: float = 1e-5) -> torch.Tensor:
    """x [B, ..., C] bf16 NDHWC -> [SiLU](GroupNorm(32)(x) [+ residual])."""
    _req(x, name="x"); _req(gamma, name="gamma"); _req(beta, name="beta")
    assert x.is_contiguous()
    B, C = x.shape[0], x.shape[-1]
    voxels = x.numel() // (B * C)
    y = torch.empty_like(x)
    if residual is not None:
        _req(residual, name="residual"); assert residual.is_contiguous() and residual.shape == x.shape
    scratch = torch.empty(B * 64 * 32 * 2, device=x.device, dtype=torch.float32)
    with _Prof('groupnorm_silu_bf16', 'byte', (6.0 if residual is None else 8.0) * x.numel()):
        rc = _lib.lib().ltxb200_groupnorm_silu_bf16(x.data_ptr(), y.data_ptr(), B, voxels, C, gamma.data_ptr(), beta.data_ptr(),
                                                    _p(residual), float(eps), 1 if silu else 0, scratch.data_ptr(), _stream())
    _lib.check(rc, "groupnorm_silu_bf16")
    return y


def adain(x: torch.Tensor, ref: torch.Tensor, factor: float = 1.0) -> torch.Tensor:
    """x [B, C, ...], ref [B, C, ...] fp32 -> per (b, c) mean / unbiased-std matching + lerp."""
    _req(x, torch.float32, "x"); _req(ref, torch.float32, "ref")
    assert x.is_contiguous() and ref.is_contiguous() and x.shape[:2] == ref.shape[:2]
    rows = x.shape[0] * x.shape[1]
    out = torch.empty_like(x)
    _lib.check(_lib.lib().ltxb200_adain_f32(x.data_ptr(), ref.data_ptr(), out.data_ptr(), rows, x.numel() // rows,
                                            ref.numel() // rows, float(factor), _stream()), "adain_f32")
    return out


def latent_from_ndhwc(x: torch.Tensor, stdv: Optional[torch.Tensor], meanv: Optional[torch.Tensor]) -> torch.Tensor:
    """x [B, F, H, W, C] bf16 -> [B, C, F, H, W] fp32, (x - mean) / std per channel."""
    _req(x, name="x"); assert x.is_contiguous() and x.dim() == 5
    B, Fr, H, W, C = x.shape
    out = torch.empty(B, C, Fr, H, W, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().ltxb200_latent_from_ndhwc(x.data_ptr(), out.data_ptr(), B, C, Fr * H * W, _p(stdv), _p(meanv), _stream()),
               "latent_from_ndhwc")
    return out


def bilinear_resize(x: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """x [..., h, w] fp32 -> [..., height, width] (F.interpolate bilinear, align_corners=False)."""
    _req(x, torch.float32, "x"); assert x.is_contiguous()
    h, w = x.shape[-2:]
    out = torch.empty(*x.shape[:-2], height, width, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().ltxb200_bilinear_resize_f32(x.data_ptr(), out.data_ptr(), x.numel() // (h * w), h, w, height, width,
                                                      _stream()), "bilinear_resize_f32")
    return out
