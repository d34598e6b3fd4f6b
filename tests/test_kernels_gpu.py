"""GPU parity tests of every C-ABI kernel against a plain PyTorch fp32 restatement of the same op
(the oracle's functions where one exists).  Tolerances are for bf16 inputs/outputs with fp32 accumulation."""
import math

import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # collected but skipped on the CPU box (-m "not gpu" deselects anyway)
    pytest.skip("needs a GPU", allow_module_level=True)

from ltx_video_gpupoor_b200 import ops  # noqa: E402
from oracle import ltx_oracle as O  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


def rnd(*shape, seed=0, scale=1.0, dtype=BF):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def rel(a, b):
    return O.rel_l2(a.float().cpu(), b.float().cpu())


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 128, 128), (300, 136, 72), (1, 512, 256), (3, 12288, 2048),
                                   (6144, 2048, 2048), (1000, 8192, 2048), (777, 2048, 8192), (6144, 128, 2048),
                                   (6144, 2048, 128), (1000, 1024, 4160), (513, 512, 8192), (5000, 1536, 8960)])   # the last three: long K (and, with LTXB200_GEMM_BN512=1, the 256 x 512 pair tiles)
def test_gemm_plain(M, N, K):
    a, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    ref = a.float() @ w.float().t() + b.float()
    out = ops.gemm(a, w, b)
    torch.cuda.synchronize()
    assert rel(out, ref) < 6e-3, rel(out, ref)


def test_gemm_epilogues():
    M, N, K = 515, 768, 320
    a, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    res = rnd(M, N, seed=4)
    rows_per_gate = 103
    gate = rnd((M + rows_per_gate - 1) // rows_per_gate, N, seed=5)
    base = a.float() @ w.float().t() + b.float()
    assert rel(ops.gemm(a, w, b, act=ops.ACT_GELU_TANH), F.gelu(base, approximate="tanh")) < 6e-3
    assert rel(ops.gemm(a, w, b, act=ops.ACT_SILU), F.silu(base)) < 6e-3
    g_full = gate.float().repeat_interleave(rows_per_gate, dim=0)[:M]
    ref = res.float() + g_full * base
    out = ops.gemm(a, w, b, residual=res, gate=gate, rows_per_gate=rows_per_gate)
    assert rel(out, ref) < 6e-3
    # in place on the residual, fp32 output, strided A
    r2 = res.clone()
    ops.gemm(a, w, b, residual=r2, out=r2)
    assert rel(r2, res.float() + base) < 6e-3
    big = rnd(M, 2 * K, seed=7)
    o32 = ops.gemm(big[:, K:], w, None, out_f32=True)
    assert o32.dtype == torch.float32
    assert rel(o32, big[:, K:].float() @ w.float().t()) < 2e-3
    # long K with the gate / residual epilogue, in place, and a partial last row tile (LTXB200_GEMM_BN512=1: the 256 x 512 pair tile)
    M, N, K = 1300, 1024, 4096
    a, w, b = rnd(M, K, seed=11), rnd(N, K, seed=12, scale=K ** -0.5), rnd(N, seed=13)
    res, gate = rnd(M, N, seed=14), rnd((M + 99) // 100, N, seed=15)
    ref = res.float() + gate.float().repeat_interleave(100, dim=0)[:M] * (a.float() @ w.float().t() + b.float())
    out = ops.gemm(a, w, b, residual=res, gate=gate, rows_per_gate=100, out=res)
    assert out.data_ptr() == res.data_ptr() and rel(out, ref) < 6e-3


# ------------------------------------------------------------------ attention
def ref_attn(q, k, v, bias=None):
    return O.attention_core(q.float().cpu(), k.float().cpu(), v.float().cpu(),
                            None if bias is None else bias.cpu()[:, None, None, :])


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("B,H,Lq,Lk", [(1, 2, 128, 128), (2, 3, 256, 384), (1, 2, 300, 333), (1, 4, 1024, 1000),
                                       (2, 2, 130, 77), (1, 2, 257, 40), (3, 40, 1500, 700), (1, 1, 64, 1)])
def test_attention(d, B, H, Lq, Lk):
    q, k, v = rnd(B, Lq, H, d, seed=1), rnd(B, Lk, H, d, seed=2), rnd(B, Lk, H, d, seed=3)
    out = ops.attention(q, k, v)
    torch.cuda.synchronize()
    assert rel(out, ref_attn(q, k, v)) < 1e-2


@pytest.mark.parametrize("d", [64, 128])
def test_attention_bias_and_fused_qkv(d):
    B, H, L, Lk = 2, 4, 384, 200
    qkv = rnd(B, L, 3 * H * d, seed=5)
    q, k, v = [qkv[:, :, i * H * d:(i + 1) * H * d].unflatten(-1, (H, d)) for i in range(3)]
    out = ops.attention(q, k, v)
    assert rel(out, ref_attn(q, k, v)) < 1e-2
    kk, vv = rnd(B, Lk, H, d, seed=6), rnd(B, Lk, H, d, seed=7)
    bias = torch.zeros(B, Lk, device=DEV)
    bias[0, 150:] = -10000.0
    bias[1, 31:] = -10000.0
    out = ops.attention(q, kk, vv, key_bias=bias)
    assert rel(out, ref_attn(q, kk, vv, bias)) < 1e-2
    # large-magnitude logits exercise the lazy rescale
    qs = (q.float() * 6).to(BF)
    out = ops.attention(qs, k, v)
    assert rel(out, ref_attn(qs, k, v)) < 1.5e-2


# ------------------------------------------------------------------ conv3d
def ref_conv(x_ndhwc, w5, b, causal):
    x = x_ndhwc.float().permute(0, 4, 1, 2, 3)
    if causal:
        x = torch.cat([x[:, :, :1].repeat(1, 1, 2, 1, 1), x], dim=2)
    else:
        x = torch.cat([x[:, :, :1], x, x[:, :, -1:]], dim=2)
    return F.conv3d(x, w5.float(), b.float(), padding=(0, 1, 1))          # NCDHW


def pack_w(w5):   # [Cout, Cin, 3,3,3] -> [Cout, 27*Cin] tap-major
    return w5.permute(0, 2, 3, 4, 1).reshape(w5.shape[0], -1).contiguous()


@pytest.mark.parametrize("B,T,H,W,Cin,Cout,causal", [(1, 3, 8, 16, 64, 128, False), (2, 2, 5, 7, 128, 64, True),
                                                     (1, 4, 16, 24, 128, 256, False), (1, 1, 3, 3, 64, 512, False)])
def test_conv3d_plain(B, T, H, W, Cin, Cout, causal):
    x = rnd(B, T, H, W, Cin, seed=1)
    w5 = rnd(Cout, Cin, 3, 3, 3, seed=2, scale=(27 * Cin) ** -0.5)
    b = rnd(Cout, seed=3)
    ref = ref_conv(x, w5, b, causal).permute(0, 2, 3, 4, 1)
    out = ops.conv3d(x, pack_w(w5), b, causal=causal)
    torch.cuda.synchronize()
    assert rel(out, ref) < 6e-3
    res = rnd(B, T, H, W, Cout, seed=4)
    out = ops.conv3d(x, pack_w(w5), b, causal=causal, residual=res)
    assert rel(out, ref + res.float()) < 6e-3


def test_conv3d_d2s_and_unpatch():
    B, T, H, W, Cin = 1, 3, 6, 10, 64
    x = rnd(B, T, H, W, Cin, seed=1)
    # depth-to-space: Cout = 8*C, reference channel order (c p1 p2 p3); kernel wants (p1 p2 p3 c)
    C = 32
    w5 = rnd(8 * C, Cin, 3, 3, 3, seed=2, scale=(27 * Cin) ** -0.5)
    b = rnd(8 * C, seed=3)
    ref = O.depth_to_space(ref_conv(x, w5, b, False))[:, :, 1:].permute(0, 2, 3, 4, 1)
    perm = torch.arange(8 * C).reshape(C, 8).t().reshape(-1).to(DEV)          # new row (p,c) <- old row c*8+p
    out = ops.conv3d(x, pack_w(w5[perm]), b[perm].contiguous(), store=ops.CONV_D2S)
    assert out.shape == ref.shape
    assert rel(out, ref) < 6e-3
    # unpatchify: Cout = 3*16, reference channel order (c r q); kernel wants (c q r)
    w5 = rnd(48, Cin, 3, 3, 3, seed=4, scale=(27 * Cin) ** -0.5)
    b = rnd(48, seed=5)
    ref = O.vae_unpatchify(ref_conv(x, w5, b, False), 4)
    perm = torch.arange(48).reshape(3, 4, 4).permute(0, 2, 1).reshape(-1).to(DEV)   # new (c,q,r) <- old (c,r,q)
    out = ops.conv3d(x, pack_w(w5[perm]), b[perm].contiguous(), store=ops.CONV_UNPATCH, out_f32=True)
    assert out.shape == ref.shape
    assert rel(out, ref) < 6e-3


# ------------------------------------------------------------------ memory-bound kernels
@pytest.mark.parametrize("D", [2048, 1536, 512, 4096, 5120])
@pytest.mark.parametrize("ln", [False, True])
def test_norm_mod(D, ln):
    M, rpg = 77, 20
    x = rnd(M, D, seed=1, scale=3.0)
    G = (M + rpg - 1) // rpg
    mod = rnd(G, 6, D, seed=2, scale=0.3)
    sc, sh = mod[:, 1], mod[:, 0]
    xf = x.float()
    base = F.layer_norm(xf, (D,), eps=1e-6) if ln else xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)
    ref = base * (1 + sc.float().repeat_interleave(rpg, 0)[:M]) + sh.float().repeat_interleave(rpg, 0)[:M]
    out = ops.norm_mod(x, sc, sh, rows_per_group=rpg, eps=1e-6, layer_norm=ln)
    assert rel(out, ref) < 8e-3
    assert rel(ops.norm_mod(x, eps=1e-6, layer_norm=ln), base) < 5e-3
    w, b = rnd(D, seed=3), rnd(D, seed=4)
    if ln:
        assert rel(ops.norm_mod(x, weight=w, bias=b, eps=1e-6, layer_norm=True),
                   F.layer_norm(xf, (D,), w.float(), b.float(), eps=1e-6)) < 5e-3


@pytest.mark.parametrize("B,N,D", [(2, 96, 2048), (3, 95, 4096), (1, 96, 1024)])
def test_qk_norm_rope_matches_oracle(B, N, D):
    """D = 2048: one warp per token (qk_norm_rope_tok_kernel); D = 4096 (LTX-Video 13B): a warp pair per token, odd token count;
    both against the fp32 reference and with the v third of the fused buffer untouched"""
    qkv = rnd(B * N, 3 * D, seed=1)
    wq, wk = (1 + 0.1 * rnd(D, seed=2).float()).to(BF), (1 + 0.1 * rnd(D, seed=3).float()).to(BF)
    coords = O.latent_to_pixel_coords(O.latent_coords(2, 6, 8, 1)).float()
    coords[:, 0] /= 25.0
    cos, sin = O.precompute_freqs_cis(coords, D, 10000.0, (20, 2048, 2048), BF)
    cos, sin = cos[0, :N].contiguous().to(DEV), sin[0, :N].contiguous().to(DEV)
    q, k = qkv[:, :D], qkv[:, D:2 * D]
    refq = O.apply_rotary_emb(O.rms_norm(q.float().view(B, N, D), 1e-5, wq.float()), cos.float(), sin.float())
    refk = O.apply_rotary_emb(O.rms_norm(k.float().view(B, N, D), 1e-5, wk.float()), cos.float(), sin.float())
    v_before = qkv[:, 2 * D:].clone()
    ops.qk_norm_rope(q, k, wq, wk, cos, sin, tokens_per_batch=N, eps=1e-5)
    assert rel(q.reshape(B, N, D), refq) < 8e-3 and rel(k.reshape(B, N, D), refk) < 8e-3
    assert torch.equal(qkv[:, 2 * D:], v_before)
    # no rope, q only (cross-attention query)
    q2 = rnd(50, D, seed=9)
    ref = O.rms_norm(q2.float(), 1e-5, wq.float())
    ops.qk_norm_rope(q2, None, wq, None)
    assert rel(q2, ref) < 6e-3


def test_small_elementwise():
    L, G, D = 3, 5, 512
    table, temb = rnd(L, 6, D, seed=1), rnd(G, 6 * D, seed=2)
    ref = (table.float()[:, None] + temb.float().view(1, G, 6, D))
    assert rel(ops.ada_add(table, temb), ref) < 5e-3
    x = rnd(40, 256, seed=3)
    assert rel(ops.act(x, ops.ACT_SILU), F.silu(x.float())) < 5e-3
    assert rel(ops.act(x, ops.ACT_GELU_TANH), F.gelu(x.float(), approximate="tanh")) < 5e-3
    t = torch.tensor([0.0, 731.1, 1000.0, 12.5], device=DEV)
    assert rel(ops.timestep_embed(t), O.timestep_sinusoid(t.cpu())) < 5e-3
    a, v = rnd(3, 10, 256, seed=4), rnd(30, 3 * 256, seed=5)
    m = torch.tensor([1.0, 0.0, 1.0], device=DEV)
    ref = a.float() * m.view(3, 1, 1) + v[:, 512:].float().view(3, 10, 256) * (1 - m.view(3, 1, 1))
    ops.stg_blend(a, v[:, 512:], m)
    assert rel(a, ref) < 5e-3
    xf = torch.randn(1024, device=DEV)
    assert torch.equal(ops.cast_bf16(xf), xf.to(BF))


@pytest.mark.parametrize("C", [128, 256, 512])
def test_pixelnorm_silu(C):
    x = rnd(2, 3, 5, 7, C, seed=1, scale=2.0)
    ref = F.silu(O.pixel_norm(x.float().permute(0, 4, 1, 2, 3))).permute(0, 2, 3, 4, 1)
    assert rel(ops.pixelnorm_silu(x), ref) < 6e-3


def test_latent_to_ndhwc():
    z = torch.randn(1, 128, 2, 3, 4, device=DEV)
    s, m = torch.rand(128, device=DEV) + 0.5, torch.randn(128, device=DEV) * 0.1
    ref = (z * s.view(1, -1, 1, 1, 1) + m.view(1, -1, 1, 1, 1)).permute(0, 2, 3, 4, 1)
    assert rel(ops.latent_to_ndhwc(z, s, m), ref) < 8e-3


@pytest.mark.parametrize("mode", ["plain", "cfg", "cfg_stg", "stg_mask"])
def test_guidance_step(mode):
    N, C, steps = 96, 128, 8
    has_cfg = mode in ("cfg", "cfg_stg")
    has_stg = mode in ("cfg_stg", "stg_mask")
    conds = 1 + has_cfg + has_stg
    pred = rnd(conds, N * C, seed=1)
    lat = torch.randn(N * C, device=DEV)
    ts = O.rf_timesteps(steps, (1, C, 2, 6, 8)).to(DEV)
    i = 3
    cmask = None
    if mode == "stg_mask":
        cmask = torch.zeros(N, device=DEV)
        cmask[:48] = 1.0
    gs, ss, rs = (3.0 if has_cfg else 1.0), (1.0 if has_stg else 0.0), (0.7 if has_stg else 1.0)
    npred = O.guidance_combine(pred.float().cpu().view(conds, N, C), conds, has_cfg, has_stg, gs, ss, rs, rs != 1.0)
    cur = ts[i].cpu()[None, None]
    if cmask is not None:
        cur = torch.min(cur, 1.0 - cmask.cpu()[None])
    ref = O.rf_step(npred, cur, lat.cpu().view(1, N, C), ts.cpu())
    if cmask is not None:
        keep = (ts[i].cpu() - 1e-6 < (1.0 - cmask.cpu()[None])).unsqueeze(-1)
        ref = torch.where(keep, ref, lat.cpu().view(1, N, C))
    scratch = torch.empty(8 * 148, device=DEV)
    lb = torch.empty(N * C, device=DEV, dtype=BF)
    ops.guidance_step(pred, lat, ts, float(ts[i]), num_conds=conds, has_cfg=has_cfg, has_stg=has_stg,
                      do_rescale=rs != 1.0, guidance_scale=gs, stg_scale=ss, rescale=rs, channels=C,
                      cond_mask=cmask, scratch=scratch, latents_bf16=lb)
    assert rel(lat.view(1, N, C), ref) < 2e-3
    assert torch.equal(lb, lat.to(BF))


# ------------------------------------------------------------------ BASELINE.json's full sizes (size-independent properties)
def test_attention_full_size_properties():
    """LTX self-attention at the bench shape (N = 6144, d = 64) and a Wan-length sequence that is not a multiple of the block
    (N = 32760, d = 128, 3 heads): (1) against torch SDPA in fp32 on the same GPU (a library reference, fast enough at this size),
    (2) softmax rows sum to one: with V = 1 the output is exactly 1 up to bf16 rounding, (3) permuting the keys (and V with them)
    leaves the output unchanged — the online softmax does not depend on the block order."""
    for (B, H, N, d) in ((1, 4, 6144, 64), (1, 3, 32760, 128)):
        g = torch.Generator().manual_seed(N)
        q, k, v = [(torch.randn(B, N, H, d, generator=g)).to(BF).to(DEV) for _ in range(3)]
        out = ops.attention(q, k, v)
        ref = F.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)
        e = float((out.float() - ref).norm() / ref.norm())
        print(f"attention full size N={N} d={d}: rel_l2 vs fp32 SDPA = {e:.3e}")
        assert e < 1e-2
        ones = torch.ones_like(v)
        o1 = ops.attention(q, k, ones).float()
        assert float((o1 - 1).abs().max()) < 1e-2
        perm = torch.randperm(N, generator=g).to(DEV)
        o2 = ops.attention(q, k[:, perm].contiguous(), v[:, perm].contiguous())
        assert float((o2.float() - out.float()).norm() / out.float().norm()) < 6e-3


def test_gemm_full_size_linearity():
    """The three LTX GEMM shapes at the bench size (M = 18432 rows = 3 conds x 6144 tokens) through the CTA-pair kernel: against
    torch's bf16 matmul with fp32 accumulation on the same GPU, and linearity gemm(a1 + a2) = gemm(a1) + gemm(a2) with fp32 outputs."""
    M = 18432
    for (N, K) in ((6144, 2048), (8192, 2048), (2048, 8192)):
        g = torch.Generator().manual_seed(N + K)
        a1 = (torch.randn(M, K, generator=g)).to(BF).to(DEV)
        a2 = (torch.randn(M, K, generator=g) * 0.5).to(BF).to(DEV)
        w = (torch.randn(N, K, generator=g) * K ** -0.5).to(BF).to(DEV)
        o1 = ops.gemm(a1, w, None, out_f32=True)
        ref = (a1.float() @ w.float().t())
        e = float((o1 - ref).norm() / ref.norm())
        print(f"gemm {M}x{N}x{K}: rel_l2 vs fp32 matmul = {e:.3e}")
        assert e < 2e-3
        a12 = (a1.float() + a2.float())
        exact = a12.to(BF).float() == a12                     # rows where the bf16 sum is exact make the identity testable
        rows = exact.all(dim=1).nonzero().flatten()[:64]
        if rows.numel():
            o12 = ops.gemm(a12.to(BF)[rows].contiguous(), w, None, out_f32=True)
            o2 = ops.gemm(a2[rows].contiguous(), w, None, out_f32=True)
            assert float((o12 - (o1[rows] + o2)).abs().max()) < 2e-2 * float(o12.abs().max())


def test_pay_attention_entry_point_vs_reference_fixture(golden_dir):
    """Row a1: `pay_attention` (the drop-in of utils/attention.py:161) against outputs of the unmodified reference function on the
    same inputs (tests/golden/pay_attention.pt), with the reference's calling conventions."""
    from ltx_video_gpupoor_b200.attention import get_attention_modes, get_supported_attention_modes, pay_attention
    assert get_attention_modes() == get_supported_attention_modes() == ["b200"]
    g = torch.load(os.path.join(golden_dir, "pay_attention.pt"), weights_only=False)
    for name, c in g.items():
        kw = {k: (v.to(DEV) if torch.is_tensor(v) and k == "attention_mask" else v) for k, v in c["kw"].items()}
        lst = [c["q"].to(DEV), c["k"].to(DEV), c["v"].to(DEV)]
        y = pay_attention(lst, **kw)
        torch.cuda.synchronize()
        assert lst == [], "ownership of q/k/v passes to the callee (utils/attention.py:185-186)"
        assert y.dtype == c["q"].dtype and tuple(y.shape) == tuple(c["out"].shape)
        e = float((y[:, : c["valid"]].float().cpu() - c["out"][:, : c["valid"]].float()).norm() / c["out"][:, : c["valid"]].float().norm())
        print(f"pay_attention[{name}] rel_l2 vs reference = {e:.3e}")
        assert e < 1e-2, name
    with pytest.raises(NotImplementedError):
        pay_attention([g["plain"]["q"].to(DEV), g["plain"]["k"].to(DEV), g["plain"]["v"].to(DEV)], causal=True)
    with pytest.raises(Exception):
        pay_attention([g["plain"]["q"], g["plain"]["k"], g["plain"]["v"]])          # CPU tensors: no fallback


def test_torch_custom_ops_match_direct_calls():
    """torch.ops.ltxb200.* (custom_ops.py) run the same kernels as ops.py; torch.compile treats them as opaque ops."""
    from ltx_video_gpupoor_b200 import custom_ops  # noqa: F401
    a, w, b = rnd(300, 128, seed=1), rnd(512, 128, seed=2, scale=0.1), rnd(512, seed=3)
    assert torch.equal(torch.ops.ltxb200.gemm(a, w, b, ops.ACT_GELU_TANH), ops.gemm(a, w, b, act=ops.ACT_GELU_TANH))
    q, k, v = rnd(1, 200, 2, 64, seed=4), rnd(1, 150, 2, 64, seed=5), rnd(1, 150, 2, 64, seed=6)
    assert torch.equal(torch.ops.ltxb200.attention(q, k, v, None, 0.0), ops.attention(q, k, v))
    res, gate = rnd(300, 512, seed=7), rnd(3, 512, seed=8)
    assert torch.equal(torch.ops.ltxb200.gemm_gate_residual(a, w, b, res, gate, 100), ops.gemm(a, w, b, residual=res, gate=gate, rows_per_gate=100))
    f = torch.compile(lambda a_, w_, b_: torch.ops.ltxb200.gemm(a_, w_, b_, 0) * 2, fullgraph=True, backend="eager")   # dynamo trace only
    assert torch.equal(f(a, w, b), ops.gemm(a, w, b) * 2)


def test_conv3d_cta_pair_matches_single_cta(monkeypatch):
    """The CTA-pair (cta_group::2) form of the implicit-GEMM convolution (LTXB200_CONV_2CTA=2 forces it) gives the same result as the
    single-CTA kernel, including an odd patch count (the last pair's second CTA is all padding) and both tile widths."""
    for (B, T, H, W, Cin, Cout, causal) in ((1, 3, 8, 16, 64, 128, True), (1, 3, 9, 20, 128, 256, False), (1, 5, 8, 16, 64, 64, True)):
        x = rnd(B, T, H, W, Cin, seed=1)
        w5 = rnd(Cout, Cin, 3, 3, 3, seed=2, scale=(27 * Cin) ** -0.5)
        b = rnd(Cout, seed=3)
        monkeypatch.setenv("LTXB200_CONV_2CTA", "0")
        one = ops.conv3d(x, pack_w(w5), b, causal=causal)
        monkeypatch.setenv("LTXB200_CONV_2CTA", "2")
        two = ops.conv3d(x, pack_w(w5), b, causal=causal)
        torch.cuda.synchronize()
        assert torch.equal(one, two)
        assert rel(two, ref_conv(x, w5, b, causal).permute(0, 2, 3, 4, 1)) < 6e-3


@pytest.mark.parametrize("B,T,H,W,Cin,Cout,causal", [(1, 3, 16, 16, 128, 128, False),      # exactly one 16x16 patch per frame
                                                     (1, 2, 21, 37, 128, 128, True),        # ragged: partial patches on both axes
                                                     (2, 3, 5, 7, 64, 48, False),           # smaller than a patch, BN = 64 tile
                                                     (1, 4, 32, 48, 256, 128, False),       # 4 channel slices, 6 patches per frame
                                                     (1, 1, 3, 3, 64, 64, True)])
def test_conv3d_halo_tiled(B, T, H, W, Cin, Cout, causal, monkeypatch):
    """The halo-tiled convolution (conv_halo.cuh; default for Cout <= 128, 3x3 spatial taps, unit stride) against fp32 torch and against
    the plain implicit-GEMM kernel (LTXB200_CONV_HALO=0) on the same inputs: same result up to the order of the fp32 accumulation."""
    x = rnd(B, T, H, W, Cin, seed=1)
    w5 = rnd(Cout, Cin, 3, 3, 3, seed=2, scale=(27 * Cin) ** -0.5)
    b = rnd(Cout, seed=3)
    res = rnd(B, T, H, W, Cout, seed=4)
    ref = ref_conv(x, w5, b, causal).permute(0, 2, 3, 4, 1)
    monkeypatch.setenv("LTXB200_CONV_HALO", "0")
    plain = ops.conv3d(x, pack_w(w5), b, causal=causal, residual=res)
    monkeypatch.setenv("LTXB200_CONV_HALO", "1")
    halo = ops.conv3d(x, pack_w(w5), b, causal=causal, residual=res)
    halo_nores = ops.conv3d(x, pack_w(w5), b, causal=causal)
    torch.cuda.synchronize()
    assert rel(halo, ref + res.float()) < 6e-3 and rel(halo_nores, ref) < 6e-3
    assert rel(halo, plain.float()) < 2e-3
    # Wan-style variants through conv_taps: zero temporal padding (causal) and a purely spatial 1x3x3 convolution
    xz = torch.cat([torch.zeros_like(x[:, :1]).repeat(1, 2, 1, 1, 1), x], dim=1).float().permute(0, 4, 1, 2, 3)
    refz = F.conv3d(xz, w5.float(), b.float(), padding=(0, 1, 1)).permute(0, 2, 3, 4, 1)
    assert rel(ops.conv_taps(x, pack_w(w5), b, 3, 3, zero_pad_t=True), refz) < 6e-3
    w2 = w5[:, :, 1].contiguous()                                                              # [Cout, Cin, 3, 3]
    ref2 = F.conv2d(x.float().reshape(B * T, H, W, Cin).permute(0, 3, 1, 2), w2.float(), b.float(), padding=1).permute(0, 2, 3, 1)
    out2 = ops.conv_taps(x, w2.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous(), b, 1, 3)
    assert rel(out2.reshape(B * T, H, W, Cout), ref2) < 6e-3


def test_conv3d_halo_unpatch_store(monkeypatch):
    """conv_out of the LTX VAE decoder (128 -> 48 channels, unpatchify store to NCFHW) on the halo-tiled kernel's 64-column tile."""
    B, T, H, W, Cin = 1, 3, 18, 20, 128
    x = rnd(B, T, H, W, Cin, seed=1)
    w5 = rnd(48, Cin, 3, 3, 3, seed=4, scale=(27 * Cin) ** -0.5)
    b = rnd(48, seed=5)
    ref = O.vae_unpatchify(ref_conv(x, w5, b, False), 4)
    perm = torch.arange(48).reshape(3, 4, 4).permute(0, 2, 1).reshape(-1).to(DEV)
    monkeypatch.setenv("LTXB200_CONV_HALO", "1")
    out = ops.conv3d(x, pack_w(w5[perm]), b[perm].contiguous(), store=ops.CONV_UNPATCH, out_f32=True)
    assert out.shape == ref.shape and rel(out, ref) < 6e-3


@pytest.mark.parametrize("Cin,Cout,H,W", [(128, 128, 16, 16), (128, 128, 9, 21), (256, 256, 8, 16), (64, 48, 5, 7)])
@pytest.mark.parametrize("keep_raw", [True, False])
def test_conv3d_fused_pixelnorm_silu_output(Cin, Cout, H, W, keep_raw):
    """ltxb200_conv3d_norm_bf16: the convolution's epilogue also writes silu(pixelnorm(y)) of its own output row (halo-tiled kernel for
    Cout <= 128, plain implicit GEMM with the 256-column tile above) == conv3d followed by the standalone pixelnorm_silu kernel."""
    B, T = 1, 3
    x = rnd(B, T, H, W, Cin, seed=1)
    w5 = rnd(Cout, Cin, 3, 3, 3, seed=2, scale=(27 * Cin) ** -0.5)
    b = rnd(Cout, seed=3)
    res = rnd(B, T, H, W, Cout, seed=4)
    y_ref = ops.conv3d(x, pack_w(w5), b, residual=res)
    if Cout % 64 == 0:
        n_ref = ops.pixelnorm_silu(y_ref)
    else:                                        # the standalone kernel has no 48-channel form: fp32 torch on the stored bf16 row
        yf = y_ref.float()
        t = (yf * torch.rsqrt(yf.pow(2).mean(-1, keepdim=True) + 1e-8)).bfloat16().float()
        n_ref = F.silu(t).bfloat16()
    y, n = ops.conv3d_norm(x, pack_w(w5), b, residual=res, keep_raw=keep_raw)
    torch.cuda.synchronize()
    assert (y is None) == (not keep_raw)
    if keep_raw:
        assert torch.equal(y, y_ref)
    assert rel(n, n_ref.float()) < 2e-3
    yf = (ref_conv(x, w5, b, False).permute(0, 2, 3, 4, 1) + res.float())
    assert rel(n, F.silu(yf * torch.rsqrt(yf.pow(2).mean(-1, keepdim=True) + 1e-8))) < 8e-3


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("B,H,Lq,Lk,lens", [(3, 4, 700, 256, [256, 77, 1]), (2, 2, 1300, 256, [129, 128]), (4, 32, 6144, 256, [200, 200, 31, 256]),
                                           (2, 3, 300, 640, [513, 60])])
def test_attention_key_lens(d, B, H, Lq, Lk, lens):
    """Per-batch key lengths (ltxb200_attention_klens_bf16): element b attends to its first lens[b] keys — the right-padded prompt mask of
    cross-attention without the bias pass and without the padded key blocks.  Same result as the (1 - mask) * -10000 bias the reference
    builds (transformer3d.py:411-415) and as the fp32 reference; every CTA walks items with DIFFERENT numbers of key blocks here."""
    q, k, v = rnd(B, Lq, H, d, seed=1), rnd(B, Lk, H, d, seed=2), rnd(B, Lk, H, d, seed=3)
    kl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    bias = torch.zeros(B, Lk, device=DEV)
    for b, n in enumerate(lens):
        bias[b, n:] = -10000.0
    out = ops.attention(q, k, v, key_lens=kl)
    out_bias = ops.attention(q, k, v, key_bias=bias)
    torch.cuda.synchronize()
    assert rel(out, out_bias.float()) < 4e-3
    if B * H * Lq * Lk <= 3 * 4 * 1300 * 640:
        assert rel(out, ref_attn(q, k, v, bias)) < 1e-2
    # run it again right behind an ordinary launch: ring / barrier phases must not depend on a launch-wide block count
    o2 = ops.attention(q, k, v, key_lens=kl)
    assert torch.equal(o2, out)


def test_attention64_pipelined_variant_parity():
    """The opt-in cross-block pipelined d = 64 kernel (csrc/attention64p.cuh, LTXB200_ATTN64P=1; the switch is read once per process, hence the
    subprocess): parity on shapes with several work items per CTA, partial query tiles and strided fused-QKV operands, deterministic."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LTXB200_ATTN64P="1")
    r = subprocess.run([sys.executable, os.path.join(root, "profiles", "scripts", "attn64p_check.py")], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
    assert "ATTN64P = 1" in r.stdout


def test_attention128_cta_pair_variant_parity():
    """The opt-in CTA-pair d = 128 kernel (csrc/attention128p2.cuh, LTXB200_ATTN128_2CTA=1; the switch is read once per process, hence the
    subprocess): cluster of 2, cta_group::2 MMAs, half of K / V per CTA, remote barrier arrivals.  Parity on partial query groups and key
    blocks, strided fused-QKV operands, key bias / key lengths, large logits; deterministic."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LTXB200_ATTN128_2CTA="1")
    r = subprocess.run([sys.executable, os.path.join(root, "profiles", "scripts", "attn128p2_check.py")], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
    assert "ATTN128_2CTA = 1" in r.stdout

