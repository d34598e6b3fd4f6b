"""CPU oracle for the LTX-Video denoising hot path (TEST INFRASTRUCTURE ONLY).

A plain-PyTorch restatement of the reference's algorithm, operating on a flat
``state_dict`` with the reference's own key names.  It is the checker for the CUDA
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it.  The product package never does.

Pinning: ``oracle/gen_golden.py`` runs the UNMODIFIED reference modules (imported
from /root/reference through ``oracle/refshim``) on the same seeded weights/inputs
and asserts this file reproduces them (fp32, CPU) before writing tests/golden/*.
The diffusers pieces the reference calls (AdaLayerNormSingle, RMSNorm, GELU,
PixArtAlphaTextProjection) are third-party (diffusers>=0.31, requirements.txt:4,
not vendored): they are restated here from the published v0.31 source; no
reference test pins them ("parity unpinned" for that boundary), except the timestep
sinusoid, which the reference vendors (ltx_video/models/transformers/embeddings.py:10-50): bit-exact fixture.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# Architecture constants (ltx_video/utils/diffusers_config_mapping.py:74-130)
# --------------------------------------------------------------------------------------
LTX_2B = dict(
    num_layers=28, num_attention_heads=32, attention_head_dim=64, in_channels=128,
    out_channels=128, cross_attention_dim=2048, caption_channels=4096,
    norm_eps=1e-6, qk_norm_eps=1e-5, rope_theta=10000.0, rope_max_pos=(20, 2048, 2048),
    timestep_scale_multiplier=1000.0, ff_mult=4,
)

# decoder block list = reversed(OURS_VAE_CONFIG["blocks"]) (causal_video_autoencoder.py:609,629)
LTX_VAE = dict(
    latent_channels=128, out_channels=3, patch_size=4, base_channels=128,
    blocks=[["res_x", 4], ["compress_all", 1], ["res_x_y", 1], ["res_x", 3],
            ["compress_all", 1], ["res_x_y", 1], ["res_x", 3], ["compress_all", 1],
            ["res_x", 3], ["res_x", 4]],
    causal_decoder=False,
)

SKIP_ATTENTION_VALUES = "attention_values"   # SkipLayerStrategy.AttentionValues
SKIP_ATTENTION_SKIP = "attention_skip"
SKIP_RESIDUAL = "residual"
SKIP_TRANSFORMER_BLOCK = "transformer_block"


# --------------------------------------------------------------------------------------
# Seeded random-init weights of the named architecture (no checkpoints: no network)
# --------------------------------------------------------------------------------------
def _uniform(gen, shape, bound, dtype=torch.float32):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1).mul_(bound).to(dtype)


def _linear(sd, gen, name, out_f, in_f, bias=True):
    b = 1.0 / math.sqrt(in_f)      # nn.Linear default: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
    sd[name + ".weight"] = _uniform(gen, (out_f, in_f), b)
    if bias:
        sd[name + ".bias"] = _uniform(gen, (out_f,), b)


def make_transformer_state_dict(cfg: dict = LTX_2B, seed: int = 0, num_layers: Optional[int] = None
                                ) -> Dict[str, Tensor]:
    """Random-init fp32 state_dict with the key names/shapes of Transformer3DModel
    (transformer3d.py:91-156, attention.py:124-185,477-558,1263-1310)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    L = cfg["num_layers"] if num_layers is None else num_layers
    sd: Dict[str, Tensor] = {}
    sd["scale_shift_table"] = torch.randn(2, D, generator=gen) / D ** 0.5
    _linear(sd, gen, "patchify_proj", D, cfg["in_channels"])
    for i in range(L):
        p = f"transformer_blocks.{i}."
        sd[p + "scale_shift_table"] = torch.randn(6, D, generator=gen) / D ** 0.5
        for a, kv_dim in (("attn1", D), ("attn2", cfg["cross_attention_dim"])):
            # RMSNorm weights init to ones in the reference; perturb so the test sees them
            sd[p + a + ".q_norm.weight"] = 1.0 + 0.1 * torch.randn(D, generator=gen)
            sd[p + a + ".k_norm.weight"] = 1.0 + 0.1 * torch.randn(D, generator=gen)
            _linear(sd, gen, p + a + ".to_q", D, D)
            _linear(sd, gen, p + a + ".to_k", D, kv_dim)
            _linear(sd, gen, p + a + ".to_v", D, kv_dim)
            _linear(sd, gen, p + a + ".to_out.0", D, D)
        _linear(sd, gen, p + "ff.net.0.proj", cfg["ff_mult"] * D, D)
        _linear(sd, gen, p + "ff.net.2", D, cfg["ff_mult"] * D)
    _linear(sd, gen, "proj_out", cfg["out_channels"], D)
    _linear(sd, gen, "adaln_single.emb.timestep_embedder.linear_1", D, 256)
    _linear(sd, gen, "adaln_single.emb.timestep_embedder.linear_2", D, D)
    _linear(sd, gen, "adaln_single.linear", 6 * D, D)
    _linear(sd, gen, "caption_projection.linear_1", D, cfg["caption_channels"])
    _linear(sd, gen, "caption_projection.linear_2", D, D)
    return sd


def vae_decoder_plan(cfg: dict = LTX_VAE) -> List[Tuple[str, int, int, int]]:
    """[(kind, index, c_in, c_out)] for decoder.up_blocks (causal_video_autoencoder.py:609-700).
    kind in {"res_x" (n layers packed as index list), "res_x_y", "d2s"}."""
    ch = cfg["base_channels"]
    for name, p in reversed(cfg["blocks"]):
        if name == "res_x_y":
            ch *= 2
    plan = []
    for idx, (name, p) in enumerate(reversed(cfg["blocks"])):
        if name == "res_x":
            plan.append(("res_x", idx, ch, ch, int(p)))
        elif name == "res_x_y":
            plan.append(("res_x_y", idx, ch, ch // 2, 1))
            ch //= 2
        elif name == "compress_all":
            plan.append(("d2s", idx, ch, ch, 1))
        else:
            raise ValueError(name)
    return plan


def vae_top_channels(cfg: dict = LTX_VAE) -> int:
    ch = cfg["base_channels"]
    for name, _ in cfg["blocks"]:
        if name == "res_x_y":
            ch *= 2
    return ch


def _conv3d(sd, gen, name, c_out, c_in, k=3):
    fan_in = c_in * k * k * k
    b = 1.0 / math.sqrt(fan_in)
    sd[name + ".weight"] = _uniform(gen, (c_out, c_in, k, k, k), b)
    sd[name + ".bias"] = _uniform(gen, (c_out,), b)


def make_vae_decoder_state_dict(cfg: dict = LTX_VAE, seed: int = 1) -> Dict[str, Tensor]:
    """Random-init decoder weights + per-channel latent statistics (keys as in
    causal_video_autoencoder.py; statistics as registered by load_state_dict :285-297)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    top = vae_top_channels(cfg)
    _conv3d(sd, gen, "decoder.conv_in.conv", top, cfg["latent_channels"])
    for kind, idx, cin, cout, n in vae_decoder_plan(cfg):
        p = f"decoder.up_blocks.{idx}."
        if kind == "res_x":
            for j in range(n):
                _conv3d(sd, gen, p + f"res_blocks.{j}.conv1.conv", cin, cin)
                _conv3d(sd, gen, p + f"res_blocks.{j}.conv2.conv", cin, cin)
        elif kind == "res_x_y":
            _conv3d(sd, gen, p + "conv1.conv", cout, cin)
            _conv3d(sd, gen, p + "conv2.conv", cout, cout)
            _conv3d(sd, gen, p + "conv_shortcut", cout, cin, k=1)
            sd[p + "norm3.norm.weight"] = 1.0 + 0.1 * torch.randn(cin, generator=gen)
            sd[p + "norm3.norm.bias"] = 0.1 * torch.randn(cin, generator=gen)
        else:
            _conv3d(sd, gen, p + "conv.conv", 8 * cin, cin)
    last = cfg["base_channels"]
    _conv3d(sd, gen, "decoder.conv_out.conv", cfg["out_channels"] * cfg["patch_size"] ** 2, last)
    sd["std_of_means"] = 0.5 + torch.rand(cfg["latent_channels"], generator=gen)
    sd["mean_of_means"] = 0.1 * torch.randn(cfg["latent_channels"], generator=gen)
    if cfg.get("timestep_conditioning", False):          # causal_video_autoencoder.py:724-733, 850-853, 1207-1210
        def temb(name, dim):
            _linear(sd, gen, name + ".timestep_embedder.linear_1", dim, 256)
            _linear(sd, gen, name + ".timestep_embedder.linear_2", dim, dim)
        sd["decoder.timestep_scale_multiplier"] = torch.tensor(1000.0)
        for kind, idx, cin, cout, n in vae_decoder_plan(cfg):
            if kind == "res_x":
                p = f"decoder.up_blocks.{idx}."
                temb(p + "time_embedder", 4 * cin)
                for j in range(n):
                    sd[p + f"res_blocks.{j}.scale_shift_table"] = torch.randn(4, cin, generator=gen) / cin ** 0.5
        temb("decoder.last_time_embedder", 2 * last)
        sd["decoder.last_scale_shift_table"] = torch.randn(2, last, generator=gen) / last ** 0.5
    return sd


# --------------------------------------------------------------------------------------
# Patchifier / coordinates — integer, bit-exact (symmetric_patchifier.py:33-84,
# vae_encode.py:190-225)
# --------------------------------------------------------------------------------------
def patchify(latents: Tensor) -> Tensor:
    """b c f h w -> b (f h w) c   (patch size 1; symmetric_patchifier.py:55-65)"""
    b, c, f, h, w = latents.shape
    return latents.permute(0, 2, 3, 4, 1).reshape(b, f * h * w, c)


def unpatchify(tokens: Tensor, f: int, h: int, w: int) -> Tensor:
    """b (f h w) c -> b c f h w   (symmetric_patchifier.py:67-84)"""
    b, n, c = tokens.shape
    return tokens.reshape(b, f, h, w, c).permute(0, 4, 1, 2, 3)


def latent_coords(f: int, h: int, w: int, batch: int, device="cpu") -> Tensor:
    """int64 [b,3,f*h*w] of (t,y,x) latent corner coords (symmetric_patchifier.py:33-51)."""
    g = torch.meshgrid(torch.arange(f, device=device), torch.arange(h, device=device),
                       torch.arange(w, device=device), indexing="ij")
    c = torch.stack(g, dim=0).reshape(3, -1)
    return c.unsqueeze(0).repeat(batch, 1, 1)


def latent_to_pixel_coords(coords: Tensor, scale=(8, 32, 32), causal_fix: bool = False) -> Tensor:
    """vae_encode.py:214-225"""
    px = coords * torch.tensor(scale, device=coords.device)[None, :, None]
    if causal_fix:
        px[:, 0] = (px[:, 0] + 1 - scale[0]).clamp(min=0)
    return px


# --------------------------------------------------------------------------------------
# RoPE table (transformer3d.py:192-255) and application (attention.py:960-975)
# --------------------------------------------------------------------------------------
def rope_freq_indices(dim: int, theta: float, device="cpu") -> Tensor:
    """theta ** linspace(0, 1, dim//6) * pi/2   (transformer3d.py:213-233, spacing 'exp')"""
    idx = theta ** torch.linspace(math.log(1, theta), math.log(theta, theta), dim // 6,
                                  device=device, dtype=torch.float32)
    return idx.to(torch.float32) * math.pi / 2


def precompute_freqs_cis(indices_grid: Tensor, dim: int, theta: float, max_pos: Sequence[int],
                         out_dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """indices_grid float [B,3,N] -> (cos, sin) [B,N,dim] (transformer3d.py:202-255)."""
    frac = torch.stack([indices_grid[:, i] / max_pos[i] for i in range(3)], dim=-1)  # [B,N,3]
    indices = rope_freq_indices(dim, theta, frac.device)
    freqs = (indices * (frac.unsqueeze(-1) * 2 - 1)).transpose(-1, -2).flatten(2)   # [B,N,F*3] (freq-major)
    cos = freqs.cos().repeat_interleave(2, dim=-1)
    sin = freqs.sin().repeat_interleave(2, dim=-1)
    if dim % 6 != 0:
        cos = torch.cat([torch.ones_like(cos[:, :, : dim % 6]), cos], dim=-1)
        sin = torch.cat([torch.zeros_like(sin[:, :, : dim % 6]), sin], dim=-1)
    return cos.to(out_dtype), sin.to(out_dtype)


def apply_rotary_emb(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """pairs (2i,2i+1) -> (-x[2i+1], x[2i]) (attention.py:960-975)"""
    x2 = x.reshape(*x.shape[:-1], -1, 2)
    rot = torch.stack((-x2[..., 1], x2[..., 0]), dim=-1).reshape(x.shape)
    return x * cos + rot * sin


# --------------------------------------------------------------------------------------
# diffusers pieces (restated; see module docstring)
# --------------------------------------------------------------------------------------
def timestep_sinusoid(t: Tensor, dim: int = 256) -> Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def rms_norm(x: Tensor, eps: float, weight: Optional[Tensor] = None) -> Tensor:
    """diffusers RMSNorm.forward: fp32 variance, x*rsqrt, cast to weight dtype, *weight."""
    dt = x.dtype
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(var + eps)
    if weight is not None:
        if weight.dtype in (torch.float16, torch.bfloat16):
            y = y.to(weight.dtype)
        return y * weight
    return y.to(dt)


def linear(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def adaln_single(sd, t_flat: Tensor, dtype) -> Tuple[Tensor, Tensor]:
    """AdaLayerNormSingle (diffusers) as used at transformer3d.py:428-433."""
    proj = timestep_sinusoid(t_flat).to(dtype)
    e = linear(sd, "adaln_single.emb.timestep_embedder.linear_1", proj)
    e = linear(sd, "adaln_single.emb.timestep_embedder.linear_2", F.silu(e))
    return linear(sd, "adaln_single.linear", F.silu(e)), e


def attention_core(q: Tensor, k: Tensor, v: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """Non-causal softmax attention, layout [B, L, H, d] in and out, scale d^-0.5
    (utils/attention.py:99-116 sdpa_wrapper semantics).  bias: additive, broadcastable
    to [B, H, Lq, Lk].  Explicit fp32 math, chunked over heads to bound memory."""
    B, Lq, H, d = q.shape
    out = torch.empty_like(q)
    scale = d ** -0.5
    hc = max(1, min(H, (1 << 28) // max(1, Lq * k.shape[1])))
    for h0 in range(0, H, hc):
        qs = q[:, :, h0:h0 + hc].permute(0, 2, 1, 3).float()
        ks = k[:, :, h0:h0 + hc].permute(0, 2, 1, 3).float()
        vs = v[:, :, h0:h0 + hc].permute(0, 2, 1, 3).float()
        s = torch.matmul(qs, ks.transpose(-1, -2)) * scale
        if bias is not None:
            b = bias.float()
            if b.shape[1] != 1:
                b = b[:, h0:h0 + hc]
            s = s + b
        p = torch.softmax(s, dim=-1)
        out[:, :, h0:h0 + hc] = torch.matmul(p, vs).permute(0, 2, 1, 3).to(q.dtype)
    return out


# --------------------------------------------------------------------------------------
# Transformer block and model forward
# --------------------------------------------------------------------------------------
def _attn(sd, p, x, ctx, heads, cos_sin, bias, eps_qk, skip_mask=None, strategy=None):
    """AttnProcessor2_0.__call__ (attention.py:986-1173); p = 'transformer_blocks.i.attn1'."""
    B = x.shape[0]
    is_self = ctx is None
    src = x if is_self else ctx
    q = rms_norm(linear(sd, p + ".to_q", x), eps_qk, sd[p + ".q_norm.weight"])
    k = rms_norm(linear(sd, p + ".to_k", src), eps_qk, sd[p + ".k_norm.weight"])
    if is_self and cos_sin is not None:
        k = apply_rotary_emb(k, *cos_sin)
        q = apply_rotary_emb(q, *cos_sin)
    v = linear(sd, p + ".to_v", src)
    d = q.shape[-1] // heads
    o = attention_core(q.reshape(B, -1, heads, d), k.reshape(B, -1, heads, d),
                       v.reshape(B, -1, heads, d), bias).reshape(B, -1, heads * d)
    if skip_mask is not None and strategy == SKIP_ATTENTION_VALUES:
        m = skip_mask.reshape(B, 1, 1).to(o.dtype)
        o = o * m + v * (1.0 - m)                       # attention.py:1134-1139
    elif skip_mask is not None and strategy == SKIP_ATTENTION_SKIP:
        m = skip_mask.reshape(B, 1, 1).to(o.dtype)
        o = o * m + x * (1.0 - m)                       # attention.py:1127-1133
    return linear(sd, p + ".to_out.0", o)


def transformer_block(sd, i, h, cos_sin, ctx, ctx_bias, temb6, cfg, skip_mask=None, strategy=None):
    """BasicTransformerBlock.forward (attention.py:205-364), adaptive_norm single_scale_shift,
    rms_norm standardisation, no affine.  temb6: [B, T, 6*D] with T = 1 or latent frames."""
    p = f"transformer_blocks.{i}."
    B, N, D = h.shape
    H = cfg["num_attention_heads"]
    if skip_mask is not None and float(skip_mask.flatten().min()) == 1.0:
        skip_mask = None
    T = temb6.shape[1]
    ada = sd[p + "scale_shift_table"][None, None] + temb6.reshape(B, T, 6, D)     # :239-241
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = [
        a.unsqueeze(2) for a in ada.unbind(dim=2)]                                  # [B,T,1,D]

    def mod(x, scale, shift):
        x = x.reshape(B, T, -1, D)
        x = x * (1 + scale) + shift
        return x.reshape(B, N, D)

    original = h
    nh = mod(rms_norm(h, cfg["norm_eps"]), scale_msa, shift_msa)
    a = _attn(sd, p + "attn1", nh, None, H, cos_sin, None, cfg["qk_norm_eps"], skip_mask, strategy)
    h = h + (a.reshape(B, T, -1, D) * gate_msa).reshape(B, N, D)                    # :282-288
    a = _attn(sd, p + "attn2", h, ctx, H, None, ctx_bias, cfg["qk_norm_eps"])       # :294-311
    h = h + a
    nh = mod(rms_norm(h, cfg["norm_eps"]), scale_mlp, shift_mlp)                    # :314-320
    ff = linear(sd, p + "ff.net.2", F.gelu(linear(sd, p + "ff.net.0.proj", nh), approximate="tanh"))
    h = h + (ff.reshape(B, T, -1, D) * gate_mlp).reshape(B, N, D)                   # :345-351
    if skip_mask is not None and strategy == SKIP_TRANSFORMER_BLOCK:
        m = skip_mask.reshape(-1, 1, 1).to(h.dtype)
        h = h * m + original * (1.0 - m)                                            # :355-362
    return h


def transformer_forward(sd: Dict[str, Tensor], cfg: dict, hidden: Tensor, cos_sin, enc: Tensor,
                        timestep: Tensor, enc_mask: Optional[Tensor] = None,
                        skip_layer_mask: Optional[Tensor] = None, strategy: Optional[str] = None,
                        latent_shape: Optional[Sequence[int]] = None,
                        collect: Optional[list] = None) -> Tensor:
    """Transformer3DModel.forward (transformer3d.py:328-507), joint_pass=True.
    hidden [B,N,128]; cos_sin (cos,sin) [1|B,N,D]; enc [B,L,4096]; timestep [B,1]|[B,N];
    enc_mask [B,L] (1 keep / 0 drop); skip_layer_mask [layers,B]."""
    dtype = hidden.dtype
    L = sum(1 for k in sd if k.endswith(".attn1.to_q.weight"))
    ctx_bias = None
    if enc_mask is not None:
        ctx_bias = ((1 - enc_mask.to(dtype)) * -10000.0)[:, None, None, :]          # :411-415 -> [B,1,1,L]
    h = linear(sd, "patchify_proj", hidden)                                         # :418
    t = cfg["timestep_scale_multiplier"] * timestep                                  # :420-421
    if t.shape[-1] > 1:                                                              # :423-425
        t = t.reshape(t.shape[0], -1, latent_shape[-2] * latent_shape[-1])[:, :, 0]
    B = h.shape[0]
    temb6, emb = adaln_single(sd, t.flatten(), dtype)
    temb6 = temb6.view(B, -1, temb6.shape[-1])
    emb = emb.view(B, -1, emb.shape[-1])
    ctx = linear(sd, "caption_projection.linear_2",
                 F.gelu(linear(sd, "caption_projection.linear_1", enc), approximate="tanh"))
    ctx = ctx.view(B, -1, h.shape[-1])                                              # :446-451
    for i in range(L):
        h = transformer_block(sd, i, h, cos_sin, ctx, ctx_bias, temb6, cfg,
                              None if skip_layer_mask is None else skip_layer_mask[i], strategy)
        if collect is not None:
            collect.append(h)
    ss = sd["scale_shift_table"][None, None] + emb[:, :, None]                      # :490-493
    shift, scale = ss[:, :, 0].unsqueeze(-2), ss[:, :, 1].unsqueeze(-2)
    D = h.shape[-1]
    h = F.layer_norm(h, (D,), eps=1e-6)                                             # :494
    T = scale.shape[1]
    h = (h.reshape(B, T, -1, D) * (1 + scale) + shift).reshape(B, -1, D)            # :498-502
    return linear(sd, "proj_out", h)


# --------------------------------------------------------------------------------------
# RectifiedFlowScheduler (rf.py)
# --------------------------------------------------------------------------------------
def rf_timesteps(num_steps: int, samples_shape: Sequence[int], terminal: Optional[float] = 0.1
                 ) -> Tensor:
    """set_timesteps with shifting='SD3' (rf.py:69-149,196-257)."""
    t = torch.linspace(1, 1 / num_steps, num_steps)
    m = samples_shape[1] if len(samples_shape) == 3 else math.prod(samples_shape[2:])
    slope = (2.05 - 0.95) / (4096 - 1024)
    mu = slope * m + (0.95 - slope * 1024)                                           # :73-82
    ts = math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** 1)                            # :69-70
    if terminal is not None:                                                         # :85-109
        one_minus = 1 - ts
        ts = 1 - (one_minus / (one_minus[-1] / (1 - terminal)))
    return ts


def rf_step(model_output: Tensor, timestep: Tensor, sample: Tensor, timesteps: Tensor) -> Tensor:
    """RectifiedFlowScheduler.step, deterministic branch (rf.py:311-380)."""
    t_eps = 1e-6
    padded = torch.cat([timesteps, torch.zeros(1, device=timesteps.device)])
    if timestep.ndim == 0:
        lower = padded[padded < timestep - t_eps][0]
        dt = timestep - lower
    else:
        assert timestep.ndim == 2
        lower_mask = padded[:, None, None] < timestep[None] - t_eps
        lower, _ = (lower_mask * padded[:, None, None]).max(dim=0)
        dt = (timestep - lower)[..., None]
    return sample - dt * model_output


def rf_step_stochastic(model_output: Tensor, timestep: Tensor, sample: Tensor, timesteps: Tensor, noise: Tensor) -> Tensor:
    """RectifiedFlowScheduler.step, stochastic_sampling branch (rf.py:369-373, add_noise :382-392): the x0 estimate x - t*v is re-noised
    to the next lower timestep with `noise` (the reference draws torch.randn_like(sample) from the global RNG)."""
    dt = sample - rf_step(torch.ones_like(model_output), timestep, sample, timesteps)      # = dt, broadcast like the deterministic branch
    t = timestep[..., None] if timestep.ndim == 2 else timestep
    x0 = sample - t * model_output
    nxt = t - dt
    return (1 - nxt) * x0 + nxt * noise


# --------------------------------------------------------------------------------------
# Guidance arithmetic + denoise loop (pipeline_ltx_video.py:1103-1256,1309-1342)
# --------------------------------------------------------------------------------------
def guidance_combine(noise_pred: Tensor, num_conds: int, do_cfg: bool, do_stg: bool,
                     guidance_scale: float, stg_scale: float, rescaling_scale: float,
                     do_rescaling: bool) -> Tensor:
    """pipeline_ltx_video.py:1183-1222 (cfg_star_rescale=True).  noise_pred [num_conds*b, N, C]."""
    chunks = noise_pred.chunk(num_conds)
    batch = chunks[0].shape[0]
    if do_stg:
        text, perturb = chunks[-2:]
    if do_cfg and guidance_scale != 0 and guidance_scale != 1:
        uncond, text = chunks[:2]
        pos = text.reshape(batch, -1)
        neg = uncond.reshape(batch, -1)
        alpha = torch.sum(pos * neg, dim=1, keepdim=True) / (torch.sum(neg ** 2, dim=1, keepdim=True) + 1e-8)
        uncond = alpha.view(batch, 1, 1) * uncond if uncond.ndim == 3 else alpha * uncond
        out = uncond + guidance_scale * (text - uncond)
    elif do_stg:
        out = text
    else:
        out = noise_pred
    if do_stg:
        out = out + stg_scale * (text - perturb)
        if do_rescaling and stg_scale > 0.0:
            f = text.reshape(batch, -1).std(dim=1, keepdim=True) / out.reshape(batch, -1).std(dim=1, keepdim=True)
            f = rescaling_scale * f + (1 - rescaling_scale)
            out = out * f.view(batch, 1, 1)
    return out


def denoise_loop(sd, cfg, latents: Tensor, enc: Tensor, enc_mask: Tensor, *, num_frames_lat: int,
                 lat_h: int, lat_w: int, frame_rate: float, num_steps: int,
                 neg_enc: Optional[Tensor] = None, neg_mask: Optional[Tensor] = None,
                 guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0,
                 skip_block_list: Optional[list] = None, strategy: Optional[str] = None,
                 conditioning_mask: Optional[Tensor] = None, model_dtype=torch.float32,
                 per_step: Optional[list] = None, timesteps: Optional[Tensor] = None,
                 guidance_timesteps: Optional[List[float]] = None, pixel_coords: Optional[Tensor] = None,
                 image_cond_noise_scale: float = 0.0, generator: Optional[torch.Generator] = None) -> Tensor:
    """LTXVideoPipeline.__call__ denoise loop (pipeline_ltx_video.py:919-1268) on patchified
    latents [b, N, C]; returns final patchified latents.  image_cond_noise_scale > 0 re-noises the hard-conditioned tokens at the
    start of every step from `generator` (add_noise_to_image_conditioning_latents, :606-629, called at :1105-1113).
    pixel_coords [b, 3, N]: the coordinates prepare_conditioning returned (:1076-1093) when extra keyframe tokens were prepended;
    default = the plain latent grid.
    guidance_scale / stg_scale / rescaling_scale / skip_block_list may be per-guidance-timestep lists (:959-1017)."""
    b, N, C = latents.shape
    D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    if timesteps is None:
        timesteps = rf_timesteps(num_steps, (b, C, num_frames_lat, lat_h, lat_w))
    timesteps = timesteps.to(latents.device)
    n = len(timesteps)
    mapping = None
    if guidance_timesteps:                                                            # :959-968
        mapping = []
        for t in timesteps:
            idx = [i for i, val in enumerate(guidance_timesteps) if val <= t]
            mapping.append(idx[0] if len(idx) > 0 else len(guidance_timesteps) - 1)
    per = lambda v: [v] * n if not isinstance(v, list) else [v[mapping[i]] for i in range(n)]
    gs = [x if x > 1.0 else 0.0 for x in per(guidance_scale)]                        # :973-990
    stg, resc = per(stg_scale), per(rescaling_scale)
    do_cfg, do_stg, do_resc = any(x > 1.0 for x in gs), any(x > 0.0 for x in stg), any(x != 1.0 for x in resc)
    num_conds = 1 + int(do_cfg) + int(do_stg)
    L = sum(1 for k in sd if k.endswith(".attn1.to_q.weight"))
    if skip_block_list is not None:                                                   # :1003-1013
        if len(skip_block_list) == 0 or not isinstance(skip_block_list[0], list):
            skip_block_list = [skip_block_list] * n
        else:
            skip_block_list = [skip_block_list[mapping[i]] for i in range(n)]
    skip_masks = None
    if do_stg and skip_block_list is not None:
        skip_masks = []
        for sb in skip_block_list:
            m = torch.ones(L, b * num_conds, dtype=model_dtype, device=latents.device)
            for bi in sb:
                m[bi, num_conds - 1::num_conds] = 0                                   # transformer3d.py:184-185
            skip_masks.append(m)
    enc_b, mask_b = enc, enc_mask
    if do_cfg:
        enc_b = torch.cat([neg_enc, enc]); mask_b = torch.cat([neg_mask, enc_mask])
    if do_stg:
        enc_b = torch.cat([enc_b, enc]); mask_b = torch.cat([mask_b, enc_mask])
    coords = pixel_coords if pixel_coords is not None else latent_to_pixel_coords(latent_coords(num_frames_lat, lat_h, lat_w, b, latents.device))
    frac = coords.to(torch.float32).clone()
    frac[:, 0] = frac[:, 0] * (1.0 / frame_rate)                                     # :1086-1087
    cos_sin = precompute_freqs_cis(frac, D, cfg["rope_theta"], cfg["rope_max_pos"], model_dtype)
    cmask = None if conditioning_mask is None else torch.cat([conditioning_mask] * num_conds)
    init_latents = latents.clone()                                                    # :1078
    for i, t in enumerate(timesteps):
        if conditioning_mask is not None and image_cond_noise_scale > 0.0:           # :606-629
            noise = torch.randn(latents.shape, generator=generator, dtype=latents.dtype, device=latents.device)
            hard = (conditioning_mask > 1.0 - 1e-6).unsqueeze(-1)
            latents = torch.where(hard, init_latents + image_cond_noise_scale * noise * (t ** 2), latents)
        x_in = torch.cat([latents] * num_conds) if num_conds > 1 else latents
        cur_t = t[None].expand(x_in.shape[0]).unsqueeze(-1)                          # [B,1]
        if cmask is not None:
            cur_t = torch.min(cur_t, 1.0 - cmask)                                    # :1145-1150 -> [B,N]
        noise_pred = transformer_forward(sd, cfg, x_in.to(model_dtype), cos_sin, enc_b.to(model_dtype),
                                         cur_t, mask_b, skip_masks[i] if skip_masks is not None else None, strategy,
                                         latent_shape=(num_frames_lat, lat_h, lat_w))
        noise_pred = guidance_combine(noise_pred, num_conds, do_cfg, do_stg, gs[i], stg[i], resc[i], do_resc)
        cur_t = cur_t[:1]
        denoised = rf_step(noise_pred, cur_t, latents, timesteps)                    # :1233, per-token branch
        if conditioning_mask is not None:
            keep = (t - 1e-6 < (1.0 - conditioning_mask)).unsqueeze(-1)              # :1341-1342
            denoised = torch.where(keep, denoised, latents)
        latents = denoised
        if per_step is not None:
            per_step.append(latents.clone())
    return latents


# --------------------------------------------------------------------------------------
# CausalVideoAutoencoder decode (vae_encode.py:94-165,239-247; causal_video_autoencoder.py)
# --------------------------------------------------------------------------------------
def causal_conv3d(sd, name, x: Tensor, causal: bool) -> Tensor:
    """CausalConv3d.forward (causal_conv3d.py:44-59): replicate-pad time, zero-pad space."""
    if causal:
        x = torch.cat([x[:, :, :1].repeat(1, 1, 2, 1, 1), x], dim=2)
    else:
        x = torch.cat([x[:, :, :1], x, x[:, :, -1:]], dim=2)
    return F.conv3d(x, sd[name + ".conv.weight"], sd[name + ".conv.bias"], padding=(0, 1, 1))


def pixel_norm(x: Tensor, eps: float = 1e-8) -> Tensor:
    """pixel_norm.py:12"""
    return x / torch.sqrt(torch.mean(x ** 2, dim=1, keepdim=True) + eps)


def _vae_time_embed(sd, name, t: Tensor) -> Tensor:
    """PixArtAlphaCombinedTimestepSizeEmbeddings(dim, 0) (diffusers): Timesteps(256) -> Linear -> SiLU -> Linear; t [B] -> [B, dim]"""
    h = F.silu(linear(sd, name + ".timestep_embedder.linear_1", timestep_sinusoid(t, 256)))
    return linear(sd, name + ".timestep_embedder.linear_2", h)


def _resnet(sd, p, x, cin, cout, causal, temb: Optional[Tensor] = None):
    """ResnetBlock3D.forward (causal_video_autoencoder.py:1197-1258), pixel_norm, no noise; `temb` [B, 4*cin] is the mid-block's
    timestep embedding when the decoder is timestep-conditioned (:1207-1237)."""
    h = pixel_norm(x)
    if temb is not None:
        ada = sd[p + "scale_shift_table"][None, :, :, None, None, None] + temb.reshape(x.shape[0], 4, -1, 1, 1, 1)
        shift1, scale1, shift2, scale2 = ada.unbind(dim=1)
        h = h * (1 + scale1) + shift1
    h = pixel_norm(causal_conv3d(sd, p + "conv1", F.silu(h), causal))
    if temb is not None:
        h = h * (1 + scale2) + shift2
    h = causal_conv3d(sd, p + "conv2", F.silu(h), causal)
    if cin != cout:
        xs = x.permute(0, 2, 3, 4, 1)
        xs = F.layer_norm(xs, (cin,), sd[p + "norm3.norm.weight"], sd[p + "norm3.norm.bias"], eps=1e-6)
        x = F.conv3d(xs.permute(0, 4, 1, 2, 3), sd[p + "conv_shortcut.weight"], sd[p + "conv_shortcut.bias"])
    return x + h


def depth_to_space(x: Tensor) -> Tensor:
    """'b (c p1 p2 p3) d h w -> b c (d p1) (h p2) (w p3)', p=2 (pixel_shuffle.py:12-20)"""
    b, c8, d, h, w = x.shape
    c = c8 // 8
    x = x.reshape(b, c, 2, 2, 2, d, h, w).permute(0, 1, 5, 2, 6, 3, 7, 4)
    return x.reshape(b, c, d * 2, h * 2, w * 2)


def vae_unpatchify(x: Tensor, p: int) -> Tensor:
    """'b (c p r q) f h w -> b c (f p) (h q) (w r)', p_t=1 (causal_video_autoencoder.py:1282-1299).
    NOTE channel order is (c, r, q): r indexes width, q indexes height."""
    b, cc, f, h, w = x.shape
    c = cc // (p * p)
    x = x.reshape(b, c, p, p, f, h, w)            # c, r, q
    x = x.permute(0, 1, 4, 5, 3, 6, 2)            # b c f h q w r
    return x.reshape(b, c, f, h * p, w * p)


def vae_decode(sd: Dict[str, Tensor], latents: Tensor, cfg: dict = LTX_VAE,
               per_channel_normalize: bool = True, collect: Optional[list] = None, timestep: Optional[Tensor] = None) -> Tensor:
    """vae_decode -> un_normalize_latents -> Decoder.forward.  latents [B,128,F,H,W] ->
    [B,3,8(F-1)+1,32H,32W].  timestep [B] for timestep-conditioned decoders (:757-795)."""
    causal = cfg["causal_decoder"]
    tc = cfg.get("timestep_conditioning", False)
    if tc:
        assert timestep is not None, "should pass timestep with timestep_conditioning=True"
        scaled_t = timestep.to(latents.dtype) * sd["decoder.timestep_scale_multiplier"]
    dt = latents.dtype
    if per_channel_normalize:
        z = latents * sd["std_of_means"].to(dt).view(1, -1, 1, 1, 1) + sd["mean_of_means"].to(dt).view(1, -1, 1, 1, 1)
    else:
        z = latents
    x = causal_conv3d(sd, "decoder.conv_in", z, causal)
    if collect is not None:
        collect.append(x)
    for kind, idx, cin, cout, n in vae_decoder_plan(cfg):
        p = f"decoder.up_blocks.{idx}."
        if kind == "res_x":
            temb = _vae_time_embed(sd, p + "time_embedder", scaled_t.flatten()) if tc else None      # UNetMidBlock3D :903-918
            for j in range(n):
                x = _resnet(sd, p + f"res_blocks.{j}.", x, cin, cin, causal, temb)
        elif kind == "res_x_y":
            x = _resnet(sd, p, x, cin, cout, causal)
        else:
            x = depth_to_space(causal_conv3d(sd, p + "conv", x, causal))[:, :, 1:]   # :1057-1062
        if collect is not None:
            collect.append(x)
    x = pixel_norm(x)
    if tc:                                                                            # :773-795
        e = _vae_time_embed(sd, "decoder.last_time_embedder", scaled_t.flatten())
        ada = sd["decoder.last_scale_shift_table"][None, :, :, None, None, None] + e.reshape(x.shape[0], 2, -1, 1, 1, 1)
        shift, scale = ada.unbind(dim=1)
        x = x * (1 + scale) + shift
    x = F.silu(x)
    x = causal_conv3d(sd, "decoder.conv_out", x, causal)
    return vae_unpatchify(x, cfg["patch_size"])


# --------------------------------------------------------------------------------------
# VAE encode (i2v / v2v conditioning): vae_encode.py:22-91 -> AutoencoderKLWrapper.encode (vae.py:265-306) ->
# Encoder.forward (causal_video_autoencoder.py:510-557), always causal, strided "compress_all" convolutions
# --------------------------------------------------------------------------------------
def vae_encoder_plan(cfg: dict = LTX_VAE) -> List[Tuple[str, int, int, int, int]]:
    """[(kind, index, c_in, c_out, n)] for encoder.down_blocks (causal_video_autoencoder.py:373-478)."""
    ch = cfg["base_channels"]
    plan = []
    for idx, (name, p) in enumerate(cfg["blocks"]):
        if name == "res_x":
            plan.append(("res_x", idx, ch, ch, int(p)))
        elif name == "res_x_y":
            plan.append(("res_x_y", idx, ch, ch * 2, 1))
            ch *= 2
        elif name == "compress_all":
            plan.append(("down", idx, ch, ch, 1))
        else:
            raise ValueError(name)
    return plan


def make_vae_encoder_state_dict(cfg: dict = LTX_VAE, seed: int = 2) -> Dict[str, Tensor]:
    """Random-init encoder weights (keys as in causal_video_autoencoder.py) + the latent statistics of the decoder dict."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    _conv3d(sd, gen, "encoder.conv_in.conv", cfg["base_channels"], cfg.get("in_channels", 3) * cfg["patch_size"] ** 2)
    ch = cfg["base_channels"]
    for kind, idx, cin, cout, n in vae_encoder_plan(cfg):
        p = f"encoder.down_blocks.{idx}."
        if kind == "res_x":
            for j in range(n):
                _conv3d(sd, gen, p + f"res_blocks.{j}.conv1.conv", cin, cin)
                _conv3d(sd, gen, p + f"res_blocks.{j}.conv2.conv", cin, cin)
        elif kind == "res_x_y":
            _conv3d(sd, gen, p + "conv1.conv", cout, cin)
            _conv3d(sd, gen, p + "conv2.conv", cout, cout)
            _conv3d(sd, gen, p + "conv_shortcut", cout, cin, k=1)
            sd[p + "norm3.norm.weight"] = 1.0 + 0.1 * torch.randn(cin, generator=gen)
            sd[p + "norm3.norm.bias"] = 0.1 * torch.randn(cin, generator=gen)
        else:
            _conv3d(sd, gen, p + "conv", cout, cin)
        ch = cout
    _conv3d(sd, gen, "encoder.conv_out.conv", cfg["latent_channels"] + 1, ch)      # latent_log_var "uniform": +1 channel
    sd["std_of_means"] = 0.5 + torch.rand(cfg["latent_channels"], generator=gen)
    sd["mean_of_means"] = 0.1 * torch.randn(cfg["latent_channels"], generator=gen)
    return sd


def vae_patchify(x: Tensor, p: int) -> Tensor:
    """'b c (f 1) (h q) (w r) -> b (c 1 r q) f h w' (causal_video_autoencoder.py:1261-1279): channel order (c, r, q)."""
    b, c, f, hh, ww = x.shape
    h, w = hh // p, ww // p
    x = x.reshape(b, c, f, h, p, w, p)              # b c f h q w r
    x = x.permute(0, 1, 6, 4, 2, 3, 5)              # b c r q f h w
    return x.reshape(b, c * p * p, f, h, w)


def strided_causal_conv3d(sd, name, x: Tensor, stride=(2, 2, 2)) -> Tensor:
    """make_conv_nd(..., stride, causal=True): replicate the first frame twice in front, zero-pad space by 1."""
    x = torch.cat([x[:, :, :1].repeat(1, 1, 2, 1, 1), x], dim=2)
    return F.conv3d(x, sd[name + ".conv.weight"], sd[name + ".conv.bias"], stride=stride, padding=(0, 1, 1))


def vae_encode_moments(sd: Dict[str, Tensor], video: Tensor, cfg: dict = LTX_VAE) -> Tuple[Tensor, Tensor]:
    """Encoder.forward + the 'uniform' log-variance expansion (:510-545) -> (mean [B,128,F',H',W'], logvar [B,1,F',H',W'])
    of the DiagonalGaussianDistribution.  video [B,3,F,H,W] in [-1, 1], F = 8k+1."""
    x = vae_patchify(video, cfg["patch_size"])
    x = causal_conv3d(sd, "encoder.conv_in", x, True)
    for kind, idx, cin, cout, n in vae_encoder_plan(cfg):
        p = f"encoder.down_blocks.{idx}."
        if kind == "res_x":
            for j in range(n):
                x = _resnet(sd, p + f"res_blocks.{j}.", x, cin, cin, True)
        elif kind == "res_x_y":
            x = _resnet(sd, p, x, cin, cout, True)
        else:
            x = strided_causal_conv3d(sd, p[:-1], x)
    x = causal_conv3d(sd, "encoder.conv_out", F.silu(pixel_norm(x)), True)
    return x[:, :-1], x[:, -1:]


def vae_encode(sd: Dict[str, Tensor], video: Tensor, cfg: dict = LTX_VAE, noise: Optional[Tensor] = None,
               per_channel_normalize: bool = True) -> Tensor:
    """vae_encode (vae_encode.py:22-91): latent_dist.sample() = mean + exp(0.5*clamp(logvar,-30,20))*noise (diffusers
    DiagonalGaussianDistribution; noise=None -> the mode), then normalize_latents (:228-237)."""
    mean, logvar = vae_encode_moments(sd, video, cfg)
    z = mean if noise is None else mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise
    if per_channel_normalize:
        z = (z - sd["mean_of_means"].to(z.dtype).view(1, -1, 1, 1, 1)) / sd["std_of_means"].to(z.dtype).view(1, -1, 1, 1, 1)
    return z * cfg.get("scaling_factor", 1.0)


# --------------------------------------------------------------------------------------
# Multi-scale flow: LatentUpsampler (latent_upsampler.py:15-149, dims=3, spatial x2), adain_filter_latent
# (pipeline_ltx_video.py:1709-1737), LTXMultiScalePipeline.__call__ (:1782-1903)
# --------------------------------------------------------------------------------------
def make_latent_upsampler_state_dict(in_channels: int = 128, mid_channels: int = 512, num_blocks_per_stage: int = 4,
                                     seed: int = 3) -> Dict[str, Tensor]:
    """LatentUpsampler(dims=3, spatial_upsample=True, temporal_upsample=False) parameters (latent_upsampler.py:56-107);
    nn.Conv default init bounds, GroupNorm affine perturbed away from (1, 0) so the test sees it."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(name, o, i, nd=3):
        b = 1.0 / math.sqrt(i * 3 ** nd)
        sd[name + ".weight"] = _uniform(gen, (o, i) + (3,) * nd, b)
        sd[name + ".bias"] = _uniform(gen, (o,), b)

    def gn(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=gen)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=gen)

    conv("initial_conv", mid_channels, in_channels); gn("initial_norm", mid_channels)
    for stage in ("res_blocks", "post_upsample_res_blocks"):
        for j in range(num_blocks_per_stage):
            conv(f"{stage}.{j}.conv1", mid_channels, mid_channels); gn(f"{stage}.{j}.norm1", mid_channels)
            conv(f"{stage}.{j}.conv2", mid_channels, mid_channels); gn(f"{stage}.{j}.norm2", mid_channels)
    conv("upsampler.0", 4 * mid_channels, mid_channels, nd=2)
    conv("final_conv", in_channels, mid_channels)
    return sd


def _up_resblock(sd, p, x):
    """ResBlock.forward (latent_upsampler.py:30-39)"""
    h = F.silu(F.group_norm(F.conv3d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1), 32,
                            sd[p + "norm1.weight"], sd[p + "norm1.bias"]))
    h = F.group_norm(F.conv3d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1), 32, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    return F.silu(h + x)


def latent_upsampler_forward(sd: Dict[str, Tensor], latent: Tensor) -> Tensor:
    """LatentUpsampler.forward, dims=3 spatial branch (latent_upsampler.py:109-149): [B,C,F,H,W] -> [B,C,F,2H,2W]."""
    b, c, f, h, w = latent.shape
    nb = sum(1 for k in sd if k.startswith("res_blocks.") and k.endswith(".conv1.weight"))
    x = F.conv3d(latent, sd["initial_conv.weight"], sd["initial_conv.bias"], padding=1)
    x = F.silu(F.group_norm(x, 32, sd["initial_norm.weight"], sd["initial_norm.bias"]))
    for j in range(nb):
        x = _up_resblock(sd, f"res_blocks.{j}.", x)
    m = x.shape[1]
    x2 = x.permute(0, 2, 1, 3, 4).reshape(b * f, m, h, w)                              # "b c f h w -> (b f) c h w"
    x2 = F.pixel_shuffle(F.conv2d(x2, sd["upsampler.0.weight"], sd["upsampler.0.bias"], padding=1), 2)   # PixelShuffleND(2)
    x = x2.reshape(b, f, m, 2 * h, 2 * w).permute(0, 2, 1, 3, 4)
    for j in range(nb):
        x = _up_resblock(sd, f"post_upsample_res_blocks.{j}.", x)
    return F.conv3d(x, sd["final_conv.weight"], sd["final_conv.bias"], padding=1)


def adain_filter_latent(latents: Tensor, reference_latents: Tensor, factor: float = 1.0) -> Tensor:
    """pipeline_ltx_video.py:1709-1737: per (batch, channel) match mean / unbiased std to the reference, then lerp."""
    dims = (2, 3, 4)
    r_sd, r_mean = torch.std_mean(reference_latents, dim=dims, keepdim=True)
    i_sd, i_mean = torch.std_mean(latents, dim=dims, keepdim=True)
    return torch.lerp(latents, (latents - i_mean) / i_sd * r_sd + r_mean, factor)


def upsample_latents(up_sd: Dict[str, Tensor], vae_sd: Dict[str, Tensor], latents: Tensor) -> Tensor:
    """LTXMultiScalePipeline._upsample_latents (:1761-1772): un-normalise, LatentUpsampler, normalise."""
    dt = latents.dtype
    std, mean = vae_sd["std_of_means"].to(dt).view(1, -1, 1, 1, 1), vae_sd["mean_of_means"].to(dt).view(1, -1, 1, 1, 1)
    return (latent_upsampler_forward(up_sd, latents * std + mean) - mean) / std


def multiscale_second_pass_init(noise: Tensor, upsampled: Tensor, t0: float) -> Tensor:
    """prepare_latents with input latents (pipeline_ltx_video.py:700-706): t0 * noise + (1 - t0) * latents."""
    return t0 * noise + (1 - t0) * upsampled


def multiscale_resize(videos: Tensor, height: int, width: int) -> Tensor:
    """LTXMultiScalePipeline.__call__ tail (:1890-1901): per-frame bilinear resize (align_corners=False) to the requested size."""
    b, c, f, h, w = videos.shape
    v = videos.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    v = F.interpolate(v, size=(height, width), mode="bilinear", align_corners=False)
    return v.reshape(b, f, c, height, width).permute(0, 2, 1, 3, 4)


def postprocess(image: Tensor) -> Tensor:
    """VaeImageProcessor.postprocess for tensors: x/2+0.5 clamped to [0,1]."""
    return (image / 2 + 0.5).clamp(0, 1)


def rel_l2(a: Tensor, b: Tensor) -> float:
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr(a: Tensor, b: Tensor, peak: float = 1.0) -> float:
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
