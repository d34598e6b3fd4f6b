"""CPU oracle for the Wan2.1 denoising hot path (TEST INFRASTRUCTURE ONLY — see ltx_oracle.py header).

Plain-PyTorch restatement of WanModel.forward (wan/modules/model.py:902-1111), WanAttentionBlock
(:397-499), WanSelfAttention / WanT2VCrossAttention (:175-274), WanRMSNorm / WanLayerNorm (:91-145),
Head (:539-573), the 3-axis RoPE (wan/modules/posemb_layers.py:222-293,299-473), the flow-matching UniPC
scheduler (wan/utils/fm_solvers_unipc.py) and the T2V denoise loop with CFG (wan/text2video.py:468-575).
Pinned by oracle/gen_golden_wan.py against the unmodified reference modules (fp32, CPU).
The Ulysses sequence-parallel forward (wan/distributed/xdit_context_parallel.py:66-192) depends on
xfuser (absent, unpinned): `ulysses_forward` below restates its published DeepSpeed-Ulysses semantics on
"virtual ranks" in one process and is checked against the single-rank forward (mathematically exact).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .ltx_oracle import attention_core, rel_l2  # noqa: F401

Tensor = torch.Tensor

# wan/configs/wan_t2v_1_3B.py:19-29, wan_t2v_14B.py:19-29, shared_config.py:9-19
WAN_1_3B = dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30, in_dim=16, out_dim=16, text_dim=4096,
                freq_dim=256, text_len=512, eps=1e-6, patch_size=(1, 2, 2))
WAN_14B = dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40, in_dim=16, out_dim=16, text_dim=4096,
               freq_dim=256, text_len=512, eps=1e-6, patch_size=(1, 2, 2))


def make_wan_state_dict(cfg: dict, seed: int = 0, num_layers: Optional[int] = None) -> Dict[str, Tensor]:
    """Seeded random init with the reference's key names (model.py:656-813).  The reference zero-inits
    head.head.weight (:1160), which would make the output identically 0: overridden."""
    g = torch.Generator().manual_seed(seed)
    D, Fd, L = cfg["dim"], cfg["ffn_dim"], cfg["num_layers"] if num_layers is None else num_layers
    sd: Dict[str, Tensor] = {}

    def lin(name, o, i, std=None):
        b = 1.0 / math.sqrt(i)
        sd[name + ".weight"] = (torch.randn(o, i, generator=g) * std) if std else ((torch.rand(o, i, generator=g) * 2 - 1) * b)
        sd[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * b

    pk = cfg["in_dim"] * math.prod(cfg["patch_size"])
    sd["patch_embedding.weight"] = ((torch.rand(D, pk, generator=g) * 2 - 1) / math.sqrt(pk)).view(D, cfg["in_dim"], *cfg["patch_size"])
    sd["patch_embedding.bias"] = (torch.rand(D, generator=g) * 2 - 1) / math.sqrt(pk)
    lin("text_embedding.0", D, cfg["text_dim"]); lin("text_embedding.2", D, D)
    lin("time_embedding.0", D, cfg["freq_dim"]); lin("time_embedding.2", D, D)
    lin("time_projection.1", 6 * D, D)
    for i in range(L):
        p = f"blocks.{i}."
        for a in ("self_attn", "cross_attn"):
            for n in ("q", "k", "v", "o"):
                lin(p + a + "." + n, D, D)
            sd[p + a + ".norm_q.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
            sd[p + a + ".norm_k.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
        sd[p + "norm3.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
        sd[p + "norm3.bias"] = 0.1 * torch.randn(D, generator=g)
        lin(p + "ffn.0", Fd, D); lin(p + "ffn.2", D, Fd)
        sd[p + "modulation"] = torch.randn(1, 6, D, generator=g) / D ** 0.5
    lin("head.head", cfg["out_dim"] * math.prod(cfg["patch_size"]), D)
    sd["head.modulation"] = torch.randn(1, 2, D, generator=g) / D ** 0.5
    if cfg.get("model_type", "t2v") == "i2v":
        # WanI2VCrossAttention k_img / v_img / norm_k_img (model.py:288-291) and MLPProj img_emb (:576-588, :768-769)
        for i in range(L):
            p = f"blocks.{i}.cross_attn."
            lin(p + "k_img", D, D); lin(p + "v_img", D, D)
            sd[p + "norm_k_img.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
        C = cfg.get("clip_dim", 1280)
        sd["img_emb.proj.0.weight"] = 1.0 + 0.1 * torch.randn(C, generator=g); sd["img_emb.proj.0.bias"] = 0.1 * torch.randn(C, generator=g)
        lin("img_emb.proj.1", C, C); lin("img_emb.proj.3", D, C)
        sd["img_emb.proj.4.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g); sd["img_emb.proj.4.bias"] = 0.1 * torch.randn(D, generator=g)
    return sd


# ----------------------------------------------------------------------------------------------------
# RoPE (posemb_layers.py): integer grid (f, h/2, w/2), dims [44, 42, 42], theta 10000, cos/sin repeat-interleaved
# ----------------------------------------------------------------------------------------------------
def rope_tables(latent_fhw: Sequence[int], rope_dims=(44, 42, 42), theta: float = 10000.0) -> Tuple[Tensor, Tensor]:
    """get_rotary_pos_embed(latents.shape[1:]) with enable_RIFLEx=False -> (cos, sin) fp32 [N, 128]
    (posemb_layers.py:432-473, 299-378, 381-430)."""
    sizes = [latent_fhw[0], latent_fhw[1] // 2, latent_fhw[2] // 2]
    grids = [torch.linspace(0, n, n + 1, dtype=torch.float32)[:n] for n in sizes]           # :138-141
    grid = torch.stack(torch.meshgrid(*grids, indexing="ij"), dim=0)
    cos, sin = [], []
    for i, d in enumerate(rope_dims):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2)[: d // 2].float() / d))                 # :419-421
        fr = torch.outer(grid[i].reshape(-1), freqs)
        cos.append(fr.cos().repeat_interleave(2, dim=1))
        sin.append(fr.sin().repeat_interleave(2, dim=1))
    return torch.cat(cos, dim=1), torch.cat(sin, dim=1)


def apply_rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """x [B, S, H, d]; fp32 math, cast back (posemb_layers.py:259-276); rotate_half on interleaved pairs (:222-226)."""
    dt = x.dtype
    xf = x.float()
    x2 = xf.reshape(*xf.shape[:-1], -1, 2)
    rot = torch.stack([-x2[..., 1], x2[..., 0]], dim=-1).flatten(3)
    return (xf * cos[None, :, None, :] + rot * sin[None, :, None, :]).to(dt)


def wan_rms_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """WanRMSNorm.forward (model.py:99-111): in-place x *= rsqrt(mean(x^2)+eps) (fp32 stat), x *= weight —
    two roundings in the activation dtype."""
    y = x.float().pow(2).mean(dim=-1, keepdim=True).add(eps).rsqrt()
    x = (x * y).to(x.dtype)
    return x * w


def sinusoidal_embedding_1d(dim: int, position: Tensor) -> Tensor:
    """model.py:18-28"""
    half = dim // 2
    position = position.type(torch.float32)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def wan_block(sd, i, x, e0, cos, sin, ctx, cfg, attn_fn=None):
    """WanAttentionBlock.forward (model.py:397-499), t2v cross-attention, no VACE/cam/audio.  x [B, N, D];
    e0 [1, 6, D]; ctx [B, 512, D].  attn_fn(q,k,v) may replace self-attention (sequence-parallel emulation)."""
    p = f"blocks.{i}."
    H, eps, D = cfg["num_heads"], cfg["eps"], cfg["dim"]
    B, N, _ = x.shape
    d = D // H
    e = (sd[p + "modulation"] + e0).chunk(6, dim=1)                                     # :436
    xm = F.layer_norm(x, (D,), eps=eps) * (1 + e[1]) + e[0]                            # :438-441
    q = wan_rms_norm(_lin(sd, p + "self_attn.q", xm), sd[p + "self_attn.norm_q.weight"], eps).view(B, N, H, d)
    k = wan_rms_norm(_lin(sd, p + "self_attn.k", xm), sd[p + "self_attn.norm_k.weight"], eps).view(B, N, H, d)
    v = _lin(sd, p + "self_attn.v", xm).view(B, N, H, d)
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    a = (attn_fn or attention_core)(q, k, v)
    y = _lin(sd, p + "self_attn.o", a.flatten(2))
    x = x + y * e[2]                                                                    # :458 addcmul_
    y = F.layer_norm(x, (D,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], eps=eps)    # :461
    q = wan_rms_norm(_lin(sd, p + "cross_attn.q", y), sd[p + "cross_attn.norm_q.weight"], eps).view(B, N, H, d)
    ctx_img = None
    if (p + "cross_attn.k_img.weight") in sd:                                           # WanI2VCrossAttention :292-344
        ctx_img, ctx = ctx[:, :257], ctx[:, 257:]
    k = wan_rms_norm(_lin(sd, p + "cross_attn.k", ctx), sd[p + "cross_attn.norm_k.weight"], eps).view(B, -1, H, d)
    v = _lin(sd, p + "cross_attn.v", ctx).view(B, -1, H, d)
    a2 = attention_core(q, k, v).flatten(2)
    if ctx_img is not None:
        ki = wan_rms_norm(_lin(sd, p + "cross_attn.k_img", ctx_img), sd[p + "cross_attn.norm_k_img.weight"], eps).view(B, -1, H, d)
        vi = _lin(sd, p + "cross_attn.v_img", ctx_img).view(B, -1, H, d)
        a2 = a2 + attention_core(q, ki, vi).flatten(2)                                   # x += img_x :337
    x = x + _lin(sd, p + "cross_attn.o", a2)                                            # :465
    y = F.layer_norm(x, (D,), eps=eps) * (1 + e[4]) + e[3]                              # :467-472
    y = _lin(sd, p + "ffn.2", F.gelu(_lin(sd, p + "ffn.0", y), approximate="tanh"))     # :479-488
    return x + y * e[5]                                                                 # :491


def patchify(x: Tensor, cfg) -> Tensor:
    """Conv3d(k=s=(1,2,2)) input as rows: [C, F, H, W] -> [N, C*1*2*2] with k = (c, pt, ph, pw) (model.py:951-954)."""
    C, Fr, H, W = x.shape
    return x.view(C, Fr, 1, H // 2, 2, W // 2, 2).permute(1, 3, 5, 0, 2, 4, 6).reshape(Fr * (H // 2) * (W // 2), C * 4)


def unpatchify(u: Tensor, grid: Sequence[int], cfg) -> Tensor:
    """model.py:1113-1136: [N, prod(patch)*c] -> [c, F, H, W] via 'fhwpqrc->cfphqwr'."""
    c = cfg["out_dim"]
    u = u[: math.prod(grid)].view(*grid, *cfg["patch_size"], c)
    u = torch.einsum("fhwpqrc->cfphqwr", u)
    return u.reshape(c, *[i * j for i, j in zip(grid, cfg["patch_size"])])


def teacache_state(coefficients, rel_l1_thresh: float, start_step: int, num_steps: int, n_seq: int = 2) -> dict:
    """The attributes text2video.py:461-464 / the UI set on the model before a TeaCache run (model.py:1029-1049)."""
    return dict(coefficients=list(coefficients), rel_l1_thresh=rel_l1_thresh, start_step=start_step, num_steps=num_steps,
                accumulated=0.0, prev_e=None, previous_residual=[None] * n_seq, skipped=0)


def wan_forward(sd: Dict[str, Tensor], cfg: dict, x_list: List[Tensor], t: Tensor, context: List[Tensor],
                cos: Tensor, sin: Tensor, attn_fn=None, clip_fea: Optional[Tensor] = None, y: Optional[Tensor] = None,
                slg_layers: Optional[Sequence[int]] = None, teacache: Optional[dict] = None, current_step: int = 0) -> List[Tensor]:
    """WanModel.forward for t2v (model.py:902-1111): x_list of [16, F, H, W]; t [1]; context list of [L<=512, 4096];
    returns list of float32 [16, F, H, W].  Sequences are batched (the reference iterates them per block)."""
    D = cfg["dim"]
    dt = sd["time_projection.1.weight"].dtype
    L = sum(1 for k in sd if k.endswith(".self_attn.q.weight"))
    w = sd["patch_embedding.weight"].flatten(1)
    xs = []
    for x in x_list:
        Fr, H, W = x.shape[1:]
        grid = (Fr, H // 2, W // 2)
        if y is not None:
            x = torch.cat([x, y.to(x.dtype)], dim=0)                                      # i2v: [mask(4) | image latent(16)] channels, model.py:948-949
        xs.append(F.linear(patchify(x.to(dt), cfg), w, sd["patch_embedding.bias"]))
    x = torch.stack(xs, 0)                                                               # [B, N, D]
    e = _lin(sd, "time_embedding.2", F.silu(_lin(sd, "time_embedding.0", sinusoidal_embedding_1d(cfg["freq_dim"], t.flatten()).to(dt))))
    e0 = _lin(sd, "time_projection.1", F.silu(e)).unflatten(1, (6, D))                    # [1, 6, D]
    ctx = torch.stack([_lin(sd, "text_embedding.2", F.gelu(_lin(sd, "text_embedding.0", torch.cat(
        [u.to(dt), u.new_zeros(cfg["text_len"] - u.size(0), u.size(1)).to(dt)])), approximate="tanh")) for u in context], 0)
    if clip_fea is not None:                                                             # img_emb MLPProj, model.py:996-998
        c = clip_fea.to(dt)
        C = c.shape[-1]
        c = F.layer_norm(c, (C,), sd["img_emb.proj.0.weight"], sd["img_emb.proj.0.bias"])
        c = _lin(sd, "img_emb.proj.3", F.gelu(_lin(sd, "img_emb.proj.1", c)))
        c = F.layer_norm(c, (D,), sd["img_emb.proj.4.weight"], sd["img_emb.proj.4.bias"])
        ctx = torch.cat([c.expand(ctx.shape[0], -1, -1), ctx], dim=1)
    should_calc = True
    if teacache is not None:                                                             # model.py:1029-1049 (joint pass, x_id = 0)
        tc = teacache
        if current_step <= tc["start_step"] or current_step == tc["num_steps"] - 1:
            tc["accumulated"] = 0.0
        else:
            rel = float(((e - tc["prev_e"]).abs().mean() / tc["prev_e"].abs().mean()).item())
            tc["accumulated"] += abs(float(np.poly1d(tc["coefficients"])(rel)))
            if tc["accumulated"] < tc["rel_l1_thresh"]:
                should_calc = False
                tc["skipped"] += 1
            else:
                tc["accumulated"] = 0.0
        tc["prev_e"] = e
    if not should_calc:
        x = x + torch.stack(teacache["previous_residual"], 0)                            # :1051-1054
    else:
        x_in = x
        for i in range(L):
            if slg_layers is not None and i in slg_layers and x.shape[0] > 1:            # :1077-1080 joint pass: cond sequence only
                x = torch.cat([wan_block(sd, i, x[:1], e0, cos, sin, ctx[:1], cfg, attn_fn), x[1:]], dim=0)
            else:
                x = wan_block(sd, i, x, e0, cos, sin, ctx, cfg, attn_fn)
        if teacache is not None:
            teacache["previous_residual"] = list((x - x_in).unbind(0))                   # :1087-1101
    eh = (sd["head.modulation"] + e.unsqueeze(1)).chunk(2, dim=1)                         # model.py:566-572
    x = F.layer_norm(x, (D,), eps=cfg["eps"]) * (1 + eh[1]) + eh[0]
    x = _lin(sd, "head.head", x)
    return [unpatchify(u, grid, cfg).float() for u in x]


# ----------------------------------------------------------------------------------------------------
# FlowUniPCMultistepScheduler (fm_solvers_unipc.py), order 2, bh2, predict_x0, flow_prediction
# ----------------------------------------------------------------------------------------------------
class UniPC:
    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2):
        self.T = num_train_timesteps
        self.order = solver_order
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()      # :109-111
        sig = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        self.sigma_min, self.sigma_max = sig[-1].item(), sig[0].item()                           # shift=1: unchanged

    def set_timesteps(self, steps: int, shift: float):
        """:160-227"""
        sigmas = np.linspace(self.sigma_max, self.sigma_min, steps + 1).copy()[:-1]
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.T
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(timesteps).to(dtype=torch.int64)
        self.model_outputs = [None] * self.order
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = 0
        self.this_order = 1

    def _coeffs(self, s_t, s_s0):
        a_t, a_s0 = 1 - s_t, 1 - s_s0
        lam_t = torch.log(a_t) - torch.log(s_t)
        lam_s0 = torch.log(a_s0) - torch.log(s_s0)
        h = lam_t - lam_s0
        hh = -h
        h_phi_1 = torch.expm1(hh)
        return a_t, lam_s0, h, hh, h_phi_1, torch.expm1(hh)

    def _rb(self, rks, hh, h_phi_1, B_h, order):
        R, b = [], []
        h_phi_k = h_phi_1 / hh - 1
        fact = 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return torch.stack(R), torch.tensor(b)

    def step(self, model_output: Tensor, sample: Tensor) -> Tensor:
        """:655-739 (convert_model_output :321, UniC :486-626, UniP :350-484)"""
        i = self.step_index
        m_t = sample - self.sigmas[i] * model_output
        if i > 0 and self.last_sample is not None:
            order = self.this_order
            s_t, s_s0 = self.sigmas[i], self.sigmas[i - 1]
            a_t, lam_s0, h, hh, h_phi_1, B_h = self._coeffs(s_t, s_s0)
            m0 = self.model_outputs[-1]
            rks, D1s = [], []
            for k in range(1, order):
                si = i - (k + 1)
                lam_si = torch.log(1 - self.sigmas[si]) - torch.log(self.sigmas[si])
                rk = (lam_si - lam_s0) / h
                rks.append(rk)
                D1s.append((self.model_outputs[-(k + 1)] - m0) / rk)
            rks.append(1.0)
            rks = torch.tensor(rks)
            R, b = self._rb(rks, hh, h_phi_1, B_h, order)
            rhos_c = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
            x_t_ = s_t / s_s0 * self.last_sample - a_t * h_phi_1 * m0
            corr = sum(rhos_c[k] * D1s[k] for k in range(len(D1s))) if D1s else 0
            sample = x_t_ - a_t * B_h * (corr + rhos_c[-1] * (m_t - m0))
        for k in range(self.order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = m_t
        this_order = min(self.order, len(self.timesteps) - i)
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        order = self.this_order
        s_t, s_s0 = self.sigmas[i + 1], self.sigmas[i]
        a_t, lam_s0, h, hh, h_phi_1, B_h = self._coeffs(s_t, s_s0)
        m0 = self.model_outputs[-1]
        x_t = s_t / s_s0 * sample - a_t * h_phi_1 * m0
        if order == 2:
            si = i - 1
            lam_si = torch.log(1 - self.sigmas[si]) - torch.log(self.sigmas[si])
            rk = (lam_si - lam_s0) / h
            x_t = x_t - a_t * B_h * (0.5 * ((self.model_outputs[-2] - m0) / rk))
        if self.lower_order_nums < self.order:
            self.lower_order_nums += 1
        self.step_index += 1
        return x_t


def get_sampling_sigmas(sampling_steps: int, shift: float) -> np.ndarray:
    """wan/utils/fm_solvers.py:22-26"""
    sigma = np.linspace(1, 0, sampling_steps + 1)[:sampling_steps]
    return shift * sigma / (1 + (shift - 1) * sigma)


class DPMpp:
    """FlowDPMSolverMultistepScheduler(solver_order=2, algorithm dpmsolver++, midpoint, flow_prediction, final sigma zero,
    lower_order_final) as `sample_solver='dpm++'` builds it (wan/utils/fm_solvers.py:69-800, text2video.py:423-432)."""

    def __init__(self, num_train_timesteps: int = 1000):
        self.T = num_train_timesteps

    def set_timesteps(self, steps: int, shift: float):
        """text2video.py:428-432 -> set_timesteps(sigmas=get_sampling_sigmas(steps, shift)) with config.shift = 1 (:226-285)"""
        sigmas = get_sampling_sigmas(steps, shift)
        sigmas = 1.0 * sigmas / (1 + (1.0 - 1) * sigmas)
        self.timesteps = torch.from_numpy(sigmas * self.T).to(dtype=torch.int64)
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0]]).astype(np.float32))
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self.step_index = 0

    def step(self, model_output: Tensor, sample: Tensor) -> Tensor:
        """:706-798 (convert_model_output :378-380, first order :415-484, second order midpoint :486-594)"""
        i, n = self.step_index, len(self.timesteps)
        lower_order_final = i == n - 1                                          # final_sigmas_type == "zero"
        lower_order_second = i == n - 2 and n < 15
        x0 = sample - self.sigmas[i] * model_output
        self.model_outputs = [self.model_outputs[1], x0]
        s_t, s_s0 = self.sigmas[i + 1], self.sigmas[i]
        a_t, a_s0 = 1 - s_t, 1 - s_s0
        lam_t, lam_s0 = torch.log(a_t) - torch.log(s_t), torch.log(a_s0) - torch.log(s_s0)
        h = lam_t - lam_s0
        if self.lower_order_nums < 1 or lower_order_final:
            x_t = (s_t / s_s0) * sample - (a_t * (torch.exp(-h) - 1.0)) * x0
        else:                                                                   # solver_order == 2 (lower_order_second is moot)
            s_s1 = self.sigmas[i - 1]
            lam_s1 = torch.log(1 - s_s1) - torch.log(s_s1)
            m0, m1 = self.model_outputs[-1], self.model_outputs[-2]
            r0 = (lam_s0 - lam_s1) / h
            D0, D1 = m0, (1.0 / r0) * (m0 - m1)
            x_t = (s_t / s_s0) * sample - (a_t * (torch.exp(-h) - 1.0)) * D0 - 0.5 * (a_t * (torch.exp(-h) - 1.0)) * D1
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        self.step_index += 1
        return x_t


def t2v_denoise(sd, cfg, noise: Tensor, context: Tensor, context_null: Tensor, steps: int, shift: float = 5.0,
                guide_scale: float = 5.0, per_step: Optional[list] = None, attn_fn=None,
                cfg_star_switch: bool = False, cfg_zero_step: int = 5, clip_fea: Optional[Tensor] = None,
                y: Optional[Tensor] = None, sample_solver: str = "unipc", slg_layers: Optional[Sequence[int]] = None,
                slg_start: float = 0.0, slg_end: float = 1.0) -> Tensor:
    """WanT2V.generate denoise loop, UniPC, plain CFG (text2video.py:399-575).  noise [16, F, H, W] fp32.
    With clip_fea / y it is the WanI2V.generate loop (image2video.py:328-414): same y and CLIP tokens for both passes.
    slg_layers: skip-layer guidance in the joint pass on the steps int(slg_start*steps) <= i < int(slg_end*steps) (text2video.py:492)."""
    sch = UniPC() if sample_solver == "unipc" else DPMpp()
    sch.set_timesteps(steps, shift)
    cos, sin = rope_tables(noise.shape[1:])
    lat = noise
    for i, t in enumerate(sch.timesteps):
        ts = torch.stack([t])
        if guide_scale == 1:
            pred = wan_forward(sd, cfg, [lat], ts, [context], cos, sin, attn_fn, clip_fea=clip_fea, y=y)[0]
        else:
            slg = slg_layers if int(slg_start * steps) <= i < int(slg_end * steps) else None
            c, u = wan_forward(sd, cfg, [lat, lat], ts, [context, context_null], cos, sin, attn_fn, clip_fea=clip_fea, y=y, slg_layers=slg)
            if cfg_star_switch and i > cfg_zero_step:                                     # :551-561 (optimized_scale :31-42)
                alpha = torch.sum(c.flatten() * u.flatten()) / (torch.sum(u.flatten() ** 2) + 1e-8)
                u = u * alpha
            pred = u + guide_scale * (c - u)                                              # :562
        lat = sch.step(pred.unsqueeze(0), lat.unsqueeze(0)).squeeze(0)
        if per_step is not None:
            per_step.append(lat.clone())
    return lat


# ----------------------------------------------------------------------------------------------------
# Ulysses sequence parallelism on virtual ranks (xdit_context_parallel.py:66-192 semantics)
# ----------------------------------------------------------------------------------------------------
def ulysses_attention_virtual(q: Tensor, k: Tensor, v: Tensor, P: int) -> Tensor:
    """q,k,v [B, N, H, d] (already RoPE'd).  Emulates: each rank holds tokens [r*N/P, (r+1)*N/P) of all H heads;
    all-to-all -> each rank holds all N tokens of heads [r*H/P, (r+1)*H/P); local attention; all-to-all back."""
    B, N, H, d = q.shape
    assert N % P == 0 and H % P == 0
    out = torch.empty_like(q)
    shards = lambda t: [t[:, r * (N // P):(r + 1) * (N // P)] for r in range(P)]
    qs, ks, vs = shards(q), shards(k), shards(v)
    for r in range(P):                                   # what rank r computes after the first all-to-all
        hs = slice(r * (H // P), (r + 1) * (H // P))
        qr = torch.cat([s[:, :, hs] for s in qs], dim=1)
        kr = torch.cat([s[:, :, hs] for s in ks], dim=1)
        vr = torch.cat([s[:, :, hs] for s in vs], dim=1)
        o = attention_core(qr, kr, vr)                   # [B, N, H/P, d]
        for dst in range(P):                             # reverse all-to-all: token shard dst gets heads hs from rank r
            out[:, dst * (N // P):(dst + 1) * (N // P), hs] = o[:, dst * (N // P):(dst + 1) * (N // P)]
    return out


def i2v_conditioning(vae_sd, vae_cfg, image: Tensor, frame_num: int, image_end: Optional[Tensor] = None,
                     add_frames_for_end_image: bool = True) -> Tensor:
    """image2video.py:232-244, 262-277: y = [conditioning-frame mask (4) | WanVAE.encode([image, zeros ..., (end image)]) (16)]
    -> [20, latent frames, H/8, W/8].  image(s): [3, H, W] in [-1, 1]; frame_num = the count after `frame_num += 1` for an added end frame."""
    from . import wan_vae_oracle as V
    h, w = image.shape[1:]
    lat_h, lat_w = h // 8, w // 8
    msk = torch.ones(1, frame_num, lat_h, lat_w)
    if image_end is not None:
        msk[:, 1:-1] = 0
        if add_frames_for_end_image:
            msk = torch.concat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:-1], torch.repeat_interleave(msk[:, -1:], repeats=4, dim=1)], dim=1)
        else:
            msk = torch.concat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:]], dim=1)
        enc = torch.concat([image[:, None], torch.zeros(3, frame_num - 2, h, w), image_end[:, None]], dim=1)
    else:
        msk[:, 1:] = 0
        msk = torch.concat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:]], dim=1)
        enc = torch.concat([image[:, None], torch.zeros(3, frame_num - 1, h, w)], dim=1)
    msk = msk.view(1, msk.shape[1] // 4, 4, lat_h, lat_w).transpose(1, 2)[0]
    lat_y = V.wan_vae_encode(vae_sd, enc, vae_cfg, torch.tensor(V.WAN_VAE_MEAN), torch.tensor(V.WAN_VAE_STD),
                             any_end_frame=image_end is not None and add_frames_for_end_image)
    return torch.concat([msk, lat_y])
