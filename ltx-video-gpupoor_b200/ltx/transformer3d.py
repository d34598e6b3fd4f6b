"""Transformer3DModel — B200-native drop-in for ltx_video/models/transformers/transformer3d.py:46-507
(+ BasicTransformerBlock / AttnProcessor2_0 / FeedForward of .../attention.py:45-364,986-1173,1263-1323).

Same constructor config keys, same state_dict layout (reference key names), same forward signature and
companion API (`precompute_freqs_cis`, `create_skip_layer_mask`, `.config`, `.dtype`, `.in_channels`).
All arithmetic runs in hand-written sm_100a kernels through the C ABI (`ops`): every Linear is the
tcgen05/TMEM GEMM with fused bias / GELU / gate / residual epilogues, attention is the TMEM flash
kernel, norms / modulation / QK-norm+RoPE are single-pass memory-bound kernels.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Any, Dict, List, Optional

import torch

from .. import ops
from ..module_like import ModuleLike
from .skip_layer_strategy import SkipLayerStrategy

BF16 = torch.bfloat16


@dataclass
class Transformer3DModelOutput:
    sample: torch.Tensor


# ltx_video/utils/diffusers_config_mapping.py:74-105 (OURS_TRANSFORMER_CONFIG), the named 2B architecture
LTX_2B_CONFIG = dict(
    num_attention_heads=32, attention_head_dim=64, in_channels=128, out_channels=128, num_layers=28,
    cross_attention_dim=2048, attention_bias=True, activation_fn="gelu-approximate", norm_elementwise_affine=False,
    norm_eps=1e-6, caption_channels=4096, qk_norm="rms_norm", standardization_norm="rms_norm",
    adaptive_norm="single_scale_shift", positional_embedding_type="rope", positional_embedding_theta=10000.0,
    positional_embedding_max_pos=[20, 2048, 2048], timestep_scale_multiplier=1000, causal_temporal_positioning=False,
)


class Transformer3DModel(ModuleLike):
    def __init__(self, **config):
        cfg = dict(LTX_2B_CONFIG)
        cfg.update({k: v for k, v in config.items() if not k.startswith("_")})
        unsupported = []
        if cfg["activation_fn"] != "gelu-approximate": unsupported.append("activation_fn")
        if cfg["qk_norm"] != "rms_norm": unsupported.append("qk_norm")
        if cfg["standardization_norm"] != "rms_norm": unsupported.append("standardization_norm")
        if cfg["adaptive_norm"] != "single_scale_shift": unsupported.append("adaptive_norm")
        if cfg["positional_embedding_type"] != "rope": unsupported.append("positional_embedding_type")
        if cfg["norm_elementwise_affine"]: unsupported.append("norm_elementwise_affine")
        if cfg["attention_head_dim"] not in (64, 128): unsupported.append("attention_head_dim")
        if unsupported:
            raise NotImplementedError(f"Transformer3DModel(B200): unsupported config values for {unsupported}")
        self.config = SimpleNamespace(**cfg)
        self.num_attention_heads = cfg["num_attention_heads"]
        self.attention_head_dim = cfg["attention_head_dim"]
        self.inner_dim = self.num_attention_heads * self.attention_head_dim
        self.in_channels = cfg["in_channels"]
        self.out_channels = cfg["out_channels"] or cfg["in_channels"]
        self.num_layers = cfg["num_layers"]
        self.positional_embedding_theta = cfg["positional_embedding_theta"]
        self.positional_embedding_max_pos = cfg["positional_embedding_max_pos"]
        self.timestep_scale_multiplier = cfg["timestep_scale_multiplier"]
        self.use_tpu_flash_attention = False
        self.use_rope = True
        self.dtype = BF16
        self.device = torch.device("cuda")
        self.w: Dict[str, torch.Tensor] = {}
        self.layers: List[Dict[str, torch.Tensor]] = []

    @classmethod
    def from_config(cls, config: dict):
        return cls(**config)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, *args, device="cuda", **kwargs):
        """transformer3d.py:271-326: a diffusers directory (config mapped, keys renamed) or a single .safetensors with its config
        in the metadata."""
        from .checkpoint_io import load_transformer_checkpoint
        config, sd = load_transformer_checkpoint(pretrained_model_path)
        model = cls.from_config(config)
        model.load_state_dict(sd, device=device)
        return model

    # ---------------------------------------------------------------------------------------------
    # weights
    # ---------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True, device="cuda", **_):
        """Takes the reference layout (transformer3d.py:257-269 incl. the 'model.diffusion_model.' prefix)
        and packs it for the kernels: bf16, q/k/v (self) and k/v (cross) projections concatenated."""
        if any(k.startswith("model.diffusion_model.") for k in state_dict):
            state_dict = {k.replace("model.diffusion_model.", ""): v for k, v in state_dict.items()
                          if k.startswith("model.diffusion_model.")}
        self.device = torch.device(device)
        used = set()

        def get(name):
            used.add(name)
            return state_dict[name].to(device=self.device, dtype=BF16).contiguous()

        def lin(name):
            return get(name + ".weight"), get(name + ".bias")

        w = {}
        w["patchify_proj.w"], w["patchify_proj.b"] = lin("patchify_proj")
        w["t_l1.w"], w["t_l1.b"] = lin("adaln_single.emb.timestep_embedder.linear_1")
        w["t_l2.w"], w["t_l2.b"] = lin("adaln_single.emb.timestep_embedder.linear_2")
        w["t_lin.w"], w["t_lin.b"] = lin("adaln_single.linear")
        w["cap1.w"], w["cap1.b"] = lin("caption_projection.linear_1")
        w["cap2.w"], w["cap2.b"] = lin("caption_projection.linear_2")
        w["proj_out.w"], w["proj_out.b"] = lin("proj_out")
        w["final_table"] = get("scale_shift_table").view(1, 2, self.inner_dim)
        tables = []
        layers = []
        for i in range(self.num_layers):
            p = f"transformer_blocks.{i}."
            tables.append(get(p + "scale_shift_table"))
            L = {}
            qw, qb = lin(p + "attn1.to_q"); kw, kb = lin(p + "attn1.to_k"); vw, vb = lin(p + "attn1.to_v")
            L["qkv.w"] = torch.cat([qw, kw, vw], 0).contiguous(); L["qkv.b"] = torch.cat([qb, kb, vb], 0).contiguous()
            L["o.w"], L["o.b"] = lin(p + "attn1.to_out.0")
            L["qn"], L["kn"] = get(p + "attn1.q_norm.weight"), get(p + "attn1.k_norm.weight")
            L["q2.w"], L["q2.b"] = lin(p + "attn2.to_q")
            kw, kb = lin(p + "attn2.to_k"); vw, vb = lin(p + "attn2.to_v")
            L["kv2.w"] = torch.cat([kw, vw], 0).contiguous(); L["kv2.b"] = torch.cat([kb, vb], 0).contiguous()
            L["o2.w"], L["o2.b"] = lin(p + "attn2.to_out.0")
            L["qn2"], L["kn2"] = get(p + "attn2.q_norm.weight"), get(p + "attn2.k_norm.weight")
            L["ff1.w"], L["ff1.b"] = lin(p + "ff.net.0.proj")
            L["ff2.w"], L["ff2.b"] = lin(p + "ff.net.2")
            layers.append(L)
        w["block_tables"] = torch.stack(tables, 0).contiguous()          # [L, 6, D]
        missing = [k for k in state_dict if k not in used]
        if strict and missing:
            raise KeyError(f"unexpected keys in state_dict: {missing[:5]} ...")
        self.w, self.layers = w, layers
        return [], missing

    # ---------------------------------------------------------------------------------------------
    # companion API used by the pipeline
    # ---------------------------------------------------------------------------------------------
    def create_skip_layer_mask(self, batch_size: int, num_conds: int, ptb_index: int,
                               skip_block_list: Optional[List[int]] = None):
        """transformer3d.py:171-186"""
        if skip_block_list is None or len(skip_block_list) == 0:
            return None
        mask = torch.ones((self.num_layers, batch_size * num_conds), device=self.device, dtype=self.dtype)
        host = torch.ones((self.num_layers, batch_size * num_conds), dtype=torch.float32)
        for block_idx in skip_block_list:
            mask[block_idx, ptb_index::num_conds] = 0
            host[block_idx, ptb_index::num_conds] = 0
        mask._ltxb200_host = host                      # lets forward() pick skipped layers without a device sync (lives and dies with the mask)
        return mask

    def get_fractional_positions(self, indices_grid):
        return torch.stack([indices_grid[:, i] / self.positional_embedding_max_pos[i] for i in range(3)], dim=-1)

    def precompute_freqs_cis(self, indices_grid: torch.Tensor, spacing: str = "exp"):
        """transformer3d.py:202-255: fp32 table, cast to the model dtype.  Index/trig work done once per
        video with torch on the device the grid lives on (not part of the per-step hot loop)."""
        if spacing != "exp":
            raise NotImplementedError("only spacing='exp' is used on the reference path")
        dim, theta = self.inner_dim, self.positional_embedding_theta
        frac = self.get_fractional_positions(indices_grid.to(torch.float32))
        idx = theta ** torch.linspace(math.log(1, theta), math.log(theta, theta), dim // 6,
                                      device=frac.device, dtype=torch.float32)
        idx = idx * math.pi / 2
        freqs = (idx * (frac.unsqueeze(-1) * 2 - 1)).transpose(-1, -2).flatten(2)
        cos = freqs.cos().repeat_interleave(2, dim=-1)
        sin = freqs.sin().repeat_interleave(2, dim=-1)
        if dim % 6 != 0:
            cos = torch.cat([torch.ones_like(cos[:, :, : dim % 6]), cos], dim=-1)
            sin = torch.cat([torch.zeros_like(sin[:, :, : dim % 6]), sin], dim=-1)
        return cos.to(self.dtype), sin.to(self.dtype)

    # ---------------------------------------------------------------------------------------------
    # forward
    # ---------------------------------------------------------------------------------------------
    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward(self, hidden_states: torch.Tensor, freqs_cis, encoder_hidden_states: Optional[torch.Tensor] = None,
                timestep: Optional[torch.Tensor] = None, class_labels=None, cross_attention_kwargs: Dict[str, Any] = None,
                attention_mask: Optional[torch.Tensor] = None, encoder_attention_mask: Optional[torch.Tensor] = None,
                skip_layer_mask: Optional[torch.Tensor] = None, skip_layer_strategy: Optional[SkipLayerStrategy] = None,
                latent_shape=None, joint_pass: bool = True, ltxv_model=None, mixed: bool = False,
                return_dict: bool = True, shared_prefix: Optional[tuple] = None, encoder_key_lens: Optional[torch.Tensor] = None):
        """transformer3d.py:328-507.  hidden_states [B,N,C_in]; freqs_cis (cos,sin) [1|B,N,D]; encoder_hidden_states
        [B,L,caption_channels]; timestep [B,1] | [B,N]; encoder_attention_mask [B,L] (1 keep / 0 drop) or bias
        [B,1,L]; skip_layer_mask [layers,B].  `joint_pass=False` (per-sample iteration for offloaded weights,
        :472-487) computes the same function and is executed as the batched pass here."""
        if attention_mask is not None:
            raise NotImplementedError("self-attention masks are never passed on the reference path")
        w, D, H, dh = self.w, self.inner_dim, self.num_attention_heads, self.attention_head_dim
        dev = self.device
        B, N, Cin = hidden_states.shape
        x_in = hidden_states.to(device=dev, dtype=BF16).reshape(B * N, Cin)
        if not x_in.is_contiguous():
            x_in = x_in.contiguous()

        # --- cross-attention key bias (transformer3d.py:411-415): (1-mask)*-10000 in the model dtype
        # extension: `encoder_key_lens` int32 [B] = number of valid prompt tokens when the mask is RIGHT-PADDED (ones then zeros, what the
        # tokenizer produces): the same attention as the (1 - mask) * -10000 bias (exp(-10000 + ...) is exactly 0 in fp32) without the bias
        # pass and without the padded key blocks.  The pipeline derives it once per call from the host mask; it replaces the mask here.
        key_lens = None
        if encoder_key_lens is not None:
            key_lens = encoder_key_lens.to(device=dev, dtype=torch.int32).contiguous()
            encoder_attention_mask = None
        key_bias = None
        if encoder_attention_mask is not None:
            m = encoder_attention_mask.to(dev)
            if m.ndim == 2:
                m = (1 - m.to(BF16)) * -10000.0
            else:
                m = m.reshape(B, -1)
            key_bias = m.to(torch.float32).contiguous()

        # --- timestep embedding (AdaLayerNormSingle, transformer3d.py:420-438)
        t = timestep.to(device=dev, dtype=torch.float32)
        if self.timestep_scale_multiplier:
            t = self.timestep_scale_multiplier * t
        if t.shape[-1] > 1:
            t = t.reshape(t.shape[0], -1, latent_shape[-2] * latent_shape[-1])[:, :, 0]
        T = t.shape[1] if t.ndim == 2 else 1
        t_flat = t.reshape(-1).contiguous()
        rows_per_group = N // T
        assert rows_per_group * T == N, "tokens must split evenly over timestep groups"
        tp = ops.timestep_embed(t_flat, 256)
        e1 = ops.gemm(tp, w["t_l1.w"], w["t_l1.b"], act=ops.ACT_SILU)
        emb = ops.gemm(e1, w["t_l2.w"], w["t_l2.b"])                               # embedded_timestep [B*T, D]
        temb6 = ops.gemm(ops.act(emb, ops.ACT_SILU), w["t_lin.w"], w["t_lin.b"])     # [B*T, 6D]
        # mixed (:439-442): `timestep.float()`, `hidden_states.float()` — the modulation tables and the residual stream are fp32,
        # the Linear layers see bf16 inputs (autocast) and their outputs are accumulated into the fp32 stream by the GEMM epilogue
        ada = ops.ada_add_f32(w["block_tables"], temb6) if mixed else ops.ada_add(w["block_tables"], temb6)     # [L, B*T, 6, D]

        # --- caption projection (:446-451)
        enc = encoder_hidden_states.to(device=dev, dtype=BF16)
        Lc = enc.shape[1]
        enc2 = enc.reshape(B * Lc, enc.shape[2])
        if not enc2.is_contiguous():
            enc2 = enc2.contiguous()
        ctx = ops.gemm(ops.gemm(enc2, w["cap1.w"], w["cap1.b"], act=ops.ACT_GELU_TANH), w["cap2.w"], w["cap2.b"])

        cos, sin = freqs_cis
        cos = cos.to(device=dev, dtype=BF16).reshape(-1, D)
        sin = sin.to(device=dev, dtype=BF16).reshape(-1, D)
        assert cos.shape[0] == N, "freqs_cis must be shared across the batch ([1, N, D])"
        cos, sin = cos.contiguous(), sin.contiguous()

        x = ops.gemm(x_in, w["patchify_proj.w"], w["patchify_proj.b"])             # [B*N, D]
        if mixed:
            x = x.float()                                                           # the bf16 projection, upcast (:418, :442)
        norm_mod = ops.norm_mod_f32in if mixed else ops.norm_mod
        gemm_res = ops.gemm_f32res if mixed else ops.gemm

        skip_host = None
        if skip_layer_mask is not None:
            skip_host = getattr(skip_layer_mask, "_ltxb200_host", None)
            if skip_host is None or tuple(skip_host.shape) != tuple(skip_layer_mask.shape):
                skip_host = skip_layer_mask.to(torch.float32).cpu()                 # foreign mask: one sync per forward
            skip_dev = skip_layer_mask.to(device=dev, dtype=torch.float32).contiguous()

        B_full, first_div = B, 0
        if shared_prefix is not None:
            n_dup, src = shared_prefix
            assert 0 < n_dup and 0 <= src and src + n_dup <= B - n_dup, "shared_prefix: duplicates must trail their sources"
            first_div = len(self.layers)
            if skip_host is not None:
                differs = (skip_host[:, B - n_dup:] != skip_host[:, src:src + n_dup]).any(dim=1)
                if bool(differs.any()):
                    first_div = int(differs.float().argmax())
            if first_div > 0:
                B = B_full - n_dup
                x = x[: B * N]

        def expand(x):
            return torch.cat([x, x[src * N:(src + n_dup) * N]], dim=0)

        for li, Lw in enumerate(self.layers):
            if B != B_full and li == first_div:
                x, B = expand(x), B_full
            if B != B_full:
                ctx_l, key_bias_l = ctx[: B * Lc], (key_bias[:B] if key_bias is not None else None)
                key_lens_l = key_lens[:B] if key_lens is not None else None
            else:
                ctx_l, key_bias_l, key_lens_l = ctx, key_bias, key_lens
            a = ada[li]                                                             # [B*T, 6, D]
            layer_skip = skip_host is not None and float(skip_host[li].min()) != 1.0
            # SkipLayerStrategy.Residual needs no branch: the reference only applies it under `attn.residual_connection`
            # (attention.py:1161-1168), which BasicTransformerBlock never enables, so the mask is ignored there as it is here.
            x_orig =x.clone() if (layer_skip and skip_layer_strategy == SkipLayerStrategy.TransformerBlock) else None
            # ---- self attention (attention.py:233-288)
            nh = norm_mod(x, a[:, 1], a[:, 0], rows_per_group=rows_per_group, eps=self.config.norm_eps)
            qkv = ops.gemm(nh, Lw["qkv.w"], Lw["qkv.b"])                            # [B*N, 3D]
            ops.qk_norm_rope(qkv[:, :D], qkv[:, D:2 * D], Lw["qn"], Lw["kn"], cos, sin, tokens_per_batch=N, eps=1e-5)
            q4 = qkv.view(B, N, 3 * D)[:, :, :D].unflatten(-1, (H, dh))
            k4 = qkv.view(B, N, 3 * D)[:, :, D:2 * D].unflatten(-1, (H, dh))
            v4 = qkv.view(B, N, 3 * D)[:, :, 2 * D:].unflatten(-1, (H, dh))
            o = ops.attention(q4, k4, v4)                                           # [B, N, H, dh]
            if layer_skip and skip_layer_strategy == SkipLayerStrategy.AttentionValues:
                ops.stg_blend(o.view(B, N, D), qkv[:, 2 * D:], skip_dev[li][:B])        # :1134-1139
            elif layer_skip and skip_layer_strategy == SkipLayerStrategy.AttentionSkip:
                ops.stg_blend(o.view(B, N, D), nh, skip_dev[li][:B])                    # :1127-1133
            gemm_res(o.view(B * N, D), Lw["o.w"], Lw["o.b"], residual=x, gate=a[:, 2], rows_per_gate=rows_per_group, out=x)
            # ---- cross attention (attention.py:294-311): query from the raw residual stream, no RoPE
            q2 = ops.gemm(ops.cast_bf16(x) if mixed else x, Lw["q2.w"], Lw["q2.b"])
            kv = ops.gemm(ctx_l, Lw["kv2.w"], Lw["kv2.b"])                          # [B*Lc, 2D]
            ops.qk_norm_rope(q2, kv[:, :D], Lw["qn2"], Lw["kn2"], None, None, eps=1e-5)
            o2 = ops.attention(q2.view(B, N, H, dh), kv.view(B, Lc, 2 * D)[:, :, :D].unflatten(-1, (H, dh)),
                               kv.view(B, Lc, 2 * D)[:, :, D:].unflatten(-1, (H, dh)), key_bias=key_bias_l, key_lens=key_lens_l)
            gemm_res(o2.view(B * N, D), Lw["o2.w"], Lw["o2.b"], residual=x, out=x)
            # ---- feed forward (attention.py:314-351)
            nh = norm_mod(x, a[:, 4], a[:, 3], rows_per_group=rows_per_group, eps=self.config.norm_eps)
            ff = ops.gemm(nh, Lw["ff1.w"], Lw["ff1.b"], act=ops.ACT_GELU_TANH)
            gemm_res(ff, Lw["ff2.w"], Lw["ff2.b"], residual=x, gate=a[:, 5], rows_per_gate=rows_per_group, out=x)
            if x_orig is not None:
                if mixed:
                    # fp32 residual stream: x * mask + x_orig * (1 - mask) (:355-362) with a 0 / 1 mask per batch row IS a row copy — the
                    # rows whose mask is 0 take their block input back, the others are untouched (1 * x + 0 * x_orig is exact in fp32)
                    mrow = skip_host[li][:B]
                    if not bool(((mrow == 0) | (mrow == 1)).all()):
                        raise NotImplementedError("mixed precision with a fractional SkipLayerStrategy.TransformerBlock mask")
                    for b_ in torch.nonzero(mrow == 0).flatten().tolist():
                        x.view(B, N, D)[b_].copy_(x_orig.view(B, N, D)[b_])
                else:
                    ops.stg_blend(x.view(B, N, D), x_orig, skip_dev[li][:B])           # :355-362
            if ltxv_model is not None and getattr(ltxv_model, "_interrupt", False):
                return [None]

        if B != B_full:
            x, B = expand(x), B_full
        # --- output head (:490-503)
        fin = (ops.ada_add_f32 if mixed else ops.ada_add)(w["final_table"], torch.cat([emb, emb], dim=1))   # [1, B*T, 2, D]
        y = norm_mod(x, fin[0][:, 1], fin[0][:, 0], rows_per_group=rows_per_group, eps=1e-6, layer_norm=True)
        out = ops.gemm(y, w["proj_out.w"], w["proj_out.b"]).view(B, N, self.out_channels)
        if not return_dict:
            return (out,)
        return Transformer3DModelOutput(sample=out)
