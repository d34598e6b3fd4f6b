"""WanModel — B200-native drop-in for wan/modules/model.py:591-1160 (t2v core: WanAttentionBlock :354-499,
WanSelfAttention :150-240, WanT2VCrossAttention :243-274, WanRMSNorm :91-111, WanLayerNorm :130-145,
Head :539-573), with the Ulysses sequence-parallel forward of wan/distributed/xdit_context_parallel.py:66-192
built in (`sp_group`).

Same constructor keys, state_dict layout and forward signature
(`forward(x: list[Tensor[C,T,H,W]], t, context: list[Tensor[L,4096]], freqs=(cos,sin), pipeline=…, …) ->
list[Tensor float32]`).  The reference iterates the sequences of a joint pass inside every block so that
offloaded weights are fetched once (:1082-1085); on a 180 GB part the sequences are simply batched.
Step skipping is built in: skip-layer guidance (`slg_layers`, model.py:1077-1080) runs the listed blocks
on the conditional sequence only, TeaCache (`enable_teacache`, model.py:854-899,1029-1101) replays the previous step's
block-stack residual when the accumulated time-embedding distance stays under `rel_l1_thresh`.
VACE / recam / fantasytalking branches are add-ons outside the benchmarked configs and raise
NotImplementedError.

Sequence parallelism (P ranks, one process per GPU): tokens are split contiguously (:131-133); everything
token-wise stays local; around self-attention q/k/v are exchanged heads<->sequence with one all-to-all each
way (DeepSpeed-Ulysses), RoPE uses the rank's global token offset (:52-57) and the full-dim QK-RMSNorm runs
before the exchange while rows still hold all heads; the head output is all-gathered at the end (:142).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from .. import ops
from ..module_like import ModuleLike

BF16 = torch.bfloat16

WAN_T2V_1_3B = dict(model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=1536, ffn_dim=8960,
                    freq_dim=256, text_dim=4096, out_dim=16, num_heads=12, num_layers=30, qk_norm=True,
                    cross_attn_norm=True, eps=1e-6)        # wan/configs/wan_t2v_1_3B.py:19-29
WAN_T2V_14B = dict(WAN_T2V_1_3B, dim=5120, ffn_dim=13824, num_heads=40, num_layers=40)   # wan_t2v_14B.py:19-29
WAN_I2V_14B = dict(WAN_T2V_14B, model_type="i2v", in_dim=36)                              # wan_i2v_14B.py:25-35
CLIP_TOKENS = 257                                                                         # WanI2VCrossAttention :306-307


class WanModel(ModuleLike):
    def __init__(self, vace_layers=None, vace_in_dim=None, model_type="t2v", patch_size=(1, 2, 2), text_len=512,
                 in_dim=16, dim=2048, ffn_dim=8192, freq_dim=256, text_dim=4096, out_dim=16, num_heads=16,
                 num_layers=32, window_size=(-1, -1), qk_norm=True, cross_attn_norm=True, eps=1e-6, recammaster=False,
                 inject_sample_info=False, fantasytalking_dim=0, sp_group=None):
        if model_type not in ("t2v", "i2v"):
            raise NotImplementedError("model_type must be 't2v' or 'i2v' (model.py:717)")
        if vace_layers is not None or recammaster or inject_sample_info or fantasytalking_dim:
            raise NotImplementedError("VACE / recam / sample-info / fantasytalking branches are out of scope")
        if not (qk_norm and cross_attn_norm) or tuple(window_size) != (-1, -1):
            raise NotImplementedError("only qk_norm=True, cross_attn_norm=True, full attention")
        if dim % num_heads or dim // num_heads != 128:
            raise NotImplementedError("head_dim must be 128 (posemb_layers.py:457-458 hard-codes [44,42,42])")
        self.config = SimpleNamespace(model_type=model_type, patch_size=tuple(patch_size), text_len=text_len, in_dim=in_dim,
                                      dim=dim, ffn_dim=ffn_dim, freq_dim=freq_dim, text_dim=text_dim, out_dim=out_dim,
                                      num_heads=num_heads, num_layers=num_layers, eps=eps)
        self.model_type, self.patch_size, self.text_len = model_type, tuple(patch_size), text_len
        self.in_dim, self.dim, self.ffn_dim, self.freq_dim = in_dim, dim, ffn_dim, freq_dim
        self.text_dim, self.out_dim, self.num_heads, self.num_layers, self.eps = text_dim, out_dim, num_heads, num_layers, eps
        # TeaCache state: same attribute names as the reference (set by the caller / text2video.py:461-464)
        self.enable_teacache = False
        self.coefficients, self.rel_l1_thresh, self.teacache_start_step, self.num_steps = [1.0, 0.0], 0.0, 0, 0
        self.teacache_multiplier = 0
        self.accumulated_rel_l1_distance, self.teacache_skipped_steps = 0, 0
        self.previous_residual, self.previous_modulated_input, self.should_calc = [None, None], None, True
        self.dtype = BF16
        self.device = torch.device("cuda")
        self.w: Dict[str, torch.Tensor] = {}
        self.layers: List[Dict[str, torch.Tensor]] = []
        self.sp_group = sp_group
        # "p2p": producing kernels store into the peers' buffers over NVLink (distributed/ulysses.py:PeerExchange);
        # "nccl": pack + all_to_all_single (kept as the comparison baseline, LTXB200_SP_EXCHANGE=nccl)
        import os
        self.sp_exchange = os.environ.get("LTXB200_SP_EXCHANGE", "p2p")
        self._sp_bufs = {}

    # ---------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True, device="cuda", **_):
        self.device = dev = torch.device(device)
        used = set()

        def get(name):
            used.add(name)
            return state_dict[name].to(device=dev, dtype=BF16).contiguous()

        def lin(name):
            return get(name + ".weight"), get(name + ".bias")

        w = {}
        w["patch.w"] = get("patch_embedding.weight").flatten(1).contiguous()        # [D, C*1*2*2]
        w["patch.b"] = get("patch_embedding.bias")
        w["text0.w"], w["text0.b"] = lin("text_embedding.0"); w["text2.w"], w["text2.b"] = lin("text_embedding.2")
        w["time0.w"], w["time0.b"] = lin("time_embedding.0"); w["time2.w"], w["time2.b"] = lin("time_embedding.2")
        w["tproj.w"], w["tproj.b"] = lin("time_projection.1")
        w["head.w"], w["head.b"] = lin("head.head")
        w["head_mod"] = get("head.modulation").view(1, 2, self.dim)
        mods, layers = [], []
        for i in range(self.num_layers):
            p = f"blocks.{i}."
            mods.append(get(p + "modulation").view(6, self.dim))
            L = {}
            qw, qb = lin(p + "self_attn.q"); kw, kb = lin(p + "self_attn.k"); vw, vb = lin(p + "self_attn.v")
            L["qkv.w"] = torch.cat([qw, kw, vw], 0).contiguous(); L["qkv.b"] = torch.cat([qb, kb, vb], 0).contiguous()
            L["o.w"], L["o.b"] = lin(p + "self_attn.o")
            L["qn"], L["kn"] = get(p + "self_attn.norm_q.weight"), get(p + "self_attn.norm_k.weight")
            L["q2.w"], L["q2.b"] = lin(p + "cross_attn.q")
            kw, kb = lin(p + "cross_attn.k"); vw, vb = lin(p + "cross_attn.v")
            L["kv2.w"] = torch.cat([kw, vw], 0).contiguous(); L["kv2.b"] = torch.cat([kb, vb], 0).contiguous()
            L["o2.w"], L["o2.b"] = lin(p + "cross_attn.o")
            L["qn2"], L["kn2"] = get(p + "cross_attn.norm_q.weight"), get(p + "cross_attn.norm_k.weight")
            if self.model_type == "i2v":        # WanI2VCrossAttention k_img / v_img / norm_k_img (:288-291)
                kw, kb = lin(p + "cross_attn.k_img"); vw, vb = lin(p + "cross_attn.v_img")
                L["kvi.w"] = torch.cat([kw, vw], 0).contiguous(); L["kvi.b"] = torch.cat([kb, vb], 0).contiguous()
                L["kni"] = get(p + "cross_attn.norm_k_img.weight")
            L["n3.w"], L["n3.b"] = get(p + "norm3.weight"), get(p + "norm3.bias")
            L["ff1.w"], L["ff1.b"] = lin(p + "ffn.0"); L["ff2.w"], L["ff2.b"] = lin(p + "ffn.2")
            layers.append(L)
        w["block_mods"] = torch.stack(mods, 0).contiguous()                          # [L, 6, D]
        if self.model_type == "i2v":            # img_emb = MLPProj(1280, dim) (:576-588, :768-769)
            w["ie.ln0.w"], w["ie.ln0.b"] = get("img_emb.proj.0.weight"), get("img_emb.proj.0.bias")
            w["ie.1.w"], w["ie.1.b"] = lin("img_emb.proj.1"); w["ie.3.w"], w["ie.3.b"] = lin("img_emb.proj.3")
            w["ie.ln4.w"], w["ie.ln4.b"] = get("img_emb.proj.4.weight"), get("img_emb.proj.4.bias")
        extra = [k for k in state_dict if k not in used]
        if strict and extra:
            raise KeyError(f"unexpected keys in state_dict: {extra[:5]} ...")
        self.w, self.layers = w, layers
        return [], extra

    # ---------------------------------------------------------------------------------------------
    def _patchify(self, x: torch.Tensor) -> torch.Tensor:
        """Conv3d(k=s=(1,2,2)) as a GEMM: [C,F,H,W] -> rows [N, C*4], k = (c, pt, ph, pw) (model.py:951-954)."""
        C, Fr, H, W = x.shape
        return x.view(C, Fr, 1, H // 2, 2, W // 2, 2).permute(1, 3, 5, 0, 2, 4, 6).reshape(Fr * (H // 2) * (W // 2), C * 4)

    def unpatchify(self, x: torch.Tensor, grid_sizes):
        """model.py:1113-1136 ('fhwpqrc->cfphqwr'); x [B, N, prod(patch)*c] -> list of [c, F, H, W]"""
        c = self.out_dim
        out = []
        for u in x:
            u = u[: math.prod(grid_sizes)].view(*grid_sizes, *self.patch_size, c)
            u = u.permute(6, 0, 3, 1, 4, 2, 5)
            out.append(u.reshape(c, *[i * j for i, j in zip(grid_sizes, self.patch_size)]))
        return out

    # ---------------------------------------------------------------------------------------------
    def _sp(self):
        g = self.sp_group
        if g is None:
            return 1, 0
        import torch.distributed as dist
        return dist.get_world_size(g), dist.get_rank(g)

    def _peer_exchange(self, B: int, n_loc: int):
        """One set of peer-mapped buffers per (sequences, local tokens) shape; every rank sees the same sequence of shapes, so
        creation and the eviction of the oldest shape beyond four are collective."""
        key = (B, n_loc)
        if key not in self._sp_bufs:
            from .distributed.ulysses import PeerExchange
            while len(self._sp_bufs) >= 4:
                self._sp_bufs.pop(next(iter(self._sp_bufs))).close()
            self._sp_bufs[key] = PeerExchange(self.sp_group, B, n_loc, self.num_heads, 128, self.device,
                                              gather_width=math.prod(self.patch_size) * self.out_dim)
        return self._sp_bufs[key]

    def close(self):
        """Release the peer-mapped exchange buffers (IPC mappings + device memory); safe to call more than once."""
        for ex in self._sp_bufs.values():
            ex.close()
        self._sp_bufs = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _self_attention_sp(self, qkv: torch.Tensor, B: int, n_loc: int, P: int) -> torch.Tensor:
        """Ulysses exchange around self-attention (see distributed/ulysses.py)."""
        from .distributed.ulysses import ulysses_self_attention
        return ulysses_self_attention(qkv, B, n_loc, self.num_heads, 128, self.sp_group,
                                      lambda q, k, v, out: ops.attention(q, k, v, out=out))

    # ---------------------------------------------------------------------------------------------
    def _time_embedding(self, tt: torch.Tensor) -> torch.Tensor:
        """time_embedding(sinusoidal_embedding_1d(freq_dim, t)) (model.py:979-981) -> [n, D] bf16"""
        w = self.w
        te = ops.timestep_embed(tt, self.freq_dim)
        return ops.gemm(ops.gemm(te, w["time0.w"], w["time0.b"], act=ops.ACT_SILU), w["time2.w"], w["time2.b"])

    def _teacache_should_calc(self, e: torch.Tensor, current_step: int, x_id: int) -> bool:
        """model.py:1029-1049: the conditional pass (x_id 0) decides, the others follow.  Under CFG-parallel the ranks of the
        unconditional half only ever run x_id 1; the decision is a function of the replicated time embedding alone, so they take
        it themselves (`_teacache_every_rank_decides`, set by WanT2V.generate) and both halves skip the same steps."""
        if x_id != 0 and not getattr(self, "_teacache_every_rank_decides", False):
            return self.should_calc
        if current_step <= self.teacache_start_step or current_step == self.num_steps - 1:
            should_calc = True
            self.accumulated_rel_l1_distance = 0
        else:
            import numpy as np
            delta = abs(np.poly1d(self.coefficients)(ops.rel_l1(e, self.previous_modulated_input)))
            self.accumulated_rel_l1_distance += delta
            if self.accumulated_rel_l1_distance < self.rel_l1_thresh:
                should_calc = False
                self.teacache_skipped_steps += 1
            else:
                should_calc = True
                self.accumulated_rel_l1_distance = 0
        self.previous_modulated_input = e
        self.should_calc = should_calc
        return should_calc

    def compute_teacache_threshold(self, start_step, timesteps=None, speed_factor=0):
        """model.py:854-899: scan thresholds 0.01..0.6 for the one whose simulated schedule is closest to
        len(timesteps)/speed_factor computed steps; sets and returns `rel_l1_thresh`."""
        import numpy as np
        rescale = np.poly1d(self.coefficients)
        tt = torch.as_tensor([float(t) for t in timesteps], device=self.device, dtype=torch.float32)
        e_all = self._time_embedding(tt)
        rel = [0.0] + [ops.rel_l1(e_all[i:i + 1].contiguous(), e_all[i - 1:i].contiguous()) for i in range(1, len(timesteps))]
        target = int(len(timesteps) / speed_factor)
        best_threshold, best_diff, best_signed, threshold = 0.01, 1000, 1000, 0.01
        while threshold <= 0.6:
            acc, nb, diff, signed = 0, 0, 1000, 1000
            for i in range(len(timesteps)):
                skip = False
                if not (i <= start_step or i == len(timesteps) - 1):
                    acc += abs(rescale(rel[i]))
                    if acc < threshold:
                        skip = True
                    else:
                        acc = 0
                if not skip:
                    nb += 1
                    signed = target - nb
                    diff = abs(signed)
            if diff < best_diff:
                best_threshold, best_diff, best_signed = threshold, diff, signed
            elif diff > best_diff:
                break
            threshold += 0.01
        self.rel_l1_thresh = best_threshold
        return best_threshold

    def _block(self, li: int, xs: torch.Tensor, st) -> None:
        """WanAttentionBlock.forward (model.py:411-499) on the rows `xs` [B*n_loc, D], updated in place."""
        w, D, H, eps = self.w, self.dim, self.num_heads, self.eps
        B, n_loc, P, rank, mods, ctx, ctx_img, cos, sin = st.B, st.n_loc, st.P, st.rank, st.mods, st.ctx, st.ctx_img, st.cos, st.sin
        M, Lc, Lw = B * n_loc, self.text_len, self.layers[li]
        m = mods[li]                                                              # [1, 6, D]
        xm = ops.norm_mod(xs, m[:, 1], m[:, 0], rows_per_group=M, eps=eps, layer_norm=True)        # :437-441
        if P > 1 and self.sp_exchange == "p2p":
            # QKV projection in token chunks, q/k norm + RoPE fused with the head scatter (overlapping the next chunk's GEMM),
            # attention epilogue fused with the return scatter
            o = self._peer_exchange(B, n_loc).self_attention(xm, Lw["qkv.w"], Lw["qkv.b"], Lw["qn"], Lw["kn"], cos, sin, eps)
        else:
            qkv = ops.gemm(xm, Lw["qkv.w"], Lw["qkv.b"])
            ops.qk_norm_rope_wan(qkv[:, :D], qkv[:, D:2 * D], Lw["qn"], Lw["kn"], cos, sin, head_dim=128,
                                     tokens_per_batch=n_loc, token_offset=rank * n_loc, eps=eps)
            if P == 1:
                q3 = qkv.view(B, n_loc, 3 * D)
                o = ops.attention(q3[:, :, :D].unflatten(-1, (H, 128)), q3[:, :, D:2 * D].unflatten(-1, (H, 128)),
                                  q3[:, :, 2 * D:].unflatten(-1, (H, 128))).view(M, D)
            else:
                o = self._self_attention_sp(qkv, B, n_loc, P)
        ops.gemm(o, Lw["o.w"], Lw["o.b"], residual=xs, gate=m[:, 2], rows_per_gate=M, out=xs)      # x.addcmul_(y, e2) :458
        y3 = ops.norm_mod(xs, weight=Lw["n3.w"], bias=Lw["n3.b"], eps=eps, layer_norm=True)         # norm3 :461
        q2 = ops.gemm(y3, Lw["q2.w"], Lw["q2.b"])
        kv = ops.gemm(ctx, Lw["kv2.w"], Lw["kv2.b"])                                                # [B*512, 2D]
        ops.qk_norm_rope_wan(q2, kv[:, :D], Lw["qn2"], Lw["kn2"], None, None, eps=eps)
        kv3 = kv.view(B, Lc, 2 * D)
        o2 = ops.attention(q2.view(B, n_loc, H, 128), kv3[:, :, :D].unflatten(-1, (H, 128)), kv3[:, :, D:].unflatten(-1, (H, 128)))
        if ctx_img is not None:             # WanI2VCrossAttention :323-337: same q over the 257 image tokens, x += img_x
            kvi = ops.gemm(ctx_img, Lw["kvi.w"], Lw["kvi.b"])                                       # [257, 2D]
            ops.qk_norm_rope_wan(None, kvi[:, :D], None, Lw["kni"], None, None, eps=eps)
            kvb = kvi.unsqueeze(0).repeat(B, 1, 1) if B > 1 else kvi.unsqueeze(0)                   # same image for every sequence
            ops.attention(q2.view(B, n_loc, H, 128), kvb[:, :, :D].unflatten(-1, (H, 128)),
                          kvb[:, :, D:].unflatten(-1, (H, 128)), out=o2, accumulate=True)
        ops.gemm(o2.view(M, D), Lw["o2.w"], Lw["o2.b"], residual=xs, out=xs)                       # x += cross_attn :465
        y2 = ops.norm_mod(xs, m[:, 4], m[:, 3], rows_per_group=M, eps=eps, layer_norm=True)        # :467-472
        ff = ops.gemm(y2, Lw["ff1.w"], Lw["ff1.b"], act=ops.ACT_GELU_TANH)
        ops.gemm(ff, Lw["ff2.w"], Lw["ff2.b"], residual=xs, gate=m[:, 5], rows_per_gate=M, out=xs) # :490-492

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def forward(self, x, t, context, vace_context=None, vace_context_scale=1.0, clip_fea=None, y=None, freqs=None,
                pipeline=None, current_step=0, x_id=0, max_steps=0, slg_layers=None, callback=None, cam_emb=None,
                fps=None, causal_block_size=1, causal_attention=False, audio_proj=None, audio_context_lens=None,
                audio_scale=None):
        """model.py:902-1111 (t2v)."""
        if vace_context is not None or cam_emb is not None or audio_proj is not None:
            raise NotImplementedError("VACE / camera / audio inputs are out of scope")
        if self.model_type == "i2v":
            assert clip_fea is not None and y is not None                             # :930-931
        elif clip_fea is not None or y is not None:
            raise ValueError("clip_fea / y are i2v inputs (model_type='i2v')")
        w, D, H, eps = self.w, self.dim, self.num_heads, self.eps
        dev = self.device
        P, rank = self._sp()
        x_list = x
        B = len(x_list)
        C, Fr, Hh, Ww = x_list[0].shape
        grid = (Fr, Hh // 2, Ww // 2)
        N = math.prod(grid)
        if N % P or H % P:
            raise ValueError(f"Ulysses needs tokens ({N}) and heads ({H}) divisible by the group size ({P}); for 12 heads on 8 GPUs use "
                             "CFG-parallel x Ulysses (wan/distributed/cfg_parallel.py)")
        n_loc = N // P
        # ---- embeddings (replicated; the token shard is taken right after the patch rows are built, :131-133)
        if y is not None:                       # i2v: [mask(4) | image latent(16)] channels appended to every sequence (:948-949)
            yd = y.to(dev)
            x_list = [torch.cat([u.to(dev), yd.to(u.dtype)], dim=0) for u in x_list]
        rows = torch.stack([self._patchify(u.to(dev)) for u in x_list], 0)            # [B, N, C*4]
        rows = rows[:, rank * n_loc:(rank + 1) * n_loc].to(BF16).reshape(B * n_loc, -1).contiguous()
        xs = ops.gemm(rows, w["patch.w"], w["patch.b"])                               # [B*n_loc, D]
        tt = t.to(device=dev, dtype=torch.float32).flatten().contiguous()
        if tt.numel() != 1:
            raise NotImplementedError("per-frame timesteps (diffusion forcing, model.py:976) are out of scope")
        e = self._time_embedding(tt)                                                  # sinusoidal_embedding_1d (:18-28) -> [1, D]
        e0 = ops.gemm(ops.act(e, ops.ACT_SILU), w["tproj.w"], w["tproj.b"])           # [1, 6D]
        mods = ops.ada_add(w["block_mods"], e0)                                       # [L, 1, 6, D] = modulation + e0 (:436)
        ctx_in = torch.zeros(B, self.text_len, self.text_dim, device=dev, dtype=BF16)  # zero-pad THEN embed (:994)
        for i, u in enumerate(context):
            ctx_in[i, : u.shape[0]] = u.to(device=dev, dtype=BF16)
        ctx = ops.gemm(ops.gemm(ctx_in.view(B * self.text_len, -1), w["text0.w"], w["text0.b"], act=ops.ACT_GELU_TANH),
                       w["text2.w"], w["text2.b"])                                    # [B*512, D]
        ctx_img = None
        if clip_fea is not None:                # img_emb: LayerNorm -> Linear -> GELU(erf) -> Linear -> LayerNorm (:996-998)
            cf = clip_fea.to(device=dev, dtype=BF16).reshape(-1, clip_fea.shape[-1]).contiguous()          # [257, 1280]
            assert cf.shape[0] == CLIP_TOKENS, "clip_fea must be [1, 257, 1280] (one image for all sequences)"
            c1 = ops.norm_mod(cf, weight=w["ie.ln0.w"], bias=w["ie.ln0.b"], eps=1e-5, layer_norm=True)
            c2 = ops.gemm(ops.gemm(c1, w["ie.1.w"], w["ie.1.b"], act=ops.ACT_GELU_ERF), w["ie.3.w"], w["ie.3.b"])
            ctx_img = ops.norm_mod(c2, weight=w["ie.ln4.w"], bias=w["ie.ln4.b"], eps=1e-5, layer_norm=True)  # [257, D]
        cos, sin = freqs
        cos = cos.to(device=dev, dtype=torch.float32).contiguous()
        sin = sin.to(device=dev, dtype=torch.float32).contiguous()
        M = B * n_loc
        Lc = self.text_len

        joint_pass = B > 1
        should_calc = self._teacache_should_calc(e, current_step, x_id) if self.enable_teacache else True
        if not should_calc:                                                            # :1051-1058 replay the cached residual
            if joint_pass:
                for i in range(B):
                    ops.axpby(xs[i * n_loc:(i + 1) * n_loc], self.previous_residual[i], out=xs[i * n_loc:(i + 1) * n_loc])
            else:
                ops.axpby(xs, self.previous_residual[x_id], out=xs)
        else:
            ori = xs.clone() if self.enable_teacache else None                         # :1065-1068
            st = SimpleNamespace(B=B, n_loc=n_loc, P=P, rank=rank, mods=mods, ctx=ctx, ctx_img=ctx_img, cos=cos, sin=sin)
            for li in range(self.num_layers):
                # mid-forward interrupts only on a single rank: with peers every rank must issue the same kernel sequence (the
                # exchange kernels wait for each other); the loops check `_interrupt` collectively at step boundaries instead
                if P == 1 and pipeline is not None and getattr(pipeline, "_interrupt", False):
                    return [None] * B
                if (x_id != 0 or joint_pass) and slg_layers is not None and li in slg_layers:      # :1077-1080
                    if not joint_pass:
                        continue                                                       # unconditional pass: block dropped
                    # joint pass: only the conditional sequence (index 0) goes through the block; the rows are updated in place
                    s1 = SimpleNamespace(**{**vars(st), "B": 1, "ctx": ctx[:Lc]})
                    self._block(li, xs[:n_loc], s1)
                else:
                    self._block(li, xs, st)
            if self.enable_teacache:                                                   # :1087-1101 residual of the block stack
                ops.axpby(xs, ori, 1.0, -1.0, out=ori)
                if joint_pass:
                    self.previous_residual = [ori[i * n_loc:(i + 1) * n_loc] for i in range(B)]
                else:
                    self.previous_residual[x_id] = ori

        eh = ops.ada_add(w["head_mod"], torch.cat([e, e], dim=1))                     # [1, 1, 2, D] = modulation + e (:566)
        yh = ops.norm_mod(xs, eh[0][:, 1], eh[0][:, 0], rows_per_group=M, eps=eps, layer_norm=True)
        out = ops.gemm(yh, w["head.w"], w["head.b"], out_f32=True).view(B, n_loc, -1)  # [B, n_loc, 64] fp32
        if P > 1 and self.sp_exchange == "p2p":                                        # :142, as peer stores + flag (no collective call)
            out = self._peer_exchange(B, n_loc).all_gather_rows(out.contiguous())
        elif P > 1:
            import torch.distributed as dist
            parts = [torch.empty_like(out) for _ in range(P)]
            dist.all_gather(parts, out.contiguous(), group=self.sp_group)
            out = torch.cat(parts, dim=1)
        return [u.float() for u in self.unpatchify(out, grid)]
