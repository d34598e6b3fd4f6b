"""Wan2.1 T2V-1.3B at FULL width and depth (BASELINE.json configs[3]'s network: dim 1536, 12 heads x 128, ffn 8960, 30 layers, 1.42 B
seeded weights) through the UNMODIFIED reference WanModel.forward and the reference's UniPC scheduler, at a small latent
(16, 3, 8, 12) -> 72 tokens, recorded as tests/golden/wan_1_3b_full.pt (TEST INFRASTRUCTURE ONLY).

Runs in float64 for the reason given in oracle/gen_golden_wan.py (WanRMSNorm's fp32-only in-place square).  Asserts that
oracle/wan_oracle.py reproduces the joint cond / uncond forward and a 3-step CFG loop, then stores the reference's outputs.
Build container only (needs /root/reference, ~30 GB of RAM):  python oracle/gen_golden_wan_full.py"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
import load_reference  # noqa: E402

load_reference.install()
from oracle import wan_oracle as W  # noqa: E402
from oracle.gen_golden_wan import GOLD, _Pipe, build_ref  # noqa: E402
from oracle.ltx_oracle import rel_l2  # noqa: E402

torch.set_grad_enabled(False)
STEPS, SHIFT, GUIDE = 3, 5.0, 5.0


def main():
    from wan.modules.posemb_layers import get_rotary_pos_embed
    from wan.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler
    cfg = dict(W.WAN_1_3B)
    t0 = time.perf_counter()
    sd = {k: v.double() for k, v in W.make_wan_state_dict(cfg, seed=0).items()}
    print(f"  weights: {sum(v.numel() for v in sd.values()) / 1e9:.2f} B parameters in {time.perf_counter() - t0:.1f} s")
    ref = build_ref(cfg, sd)
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(16, 3, 8, 12, generator=g).double()
    ctx = torch.randn(20, 4096, generator=g).double()
    ctx0 = torch.randn(11, 4096, generator=g).double()
    cos_r, sin_r = get_rotary_pos_embed(lat.shape[1:], enable_RIFLEx=False)
    cos, sin = W.rope_tables(lat.shape[1:])
    assert torch.equal(cos, cos_r) and torch.equal(sin, sin_r)
    t = torch.tensor([937])
    y_ref = ref([lat.clone(), lat.clone()], t=t, context=[ctx, ctx0], freqs=(cos_r, sin_r), pipeline=_Pipe())
    y = W.wan_forward(sd, cfg, [lat, lat], t, [ctx, ctx0], cos, sin)
    for a, b in zip(y, y_ref):
        e = rel_l2(a, b)
        print(f"  wan_forward (30 layers): rel_l2(oracle, reference) = {e:.3e}")
        assert e < 2e-5
    s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    s.set_timesteps(STEPS, device="cpu", shift=SHIFT)
    latents, ref_steps = lat.clone(), []
    for tt in s.timesteps:                                  # text2video.py:468-575, plain CFG
        c, u = ref([latents, latents], t=torch.stack([tt]), context=[ctx, ctx0], freqs=(cos_r, sin_r), pipeline=_Pipe())
        latents = s.step((u + GUIDE * (c - u)).unsqueeze(0), tt, latents.unsqueeze(0), return_dict=False)[0].squeeze(0)
        ref_steps.append(latents.clone())
    mine = []
    W.t2v_denoise(sd, cfg, lat, ctx, ctx0, steps=STEPS, shift=SHIFT, guide_scale=GUIDE, per_step=mine)
    for i, (a, b) in enumerate(zip(mine, ref_steps)):
        e = rel_l2(a, b)
        print(f"  t2v loop step {i}: rel_l2 = {e:.3e}")
        assert e < 5e-5
    torch.save(dict(cfg=cfg, seed_weights=0, lat=lat.float(), ctx=ctx.float(), ctx0=ctx0.float(), t=t, steps=STEPS, shift=SHIFT, guide=GUIDE,
                    fwd=[a.float().clone() for a in y_ref], loop=[a.float() for a in ref_steps]),
               os.path.join(GOLD, "wan_1_3b_full.pt"))
    print("written", os.path.join(GOLD, "wan_1_3b_full.pt"), os.path.getsize(os.path.join(GOLD, "wan_1_3b_full.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
