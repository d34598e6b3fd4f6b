"""Full-depth / full-call GPU parity tests added at the end of round 1 and first run (and fixed) in round 2: Wan2.1-1.3B at all 30
layers against the fp64 reference fixture with the reference's own bf16 path as the noise floor, and the LTX i2v call with
image_cond_noise_scale > 0 against the latents of the reference's own __call__.
    python -m pytest tests/test_zz_full_depth_gpu.py -m gpu -x -q -s"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

DEV = "cuda"


def _sdpa_core(q, k, v, bias=None):
    """utils/attention.py:99-116 on the GPU: torch SDPA on [B, H, L, d]"""
    import torch.nn.functional as F
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                       attn_mask=None if bias is None else bias.to(q.dtype))
    return o.transpose(1, 2)


def test_wan_1_3b_full_depth_vs_reference_fixture(golden_dir, monkeypatch):
    """Wan2.1-1.3B at full width and depth (30 layers, 1.42 B seeded weights) against the fixture recorded from the unmodified
    reference in fp64 (oracle/gen_golden_wan_full.py).  With random weights, 30 layers and guide scale 5 the bf16 arithmetic the
    reference itself ships with is NOT within 2e-2 of the fp64 result after a few steps, so the contract is checked the way
    BASELINE.json words it — "match the reference PyTorch path ... in bf16": next to the fixture the reference's own bf16 path
    (the pinned restatement in bf16 on this GPU with torch SDPA) is run as the noise floor, and the drop-in has to be within
    2e-2 of the fp64 truth OR no further from it than 1.25x the reference's own bf16 error at that step."""
    from ltx_video_gpupoor_b200.wan.model import WanModel
    from ltx_video_gpupoor_b200.wan.posemb_layers import get_rotary_pos_embed
    from ltx_video_gpupoor_b200.wan.text2video import WanT2V
    from oracle import wan_oracle as W
    g = torch.load(os.path.join(golden_dir, "wan_1_3b_full.pt"), weights_only=False)
    cfg = g["cfg"]
    sd = W.make_wan_state_dict(cfg, seed=g["seed_weights"])
    m = WanModel(dim=cfg["dim"], ffn_dim=cfg["ffn_dim"], num_heads=cfg["num_heads"], num_layers=cfg["num_layers"])
    m.load_state_dict(sd)
    cos, sin = get_rotary_pos_embed(g["lat"].shape[1:])
    y = m([g["lat"].to(DEV), g["lat"].to(DEV)], t=g["t"].to(DEV), context=[g["ctx"].to(DEV), g["ctx0"].to(DEV)], freqs=(cos, sin))
    torch.cuda.synchronize()
    for a, b in zip(y, g["fwd"]):
        e = W.rel_l2(a.cpu(), b)
        print(f"wan-1.3B (30 layers) forward rel_l2 vs reference = {e:.3e}")
        assert e < 2e-2
    steps = []
    _, F_, H_, W_ = g["lat"].shape
    WanT2V(m).generate(width=W_ * 8, height=H_ * 8, frame_num=(F_ - 1) * 4 + 1, shift=g["shift"], sampling_steps=g["steps"],
                       guide_scale=g["guide"], cfg_star_switch=False, context=g["ctx"], context_null=g["ctx0"], noise=g["lat"],
                       _per_step_latents=steps)
    torch.cuda.synchronize()
    # the reference's own bf16 path on this GPU: same loop, oracle modules in bf16 + SDPA
    monkeypatch.setattr(W, "attention_core", _sdpa_core)
    sd16 = {k: v.to(DEV, torch.bfloat16) for k, v in sd.items()}
    sch = W.UniPC()
    sch.set_timesteps(g["steps"], g["shift"])
    lat, floor = g["lat"].to(DEV), []
    ctx, ctx0, cosd, sind = g["ctx"].to(DEV, torch.bfloat16), g["ctx0"].to(DEV, torch.bfloat16), cos.to(DEV), sin.to(DEV)
    with torch.no_grad():
        for t in sch.timesteps:
            c, u = W.wan_forward(sd16, cfg, [lat, lat], torch.stack([t]).to(DEV), [ctx, ctx0], cosd, sind)
            pred = (u.float() + g["guide"] * (c.float() - u.float())).cpu()
            lat = sch.step(pred.unsqueeze(0), lat.cpu().unsqueeze(0)).squeeze(0).to(DEV)
            floor.append(lat.cpu())
    for i, (a, f, b) in enumerate(zip(steps, floor, g["loop"])):
        e, ef = W.rel_l2(a.cpu(), b), W.rel_l2(f, b)
        print(f"wan-1.3B (30 layers) loop step {i}: latents rel_l2 vs the fp64 reference: drop-in {e:.3e}, the reference's own bf16 path {ef:.3e}")
        assert e < max(2e-2, 1.25 * ef)


def test_pipeline_i2v_image_cond_noise_vs_reference_fixture(golden_dir):
    """image_cond_noise_scale = 0.15 through the CUDA pipeline (pixel-space first-frame conditioning, fp32 prompt embeddings so that the
    per-step noise is drawn in fp32 from the same CPU generator as in the reference) vs the latents of the reference's own __call__."""
    from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import ConditioningItem, LTXVideoPipeline
    from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier
    from ltx_video_gpupoor_b200.ltx.transformer3d import Transformer3DModel
    from oracle import ltx_oracle as O
    g = torch.load(os.path.join(golden_dir, "ltx_pipeline_i2v.pt"), weights_only=False)
    m = g["meta"]
    tr = Transformer3DModel(num_layers=m["num_layers"])
    tr.load_state_dict(O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=m["num_layers"]))
    vae = CausalVideoAutoencoder()
    vsd = dict(O.make_vae_decoder_state_dict(seed=1))
    vsd.update(O.make_vae_encoder_state_dict(seed=2))
    vae.load_state_dict(vsd)
    pipe = LTXVideoPipeline(vae=vae, transformer=tr, scheduler=RectifiedFlowScheduler(), patchifier=SymmetricPatchifier(1))
    outs = {}
    for scale, key in ((0.0, "latents"), (0.15, "latents_cond_noise_0p15")):
        lat = pipe(height=m["H"], width=m["W"], num_frames=m["F"], frame_rate=m["fps"], prompt_embeds=g["pe"], prompt_attention_mask=g["pm"],
                   num_inference_steps=m["steps"], guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0,
                   generator=torch.Generator().manual_seed(g["noise_seed"]), output_type="latent", return_dict=False, is_video=True,
                   vae_per_channel_normalize=True, image_cond_noise_scale=scale,
                   conditioning_items=[ConditioningItem(media_item=g["image"], media_frame_number=0, conditioning_strength=1.0,
                                                        encode_noise=g["noise_e"])])[0]
        torch.cuda.synchronize()
        outs[key] = lat.float().cpu()
        e = O.rel_l2(outs[key], g[key])
        print(f"i2v from pixels, image_cond_noise_scale {scale}: latents rel_l2 vs the reference's own __call__ = {e:.3e}")
        assert e < 2e-2
    # the hard-conditioned first latent frame is never touched by the model update: what the noise leaves there is
    # 0.15 * (last draw) * t_last^2 exactly, so the DIFFERENCE of the two runs must be the reference's difference (same generator stream)
    d_ours = (outs["latents_cond_noise_0p15"] - outs["latents"])[:, :, :1]
    d_ref = (g["latents_cond_noise_0p15"] - g["latents"])[:, :, :1]
    e = O.rel_l2(d_ours, d_ref)
    print(f"first-frame noise term vs the reference's: rel_l2 = {e:.3e}")
    assert e < 2e-2
