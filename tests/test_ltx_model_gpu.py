"""GPU parity of the LTX drop-ins (Transformer3DModel.forward, LTXVideoPipeline.__call__,
CausalVideoAutoencoder decode) against (a) the fixtures recorded from the unmodified reference and
(b) the oracle run live on the same seeded inputs.  Tolerances from BASELINE.json north_star: per-step
latents <= 2e-2 relative L2 in bf16, decoded frames PSNR >= 40 dB, index work bit-exact."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder, vae_decode  # noqa: E402
from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import ConditioningItem, LTXVideoPipeline  # noqa: E402
from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler  # noqa: E402
from ltx_video_gpupoor_b200.ltx.skip_layer_strategy import SkipLayerStrategy  # noqa: E402
from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier  # noqa: E402
from ltx_video_gpupoor_b200.ltx.transformer3d import Transformer3DModel  # noqa: E402
from oracle import ltx_oracle as O  # noqa: E402

DEV = "cuda"
# bf16 weights + bf16 activations vs the fp32 reference: the reference's own bf16-vs-fp32 noise floor on the
# raw model output is 1.5e-2; the 2e-2 contract is on LATENTS.
TOL_MODEL_OUT = 2e-2
TOL_LATENTS = 2e-2


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _model(num_layers, seed=0):
    sd = O.make_transformer_state_dict(O.LTX_2B, seed=seed, num_layers=num_layers)
    m = Transformer3DModel(num_layers=num_layers)
    m.load_state_dict(sd)
    return m, sd


def test_rope_table_and_coords_bit_exact():
    m, _ = _model(1)
    p = SymmetricPatchifier(1)
    x = torch.randn(1, 128, 3, 4, 6, device=DEV)
    tok, coords = p.patchify(x)
    assert torch.equal(coords.cpu(), O.latent_coords(3, 4, 6, 1))
    assert torch.equal(tok.cpu(), O.patchify(x.cpu()))
    assert torch.equal(p.unpatchify(tok, 4, 6, 128), x)
    px = O.latent_to_pixel_coords(coords.cpu()).float()
    px[:, 0] /= 25.0
    cos, sin = m.precompute_freqs_cis(px)             # same device as the oracle (cpu) -> bit-exact
    c2, s2 = O.precompute_freqs_cis(px, 2048, 10000.0, (20, 2048, 2048), torch.bfloat16)
    assert torch.equal(cos, c2) and torch.equal(sin, s2)


def test_transformer_vs_reference_fixture(golden_dir):
    g = _load(golden_dir, "ltx_transformer.pt")
    meta = g["meta"]
    m, sd = _model(meta["num_layers"], meta["seed_weights"])
    f, h, w = meta["f"], meta["h"], meta["w"]
    coords = O.latent_to_pixel_coords(O.latent_coords(f, h, w, 1)).float()
    coords[:, 0] *= 1.0 / 25.0
    fc = m.precompute_freqs_cis(coords.to(DEV))
    for tag, strat in (("t2v", None), ("stg", SkipLayerStrategy.AttentionValues)):
        c = g[tag]
        skip = None if c["skip"] is None else m.create_skip_layer_mask(1, 3, 2, [1])
        if skip is not None:
            assert torch.equal(skip.float().cpu(), c["skip"].float())
        y = m(c["hidden"].to(DEV), freqs_cis=fc, encoder_hidden_states=c["enc"].to(DEV), timestep=c["timestep"].to(DEV),
              encoder_attention_mask=c["mask"].to(DEV), skip_layer_mask=skip, skip_layer_strategy=strat,
              latent_shape=(f, h, w), return_dict=False)[0]
        torch.cuda.synchronize()
        err = O.rel_l2(y.float().cpu(), c["out"])
        print(f"transformer[{tag}] rel_l2 vs reference fp32 = {err:.3e}")
        assert err < TOL_MODEL_OUT


def _pipe(num_layers):
    m, sd = _model(num_layers)
    vsd = O.make_vae_decoder_state_dict(seed=1)
    vae = CausalVideoAutoencoder()
    vae.load_state_dict(vsd)
    pipe = LTXVideoPipeline(vae=vae, transformer=m, scheduler=RectifiedFlowScheduler(), patchifier=SymmetricPatchifier(1))
    return pipe, sd, vsd


def test_pipeline_latents_vs_reference_fixture(golden_dir):
    g = _load(golden_dir, "ltx_pipeline.pt")
    meta = g["meta"]
    pipe, sd, _ = _pipe(meta["num_layers"])
    for tag in ("plain", "cfg_stg"):
        kw = dict(g[tag]["kw"])
        if "skip_layer_strategy" in kw:
            kw["skip_layer_strategy"] = SkipLayerStrategy[kw["skip_layer_strategy"]]
        per_step = []
        lat = pipe(height=meta["H"], width=meta["W"], num_frames=meta["F"], frame_rate=meta["fps"],
                   prompt_embeds=g["pe"], prompt_attention_mask=g["pm"], negative_prompt_embeds=g["ne"],
                   negative_prompt_attention_mask=g["nm"], num_inference_steps=meta["steps"],
                   generator=torch.Generator().manual_seed(g["noise_seed"]), output_type="latent", return_dict=False,
                   is_video=True, vae_per_channel_normalize=True, _per_step_latents=per_step, **kw)[0]
        torch.cuda.synchronize()
        # noise is drawn in prompt_embeds' dtype (fp32 here) exactly as the reference does -> same initial latents
        noise = torch.randn(1, 72, 128, generator=torch.Generator().manual_seed(g["noise_seed"]))
        ref_steps = []
        O.denoise_loop(sd, O.LTX_2B, noise, g["pe"], g["pm"], num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=meta["fps"],
                       num_steps=meta["steps"], neg_enc=g["ne"], neg_mask=g["nm"], guidance_scale=kw["guidance_scale"],
                       stg_scale=kw["stg_scale"], rescaling_scale=kw["rescaling_scale"],
                       skip_block_list=kw.get("skip_block_list"),
                       strategy=O.SKIP_ATTENTION_VALUES if "skip_block_list" in kw else None, per_step=ref_steps)
        for i, (a, b) in enumerate(zip(per_step, ref_steps)):
            err = O.rel_l2(a.cpu(), b)
            print(f"pipeline[{tag}] step {i}: latents rel_l2 vs oracle = {err:.3e}")
            assert err < TOL_LATENTS
        err = O.rel_l2(lat.float().cpu(), g[tag]["latents"])
        print(f"pipeline[{tag}] final latents rel_l2 vs reference fixture (fp32 noise) = {err:.3e}")
        assert err < TOL_LATENTS


def test_pipeline_shared_stg_prefix_is_bit_identical(golden_dir):
    """share_stg_prefix=True (the perturbed condition's rows are copied from the text condition's rows up to the first skipped
    block instead of being recomputed) must not change a single bit of any step's latents; every skip strategy."""
    g = _load(golden_dir, "ltx_pipeline.pt")
    meta = g["meta"]
    pipe, _, _ = _pipe(meta["num_layers"])
    kw = dict(g["cfg_stg"]["kw"])
    for strat in (SkipLayerStrategy.AttentionValues, SkipLayerStrategy.AttentionSkip, SkipLayerStrategy.TransformerBlock):
        kw["skip_layer_strategy"] = strat
        for blocks in ([meta["num_layers"] - 1], [0], [1, meta["num_layers"] - 1]):
            kw["skip_block_list"] = blocks
            runs = []
            for share in (False, True):
                per_step = []
                pipe(height=meta["H"], width=meta["W"], num_frames=meta["F"], frame_rate=meta["fps"], prompt_embeds=g["pe"],
                     prompt_attention_mask=g["pm"], negative_prompt_embeds=g["ne"], negative_prompt_attention_mask=g["nm"],
                     num_inference_steps=meta["steps"], generator=torch.Generator().manual_seed(g["noise_seed"]),
                     output_type="latent", return_dict=False, is_video=True, vae_per_channel_normalize=True,
                     _per_step_latents=per_step, share_stg_prefix=share, **kw)
                runs.append(per_step)
            assert len(runs[0]) == len(runs[1]) == meta["steps"]
            for a, b in zip(*runs):
                assert torch.equal(a, b), (strat, blocks)


def test_pipeline_i2v_conditioning_vs_oracle():
    pipe, sd, _ = _pipe(2)
    g = torch.Generator().manual_seed(3)
    pe = torch.randn(1, 16, 4096, generator=g)
    pm = torch.ones(1, 16)
    cond_lat = torch.randn(1, 128, 1, 4, 6, generator=g)
    per_step = []
    pipe(height=128, width=192, num_frames=17, frame_rate=25.0, prompt_embeds=pe, prompt_attention_mask=pm,
         num_inference_steps=3, guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0,
         generator=torch.Generator().manual_seed(5), output_type="latent", return_dict=False, is_video=True,
         conditioning_items=[ConditioningItem(latents=cond_lat, media_frame_number=0, conditioning_strength=1.0)],
         _per_step_latents=per_step)
    noise = torch.randn(1, 72, 128, generator=torch.Generator().manual_seed(5))
    init = O.unpatchify(noise, 3, 4, 6).clone()
    init[:, :, :1] = cond_lat
    cmask = torch.zeros(1, 3, 4, 6); cmask[:, :1] = 1.0
    ref_steps = []
    O.denoise_loop(sd, O.LTX_2B, O.patchify(init), pe, pm, num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=25.0,
                   num_steps=3, conditioning_mask=cmask.reshape(1, -1), per_step=ref_steps)
    for i, (a, b) in enumerate(zip(per_step, ref_steps)):
        err = O.rel_l2(a.cpu(), b)
        print(f"i2v step {i}: rel_l2 = {err:.3e}")
        assert err < TOL_LATENTS
    # hard-conditioned tokens are never touched
    assert torch.equal(per_step[-1][:, :24].cpu(), O.patchify(init)[:, :24])


def test_vae_decode_vs_reference_fixture(golden_dir):
    g = _load(golden_dir, "ltx_vae_decode.pt")
    vsd = O.make_vae_decoder_state_dict(seed=g["seed_weights"])
    vae = CausalVideoAutoencoder()
    vae.load_state_dict(vsd)
    y = vae_decode(g["z"].to(DEV), vae, is_video=True, vae_per_channel_normalize=True)
    torch.cuda.synchronize()
    assert tuple(y.shape) == (1, 3, 9, 96, 128)
    a = O.postprocess(y.float().cpu())
    b = O.postprocess(g["out"].float())
    ps = O.psnr(a, b)
    print(f"vae decode PSNR vs reference = {ps:.1f} dB, rel_l2 = {O.rel_l2(y.float().cpu(), g['out'].float()):.3e}")
    assert ps >= 40.0


def test_vae_decode_timestep_conditioned_vs_reference_fixture(golden_dir):
    g = _load(golden_dir, "ltx_vae_decode_timestep.pt")
    cfg = dict(O.LTX_VAE, timestep_conditioning=True)
    vsd = O.make_vae_decoder_state_dict(cfg, seed=g["seed_weights"])
    vae = CausalVideoAutoencoder(timestep_conditioning=True)
    vae.load_state_dict(vsd)
    assert vae.decoder.timestep_conditioning
    y = vae_decode(g["z"].to(DEV), vae, is_video=True, vae_per_channel_normalize=True, timestep=g["timestep"])
    torch.cuda.synchronize()
    ps = O.psnr(O.postprocess(y.float().cpu()), O.postprocess(g["out"].float()))
    print(f"timestep-conditioned vae decode PSNR vs reference = {ps:.1f} dB")
    assert tuple(y.shape) == (1, 3, 9, 96, 128) and ps >= 40.0
    y2 = vae_decode(g["z"].to(DEV), vae, is_video=True, vae_per_channel_normalize=True, timestep=torch.tensor([0.5]))
    assert O.psnr(O.postprocess(y2.float().cpu()), O.postprocess(g["out"].float())) < ps - 3      # the timestep is used
    # two videos per decode (one timestep per video, all equal, as the reference's callers pass them): every video = its own decode
    z2 = torch.cat([g["z"], g["z"].flip(-1)], dim=0).to(DEV)
    yb = vae_decode(z2, vae, is_video=True, vae_per_channel_normalize=True, timestep=g["timestep"].flatten()[:1].repeat(2))
    y1 = vae_decode(z2[1:], vae, is_video=True, vae_per_channel_normalize=True, timestep=g["timestep"])
    assert tuple(yb.shape) == (2, 3, 9, 96, 128) and torch.equal(yb[:1], y) and torch.equal(yb[1:], y1)
    with pytest.raises(NotImplementedError):
        vae_decode(z2, vae, is_video=True, vae_per_channel_normalize=True, timestep=torch.tensor([0.05, 0.5]))


# ------------------------------------------------------------------ VAE encode (i2v / v2v conditioning from pixels)
@pytest.mark.parametrize("st,shw,T,H,W", [(2, 2, 9, 16, 24), (1, 2, 3, 7, 10), (2, 1, 5, 6, 9)])
def test_conv3d_strided(st, shw, T, H, W):
    import torch.nn.functional as F
    from ltx_video_gpupoor_b200 import ops
    g = torch.Generator().manual_seed(3)
    Cin, Cout = 64, 128
    x = torch.randn(1, T, H, W, Cin, generator=g).bfloat16().cuda()
    w5 = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) * (27 * Cin) ** -0.5).bfloat16().cuda()
    b = torch.randn(Cout, generator=g).bfloat16().cuda()
    xr = x.float().permute(0, 4, 1, 2, 3)
    xr = torch.cat([xr[:, :, :1].repeat(1, 1, 2, 1, 1), xr], dim=2)
    ref = F.conv3d(xr, w5.float(), b.float(), stride=(st, shw, shw), padding=(0, 1, 1)).permute(0, 2, 3, 4, 1)
    out = ops.conv3d_strided(x, w5.permute(0, 2, 3, 4, 1).reshape(Cout, -1).contiguous(), b, st, shw)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    assert O.rel_l2(out.float().cpu(), ref.cpu()) < 6e-3


def test_vae_encode_vs_reference_fixture(golden_dir):
    from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder, vae_encode
    g = torch.load(os.path.join(golden_dir, "ltx_vae_encode.pt"), weights_only=False)
    sd = O.make_vae_encoder_state_dict(seed=g["seed_weights"])
    sd.update({k: v for k, v in O.make_vae_decoder_state_dict(seed=1).items() if k.startswith("decoder.")})
    vae = CausalVideoAutoencoder()
    vae.load_state_dict(sd)
    for tag in ("video", "image"):
        c = g[tag]
        post = vae.encode(c["x"].float().cuda()).latent_dist
        e = O.rel_l2(post.mean.cpu(), c["mean"])
        print(f"vae encode ({tag}): mean rel_l2 vs reference = {e:.3e}")
        assert e < 2e-2
        assert O.rel_l2(post.logvar[:, :1].cpu(), c["logvar"].clamp(-30, 20)) < 2e-2
        z = vae_encode(c["x"].float().cuda(), vae, vae_per_channel_normalize=True, noise=c["noise"])
        assert O.rel_l2(z.cpu(), c["z"]) < 2e-2


def test_pipeline_i2v_from_pixels_vs_oracle():
    """BASELINE config 3 in small: the conditioning image goes through the VAE ENCODER (vae_encode.py:22-91), the loop runs with
    the per-token timesteps / conditioning mask, and the result is compared with the oracle driven by the oracle's own encode."""
    pipe, sd, _ = _pipe(2)
    esd = O.make_vae_encoder_state_dict(seed=2)
    vsd = dict(O.make_vae_decoder_state_dict(seed=1))
    vsd.update({k: v for k, v in esd.items()})                      # encoder.* + the latent statistics of the encoder dict
    pipe.vae.load_state_dict(vsd)
    g = torch.Generator().manual_seed(4)
    pe, pm = torch.randn(1, 16, 4096, generator=g), torch.ones(1, 16)
    image = torch.rand(1, 3, 1, 128, 192, generator=g) * 2 - 1
    noise_e = torch.zeros(1, 128, 1, 4, 6)                             # latent_dist.mode(): deterministic
    per_step = []
    pipe(height=128, width=192, num_frames=17, frame_rate=25.0, prompt_embeds=pe, prompt_attention_mask=pm,
         num_inference_steps=3, guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0,
         generator=torch.Generator().manual_seed(5), output_type="latent", return_dict=False, is_video=True,
         vae_per_channel_normalize=True,
         conditioning_items=[ConditioningItem(media_item=image, media_frame_number=0, conditioning_strength=1.0, encode_noise=noise_e)],
         _per_step_latents=per_step)
    cond_lat = O.vae_encode(esd, image, noise=None)
    noise = torch.randn(1, 72, 128, generator=torch.Generator().manual_seed(5))
    init = O.unpatchify(noise, 3, 4, 6).clone()
    init[:, :, :1] = cond_lat
    cmask = torch.zeros(1, 3, 4, 6); cmask[:, :1] = 1.0
    ref_steps = []
    O.denoise_loop(sd, O.LTX_2B, O.patchify(init), pe, pm, num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=25.0,
                   num_steps=3, conditioning_mask=cmask.reshape(1, -1), per_step=ref_steps)
    for i, (a, b) in enumerate(zip(per_step, ref_steps)):
        err = O.rel_l2(a.cpu(), b)
        print(f"i2v-from-pixels step {i}: rel_l2 = {err:.3e}")
        assert err < TOL_LATENTS


# ------------------------------------------------------------------ multi-scale flow
def test_multiscale_kernels():
    import torch.nn.functional as F
    from ltx_video_gpupoor_b200 import ops
    g = torch.Generator().manual_seed(3)
    # GroupNorm(32) + SiLU (+ residual) on NDHWC
    for C in (256, 512):
        x = (torch.randn(2, 3, 5, 7, C, generator=g) * 1.5 + 0.3).bfloat16().to(DEV)
        r = torch.randn(2, 3, 5, 7, C, generator=g).bfloat16().to(DEV)
        ga = (1 + 0.1 * torch.randn(C, generator=g)).bfloat16().to(DEV); be = (0.1 * torch.randn(C, generator=g)).bfloat16().to(DEV)
        xc = x.float().permute(0, 4, 1, 2, 3)
        ref = F.group_norm(xc, 32, ga.float(), be.float(), eps=1e-5)
        y = ops.groupnorm_silu(x, ga, be)
        assert O.rel_l2(y.float().permute(0, 4, 1, 2, 3).cpu(), F.silu(ref).cpu()) < 6e-3
        y = ops.groupnorm_silu(x, ga, be, residual=r)
        assert O.rel_l2(y.float().permute(0, 4, 1, 2, 3).cpu(), F.silu(ref + r.float().permute(0, 4, 1, 2, 3)).cpu()) < 6e-3
        y = ops.groupnorm_silu(x, ga, be, silu=False)
        assert O.rel_l2(y.float().permute(0, 4, 1, 2, 3).cpu(), ref.cpu()) < 6e-3
    # centred zero-padded 3x3x3 convolution and per-frame 3x3 convolution
    x = torch.randn(1, 4, 6, 9, 128, generator=g).bfloat16().to(DEV)
    w = (torch.randn(256, 128, 3, 3, 3, generator=g) / 50).bfloat16().to(DEV)
    b = torch.randn(256, generator=g).bfloat16().to(DEV)
    y = ops.conv_taps(x, w.permute(0, 2, 3, 4, 1).reshape(256, -1).contiguous(), b, 3, 3, centered=True)
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.float(), b.float(), padding=1)
    assert O.rel_l2(y.float().permute(0, 4, 1, 2, 3).cpu(), ref.cpu()) < 6e-3
    w2 = (torch.randn(256, 128, 3, 3, generator=g) / 30).bfloat16().to(DEV)
    y = ops.conv_taps(x, w2.permute(0, 2, 3, 1).reshape(256, -1).contiguous(), b, 1, 3, centered=True)
    ref = F.conv2d(x[0].float().permute(0, 3, 1, 2), w2.float(), b.float(), padding=1)
    assert O.rel_l2(y[0].float().permute(0, 3, 1, 2).cpu(), ref.cpu()) < 6e-3
    # AdaIN, latent re-normalisation, bilinear resize (fp32 kernels)
    a = (torch.randn(1, 128, 3, 8, 12, generator=g) * 0.7 - 0.2).to(DEV); r = (torch.randn(1, 128, 3, 4, 6, generator=g) * 1.9 + 0.4).to(DEV)
    for f in (1.0, 0.25):
        assert O.rel_l2(ops.adain(a, r, f).cpu(), O.adain_filter_latent(a.cpu(), r.cpu(), f)) < 1e-5
    z = torch.randn(2, 3, 4, 5, 128, generator=g).bfloat16().to(DEV)
    sd_, mu_ = (0.5 + torch.rand(128, generator=g)).to(DEV), torch.randn(128, generator=g).to(DEV)
    ref = (z.float().permute(0, 4, 1, 2, 3) - mu_.view(1, -1, 1, 1, 1)) / sd_.view(1, -1, 1, 1, 1)
    assert torch.allclose(ops.latent_from_ndhwc(z, sd_, mu_), ref, rtol=1e-6, atol=1e-6)
    v = torch.rand(1, 3, 5, 48, 80, generator=g).to(DEV)
    for (hh, ww) in ((40, 64), (96, 160), (50, 70)):
        assert torch.allclose(ops.bilinear_resize(v, hh, ww).cpu(), O.multiscale_resize(v.cpu(), hh, ww), rtol=1e-5, atol=1e-5)


def test_latent_upsampler_vs_reference_fixture(golden_dir):
    from ltx_video_gpupoor_b200.ltx.latent_upsampler import LatentUpsampler, adain_filter_latent
    u = _load(golden_dir, "ltx_multiscale.pt")["upsampler"]
    up = LatentUpsampler(in_channels=128, mid_channels=u["mid"], num_blocks_per_stage=u["nb"], dims=3)
    up.load_state_dict(O.make_latent_upsampler_state_dict(128, u["mid"], u["nb"], seed=u["seed"]))
    y = up(u["z"].to(DEV))
    torch.cuda.synchronize()
    assert tuple(y.shape) == (1, 128, 3, 8, 12) and y.dtype == torch.float32
    e = O.rel_l2(y.cpu(), u["out"])
    print(f"latent upsampler rel_l2 vs reference = {e:.3e}")
    assert e < 2e-2
    assert O.rel_l2(adain_filter_latent(u["out"].to(DEV), u["ref_lat"].to(DEV)).cpu(), u["adain"]) < 1e-5


def test_multiscale_pipeline_vs_reference_fixture(golden_dir):
    from ltx_video_gpupoor_b200.ltx.latent_upsampler import LatentUpsampler
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import LTXMultiScalePipeline
    g = _load(golden_dir, "ltx_multiscale.pt")
    u, p = g["upsampler"], g["pipeline"]
    m = p["meta"]
    pipe, sd, vsd = _pipe(m["num_layers"])
    up = LatentUpsampler(in_channels=128, mid_channels=u["mid"], num_blocks_per_stage=u["nb"], dims=3)
    up.load_state_dict(O.make_latent_upsampler_state_dict(128, u["mid"], u["nb"], seed=u["seed"]))
    multi = LTXMultiScalePipeline(pipe, up)
    common = dict(downscale_factor=m["downscale_factor"], first_pass=dict(p["first_pass"]), second_pass=dict(p["second_pass"]),
                  height=m["H"], width=m["W"], num_frames=m["F"], frame_rate=m["fps"], prompt_embeds=p["pe"],
                  prompt_attention_mask=p["pm"], negative_prompt_embeds=p["ne"], negative_prompt_attention_mask=p["nm"],
                  num_inference_steps1=m["steps"], num_inference_steps2=m["steps"],
                  skip_layer_strategy=SkipLayerStrategy.AttentionValues, VAE_tile_size=(0, 0), is_video=True,
                  vae_per_channel_normalize=True)
    lat = multi(**common, output_type="latent", generator=torch.Generator().manual_seed(m["noise_seed"]))
    torch.cuda.synchronize()
    e = O.rel_l2(lat.float().cpu(), p["latents"])
    print(f"multi-scale final latents rel_l2 vs reference fixture = {e:.3e}")
    assert tuple(lat.shape) == (1, 128, 3, 6, 10) and e < TOL_LATENTS
    img = multi(**common, output_type="pt", generator=torch.Generator().manual_seed(m["noise_seed"]))
    torch.cuda.synchronize()
    assert tuple(img.shape) == (1, 3, m["F"], m["H"], m["W"])
    ps = O.psnr(img.float().cpu()[:, :, ::4, ::8, ::8], p["frames_sub"].float())
    print(f"multi-scale decoded + resized frames PSNR vs reference = {ps:.1f} dB")
    assert ps >= 40.0


def test_from_pretrained_single_file(tmp_path, golden_dir):
    """transformer3d.py:313-325 / causal_video_autoencoder.py:104-114 / latent_upsampler.py:183-199: one .safetensors holding
    `model.diffusion_model.*` and `vae.*` tensors with the configs in its metadata -> identical to loading the state dicts directly."""
    import json
    from safetensors.torch import save_file
    from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import LTX_VAE_CONFIG
    from ltx_video_gpupoor_b200.ltx.latent_upsampler import LatentUpsampler
    from ltx_video_gpupoor_b200.ltx.transformer3d import LTX_2B_CONFIG
    m0, sd = _model(1)
    vsd = O.make_vae_decoder_state_dict(seed=1)
    one = {"model.diffusion_model." + k: v.contiguous() for k, v in sd.items()}
    one.update({"vae." + k: v.contiguous() for k, v in vsd.items() if k.startswith("decoder.")})
    one["vae.per_channel_statistics.std-of-means"] = vsd["std_of_means"]
    one["vae.per_channel_statistics.mean-of-means"] = vsd["mean_of_means"]
    f = tmp_path / "ltxv-test.safetensors"
    save_file(one, str(f), metadata={"config": json.dumps({"transformer": dict(LTX_2B_CONFIG, num_layers=1), "vae": dict(LTX_VAE_CONFIG)})})
    m1 = Transformer3DModel.from_pretrained(f)
    assert m1.num_layers == 1 and all(torch.equal(m0.w[k], m1.w[k]) for k in m0.w)
    vae = CausalVideoAutoencoder.from_pretrained(f)
    g = _load(golden_dir, "ltx_vae_decode.pt")
    y = vae_decode(g["z"].to(DEV), vae, is_video=True, vae_per_channel_normalize=True)
    assert O.psnr(O.postprocess(y.float().cpu()), O.postprocess(g["out"].float())) >= 40.0
    u = _load(golden_dir, "ltx_multiscale.pt")["upsampler"]
    usd = O.make_latent_upsampler_state_dict(128, u["mid"], u["nb"], seed=u["seed"])
    uf = tmp_path / "up.safetensors"
    save_file({k: v.contiguous() for k, v in usd.items()}, str(uf), metadata={"config": json.dumps(
        {"_class_name": "LatentUpsampler", "in_channels": 128, "mid_channels": u["mid"], "num_blocks_per_stage": u["nb"], "dims": 3,
         "spatial_upsample": True, "temporal_upsample": False})})
    up = LatentUpsampler.from_pretrained(uf)
    assert O.rel_l2(up(u["z"].to(DEV)).cpu(), u["out"]) < 2e-2


def test_rf_stochastic_step_vs_reference_formula():
    """rf.py:370-373: x0 = x - t*v; prev = (1 - t_next) * x0 + t_next * noise — scheduler.step and the fused pipeline step."""
    s = RectifiedFlowScheduler()
    s.set_timesteps(5, samples_shape=(1, 128, 3, 4, 6), device=DEV)
    g = torch.Generator().manual_seed(9)
    x, v, nz = [torch.randn(1, 72, 128, generator=g) for _ in range(3)]
    ts = s.timesteps_host
    for i in (0, 2, 4):
        t, tn = float(ts[i]), float(ts[i + 1]) if i + 1 < len(ts) else 0.0
        out = s.step(v.to(DEV), ts[i], x.to(DEV), return_dict=False, stochastic_sampling=True, noise=nz.to(DEV))[0]
        ref = (1 - tn) * (x - t * v) + tn * nz                # fp32 predictions stay fp32 (the reference's tensor arithmetic)
        assert out.dtype == torch.float32 and O.rel_l2(out.cpu(), ref) < 1e-5
        det = s.step(v.to(DEV), ts[i], x.to(DEV), return_dict=False)[0]
        assert O.rel_l2(det.cpu(), x - (t - tn) * v) < 1e-5
        # bf16 model output and bf16 sample: the result comes back in the promoted dtype, as `sample - dt * model_output` would
        det16 = s.step(v.to(DEV, torch.bfloat16), ts[i], x.to(DEV, torch.bfloat16), return_dict=False)[0]
        assert det16.dtype == torch.bfloat16
        assert O.rel_l2(det16.float().cpu(), x.bfloat16().float() - (t - tn) * v.bfloat16().float()) < 8e-3
    # per-token [B, N] timesteps (rf.py:361-367): every token looks up its own next-lower schedule entry
    tok = torch.rand(1, 72, generator=g)
    tok[0, :3] = torch.tensor([float(ts[1]), float(ts[-1]), 1e-7])          # exactly on the schedule / on its last entry / below everything
    out = s.step(v.to(DEV), tok.to(DEV), x.to(DEV), return_dict=False)[0]
    assert out.dtype == torch.float32 and torch.equal(out.cpu(), O.rf_step(v, tok, x, ts))
    out = s.step(v.to(DEV), tok.to(DEV), x.to(DEV), return_dict=False, stochastic_sampling=True, noise=nz.to(DEV))[0]
    assert O.rel_l2(out.cpu(), O.rf_step_stochastic(v, tok, x, ts, nz)) < 1e-6      # the oracle forms dt another way (itself 1e-6 from the reference);
                                                                                    # bit-exactness is checked against the reference's own output below
    # through the pipeline: runs, is reproducible for a seeded generator, and differs from the deterministic sampler
    pipe, sd, _ = _pipe(1)
    kw = dict(height=128, width=192, num_frames=17, frame_rate=25.0, prompt_embeds=torch.randn(1, 16, 4096, generator=g),
              prompt_attention_mask=torch.ones(1, 16), num_inference_steps=3, guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0,
              output_type="latent", return_dict=False, is_video=True, vae_per_channel_normalize=True)
    a = pipe(**kw, generator=torch.Generator().manual_seed(1), stochastic_sampling=True)[0]
    b = pipe(**kw, generator=torch.Generator().manual_seed(1), stochastic_sampling=True)[0]
    c = pipe(**kw, generator=torch.Generator().manual_seed(1))[0]
    assert torch.equal(a, b) and O.rel_l2(a.cpu(), c.cpu()) > 1e-2


def test_rf_step_per_token_timesteps_vs_reference_fixture(golden_dir):
    """`RectifiedFlowScheduler.step` with per-token [B, N] timesteps against what the unmodified reference scheduler returned
    (`rf_scheduler.pt`: deterministic and stochastic, the noise the reference drew replayed): fp32, bit for bit."""
    g = _load(golden_dir, "rf_scheduler.pt")
    for case in g.values():
        s = RectifiedFlowScheduler()
        s.set_timesteps(case["steps"], samples_shape=case["shape"], device=DEV)
        assert torch.equal(s.timesteps_host, case["timesteps"])
        out = s.step(case["v"].to(DEV), case["tt"].to(DEV), case["x"].to(DEV), return_dict=False)[0]
        assert torch.equal(out.cpu(), case["stepped"])
        st = case["stochastic"][-1]
        assert st["t"].shape == case["tt"].shape
        out = s.step(case["v"].to(DEV), st["t"].to(DEV), case["x"].to(DEV), return_dict=False, stochastic_sampling=True, noise=st["noise"].to(DEV))[0]
        assert torch.equal(out.cpu(), st["out"])


def test_reference_selfcheck_encoder_first_frame_causality_gpu():
    """The reference's in-file check (causal_video_autoencoder.py:1384-1389) on the CUDA path: encoding the first frame alone gives the
    first latent frame of the whole video (bf16 tolerance instead of atol 1e-6)."""
    from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder
    sd = O.make_vae_encoder_state_dict(seed=2)
    sd.update({k: v for k, v in O.make_vae_decoder_state_dict(seed=1).items() if k.startswith("decoder.")})
    vae = CausalVideoAutoencoder()
    vae.load_state_dict(sd)
    video = (torch.rand(1, 3, 17, 64, 64, generator=torch.Generator().manual_seed(3)) * 2 - 1).cuda()
    lat_v = vae.encode(video).latent_dist.mode()
    lat_i = vae.encode(video[:, :, :1]).latent_dist.mode()
    assert tuple(lat_v.shape) == (1, 128, 3, 2, 2) and tuple(lat_i.shape) == (1, 128, 1, 2, 2)
    assert O.rel_l2(lat_i.float().cpu(), lat_v[:, :, :1].float().cpu()) < 1e-2


def test_pipeline_callbacks_and_interrupt(golden_dir):
    """Drop-in boundary: progress `callback(step_idx, latents|None, is_start, ...)` (pipeline_ltx_video.py:1100-1101,
    1243-1248), `callback_on_step_end(pipe, i, t, {})` (:1255-1256), and `ltxv_model._interrupt` polled between transformer blocks
    (transformer3d.py:468-469 -> forward returns [None]; the pipeline returns None)."""
    from types import SimpleNamespace
    g = _load(golden_dir, "ltx_pipeline.pt")
    meta = g["meta"]
    pipe, _, _ = _pipe(meta["num_layers"])
    seen, ends = [], []
    kw = dict(height=meta["H"], width=meta["W"], num_frames=meta["F"], frame_rate=meta["fps"], prompt_embeds=g["pe"], prompt_attention_mask=g["pm"],
              num_inference_steps=3, guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0, generator=torch.Generator().manual_seed(1),
              output_type="latent", return_dict=False, is_video=True, vae_per_channel_normalize=True)
    model = SimpleNamespace(_interrupt=False)
    out = pipe(callback=lambda i, lat, is_start, **k: seen.append((i, None if lat is None else tuple(lat.shape), is_start, dict(k))),
               callback_on_step_end=lambda p, i, t, d: ends.append((i, float(t))), ltxv_model=model, pass_no=2, **kw)
    assert out is not None
    assert seen[0] == (-1, None, True, {"override_num_inference_steps": 3, "pass_no": 2})
    assert [s[0] for s in seen[1:]] == [0, 1, 2] and all(s[1] == (128, 3, 4, 6) and s[2] is False and s[3] == {"pass_no": 2} for s in seen[1:])
    assert [e[0] for e in ends] == [0, 1, 2] and ends[0][1] > ends[1][1] > ends[2][1]
    # interrupt raised by the UI thread after the first step: the next forward returns [None], the call returns None
    def stop_after_first(i, lat, is_start, **k):
        if i == 0:
            model._interrupt = True
    assert pipe(callback=stop_after_first, ltxv_model=model, **kw) is None
    tr = pipe.transformer
    y = tr(torch.zeros(1, 72, 128), freqs_cis=tr.precompute_freqs_cis(torch.zeros(1, 3, 72, device="cuda")), encoder_hidden_states=g["pe"],
           timestep=torch.ones(1, 1), encoder_attention_mask=g["pm"], latent_shape=(3, 4, 6), ltxv_model=model, return_dict=False)
    assert y == [None]


def test_transformer_13b_geometry_head_dim_128():
    """The LTX-Video 13B (0.9.7) transformer uses 32 heads x 128 = 4096 channels, 48 layers; what the reference app runs for the
    multi-scale flow.  Same code path at a reduced width (8 heads x 128 = 1024, 2 layers, 1024 % 6 == 4096 % 6 == 4 for the RoPE
    padding columns): the d = 128 attention kernel and the wider-head RoPE/norm kernels inside Transformer3DModel vs the fp32 oracle."""
    cfg = dict(O.LTX_2B, num_layers=2, num_attention_heads=8, attention_head_dim=128, cross_attention_dim=1024)
    sd = O.make_transformer_state_dict(cfg, seed=3)
    m = Transformer3DModel(num_layers=2, num_attention_heads=8, attention_head_dim=128, cross_attention_dim=1024)
    m.load_state_dict(sd)
    f, h, w = 3, 6, 8
    g = torch.Generator().manual_seed(9)
    hidden = torch.randn(2, f * h * w, 128, generator=g)
    enc = torch.randn(2, 24, 4096, generator=g)
    mask = torch.ones(2, 24)
    mask[1, 17:] = 0
    t = torch.tensor([[0.9], [0.4]])
    coords = O.latent_to_pixel_coords(O.latent_coords(f, h, w, 1)).float()
    coords[:, 0] *= 1.0 / 25.0
    cos_sin = O.precompute_freqs_cis(coords, 1024, cfg["rope_theta"], cfg["rope_max_pos"])
    ref = O.transformer_forward(sd, cfg, hidden, cos_sin, enc, t, mask, latent_shape=(f, h, w))
    fc = m.precompute_freqs_cis(coords.to(DEV))
    assert O.rel_l2(fc[0].float().cpu(), cos_sin[0]) < 4e-3
    y = m(hidden.to(DEV), freqs_cis=fc, encoder_hidden_states=enc.to(DEV), timestep=t.to(DEV), encoder_attention_mask=mask.to(DEV),
          latent_shape=(f, h, w), return_dict=False)[0]
    torch.cuda.synchronize()
    err = O.rel_l2(y.float().cpu(), ref)
    print(f"transformer[13B geometry, d=128] rel_l2 vs fp32 oracle = {err:.3e}")
    assert err < TOL_MODEL_OUT


def test_pipeline_media_items_vid2vid():
    """`media_items` (pipeline_ltx_video.py:682-710): the video is encoded with the VAE, noised to the first kept timestep and denoised from
    there — identical to handing the same encoded latents in through `latents=` (the reference treats both alike, :688-710)."""
    from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import vae_encode
    pipe, _, _ = _pipe(2)
    vsd = O.make_vae_encoder_state_dict(seed=2)
    vsd.update({k: v for k, v in O.make_vae_decoder_state_dict(seed=1).items() if k.startswith("decoder.")})
    pipe.vae.load_state_dict(vsd)
    g = torch.Generator().manual_seed(4)
    video = (torch.rand(1, 3, 17, 128, 192, generator=g) * 2 - 1)
    pe, pm = torch.randn(1, 16, 4096, generator=g), torch.ones(1, 16)
    kw = dict(height=128, width=192, num_frames=17, frame_rate=25.0, prompt_embeds=pe, prompt_attention_mask=pm, num_inference_steps=4,
              skip_initial_inference_steps=2, guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0, output_type="latent", return_dict=False,
              is_video=True, vae_per_channel_normalize=True)
    torch.manual_seed(11)
    a = pipe(media_items=video, generator=torch.Generator().manual_seed(7), **kw)[0]
    torch.manual_seed(11)
    lat = vae_encode(video.cuda(), pipe.vae, vae_per_channel_normalize=True)
    b = pipe(latents=lat, generator=torch.Generator().manual_seed(7), **kw)[0]
    assert tuple(a.shape) == (1, 128, 3, 4, 6) and torch.equal(a, b)
    with pytest.raises(AssertionError):
        pipe(media_items=video, latents=lat, generator=torch.Generator().manual_seed(7), **kw)
    with pytest.raises(AssertionError):                       # no skipped steps: the first timestep is 1.0 and the media would be replaced by noise
        pipe(media_items=video, generator=torch.Generator().manual_seed(7), **{**kw, "skip_initial_inference_steps": 0})


def test_pipeline_keyframe_conditioning_vs_oracle():
    """A conditioning frame in the MIDDLE of the video (media_frame_number = 16, pipeline_ltx_video.py:1449-1503): its noised latent is
    prepended as 24 extra tokens with their own pixel coordinates and per-token timesteps, and dropped again before unpatchify; the loop
    is compared with the oracle driven by the same tokens / coordinates / mask."""
    pipe, sd, _ = _pipe(2)
    g = torch.Generator().manual_seed(3)
    pe, pm = torch.randn(1, 16, 4096, generator=g), torch.ones(1, 16)
    key = torch.randn(1, 128, 1, 4, 6, generator=g)
    items = lambda: [ConditioningItem(latents=key.clone(), media_frame_number=16, conditioning_strength=1.0)]
    per_step = []
    out = pipe(height=128, width=192, num_frames=33, frame_rate=25.0, prompt_embeds=pe, prompt_attention_mask=pm, num_inference_steps=3,
               guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0, generator=torch.Generator().manual_seed(5), output_type="latent",
               return_dict=False, is_video=True, conditioning_items=items(), _per_step_latents=per_step)[0]
    assert tuple(out.shape) == (1, 128, 5, 4, 6) and per_step[0].shape[1] == 120 + 24
    # the same RNG stream on the host: initial noise, then the keyframe noise inside prepare_conditioning
    gen = torch.Generator().manual_seed(5)
    init = O.unpatchify(torch.randn(1, 120, 128, generator=gen), 5, 4, 6)
    tok, px, cm, extra = pipe.prepare_conditioning(items(), init.clone(), 33, 128, 192, vae_per_channel_normalize=True, generator=gen)
    assert extra == 24
    ref_steps = []
    O.denoise_loop(sd, O.LTX_2B, tok.float(), pe, pm, num_frames_lat=5, lat_h=4, lat_w=6, frame_rate=25.0, num_steps=3,
                   conditioning_mask=cm, per_step=ref_steps, pixel_coords=px)
    for i, (a, b) in enumerate(zip(per_step, ref_steps)):
        err = O.rel_l2(a.cpu(), b)
        print(f"keyframe conditioning step {i}: rel_l2 = {err:.3e}")
        assert err < TOL_LATENTS
    assert torch.equal(per_step[-1][:, :24].cpu(), tok[:, :24].float())          # the hard-conditioned keyframe tokens are never touched


def test_ltx_transformer_without_prompt_mask_equals_all_ones_mask():
    """`encoder_attention_mask=None` must give the all-ones-mask result (the unbiased cross-attention path differs only
    in which exponentials come from the polynomial), so that the pipeline may drop an all-ones mask."""
    from ltx_video_gpupoor_b200.ltx.transformer3d import Transformer3DModel
    from oracle import ltx_oracle as O
    f, h, w, Lc = 3, 4, 6, 32
    m = Transformer3DModel(num_layers=2)
    m.load_state_dict(O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=2))
    g = torch.Generator().manual_seed(11)
    hidden = torch.randn(2, f * h * w, 128, generator=g)
    enc = torch.randn(2, Lc, 4096, generator=g)
    coords = O.latent_to_pixel_coords(O.latent_coords(f, h, w, 1)).float()
    coords[:, 0] /= 25.0
    fc = m.precompute_freqs_cis(coords.to(DEV))
    kw = dict(freqs_cis=fc, encoder_hidden_states=enc.to(DEV), timestep=torch.full((2, 1), 0.7, device=DEV), latent_shape=(f, h, w),
              return_dict=False)
    y1 = m(hidden.to(DEV), encoder_attention_mask=torch.ones(2, Lc, device=DEV), **kw)[0]
    y0 = m(hidden.to(DEV), encoder_attention_mask=None, **kw)[0]
    torch.cuda.synchronize()
    e = O.rel_l2(y0.float().cpu(), y1.float().cpu())
    print(f"transformer without mask vs all-ones mask: rel_l2 = {e:.3e}")
    assert e < 5e-3


def test_pipeline_resizes_conditioning_media_like_the_reference():
    """pipeline_ltx_video.py:748-760, 1402-1404: a conditioning image of another size is bilinearly resized (align_corners=False) to the
    target size before the VAE encoder — what the multi-scale first pass relies on.  Passing the image at 192x288 must give the latents
    of passing F.interpolate(image, (128, 192)) directly."""
    import torch.nn.functional as F
    from ltx_video_gpupoor_b200.ltx.causal_video_autoencoder import CausalVideoAutoencoder
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import ConditioningItem, LTXVideoPipeline
    from ltx_video_gpupoor_b200.ltx.rf import RectifiedFlowScheduler
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier
    from ltx_video_gpupoor_b200.ltx.transformer3d import Transformer3DModel
    from oracle import ltx_oracle as O
    m = Transformer3DModel(num_layers=2)
    m.load_state_dict(O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=2))
    vae = CausalVideoAutoencoder()
    vsd = dict(O.make_vae_decoder_state_dict(seed=1))
    vsd.update(O.make_vae_encoder_state_dict(seed=2))
    vae.load_state_dict(vsd)
    pipe = LTXVideoPipeline(vae=vae, transformer=m, scheduler=RectifiedFlowScheduler(), patchifier=SymmetricPatchifier(1))
    g = torch.Generator().manual_seed(4)
    pe, pm = torch.randn(1, 16, 4096, generator=g), torch.ones(1, 16)
    big = torch.rand(1, 3, 1, 192, 288, generator=g) * 2 - 1
    small = F.interpolate(big[:, :, 0], size=(128, 192), mode="bilinear", align_corners=False)[:, :, None]
    noise_e = torch.zeros(1, 128, 1, 4, 6)
    outs = []
    for image in (big, small):
        outs.append(pipe(height=128, width=192, num_frames=17, frame_rate=25.0, prompt_embeds=pe, prompt_attention_mask=pm,
                         num_inference_steps=2, guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0,
                         generator=torch.Generator().manual_seed(5), output_type="latent", return_dict=False, is_video=True,
                         vae_per_channel_normalize=True,
                         conditioning_items=[ConditioningItem(media_item=image, media_frame_number=0, conditioning_strength=1.0,
                                                              encode_noise=noise_e)])[0].float().cpu())
    e = O.rel_l2(outs[0], outs[1])
    print(f"resized-in-pipeline vs pre-resized conditioning image: latents rel_l2 = {e:.3e}")
    assert e < 2e-3


def test_pipeline_two_videos_per_call():
    """pipeline_ltx_video.py:632-710, 1034-1051: more than one video per call (batch = prompts x num_images_per_prompt).  The cond batch is
    cond-major ([negative x b, positive x b]); guidance statistics are per sample.  Sample 0 of the 2-video call is bit-identical to the
    1-video call with the same seed (the generator's stream starts with its noise), and both samples match the oracle loop run on the batch."""
    pipe, sd, _ = _pipe(2)
    g = torch.Generator().manual_seed(21)
    pe, ne = torch.randn(2, 24, 4096, generator=g), torch.randn(2, 24, 4096, generator=g)
    pm = torch.ones(2, 24)
    pm[1, 17:] = 0
    kw = dict(height=128, width=192, num_frames=17, frame_rate=25.0, num_inference_steps=3, guidance_scale=3.0, stg_scale=0.0, rescaling_scale=1.0,
              output_type="latent", return_dict=False, is_video=True, vae_per_channel_normalize=True)
    two = pipe(prompt_embeds=pe, prompt_attention_mask=pm, negative_prompt_embeds=ne, negative_prompt_attention_mask=pm,
               generator=torch.Generator().manual_seed(5), **kw)[0]
    one = pipe(prompt_embeds=pe[:1], prompt_attention_mask=pm[:1], negative_prompt_embeds=ne[:1], negative_prompt_attention_mask=pm[:1],
               generator=torch.Generator().manual_seed(5), **kw)[0]
    torch.cuda.synchronize()
    assert tuple(two.shape) == (2, 128, 3, 4, 6) and tuple(one.shape) == (1, 128, 3, 4, 6)
    assert torch.equal(two[:1], one)
    noise = torch.randn(2, 72, 128, generator=torch.Generator().manual_seed(5))
    for j in range(2):                # the oracle loop handles one video at a time: sample j = its slice of the noise and its prompts
        ref = O.denoise_loop(sd, O.LTX_2B, noise[j:j + 1], pe[j:j + 1], pm[j:j + 1], num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=25.0,
                             num_steps=3, neg_enc=ne[j:j + 1], neg_mask=pm[j:j + 1], guidance_scale=3.0)
        e = O.rel_l2(two[j:j + 1].float().cpu(), O.unpatchify(ref, 3, 4, 6))
        print(f"2 videos per call, sample {j}: final latents rel_l2 vs the oracle loop = {e:.3e}")
        assert e < TOL_LATENTS
    # num_images_per_prompt repeats every prompt (:849-872)
    rep = pipe(prompt_embeds=pe[:1], prompt_attention_mask=pm[:1], negative_prompt_embeds=ne[:1], negative_prompt_attention_mask=pm[:1],
               num_images_per_prompt=2, generator=torch.Generator().manual_seed(5), **kw)[0]
    assert tuple(rep.shape) == (2, 128, 3, 4, 6) and torch.equal(rep[:1], one)


def test_pipeline_two_videos_with_conditioning():
    """More than one video per call WITH conditioning (pipeline_ltx_video.py:632-710, 1344-1548 on a batch): per-sample conditioning masks and
    per-token timesteps, a first-frame latent per video (and one shared by both).  Sample 0 equals the one-video call bit for bit; both
    samples match the oracle loop run per video; hard-conditioned tokens are never touched."""
    pipe, sd, _ = _pipe(2)
    g = torch.Generator().manual_seed(23)
    pe, pm = torch.randn(2, 16, 4096, generator=g), torch.ones(2, 16)
    cond = torch.randn(2, 128, 1, 4, 6, generator=g)
    kw = dict(height=128, width=192, num_frames=17, frame_rate=25.0, num_inference_steps=3, guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0,
              output_type="latent", return_dict=False, is_video=True, vae_per_channel_normalize=True)
    steps2 = []
    two = pipe(prompt_embeds=pe, prompt_attention_mask=pm, generator=torch.Generator().manual_seed(5), _per_step_latents=steps2,
               conditioning_items=[ConditioningItem(latents=cond, media_frame_number=0, conditioning_strength=1.0)], **kw)[0]
    one = pipe(prompt_embeds=pe[:1], prompt_attention_mask=pm[:1], generator=torch.Generator().manual_seed(5),
               conditioning_items=[ConditioningItem(latents=cond[:1], media_frame_number=0, conditioning_strength=1.0)], **kw)[0]
    torch.cuda.synchronize()
    assert tuple(two.shape) == (2, 128, 3, 4, 6) and torch.equal(two[:1], one)
    noise = torch.randn(2, 72, 128, generator=torch.Generator().manual_seed(5))
    cmask = torch.zeros(1, 3, 4, 6); cmask[:, :1] = 1.0
    for j in range(2):
        init = O.unpatchify(noise[j:j + 1], 3, 4, 6).clone()
        init[:, :, :1] = cond[j:j + 1]
        ref = O.denoise_loop(sd, O.LTX_2B, O.patchify(init), pe[j:j + 1], pm[j:j + 1], num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=25.0,
                             num_steps=3, conditioning_mask=cmask.reshape(1, -1))
        e = O.rel_l2(two[j:j + 1].float().cpu(), O.unpatchify(ref, 3, 4, 6))
        print(f"2 videos with conditioning, sample {j}: final latents rel_l2 vs the oracle loop = {e:.3e}")
        assert e < TOL_LATENTS
        assert torch.equal(steps2[-1][j, :24].cpu(), O.patchify(init)[0, :24])           # the conditioned first frame is never touched
    # one conditioning item shared by both videos, and the stochastic sampler on a batch (runs, reproducible, differs from the Euler step)
    shared = pipe(prompt_embeds=pe, prompt_attention_mask=pm, generator=torch.Generator().manual_seed(5),
                  conditioning_items=[ConditioningItem(latents=cond[:1], media_frame_number=0, conditioning_strength=1.0)], **kw)[0]
    assert torch.equal(shared[:1], one) and torch.equal(shared[1, :, 0], shared[0, :, 0])
    a = pipe(prompt_embeds=pe, prompt_attention_mask=pm, generator=torch.Generator().manual_seed(6), stochastic_sampling=True, **kw)[0]
    b = pipe(prompt_embeds=pe, prompt_attention_mask=pm, generator=torch.Generator().manual_seed(6), stochastic_sampling=True, **kw)[0]
    c = pipe(prompt_embeds=pe, prompt_attention_mask=pm, generator=torch.Generator().manual_seed(6), **kw)[0]
    assert torch.equal(a, b) and O.rel_l2(a.cpu(), c.cpu()) > 1e-2 and O.rel_l2(a[:1].cpu(), a[1:].cpu()) > 1e-2


def test_transformer_and_pipeline_mixed_precision(golden_dir):
    """`mixed=True` (transformer3d.py:343,439-442 / `mixed_precision=True`, pipeline_ltx_video.py:1061,1152-1177): fp32 residual stream and
    fp32 AdaLN tables, bf16 Linear inputs.  Against the fp32 reference fixture the mixed forward must be inside the contract and not
    further away than the plain bf16 forward; the pipeline call with mixed_precision=True must match the oracle loop."""
    g = _load(golden_dir, "ltx_transformer.pt")
    meta = g["meta"]
    m, sd = _model(meta["num_layers"], meta["seed_weights"])
    f, h, w = meta["f"], meta["h"], meta["w"]
    coords = O.latent_to_pixel_coords(O.latent_coords(f, h, w, 1)).float()
    coords[:, 0] *= 1.0 / 25.0
    fc = m.precompute_freqs_cis(coords.to(DEV))
    for tag, strat in (("t2v", None), ("stg", SkipLayerStrategy.AttentionValues)):
        c = g[tag]
        skip = None if c["skip"] is None else m.create_skip_layer_mask(1, 3, 2, [1])
        kw = dict(freqs_cis=fc, encoder_hidden_states=c["enc"].to(DEV), timestep=c["timestep"].to(DEV), encoder_attention_mask=c["mask"].to(DEV),
                  skip_layer_mask=skip, skip_layer_strategy=strat, latent_shape=(f, h, w), return_dict=False)
        y16 = m(c["hidden"].to(DEV), **kw)[0]
        y32 = m(c["hidden"].to(DEV), mixed=True, **kw)[0]
        torch.cuda.synchronize()
        e16, e32 = O.rel_l2(y16.float().cpu(), c["out"]), O.rel_l2(y32.float().cpu(), c["out"])
        print(f"transformer[{tag}] rel_l2 vs reference fp32: bf16 stream {e16:.3e}, mixed (fp32 stream) {e32:.3e}")
        assert y32.dtype == torch.bfloat16 and e32 < TOL_MODEL_OUT and e32 < 1.05 * e16
    # SkipLayerStrategy.TransformerBlock under mixed precision: the fp32 blend of the residual stream (attention.py:355-362) is a row copy
    c = g["stg"]
    kw = dict(freqs_cis=fc, encoder_hidden_states=c["enc"].to(DEV), timestep=c["timestep"].to(DEV), encoder_attention_mask=c["mask"].to(DEV),
              skip_layer_mask=m.create_skip_layer_mask(1, 3, 2, [1]), skip_layer_strategy=SkipLayerStrategy.TransformerBlock,
              latent_shape=(f, h, w), return_dict=False)
    yb = m(c["hidden"].to(DEV), mixed=True, **kw)[0]
    cos, sin = O.precompute_freqs_cis(coords, 2048, O.LTX_2B["rope_theta"], O.LTX_2B["rope_max_pos"])
    ref = O.transformer_forward(sd, O.LTX_2B, c["hidden"], (cos, sin), c["enc"], c["timestep"], c["mask"], c["skip"], O.SKIP_TRANSFORMER_BLOCK, (f, h, w))
    e = O.rel_l2(yb.float().cpu(), ref)
    print(f"transformer[stg, TransformerBlock] mixed rel_l2 vs the fp32 oracle: {e:.3e}")
    assert e < TOL_MODEL_OUT
    assert not torch.equal(yb, m(c["hidden"].to(DEV), mixed=True, **dict(kw, skip_layer_strategy=SkipLayerStrategy.AttentionValues))[0])
    pipe, sd, _ = _pipe(2)
    gen = torch.Generator().manual_seed(31)
    pe, pm = torch.randn(1, 24, 4096, generator=gen), torch.ones(1, 24)
    steps = []
    pipe(height=128, width=192, num_frames=17, frame_rate=25.0, prompt_embeds=pe, prompt_attention_mask=pm, num_inference_steps=3,
         guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0, generator=torch.Generator().manual_seed(8), output_type="latent",
         return_dict=False, is_video=True, mixed_precision=True, _per_step_latents=steps)
    noise = torch.randn(1, 72, 128, generator=torch.Generator().manual_seed(8))
    ref = []
    O.denoise_loop(sd, O.LTX_2B, noise, pe, pm, num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=25.0, num_steps=3, per_step=ref)
    for i, (a, b) in enumerate(zip(steps, ref)):
        e = O.rel_l2(a.cpu(), b)
        print(f"pipeline mixed_precision step {i}: latents rel_l2 vs the fp32 oracle = {e:.3e}")
        assert e < TOL_LATENTS


def test_pipeline_right_padded_prompt_mask_uses_key_lengths():
    """A right-padded prompt mask (what the tokenizer produces) is turned into per-sample key lengths once per call (no bias pass, padded
    key blocks skipped); a mask with a hole in it keeps the additive bias.  Both must give the latents of the oracle loop, and the
    key-length path must agree with the bias path on the same mask."""
    pipe, sd, _ = _pipe(2)
    g = torch.Generator().manual_seed(17)
    pe, ne = torch.randn(1, 200, 4096, generator=g), torch.randn(1, 200, 4096, generator=g)
    pm, nm = torch.ones(1, 200), torch.ones(1, 200)
    pm[:, 141:] = 0
    nm[:, 9:] = 0
    kw = dict(height=128, width=192, num_frames=17, frame_rate=25.0, prompt_embeds=pe, negative_prompt_embeds=ne, num_inference_steps=3,
              guidance_scale=3.0, stg_scale=0.0, rescaling_scale=1.0, output_type="latent", return_dict=False, is_video=True)
    a = pipe(prompt_attention_mask=pm, negative_prompt_attention_mask=nm, generator=torch.Generator().manual_seed(4), **kw)[0]
    assert pipe._state.key_lens_b is not None and pipe._state.mask_b is not None and pipe._state.key_lens_b.tolist() == [9, 141]
    saved = pipe._state.key_lens_b
    st = pipe(prompt_attention_mask=pm, negative_prompt_attention_mask=nm, generator=torch.Generator().manual_seed(4), _prepare_only=True, **kw)
    st.key_lens_b = None                                     # the same call on the additive-bias path
    for i in range(3):
        pipe.denoise_step(st, i)
    b = pipe.patchifier.unpatchify(st.lat32.view(1, st.N, st.C), 4, 6, 128)
    torch.cuda.synchronize()
    e_ab = O.rel_l2(a.float().cpu(), b.float().cpu())
    noise = torch.randn(1, 72, 128, generator=torch.Generator().manual_seed(4))
    ref = O.denoise_loop(sd, O.LTX_2B, noise, pe, pm, num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=25.0, num_steps=3, neg_enc=ne, neg_mask=nm,
                         guidance_scale=3.0)
    e = O.rel_l2(a.float().cpu(), O.unpatchify(ref, 3, 4, 6))
    print(f"right-padded masks as key lengths {saved.tolist()}: final latents rel_l2 vs oracle {e:.3e}; vs the additive-bias path {e_ab:.3e}")
    assert e < TOL_LATENTS and e_ab < 5e-3
    pm2 = pm.clone()
    pm2[:, 20] = 0                                           # a hole: not a key length
    pipe(prompt_attention_mask=pm2, negative_prompt_attention_mask=nm, generator=torch.Generator().manual_seed(4), **kw)
    assert pipe._state.key_lens_b is None and pipe._state.mask_b is not None


# ------------------------------------------------------------------ guidance-condition parallelism (ltx/distributed/cond_parallel.py)
def _cond_parallel_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pipe, _, _ = _pipe(2)
        g = torch.Generator().manual_seed(31)
        pe, ne = torch.randn(1, 24, 4096, generator=g), torch.randn(1, 24, 4096, generator=g)
        pm = torch.ones(1, 24)
        pm[0, 19:] = 0                                                     # right-padded prompt: the key-length path, sliced per rank
        kw = dict(height=128, width=192, num_frames=17, frame_rate=25.0, num_inference_steps=3, guidance_scale=3.0, stg_scale=1.0,
                  rescaling_scale=0.7, skip_block_list=[1], skip_layer_strategy=SkipLayerStrategy.AttentionValues, output_type="latent",
                  return_dict=False, is_video=True, vae_per_channel_normalize=True, prompt_embeds=pe, prompt_attention_mask=pm,
                  negative_prompt_embeds=ne, negative_prompt_attention_mask=pm)
        steps_cp, steps_one = [], []
        cp = pipe(generator=torch.Generator().manual_seed(5), cond_parallel_group=dist.group.WORLD, _per_step_latents=steps_cp, **kw)[0]
        one = pipe(generator=torch.Generator().manual_seed(5), _per_step_latents=steps_one, **kw)[0]       # all three conditions on this GPU
        torch.cuda.synchronize()
        # an interrupt seen by ONE rank ends the call on every rank at the same step boundary (OR over the group)
        from types import SimpleNamespace
        stopped = pipe(generator=torch.Generator().manual_seed(5), cond_parallel_group=dist.group.WORLD,
                       ltxv_model=SimpleNamespace(_interrupt=(rank == world - 1)), **kw)
        running = pipe(generator=torch.Generator().manual_seed(5), cond_parallel_group=dist.group.WORLD,
                       ltxv_model=SimpleNamespace(_interrupt=False), **kw)[0]
        torch.cuda.synchronize()
        ret[rank] = dict(equal=bool(torch.equal(cp, one)) and all(torch.equal(a, b) for a, b in zip(steps_cp, steps_one)),
                         steps=len(steps_cp), latents=cp.float().cpu(), stopped=stopped is None, running=bool(torch.equal(running, one)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pipeline_cond_parallel(world):
    """The [uncond, text, perturbed] rows of a denoise step on different GPUs (2 ranks: 2 + 1 conditions, 3 ranks: one each), predictions
    exchanged once per step, guidance + scheduler replicated: every rank ends every step with the latents of the single-GPU call, bit for bit
    (a row's GEMM / attention / norm results do not depend on the other rows of the batch)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world if world != 3 else 4})")
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_cond_parallel_worker, args=(world, 30500 + os.getpid() % 300, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r]["steps"] == 3 and ret[r]["equal"], f"rank {r}: cond-parallel latents differ from the single-GPU call"
        assert ret[r]["stopped"] and ret[r]["running"], f"rank {r}: interrupt handling under cond-parallel"
        assert torch.equal(ret[r]["latents"], ret[0]["latents"])
