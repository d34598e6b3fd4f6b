"""CPU: the oracle restatement must reproduce the fixtures that oracle/gen_golden.py
recorded from the UNMODIFIED reference modules (fp32, CPU)."""
import os

import pytest
import torch

from oracle import ltx_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_patchifier_coords_bit_exact(golden_dir):
    g = _load(golden_dir, "patchifier.pt")
    c = O.latent_coords(3, 4, 6, 2)
    assert torch.equal(c, g["coords"])
    assert torch.equal(O.latent_to_pixel_coords(c), g["px"])
    assert torch.equal(O.latent_to_pixel_coords(c, causal_fix=True), g["px_fix"])
    x = torch.randn(2, 5, 3, 4, 6)
    assert torch.equal(O.unpatchify(O.patchify(x), 3, 4, 6), x)


def test_rf_scheduler_bit_exact(golden_dir):
    g = _load(golden_dir, "rf_scheduler.pt")
    for case in g.values():
        ts = O.rf_timesteps(case["steps"], case["shape"])
        assert torch.equal(ts, case["timesteps"])
        assert torch.equal(O.rf_step(case["v"], case["tt"], case["x"], ts), case["stepped"])
        for st in case["stochastic"]:                       # stochastic sampler (rf.py:369-373) with the noise the reference drew
            assert O.rel_l2(O.rf_step_stochastic(case["v"], st["t"], case["x"], ts, st["noise"]), st["out"]) < 1e-6


def test_known_timesteps():
    # latent (1,128,2,8,8), 4 steps
    ts = O.rf_timesteps(4, (1, 128, 2, 8, 8))
    assert torch.allclose(ts, torch.tensor([1.0, 0.7793, 0.4914, 0.1]), atol=1e-4)


def test_transformer_matches_reference(golden_dir):
    g = _load(golden_dir, "ltx_transformer.pt")
    m = g["meta"]
    cfg = O.LTX_2B
    sd = O.make_transformer_state_dict(cfg, seed=m["seed_weights"], num_layers=m["num_layers"])
    f, h, w = m["f"], m["h"], m["w"]
    coords = O.latent_to_pixel_coords(O.latent_coords(f, h, w, 1)).float()
    coords[:, 0] *= 1.0 / 25.0
    cos, sin = O.precompute_freqs_cis(coords, 2048, cfg["rope_theta"], cfg["rope_max_pos"])
    assert torch.equal(cos[0, 5], g["rope_cos_row5"]) and torch.equal(sin[0, 5], g["rope_sin_row5"])
    for tag, strat in (("t2v", None), ("stg", O.SKIP_ATTENTION_VALUES)):
        c = g[tag]
        y = O.transformer_forward(sd, cfg, c["hidden"], (cos, sin), c["enc"], c["timestep"], c["mask"],
                                  c["skip"], strat, (f, h, w))
        assert O.rel_l2(y, c["out"]) < 2e-5


def test_vae_decode_matches_reference(golden_dir):
    g = _load(golden_dir, "ltx_vae_decode.pt")
    sd = O.make_vae_decoder_state_dict(seed=g["seed_weights"])
    y = O.vae_decode(sd, g["z"])
    assert y.shape == (1, 3, 9, 96, 128)
    assert O.rel_l2(y, g["out"].float()) < 2e-3      # fixture stored as fp16


def test_pipeline_matches_reference(golden_dir):
    g = _load(golden_dir, "ltx_pipeline.pt")
    m = g["meta"]
    cfg = O.LTX_2B
    sd = O.make_transformer_state_dict(cfg, seed=0, num_layers=m["num_layers"])
    for tag in ("plain", "cfg_stg"):
        kw = g[tag]["kw"]
        noise = torch.randn(1, 72, 128, generator=torch.Generator().manual_seed(g["noise_seed"]))
        lat = O.denoise_loop(sd, cfg, noise, g["pe"], g["pm"], num_frames_lat=3, lat_h=4, lat_w=6,
                             frame_rate=m["fps"], num_steps=m["steps"], neg_enc=g["ne"], neg_mask=g["nm"],
                             guidance_scale=kw["guidance_scale"], stg_scale=kw["stg_scale"],
                             rescaling_scale=kw["rescaling_scale"], skip_block_list=kw.get("skip_block_list"),
                             strategy=O.SKIP_ATTENTION_VALUES if "skip_block_list" in kw else None)
        assert O.rel_l2(O.unpatchify(lat, 3, 4, 6), g[tag]["latents"]) < 5e-5


def test_wan_i2v_oracle_matches_reference(golden_dir):
    """wan_oracle with the i2v branch (y channels, img_emb MLPProj, WanI2VCrossAttention) vs the fixture recorded from
    the unmodified reference WanModel(model_type='i2v') in fp64 (oracle/gen_golden_wan.py:main_i2v)."""
    from oracle import wan_oracle as W
    g = _load(golden_dir, "wan_i2v.pt")
    cfg = g["cfg"]
    sd = W.make_wan_state_dict(cfg, seed=1)
    cos, sin = W.rope_tables(g["lat"].shape[1:])
    y = W.wan_forward(sd, cfg, [g["lat"], g["lat"]], g["t"], [g["ctx"], g["ctx0"]], cos, sin, clip_fea=g["clip"], y=g["y"])
    for a, b in zip(y, g["fwd"]):
        assert W.rel_l2(a, b.float()) < 5e-5          # fp32 oracle vs fp64 reference
    steps = []
    W.t2v_denoise(sd, cfg, g["lat"], g["ctx"], g["ctx0"], steps=4, shift=5.0, guide_scale=5.0, per_step=steps, clip_fea=g["clip"], y=g["y"])
    for a, b in zip(steps, g["loop"]):
        assert W.rel_l2(a, b) < 2e-4


def test_wan_vae_decode_oracle_matches_reference(golden_dir):
    """The one-pass Wan VAE decode oracle vs the fixture recorded from the unmodified reference's frame-by-frame streaming decode
    with its CACHE_T feature cache (oracle/gen_golden_wan_vae.py: identical in fp64, 1.7e-6 in fp32)."""
    from oracle import wan_vae_oracle as V
    g = _load(golden_dir, "wan_vae_decode.pt")
    sd = V.make_wan_vae_decoder_state_dict(g["cfg"], seed=g["seed_weights"])
    y = V.wan_vae_decode(sd, g["z"], g["cfg"], torch.tensor(V.WAN_VAE_MEAN), torch.tensor(V.WAN_VAE_STD))
    assert y.shape == (3, 13, 48, 80)
    assert O.rel_l2(y, g["out"].float()) < 2e-3      # fixture stored as fp16


def test_wan_vae_encode_oracle_matches_reference(golden_dir):
    """The one-pass Wan VAE encode oracle vs the fixture recorded from the unmodified reference's chunked (1, 4, 4, ...) streaming
    encode (oracle/gen_golden_wan_vae.py: identical in fp64, 1.7e-7 in fp32); a single image and odd spatial sizes too."""
    from oracle import wan_vae_oracle as V
    g = _load(golden_dir, "wan_vae_encode.pt")
    sd = V.make_wan_vae_encoder_state_dict(g["cfg"], seed=g["seed_weights"])
    mean, std = torch.tensor(V.WAN_VAE_MEAN), torch.tensor(V.WAN_VAE_STD)
    mu = V.wan_vae_encode(sd, g["video"], g["cfg"], mean, std)
    assert mu.shape == (16, 3, 6, 10)
    assert O.rel_l2(mu, g["mu"]) < 1e-5
    # causality of the one-pass form: the first latent frame depends on the first video frame only (the reference's first chunk)
    mu1 = V.wan_vae_encode(sd, g["video"][:, :1], g["cfg"], mean, std)
    assert O.rel_l2(mu1, mu[:, :1]) < 1e-5
    with pytest.raises(AssertionError):
        V.wan_vae_encode(sd, g["video"][:, :7], g["cfg"], mean, std)


def test_vae_encode_oracle_matches_reference(golden_dir):
    """Encoder.forward + latent_dist.sample() + normalize_latents (oracle) vs the fixture recorded from the unmodified reference
    (oracle/gen_golden.py:case_vae_encode — bit-identical there)."""
    g = _load(golden_dir, "ltx_vae_encode.pt")
    sd = O.make_vae_encoder_state_dict(seed=g["seed_weights"])
    for tag in ("video", "image"):
        c = g[tag]
        mean, logvar = O.vae_encode_moments(sd, c["x"].float())
        assert O.rel_l2(mean, c["mean"]) < 2e-3                         # input stored as fp16
        z = O.vae_encode(sd, c["x"].float(), noise=c["noise"])
        assert O.rel_l2(z, c["z"]) < 2e-3


def test_wan_step_skipping_oracle_matches_reference(golden_dir):
    """Skip-layer guidance (joint pass) and TeaCache control flow of WanModel.forward (model.py:1029-1101): oracle vs the fixture
    recorded from the unmodified reference in fp64 (oracle/gen_golden_wan.py:main_skip)."""
    from oracle import wan_oracle as W
    g = _load(golden_dir, "wan_skip.pt")
    cfg = g["cfg"]
    sd = W.make_wan_state_dict(cfg, seed=0)
    cos, sin = W.rope_tables(g["lat"].shape[1:])
    y = W.wan_forward(sd, cfg, [g["lat"], g["lat"]], g["t"], [g["ctx"], g["ctx0"]], cos, sin, slg_layers=g["slg_layers"])
    for a, b in zip(y, g["slg_fwd"]):
        assert W.rel_l2(a, b.float()) < 5e-5
    plain = W.wan_forward(sd, cfg, [g["lat"], g["lat"]], g["t"], [g["ctx"], g["ctx0"]], cos, sin)
    assert W.rel_l2(y[0], plain[0]) < 1e-6 and W.rel_l2(y[1], plain[1]) > 1e-3      # only the unconditional sequence skips
    tcg = g["teacache"]
    tc = W.teacache_state(tcg["coefficients"], tcg["rel_l1_thresh"], tcg["start_step"], tcg["steps"])
    so = W.UniPC(); so.set_timesteps(tcg["steps"], 5.0)
    lat = g["lat"].clone()
    for i, tt in enumerate(so.timesteps):
        c, u = W.wan_forward(sd, cfg, [lat, lat], torch.stack([tt]), [g["ctx"], g["ctx0"]], cos, sin, teacache=tc, current_step=i)
        lat = so.step((u + 5.0 * (c - u)).unsqueeze(0), lat.unsqueeze(0)).squeeze(0)
        assert W.rel_l2(lat, g["teacache_loop"][i]) < 2e-4
    assert tc["skipped"] == tcg["skipped"] and 0 < tc["skipped"] < tcg["steps"] - 2


def test_multiscale_oracle_matches_reference(golden_dir):
    """LatentUpsampler, adain_filter_latent and the LTXMultiScalePipeline flow (oracle) vs the fixture recorded from the unmodified
    reference (oracle/gen_golden.py:case_multiscale — upsampler / AdaIN bit-identical, final latents 5e-7)."""
    g = _load(golden_dir, "ltx_multiscale.pt")
    u = g["upsampler"]
    usd = O.make_latent_upsampler_state_dict(128, u["mid"], u["nb"], seed=u["seed"])
    y = O.latent_upsampler_forward(usd, u["z"])
    assert y.shape == (1, 128, 3, 8, 12) and O.rel_l2(y, u["out"]) < 1e-5
    assert O.rel_l2(O.adain_filter_latent(u["out"], u["ref_lat"]), u["adain"]) < 1e-6
    p = g["pipeline"]
    m = p["meta"]
    sd = O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=m["num_layers"])
    vsd = O.make_vae_decoder_state_dict(seed=1)
    # second pass from the recorded first-pass latents (the first pass is the plain pipeline, covered above)
    gen = torch.Generator().manual_seed(m["noise_seed"])
    f, h1, w1, h2, w2 = 3, 3, 5, 6, 10
    torch.randn(1, f * h1 * w1, 128, generator=gen)                              # the first pass' draw
    upl = O.adain_filter_latent(O.upsample_latents(usd, vsd, p["first_latents"]), p["first_latents"])
    n2 = O.unpatchify(torch.randn(1, f * h2 * w2, 128, generator=gen), f, h2, w2)
    ts2 = O.rf_timesteps(m["steps"], (1, 128, f, h2, w2))[p["second_pass"]["skip_initial_inference_steps"]:]
    sp = p["second_pass"]
    l2 = O.denoise_loop(sd, O.LTX_2B, O.patchify(O.multiscale_second_pass_init(n2, upl, float(ts2[0]))), p["pe"], p["pm"],
                        num_frames_lat=f, lat_h=h2, lat_w=w2, frame_rate=m["fps"], num_steps=m["steps"], timesteps=ts2,
                        neg_enc=p["ne"], neg_mask=p["nm"], guidance_scale=sp["guidance_scale"], stg_scale=sp["stg_scale"],
                        rescaling_scale=sp["rescaling_scale"], guidance_timesteps=sp["guidance_timesteps"],
                        skip_block_list=sp["skip_block_list"], strategy=O.SKIP_ATTENTION_VALUES)
    assert O.rel_l2(O.unpatchify(l2, f, h2, w2), p["latents"]) < 5e-5


def test_vae_decode_timestep_conditioned_oracle_matches_reference(golden_dir):
    """Timestep-conditioned decoder (causal_video_autoencoder.py:724-733,757-795,1207-1237) vs the fixture recorded from the
    unmodified reference (oracle/gen_golden.py:case_vae — bit-identical there)."""
    g = _load(golden_dir, "ltx_vae_decode_timestep.pt")
    cfg = dict(O.LTX_VAE, timestep_conditioning=True)
    sd = O.make_vae_decoder_state_dict(cfg, seed=g["seed_weights"])
    y = O.vae_decode(sd, g["z"], cfg, timestep=g["timestep"])
    assert y.shape == (1, 3, 9, 96, 128)
    assert O.rel_l2(y, g["out"].float()) < 2e-3      # fixture stored as fp16


def test_wan_dpmpp_oracle_matches_reference(golden_dir):
    """sample_solver='dpm++' (wan/utils/fm_solvers.py via text2video.py:423-432): oracle DPMpp vs the trajectories recorded from the
    unmodified FlowDPMSolverMultistepScheduler (oracle/gen_golden_wan.py:main_dpm)."""
    from oracle import wan_oracle as W
    g = _load(golden_dir, "wan_dpmpp.pt")
    for (steps, shift), c in g.items():
        o = W.DPMpp(); o.set_timesteps(steps, shift)
        assert torch.equal(o.timesteps, c["timesteps"]) and torch.equal(o.sigmas, c["sigmas"])
        if "x0" in c:
            x = c["x0"]
            for i in range(steps):
                x = o.step(c["v"][i], x)
                assert W.rel_l2(x, c["x"][i]) < 1e-5


# ---- the reference's own in-file checks, restated on the oracle -------------------------------------------------
def test_reference_selfcheck_vae_patchify_roundtrip():
    """causal_video_autoencoder.py:1341-1347 `test_vae_patchify_unpatchify`: unpatchify(patchify(x)) == x (here: spatial patch 4,
    the configuration the shipped VAE uses; `vae_unpatchify` / `vae_patchify` are each other's inverse)."""
    x = torch.randn(2, 3, 8, 64, 64, generator=torch.Generator().manual_seed(0))
    assert torch.equal(O.vae_unpatchify(O.vae_patchify(x, 4), 4), x)
    lat = torch.randn(2, 128, 3, 4, 6, generator=torch.Generator().manual_seed(1))
    assert torch.equal(O.unpatchify(O.patchify(lat), 3, 4, 6), lat)          # SymmetricPatchifier round trip (symmetric_patchifier.py:33-84)


def test_reference_selfcheck_encoder_first_frame_causality():
    """causal_video_autoencoder.py:1350-1400 `demo_video_autoencoder_forward_backward`: the latent of the first frame alone equals
    the first latent frame of the whole video (atol 1e-6 there, :1389) — for the LTX encoder and, same property, the Wan encoder."""
    sd = O.make_vae_encoder_state_dict(seed=2)
    video = torch.rand(1, 3, 9, 32, 32, generator=torch.Generator().manual_seed(3)) * 2 - 1
    mean_v, _ = O.vae_encode_moments(sd, video)
    mean_i, _ = O.vae_encode_moments(sd, video[:, :, :1])
    assert torch.allclose(mean_i, mean_v[:, :, :1], atol=1e-5)


def test_pay_attention_oracle_matches_reference(golden_dir):
    """Row a1 of the scope table: the attention entry point.  oracle/attention_oracle.py vs the fixture recorded from the unmodified
    `pay_attention` (utils/attention.py:161-398, sdpa path): plain, additive key mask, k_lens runs in a batch, q_lens/k_lens with the
    uninitialised tail, dtype contract; and the ownership convention (the caller's list is emptied)."""
    from oracle import attention_oracle as A
    g = _load(golden_dir, "pay_attention.pt")
    assert set(g) == {"plain", "mask", "k_lens_batch", "q_k_lens_single", "dtypes"}
    for name, c in g.items():
        lst = [c["q"].clone(), c["k"].clone(), c["v"].clone()]
        y = A.pay_attention(lst, **{k: v for k, v in c["kw"].items() if k != "softmax_scale"})
        assert lst == []
        assert y.dtype == c["out"].dtype == c["q"].dtype and y.shape == c["out"].shape
        assert O.rel_l2(y[:, : c["valid"]].float(), c["out"][:, : c["valid"]].float()) < (1e-2 if name == "dtypes" else 1e-5), name


def test_wan_vae_any_end_frame_oracle_matches_reference(golden_dir):
    """any_end_frame (wan/modules/vae.py:541-557, 597-601): the last (latent) frame is coded on its own, without the feature caches.  The
    one-pass oracle vs fixtures recorded from the unmodified reference's streaming encode / decode (identical in fp64)."""
    from oracle import wan_vae_oracle as V
    mean, std = torch.tensor(V.WAN_VAE_MEAN), torch.tensor(V.WAN_VAE_STD)
    g = _load(golden_dir, "wan_vae_encode.pt")
    sd = V.make_wan_vae_encoder_state_dict(g["cfg"], seed=g["seed_weights"])
    mu = V.wan_vae_encode(sd, g["video_end_frame"], g["cfg"], mean, std, any_end_frame=True)
    assert mu.shape == (16, 4, 6, 10) and O.rel_l2(mu, g["mu_end_frame"]) < 1e-5
    assert O.rel_l2(mu[:, :3], g["mu"]) < 1e-5                      # the prefix is the plain encode of the first 9 frames
    d = _load(golden_dir, "wan_vae_decode.pt")
    sdd = V.make_wan_vae_decoder_state_dict(d["cfg"], seed=d["seed_weights"])
    y = V.wan_vae_decode(sdd, d["z"], d["cfg"], mean, std, any_end_frame=True)
    assert y.shape == (3, 10, 48, 80) and O.rel_l2(y, d["out_end_frame"].float()) < 2e-3


def test_baseline_config0_full_depth_matches_reference(golden_dir):
    """BASELINE.json configs[0] exactly (LTX-2B, 28 layers, 256x256x9, 4 steps + VAE decode, fp32 on CPU): the oracle's
    denoise loop and decode vs the latents / frames the unmodified reference pipeline produced (oracle/gen_golden_config0.py)."""
    g = _load(golden_dir, "ltx_config0.pt")
    m = g["meta"]
    f, h, w = m["latent"]
    sd = O.make_transformer_state_dict(O.LTX_2B, seed=m["seed_weights"], num_layers=m["num_layers"])
    pe = torch.randn(1, 32, 4096, generator=torch.Generator().manual_seed(m["seed_prompt"]))
    noise = torch.randn(1, f * h * w, 128, generator=torch.Generator("cpu").manual_seed(m["seed_noise"]))
    per_step = []
    with torch.no_grad():
        lat = O.denoise_loop(sd, O.LTX_2B, noise, pe, torch.ones(1, 32), num_frames_lat=f, lat_h=h, lat_w=w,
                             frame_rate=m["fps"], num_steps=m["steps"], per_step=per_step)
        del sd
        assert O.rel_l2(O.unpatchify(lat, f, h, w), g["latents"]) < 2e-4
        for a, b in zip(per_step, g["per_step_oracle"]):
            assert O.rel_l2(a, b) < 2e-4
        img = O.postprocess(O.vae_decode(O.make_vae_decoder_state_dict(seed=m["seed_vae"]), O.unpatchify(lat, f, h, w)))
    assert tuple(img.shape) == (1, 3, m["F"], m["H"], m["W"])
    assert O.psnr(img, g["frames"].float()) > 60          # fixture frames are stored in fp16


def test_timestep_sinusoid_bit_exact(golden_dir):
    """The oracle's timestep sinusoid vs the reference's vendored get_timestep_embedding (ltx_video/models/transformers/embeddings.py:10-50),
    the one diffusers function on the path that the reference tree itself contains."""
    g = _load(golden_dir, "timestep_sinusoid.pt")
    assert torch.equal(O.timestep_sinusoid(g["t"], 256), g["emb"])


def test_pipeline_i2v_from_pixels_matches_reference_call(golden_dir):
    """The oracle's composition for pixel-space first-frame conditioning (vae_encode -> blend into the noise -> loop with the conditioning
    mask) vs the latents of the reference's own LTXVideoPipeline.__call__ with a ConditioningItem(media_item=...) (gen_golden.py:case_pipeline_i2v)."""
    g = _load(golden_dir, "ltx_pipeline_i2v.pt")
    m = g["meta"]
    sd = O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=m["num_layers"])
    esd = O.make_vae_encoder_state_dict(seed=2)
    with torch.no_grad():
        cond_lat = O.vae_encode(esd, g["image"], noise=g["noise_e"])
        init = O.unpatchify(torch.randn(1, 72, 128, generator=torch.Generator().manual_seed(g["noise_seed"])), 3, 4, 6).clone()
        init[:, :, :1] = cond_lat
        cmask = torch.zeros(1, 3, 4, 6)
        cmask[:, :1] = 1.0
        lat = O.denoise_loop(sd, O.LTX_2B, O.patchify(init), g["pe"], g["pm"], num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=m["fps"],
                             num_steps=m["steps"], conditioning_mask=cmask.reshape(1, -1))
    assert O.rel_l2(O.unpatchify(lat, 3, 4, 6), g["latents"]) < 5e-5


def test_keyframe_call_matches_reference_call(golden_dir):
    """A keyframe in the middle of the video through the reference's own __call__ (oracle/gen_golden_conditioning.py:keyframe_loop) vs the
    composition the GPU test uses as its checker: product prepare_conditioning (host torch code) -> oracle loop with the extra tokens'
    pixel coordinates and per-token timesteps -> extra tokens dropped."""
    from types import SimpleNamespace
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import ConditioningItem, LTXVideoPipeline
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier
    g = _load(golden_dir, "ltx_keyframe_loop.pt")
    m = g["meta"]
    sd = O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=m["num_layers"])
    pipe = LTXVideoPipeline.__new__(LTXVideoPipeline)
    pipe.vae = SimpleNamespace(spatial_downscale_factor=32, temporal_downscale_factor=8)
    pipe.patchifier = SymmetricPatchifier(1)
    pipe.transformer = SimpleNamespace(config=SimpleNamespace(causal_temporal_positioning=False))
    gen = torch.Generator().manual_seed(g["noise_seed"])
    init = O.unpatchify(torch.randn(1, 120, 128, generator=gen), 5, 4, 6)
    tok, px, cm, extra = pipe.prepare_conditioning([ConditioningItem(latents=g["key"].clone(), media_frame_number=m["frame"], conditioning_strength=1.0)],
                                                   init.clone(), m["F"], m["H"], m["W"], vae_per_channel_normalize=True, generator=gen)
    assert extra == 24
    with torch.no_grad():
        lat = O.denoise_loop(sd, O.LTX_2B, tok.float(), g["pe"], g["pm"], num_frames_lat=5, lat_h=4, lat_w=6, frame_rate=m["fps"],
                             num_steps=m["steps"], conditioning_mask=cm, pixel_coords=px)
    assert O.rel_l2(O.unpatchify(lat[:, extra:], 5, 4, 6), g["latents"]) < 5e-5


def test_pipeline_i2v_image_cond_noise_matches_reference_call(golden_dir):
    """image_cond_noise_scale = 0.15 (pipeline_ltx_video.py:606-629, 1105-1113): the hard-conditioned tokens are re-noised from the call's
    generator at the start of every step.  Oracle loop vs the reference's own __call__ (gen_golden.py:case_pipeline_i2v)."""
    g = _load(golden_dir, "ltx_pipeline_i2v.pt")
    m = g["meta"]
    sd = O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=m["num_layers"])
    with torch.no_grad():
        cond_lat = O.vae_encode(O.make_vae_encoder_state_dict(seed=2), g["image"], noise=g["noise_e"])
        gen = torch.Generator().manual_seed(g["noise_seed"])
        init = O.unpatchify(torch.randn(1, 72, 128, generator=gen), 3, 4, 6).clone()
        init[:, :, :1] = cond_lat
        cmask = torch.zeros(1, 3, 4, 6)
        cmask[:, :1] = 1.0
        lat = O.denoise_loop(sd, O.LTX_2B, O.patchify(init), g["pe"], g["pm"], num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=m["fps"],
                             num_steps=m["steps"], conditioning_mask=cmask.reshape(1, -1), image_cond_noise_scale=0.15, generator=gen)
    assert O.rel_l2(O.unpatchify(lat, 3, 4, 6), g["latents_cond_noise_0p15"]) < 5e-5
    assert O.rel_l2(g["latents_cond_noise_0p15"], g["latents"]) > 1e-4          # 7.5e-4: the noise is scaled by t^2 and only the last draw survives
