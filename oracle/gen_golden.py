"""Pin the oracle against the UNMODIFIED reference and write tests/golden/*.pt.

Run in the build container only (needs /root/reference):
    python oracle/gen_golden.py
For every case it (1) builds the reference module from /root/reference under the
diffusers/mmgp shims, (2) loads the oracle's seeded state_dict into it, (3) runs both on
the same seeded inputs in fp32 on CPU, (4) asserts they agree to float round-off, and
(5) stores inputs' seeds + the REFERENCE outputs as small fixtures.  tests/test_oracle_golden.py
re-checks the oracle against these fixtures on any machine (no reference needed).
"""
import contextlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))

import load_reference  # noqa: E402

load_reference.install()

from oracle import ltx_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
torch.manual_seed(0)
torch.set_grad_enabled(False)


class _NoInterrupt:
    _interrupt = False


def _check(name, a, b, tol=2e-5):
    err = O.rel_l2(a, b)
    print(f"  {name}: rel_l2(oracle, reference) = {err:.3e}  max|d|={float((a - b).abs().max()):.3e}")
    assert err < tol, f"{name}: oracle disagrees with the reference ({err})"
    return err


def build_ref_transformer(num_layers, sd):
    from ltx_video.models.transformers.transformer3d import Transformer3DModel
    from ltx_video.utils.diffusers_config_mapping import OURS_TRANSFORMER_CONFIG
    cfg = dict(OURS_TRANSFORMER_CONFIG)
    cfg["num_layers"] = num_layers
    m = Transformer3DModel.from_config(cfg)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.eval()


def case_transformer():
    """Transformer3DModel.forward: t2v timestep, per-token (i2v) timestep, STG skip mask."""
    from ltx_video.utils.skip_layer_strategy import SkipLayerStrategy
    cfg = O.LTX_2B
    L = 2
    sd = O.make_transformer_state_dict(cfg, seed=0, num_layers=L)
    ref = build_ref_transformer(L, sd)
    f, h, w = 3, 4, 6
    N = f * h * w
    g = torch.Generator().manual_seed(11)
    out = {"meta": dict(num_layers=L, f=f, h=h, w=w, seed_weights=0, seed_inputs=11)}
    for B, tag in ((1, "t2v"), (3, "stg")):
        hidden = torch.randn(B, N, 128, generator=g)
        enc = torch.randn(B, 24, 4096, generator=g)
        mask = torch.ones(B, 24)
        mask[:, 17:] = 0
        coords = O.latent_to_pixel_coords(O.latent_coords(f, h, w, 1)).float()
        coords[:, 0] *= 1.0 / 25.0
        cos_ref, sin_ref = ref.precompute_freqs_cis(coords)
        cos, sin = O.precompute_freqs_cis(coords, 2048, cfg["rope_theta"], cfg["rope_max_pos"])
        assert torch.equal(cos, cos_ref) and torch.equal(sin, sin_ref), "rope table not bit-exact"
        if tag == "t2v":
            ts = torch.full((B, 1), 0.7311)
            skip, strat_ref, strat = None, None, None
        else:
            # per-token timesteps: first latent frame conditioned (t = 0), rest 0.61
            ts = torch.full((B, N), 0.61)
            ts[:, : h * w] = 0.0
            skip = ref.create_skip_layer_mask(1, 3, 2, [1])
            strat_ref, strat = SkipLayerStrategy.AttentionValues, O.SKIP_ATTENTION_VALUES
        y_ref = ref(hidden.clone(), freqs_cis=(cos_ref, sin_ref), encoder_hidden_states=enc,
                    timestep=ts, encoder_attention_mask=mask, skip_layer_mask=skip,
                    skip_layer_strategy=strat_ref, latent_shape=(f, h, w), joint_pass=True,
                    ltxv_model=_NoInterrupt(), return_dict=False)[0]
        y = O.transformer_forward(sd, cfg, hidden, (cos, sin), enc, ts, mask, skip, strat, (f, h, w))
        _check(f"transformer[{tag}]", y, y_ref)
        # joint_pass=False (transformer3d.py:472-487) walks the blocks sample by sample to save memory: same arithmetic per sample, so
        # the drop-in runs every sample in one batch whatever the flag says
        y_seq = ref(hidden.clone(), freqs_cis=(cos_ref, sin_ref), encoder_hidden_states=enc,
                    timestep=ts, encoder_attention_mask=mask, skip_layer_mask=skip,
                    skip_layer_strategy=strat_ref, latent_shape=(f, h, w), joint_pass=False,
                    ltxv_model=_NoInterrupt(), return_dict=False)[0]
        _check(f"transformer[{tag}] joint_pass=False vs True (reference)", y_seq, y_ref, tol=2e-6)
        out[tag] = dict(hidden=hidden, enc=enc, mask=mask, timestep=ts, out=y_ref,
                        skip=skip)
    out["rope_cos_row5"] = cos_ref[0, 5].clone()
    out["rope_sin_row5"] = sin_ref[0, 5].clone()
    torch.save(out, os.path.join(GOLD, "ltx_transformer.pt"))


def case_scheduler():
    from ltx_video.schedulers.rf import RectifiedFlowScheduler
    from ltx_video.utils.diffusers_config_mapping import OURS_SCHEDULER_CONFIG
    out = {}
    for steps, shape in ((4, (1, 128, 2, 8, 8)), (30, (1, 128, 16, 16, 24)), (40, (2, 128, 4, 16, 16)), (7, (1, 77, 128))):
        s = RectifiedFlowScheduler.from_config(dict(OURS_SCHEDULER_CONFIG))
        s.set_timesteps(steps, samples_shape=shape, device="cpu")
        mine = O.rf_timesteps(steps, shape)
        assert torch.equal(mine, s.timesteps), (mine, s.timesteps)
        g = torch.Generator().manual_seed(steps)
        x = torch.randn(1, 24, 128, generator=g)
        v = torch.randn(1, 24, 128, generator=g)
        # per-token branch as the pipeline calls it ([1,1] timestep) and a ragged per-token one
        for i in (0, steps // 2, steps - 1):
            t = s.timesteps[i][None, None]
            a = s.step(v, t, x, return_dict=False)[0]
            b = O.rf_step(v, t, x, s.timesteps)
            assert torch.equal(a, b)
        tt = torch.full((1, 24), float(s.timesteps[1]))
        tt[:, :5] = 0.0
        a = s.step(v, tt, x, return_dict=False)[0]
        b = O.rf_step(v, tt, x, s.timesteps)
        assert torch.equal(a, b)
        # stochastic sampler (:369-373): the reference draws its noise from the global RNG
        stoch = []
        for tq in (s.timesteps[0][None, None], s.timesteps[steps // 2][None, None], s.timesteps[-1][None, None], tt):
            torch.manual_seed(77)
            a_s = s.step(v, tq, x, return_dict=False, stochastic_sampling=True)[0]
            torch.manual_seed(77)
            nz = torch.randn_like(x)
            b_s = O.rf_step_stochastic(v, tq, x, s.timesteps, nz)
            assert O.rel_l2(b_s, a_s) < 1e-6, O.rel_l2(b_s, a_s)
            stoch.append(dict(t=tq.clone(), noise=nz, out=a_s.clone()))
        a0 = s.step(v, s.timesteps[1], x, return_dict=False)[0]
        assert torch.equal(a0, O.rf_step(v, s.timesteps[1], x, s.timesteps))
        out[f"{steps}_{'x'.join(map(str, shape))}"] = dict(steps=steps, shape=shape, timesteps=s.timesteps.clone(),
                                                           x=x, v=v, tt=tt, stepped=a, stochastic=stoch)
    print("  scheduler: timesteps + step bit-exact; stochastic step <= 1e-6")
    torch.save(out, os.path.join(GOLD, "rf_scheduler.pt"))


def case_patchifier():
    from ltx_video.models.transformers.symmetric_patchifier import SymmetricPatchifier
    from ltx_video.models.autoencoders.vae_encode import latent_to_pixel_coords_from_factors
    p = SymmetricPatchifier(patch_size=1)
    x = torch.arange(2 * 5 * 3 * 4 * 6, dtype=torch.float32).reshape(2, 5, 3, 4, 6)
    tok, coords = p.patchify(x)
    assert torch.equal(tok, O.patchify(x))
    assert torch.equal(coords, O.latent_coords(3, 4, 6, 2))
    back = p.unpatchify(tok, 4, 6, 5)
    assert torch.equal(back, O.unpatchify(tok, 3, 4, 6)) and torch.equal(back, x)
    for fix in (False, True):
        a = latent_to_pixel_coords_from_factors(coords, (8, 32, 32), fix)
        assert torch.equal(a, O.latent_to_pixel_coords(coords, (8, 32, 32), fix))
    print("  patchifier/coords: bit-exact")
    torch.save(dict(coords=coords, px=O.latent_to_pixel_coords(coords), px_fix=O.latent_to_pixel_coords(coords, causal_fix=True)),
               os.path.join(GOLD, "patchifier.pt"))


def build_ref_vae(sd, **cfg_overrides):
    from ltx_video.models.autoencoders.causal_video_autoencoder import CausalVideoAutoencoder
    from ltx_video.utils.diffusers_config_mapping import OURS_VAE_CONFIG
    vae = CausalVideoAutoencoder.from_config(dict(OURS_VAE_CONFIG, **cfg_overrides))
    dec_sd = {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}
    vae.decoder.load_state_dict(dec_sd, strict=True)
    vae.register_buffer("std_of_means", sd["std_of_means"])
    vae.register_buffer("mean_of_means", sd["mean_of_means"])
    return vae.eval()


def case_vae_encode():
    """Encoder.forward + DiagonalGaussianDistribution + normalize_latents as driven by vae_encode (vae_encode.py:22-91)."""
    from ltx_video.models.autoencoders.causal_video_autoencoder import CausalVideoAutoencoder
    from ltx_video.models.autoencoders.vae_encode import vae_encode
    from ltx_video.utils.diffusers_config_mapping import OURS_VAE_CONFIG
    sd = O.make_vae_encoder_state_dict(seed=2)
    vae = CausalVideoAutoencoder.from_config(dict(OURS_VAE_CONFIG))
    vae.encoder.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    vae.register_buffer("std_of_means", sd["std_of_means"])
    vae.register_buffer("mean_of_means", sd["mean_of_means"])
    vae = vae.eval()
    g = torch.Generator().manual_seed(11)
    video = (torch.rand(1, 3, 9, 64, 96, generator=g) * 2 - 1)
    image = video[:, :, :1].contiguous()                                   # single conditioning frame (i2v)
    out = {}
    for tag, x in (("video", video), ("image", image)):
        post = vae.encode(x).latent_dist
        mean, logvar = O.vae_encode_moments(sd, x)
        _check(f"vae_encode moments mean ({tag})", mean, post.mean, tol=1e-4)
        assert O.rel_l2(logvar.expand_as(post.logvar).clamp(-30, 20), post.logvar) < 1e-4
        torch.manual_seed(123)
        z_ref = vae_encode(x, vae, vae_per_channel_normalize=True)
        torch.manual_seed(123)
        noise = torch.randn(mean.shape)
        z = O.vae_encode(sd, x, noise=noise)
        _check(f"vae_encode sample+normalize ({tag})", z, z_ref, tol=1e-4)
        out[tag] = dict(x=x.half(), mean=post.mean.clone(), logvar=post.logvar[:, :1].clone(), z=z_ref.clone(), noise=noise)
    assert O.vae_encode_moments(sd, video)[0].shape == (1, 128, 2, 2, 3)
    torch.save(dict(seed_weights=2, **out), os.path.join(GOLD, "ltx_vae_encode.pt"))


def case_vae():
    from ltx_video.models.autoencoders.vae_encode import vae_decode
    sd = O.make_vae_decoder_state_dict(seed=1)
    vae = build_ref_vae(sd)
    g = torch.Generator().manual_seed(5)
    z = torch.randn(1, 128, 2, 3, 4, generator=g)
    y_ref = vae_decode(z, vae, is_video=True, vae_per_channel_normalize=True)
    y = O.vae_decode(sd, z)
    _check("vae_decode", y, y_ref, tol=1e-4)
    assert y_ref.shape == (1, 3, 9, 96, 128)
    # the reference's own in-file check: patchify∘unpatchify == identity (causal_video_autoencoder.py:1341-1347)
    from ltx_video.models.autoencoders.causal_video_autoencoder import patchify as rp, unpatchify as ru
    xx = torch.randn(2, 3, 8, 64, 64, generator=g)
    assert torch.equal(ru(rp(xx, 4, 1), 4, 1), xx)
    assert torch.equal(ru(rp(xx, 4, 1), 4, 1), O.vae_unpatchify(rp(xx, 4, 1), 4))
    torch.save(dict(z=z, out=y_ref.to(torch.float16), seed_weights=1), os.path.join(GOLD, "ltx_vae_decode.pt"))
    # timestep-conditioned decoder (causal_video_autoencoder.py:724-733,757-795,1207-1237; the 0.9.x VAEs, decode_timestep 0.05)
    tcfg = dict(O.LTX_VAE, timestep_conditioning=True)
    tsd = O.make_vae_decoder_state_dict(tcfg, seed=4)
    tvae = build_ref_vae(tsd, timestep_conditioning=True)
    t = torch.tensor([0.05])
    yt_ref = vae_decode(z, tvae, is_video=True, vae_per_channel_normalize=True, timestep=t)
    yt = O.vae_decode(tsd, z, tcfg, timestep=t)
    _check("vae_decode(timestep-conditioned)", yt, yt_ref, tol=1e-4)
    assert O.rel_l2(O.vae_decode(tsd, z, tcfg, timestep=torch.tensor([0.5])), yt_ref) > 1e-3      # the timestep matters
    torch.save(dict(z=z, out=yt_ref.to(torch.float16), seed_weights=4, timestep=t), os.path.join(GOLD, "ltx_vae_decode_timestep.pt"))


@contextlib.contextmanager
def _cuda_to_cpu():
    """pipeline_ltx_video.py:1041 hard-codes .to("cuda"); map it to cpu for the CPU run."""
    orig = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x == "cuda") else x for x in a)
        return orig(self, *a, **k)

    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.to = orig


def case_pipeline():
    """LTXVideoPipeline.__call__ (output_type='latent'): guidance off, CFG+STG+rescale preset,
    and the i2v per-token-timestep path via a conditioning mask (driven through latents)."""
    from ltx_video.pipelines.pipeline_ltx_video import LTXVideoPipeline
    from ltx_video.schedulers.rf import RectifiedFlowScheduler
    from ltx_video.models.transformers.symmetric_patchifier import SymmetricPatchifier
    from ltx_video.utils.diffusers_config_mapping import OURS_SCHEDULER_CONFIG
    from ltx_video.utils.skip_layer_strategy import SkipLayerStrategy
    cfg = O.LTX_2B
    L = 2
    sd = O.make_transformer_state_dict(cfg, seed=0, num_layers=L)
    tr = build_ref_transformer(L, sd)
    vsd = O.make_vae_decoder_state_dict(seed=1)
    vae = build_ref_vae(vsd)
    pipe = LTXVideoPipeline(tokenizer=None, text_encoder=None, vae=vae, transformer=tr,
                            scheduler=RectifiedFlowScheduler.from_config(dict(OURS_SCHEDULER_CONFIG)),
                            patchifier=SymmetricPatchifier(patch_size=1),
                            prompt_enhancer_image_caption_model=None,
                            prompt_enhancer_image_caption_processor=None,
                            prompt_enhancer_llm_model=None, prompt_enhancer_llm_tokenizer=None)
    H, W, F_, fps, steps = 128, 192, 17, 25.0, 4          # latent (1,128,3,4,6), N=72
    g = torch.Generator().manual_seed(42)
    pe = torch.randn(1, 32, 4096, generator=g)
    ne = torch.randn(1, 32, 4096, generator=g)
    pm = torch.ones(1, 32); pm[:, 20:] = 0
    nm = torch.ones(1, 32); nm[:, 9:] = 0
    out = {"meta": dict(H=H, W=W, F=F_, fps=fps, steps=steps, num_layers=L)}
    cwd = os.getcwd()
    os.chdir("/tmp")                                          # 'lala.pt' side effect (:1288)
    try:
        for tag, kw in (("plain", dict(guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0)),
                        ("cfg_stg", dict(guidance_scale=3.0, stg_scale=1.0, rescaling_scale=0.7,
                                         skip_block_list=[1],
                                         skip_layer_strategy=SkipLayerStrategy.AttentionValues))):
            per_step = []
            gen = torch.Generator().manual_seed(7)
            with _cuda_to_cpu():
                lat = pipe(height=H, width=W, num_frames=F_, frame_rate=fps, prompt_embeds=pe,
                           prompt_attention_mask=pm, negative_prompt_embeds=ne,
                           negative_prompt_attention_mask=nm, num_inference_steps=steps,
                           generator=gen, output_type="latent", return_dict=False, joint_pass=True,
                           ltxv_model=_NoInterrupt(), is_video=True, vae_per_channel_normalize=True,
                           callback_on_step_end=None, **kw)[0]
            # same initial noise the pipeline drew (prepare_latents :696-699)
            gen = torch.Generator().manual_seed(7)
            noise = torch.randn(1, 3 * 4 * 6, 128, generator=gen)
            mine = O.denoise_loop(sd, cfg, noise, pe, pm, num_frames_lat=3, lat_h=4, lat_w=6,
                                  frame_rate=fps, num_steps=steps, neg_enc=ne, neg_mask=nm,
                                  guidance_scale=kw["guidance_scale"], stg_scale=kw["stg_scale"],
                                  rescaling_scale=kw["rescaling_scale"],
                                  skip_block_list=kw.get("skip_block_list"),
                                  strategy=O.SKIP_ATTENTION_VALUES if "skip_block_list" in kw else None,
                                  per_step=per_step)
            mine5 = O.unpatchify(mine, 3, 4, 6)
            _check(f"pipeline[{tag}] final latents", mine5, lat, tol=5e-5)
            out[tag] = dict(latents=lat, kw={k: (v if not hasattr(v, "name") else v.name) for k, v in kw.items()})
        # decoded frames for the plain case through the reference pipeline (full path)
        gen = torch.Generator().manual_seed(7)
        with _cuda_to_cpu():
            img = pipe(height=H, width=W, num_frames=F_, frame_rate=fps, prompt_embeds=pe,
                       prompt_attention_mask=pm, negative_prompt_embeds=ne,
                       negative_prompt_attention_mask=nm, num_inference_steps=steps,
                       generator=gen, output_type="pt", return_dict=False, joint_pass=True,
                       ltxv_model=_NoInterrupt(), is_video=True, vae_per_channel_normalize=True,
                       guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0)[0]
        mine_img = O.postprocess(O.vae_decode(vsd, O.unpatchify(
            O.denoise_loop(sd, cfg, torch.randn(1, 72, 128, generator=torch.Generator().manual_seed(7)), pe, pm,
                           num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=fps, num_steps=steps), 3, 4, 6)))
        print(f"  pipeline decoded frames: PSNR(oracle, reference) = {O.psnr(mine_img, img):.1f} dB")
        assert O.psnr(mine_img, img) > 80
        out["plain"]["frames_sub"] = img[:, :, ::4, ::8, ::8].to(torch.float16).clone()
    finally:
        os.chdir(cwd)
    out.update(pe=pe, ne=ne, pm=pm, nm=nm, noise_seed=7)
    torch.save(out, os.path.join(GOLD, "ltx_pipeline.pt"))


def case_pipeline_i2v():
    """BASELINE configs[2] in small, end to end through the reference's OWN __call__: a pixel-space ConditioningItem goes through the
    reference VAE ENCODER inside prepare_conditioning (:1396-1448), the loop runs with the per-token timesteps / conditioning mask and
    image_cond_noise_scale = 0; the oracle composes vae_encode + first-frame blend + denoise_loop and must give the same latents."""
    from ltx_video.pipelines.pipeline_ltx_video import ConditioningItem, LTXVideoPipeline
    from ltx_video.schedulers.rf import RectifiedFlowScheduler
    from ltx_video.models.transformers.symmetric_patchifier import SymmetricPatchifier
    from ltx_video.utils.diffusers_config_mapping import OURS_SCHEDULER_CONFIG
    cfg, L = O.LTX_2B, 2
    sd = O.make_transformer_state_dict(cfg, seed=0, num_layers=L)
    tr = build_ref_transformer(L, sd)
    esd = O.make_vae_encoder_state_dict(seed=2)                     # encoder.* + the latent statistics
    vae = build_ref_vae(dict(O.make_vae_decoder_state_dict(seed=1), std_of_means=esd["std_of_means"], mean_of_means=esd["mean_of_means"]))
    vae.encoder.load_state_dict({k[len("encoder."):]: v for k, v in esd.items() if k.startswith("encoder.")}, strict=True)
    pipe = LTXVideoPipeline(tokenizer=None, text_encoder=None, vae=vae, transformer=tr,
                            scheduler=RectifiedFlowScheduler.from_config(dict(OURS_SCHEDULER_CONFIG)),
                            patchifier=SymmetricPatchifier(patch_size=1), prompt_enhancer_image_caption_model=None,
                            prompt_enhancer_image_caption_processor=None, prompt_enhancer_llm_model=None, prompt_enhancer_llm_tokenizer=None)
    H, W, F_, fps, steps = 128, 192, 17, 25.0, 3                   # latent (1,128,3,4,6)
    g = torch.Generator().manual_seed(4)
    pe, pm = torch.randn(1, 16, 4096, generator=g), torch.ones(1, 16)
    image = torch.rand(1, 3, 1, H, W, generator=g) * 2 - 1
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        torch.manual_seed(123)                                       # latent_dist.sample() draws from the global RNG (vae_encode.py:77)
        with _cuda_to_cpu():
            lat = pipe(height=H, width=W, num_frames=F_, frame_rate=fps, prompt_embeds=pe, prompt_attention_mask=pm,
                       negative_prompt_embeds=None, negative_prompt_attention_mask=None, num_inference_steps=steps,
                       generator=torch.Generator().manual_seed(5), output_type="latent", return_dict=False, joint_pass=True,
                       ltxv_model=_NoInterrupt(), is_video=True, vae_per_channel_normalize=True, guidance_scale=1.0, stg_scale=0.0,
                       rescaling_scale=1.0, image_cond_noise_scale=0.0,
                       conditioning_items=[ConditioningItem(media_item=image, media_frame_number=0, conditioning_strength=1.0)])[0]
    finally:
        os.chdir(cwd)
    torch.manual_seed(123)
    noise_e = torch.randn(1, 128, 1, 4, 6)
    cond_lat = O.vae_encode(esd, image, noise=noise_e)
    init = O.unpatchify(torch.randn(1, 72, 128, generator=torch.Generator().manual_seed(5)), 3, 4, 6).clone()
    init[:, :, :1] = cond_lat
    cmask = torch.zeros(1, 3, 4, 6)
    cmask[:, :1] = 1.0
    mine = O.denoise_loop(sd, cfg, O.patchify(init), pe, pm, num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=fps, num_steps=steps,
                          conditioning_mask=cmask.reshape(1, -1))
    _check("pipeline i2v from pixels (reference __call__ vs oracle composition)", O.unpatchify(mine, 3, 4, 6), lat, tol=5e-5)
    # the same call with image_cond_noise_scale = 0.15 (the app's value, ltxv.py; :606-629): the hard-conditioned first-frame tokens are
    # re-noised from the call's generator at the start of every step
    os.chdir("/tmp")
    try:
        torch.manual_seed(123)
        with _cuda_to_cpu():
            lat_n = pipe(height=H, width=W, num_frames=F_, frame_rate=fps, prompt_embeds=pe, prompt_attention_mask=pm,
                         negative_prompt_embeds=None, negative_prompt_attention_mask=None, num_inference_steps=steps,
                         generator=torch.Generator().manual_seed(5), output_type="latent", return_dict=False, joint_pass=True,
                         ltxv_model=_NoInterrupt(), is_video=True, vae_per_channel_normalize=True, guidance_scale=1.0, stg_scale=0.0,
                         rescaling_scale=1.0, image_cond_noise_scale=0.15,
                         conditioning_items=[ConditioningItem(media_item=image, media_frame_number=0, conditioning_strength=1.0)])[0]
    finally:
        os.chdir(cwd)
    gen = torch.Generator().manual_seed(5)                          # one stream: initial noise, then one draw per step
    init = O.unpatchify(torch.randn(1, 72, 128, generator=gen), 3, 4, 6).clone()
    init[:, :, :1] = cond_lat
    mine_n = O.denoise_loop(sd, cfg, O.patchify(init), pe, pm, num_frames_lat=3, lat_h=4, lat_w=6, frame_rate=fps, num_steps=steps,
                            conditioning_mask=cmask.reshape(1, -1), image_cond_noise_scale=0.15, generator=gen)
    _check("pipeline i2v, image_cond_noise_scale 0.15", O.unpatchify(mine_n, 3, 4, 6), lat_n, tol=5e-5)
    print(f"  (image_cond_noise 0.15 vs 0: final latents differ by {O.rel_l2(lat_n, lat):.3e}; first frame {O.rel_l2(lat_n[:, :, :1], lat[:, :, :1]):.3e})")
    torch.save(dict(meta=dict(H=H, W=W, F=F_, fps=fps, steps=steps, num_layers=L), pe=pe, pm=pm, image=image, noise_e=noise_e,
                    noise_seed=5, latents=lat.clone(), latents_cond_noise_0p15=lat_n.clone()), os.path.join(GOLD, "ltx_pipeline_i2v.pt"))


def build_ref_upsampler(sd, in_channels, mid_channels, nb):
    from ltx_video.models.autoencoders.latent_upsampler import LatentUpsampler
    m = LatentUpsampler(in_channels=in_channels, mid_channels=mid_channels, num_blocks_per_stage=nb, dims=3,
                        spatial_upsample=True, temporal_upsample=False)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def case_multiscale():
    """LatentUpsampler.forward, adain_filter_latent and the whole LTXMultiScalePipeline.__call__ (first pass at
    2/3 resolution -> upsample -> AdaIN -> second pass from the partially re-noised latents -> decode -> bilinear resize), with
    per-guidance-timestep guidance tables like ltxv-13b-0.9.7-dev.yaml."""
    from ltx_video.pipelines.pipeline_ltx_video import LTXVideoPipeline, LTXMultiScalePipeline, adain_filter_latent
    from ltx_video.schedulers.rf import RectifiedFlowScheduler
    from ltx_video.models.transformers.symmetric_patchifier import SymmetricPatchifier
    from ltx_video.utils.diffusers_config_mapping import OURS_SCHEDULER_CONFIG
    from ltx_video.utils.skip_layer_strategy import SkipLayerStrategy
    MID, NB = 256, 1
    usd = O.make_latent_upsampler_state_dict(128, MID, NB, seed=3)
    up = build_ref_upsampler(usd, 128, MID, NB)
    g = torch.Generator().manual_seed(11)
    z = torch.randn(1, 128, 3, 4, 6, generator=g)
    y_ref = up(z)
    _check("latent upsampler", O.latent_upsampler_forward(usd, z), y_ref, tol=2e-5)
    ref_lat = torch.randn(1, 128, 3, 2, 3, generator=g) * 1.7 + 0.3
    a_ref = adain_filter_latent(y_ref, ref_lat)
    _check("adain_filter_latent", O.adain_filter_latent(y_ref, ref_lat), a_ref, tol=2e-6)
    _check("adain_filter_latent(0.25)", O.adain_filter_latent(y_ref, ref_lat, 0.25), adain_filter_latent(y_ref, ref_lat, 0.25), tol=2e-6)
    out = dict(upsampler=dict(mid=MID, nb=NB, seed=3, z=z, out=y_ref, ref_lat=ref_lat, adain=a_ref))

    cfg = O.LTX_2B
    L = 2
    sd = O.make_transformer_state_dict(cfg, seed=0, num_layers=L)
    tr = build_ref_transformer(L, sd)
    vsd = O.make_vae_decoder_state_dict(seed=1)
    vae = build_ref_vae(vsd)
    pipe = LTXVideoPipeline(tokenizer=None, text_encoder=None, vae=vae, transformer=tr,
                            scheduler=RectifiedFlowScheduler.from_config(dict(OURS_SCHEDULER_CONFIG)),
                            patchifier=SymmetricPatchifier(patch_size=1), prompt_enhancer_image_caption_model=None,
                            prompt_enhancer_image_caption_processor=None, prompt_enhancer_llm_model=None,
                            prompt_enhancer_llm_tokenizer=None)
    H, W, F_, fps = 160, 256, 17, 25.0          # x 0.6666666 -> 96 x 160 (latent 3x3x5); second pass 192 x 320 (latent 3x6x10)
    pe = torch.randn(1, 32, 4096, generator=g)
    ne = torch.randn(1, 32, 4096, generator=g)
    pm = torch.ones(1, 32); pm[:, 20:] = 0
    nm = torch.ones(1, 32); nm[:, 9:] = 0
    pipe.encode_prompt = lambda *a, **k: (pe, pm, ne, nm)     # the T5 encoder is out of scope: hand the embeddings over
    multi = LTXMultiScalePipeline(pipe, up)
    first = dict(guidance_scale=[1, 3, 1], stg_scale=[0, 1, 1], rescaling_scale=[1, 0.7, 1], guidance_timesteps=[1.0, 0.95, 0.6],
                 skip_block_list=[[], [1], [0]], skip_final_inference_steps=1, cfg_star_rescale=True)
    second = dict(guidance_scale=[1], stg_scale=[1], rescaling_scale=[1], guidance_timesteps=[1.0], skip_block_list=[1],
                  skip_initial_inference_steps=2, cfg_star_rescale=True)
    common = dict(downscale_factor=0.6666666, first_pass=first, second_pass=second, height=H, width=W, num_frames=F_,
                  frame_rate=fps, prompt="p", negative_prompt="n", num_inference_steps1=5, num_inference_steps2=5,
                  skip_layer_strategy=SkipLayerStrategy.AttentionValues, VAE_tile_size=(0, 0), ltxv_model=_NoInterrupt(),
                  device="cpu", return_dict=True, is_video=True, vae_per_channel_normalize=True, enhance_prompt=False)
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        with _cuda_to_cpu():
            lat = multi(**common, output_type="latent", generator=torch.Generator().manual_seed(7))
            img = multi(**common, output_type="pt", generator=torch.Generator().manual_seed(7))
    finally:
        os.chdir(cwd)
    # ---- the same flow on the oracle
    gen = torch.Generator().manual_seed(7)
    f, h1, w1, h2, w2 = 3, 3, 5, 6, 10
    kw = dict(num_frames_lat=f, frame_rate=fps, neg_enc=ne, neg_mask=nm, strategy=O.SKIP_ATTENTION_VALUES)
    n1 = torch.randn(1, f * h1 * w1, 128, generator=gen)
    ts1 = O.rf_timesteps(5, (1, 128, f, h1, w1))[:4]
    l1 = O.denoise_loop(sd, cfg, n1, pe, pm, lat_h=h1, lat_w=w1, num_steps=5, timesteps=ts1, guidance_scale=first["guidance_scale"],
                        stg_scale=first["stg_scale"], rescaling_scale=first["rescaling_scale"],
                        guidance_timesteps=first["guidance_timesteps"], skip_block_list=first["skip_block_list"], **kw)
    l1 = O.unpatchify(l1, f, h1, w1)
    upl = O.adain_filter_latent(O.upsample_latents(usd, vsd, l1), l1)
    n2 = O.unpatchify(torch.randn(1, f * h2 * w2, 128, generator=gen), f, h2, w2)
    ts2 = O.rf_timesteps(5, (1, 128, f, h2, w2))[2:]
    init2 = O.multiscale_second_pass_init(n2, upl, float(ts2[0]))
    l2 = O.denoise_loop(sd, cfg, O.patchify(init2), pe, pm, lat_h=h2, lat_w=w2, num_steps=5, timesteps=ts2,
                        guidance_scale=second["guidance_scale"], stg_scale=second["stg_scale"],
                        rescaling_scale=second["rescaling_scale"], guidance_timesteps=second["guidance_timesteps"],
                        skip_block_list=second["skip_block_list"], **kw)
    l2 = O.unpatchify(l2, f, h2, w2)
    _check("multi-scale final latents", l2, lat, tol=5e-5)
    mine_img = O.multiscale_resize(O.postprocess(O.vae_decode(vsd, l2)), H, W)
    print(f"  multi-scale decoded + resized frames: PSNR(oracle, reference) = {O.psnr(mine_img, img):.1f} dB")
    assert tuple(img.shape) == (1, 3, F_, H, W) and O.psnr(mine_img, img) > 70
    strat = lambda d: {k: v for k, v in d.items()}
    out["pipeline"] = dict(meta=dict(H=H, W=W, F=F_, fps=fps, num_layers=L, steps=5, downscale_factor=0.6666666, noise_seed=7),
                           first_pass=strat(first), second_pass=strat(second), pe=pe, ne=ne, pm=pm, nm=nm,
                           first_latents=l1, latents=lat, frames_sub=img[:, :, ::4, ::8, ::8].to(torch.float16).clone())
    torch.save(out, os.path.join(GOLD, "ltx_multiscale.pt"))
    print("written", os.path.join(GOLD, "ltx_multiscale.pt"))


def case_sinusoid():
    """The timestep sinusoid of AdaLayerNormSingle: diffusers' get_timestep_embedding, which the reference VENDORS
    (ltx_video/models/transformers/embeddings.py:10-50) — so this piece of the diffusers boundary is pinned by reference code.
    The oracle's restatement must match it bit for bit (flip_sin_to_cos=True, downscale_freq_shift=0, 256 channels)."""
    from ltx_video.models.transformers.embeddings import get_timestep_embedding
    t = torch.cat([torch.tensor([0.0, 1.0, 0.5, 999.0, 1000.0]), torch.rand(59, generator=torch.Generator().manual_seed(9)) * 1000])
    ref = get_timestep_embedding(t, 256, flip_sin_to_cos=True, downscale_freq_shift=0)
    assert torch.equal(O.timestep_sinusoid(t, 256), ref), "timestep sinusoid not bit-exact"
    print("  timestep sinusoid: bit-exact against the reference's vendored get_timestep_embedding", tuple(ref.shape))
    torch.save(dict(t=t, emb=ref), os.path.join(GOLD, "timestep_sinusoid.pt"))


if __name__ == "__main__":
    which = sys.argv[1:] or ["patchifier", "scheduler", "sinusoid", "transformer", "vae", "vae_encode", "pipeline", "pipeline_i2v", "multiscale"]
    for w in which:
        print(f"[{w}]")
        globals()["case_" + w]()
    print("golden fixtures written to", GOLD)
