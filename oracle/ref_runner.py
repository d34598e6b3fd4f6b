"""Run the UNMODIFIED reference (oracle/_ref, populated by oracle/build_ref.py; /root/reference in the build container) on
the host CPU cores.  BASELINE INFRASTRUCTURE ONLY: imported by bench.py's `--impl reference` arm and its `cpu_baseline` leg,
never by the product path.

Two measurements:
  * `config0()`         — BASELINE.json configs[0] for real: the reference LTXVideoPipeline.__call__
                          (pipeline_ltx_video.py:763) with its own Transformer3DModel (28 layers) and CausalVideoAutoencoder,
                          t2v 256x256x9, 4 denoise steps + VAE decode, fp32;
  * `full_step_real()`   — configs[1] for real: ONE whole denoise step of the bench workload (768x512x121 -> 6144 tokens, all 28
                          layers, all num_conds guidance conditions in one batched forward incl. the STG skip-layer mask — the
                          call pipeline_ltx_video.py:1153-1177 makes once per step), timed, nothing extrapolated (30-60 s);
  * `full_size_sample()` — a bounded sample of configs[1] (768x512x121 -> 6144 tokens): the reference
                          Transformer3DModel.forward (transformer3d.py:328) at full size with 1 and 3 layers; per-layer and
                          fixed costs are separated and EXTRAPOLATED to 28 layers x num_conds (one full step takes minutes).
The third-party `diffusers` / `mmgp` imports of the reference are answered by oracle/refshim (neither is installed).
"""
import contextlib
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_root():
    for p in (os.path.join(HERE, "_ref"), os.environ.get("LTX_REFERENCE_ROOT", ""), "/root/reference"):
        if p and os.path.isdir(os.path.join(p, "ltx_video")):
            return p
    return None


def available() -> bool:
    return reference_root() is not None


_installed = False


def _install():
    global _installed
    if _installed:
        return
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    shim = os.path.join(HERE, "refshim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    import load_reference
    load_reference.REFERENCE_ROOT = reference_root()
    load_reference.install()
    torch.set_grad_enabled(False)
    _installed = True


class _NoInterrupt:
    _interrupt = False


@contextlib.contextmanager
def _cuda_to_cpu():
    """pipeline_ltx_video.py:1041 hard-codes .to("cuda"); the CPU arm maps it to cpu."""
    orig = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x == "cuda") else x for x in a)
        return orig(self, *a, **k)

    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.to = orig


def _ref_transformer(num_layers, sd):
    from ltx_video.models.transformers.transformer3d import Transformer3DModel
    from ltx_video.utils.diffusers_config_mapping import OURS_TRANSFORMER_CONFIG
    cfg = dict(OURS_TRANSFORMER_CONFIG)
    cfg["num_layers"] = num_layers
    m = Transformer3DModel.from_config(cfg)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def full_size_sample(wl: dict, layers=(1, 3)) -> dict:
    """One bounded sample of a full-size denoise step through the reference's own Transformer3DModel.forward."""
    _install()
    from oracle import ltx_oracle as O
    torch.manual_seed(0)
    f, h, w = wl["num_frames"] // 8 + 1, wl["height"] // 32, wl["width"] // 32
    N = f * h * w
    hidden = torch.randn(1, N, 128)
    enc = torch.randn(1, wl["prompt_tokens"], 4096)
    mask = torch.ones(1, wl["prompt_tokens"])
    coords = O.latent_to_pixel_coords(O.latent_coords(f, h, w, 1)).float()
    coords[:, 0] /= wl["frame_rate"]
    ts = torch.full((1, 1), 0.5)
    times = {}
    for L in layers:
        sd = O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=L)
        m = _ref_transformer(L, sd)
        freqs = m.precompute_freqs_cis(coords)
        t0 = time.perf_counter()
        m(hidden, freqs_cis=freqs, encoder_hidden_states=enc, timestep=ts, encoder_attention_mask=mask, latent_shape=(f, h, w),
          joint_pass=True, ltxv_model=_NoInterrupt(), return_dict=False)
        times[L] = time.perf_counter() - t0
        del m, sd
    t_layer = max(times[layers[1]] - times[layers[0]], 1e-9) / (layers[1] - layers[0])
    t_fixed = max(times[layers[0]] - layers[0] * t_layer, 0.0)
    step_s = wl["num_conds"] * (t_fixed + 28 * t_layer)
    return dict(step_s=step_s, t_layer=t_layer, t_fixed=t_fixed, raw=times, tokens=N, sample_s=sum(times.values()))


def _full_size_inputs(wl, batch):
    from oracle import ltx_oracle as O
    torch.manual_seed(0)
    f, h, w = wl["num_frames"] // 8 + 1, wl["height"] // 32, wl["width"] // 32
    N = f * h * w
    hidden = torch.randn(batch, N, 128)
    enc = torch.randn(batch, wl["prompt_tokens"], 4096)
    mask = torch.ones(batch, wl["prompt_tokens"])
    coords = O.latent_to_pixel_coords(O.latent_coords(f, h, w, 1)).float()
    coords[:, 0] /= wl["frame_rate"]
    ts = torch.full((batch, 1), 0.5)
    return (f, h, w), N, hidden, enc, mask, coords.expand(batch, -1, -1).contiguous(), ts


def reference_2b_transformer():
    """The unmodified reference Transformer3DModel at the full 28 layers (1.92 B seeded fp32 weights)."""
    _install()
    from oracle import ltx_oracle as O
    sd = O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=28)
    return _ref_transformer(28, sd)


def full_step_real(wl: dict, tr=None) -> dict:
    """ONE whole denoise step of the bench workload through the reference's own Transformer3DModel.forward: every layer, every
    guidance condition (batched the way pipeline_ltx_video.py:1034-1051 batches them, STG skip-layer mask :1137-1150 included).
    The guidance arithmetic and the scheduler update that complete a step (:1189-1235) are elementwise work on 3 MB tensors."""
    _install()
    from ltx_video.utils.skip_layer_strategy import SkipLayerStrategy
    if tr is None:
        tr = reference_2b_transformer()
    conds = wl["num_conds"]
    latent_shape, N, hidden, enc, mask, coords, ts = _full_size_inputs(wl, conds)
    freqs = tr.precompute_freqs_cis(coords)
    kw = {}
    if conds == 3:                                 # the 2B preset: STG on block 19, perturbed condition last
        kw = dict(skip_layer_mask=tr.create_skip_layer_mask(1, conds, conds - 1, [19]), skip_layer_strategy=SkipLayerStrategy.AttentionValues)
    t0 = time.perf_counter()
    out = tr(hidden, freqs_cis=freqs, encoder_hidden_states=enc, timestep=ts, encoder_attention_mask=mask, latent_shape=latent_shape,
             joint_pass=True, ltxv_model=_NoInterrupt(), return_dict=False, **kw)[0]
    step_s = time.perf_counter() - t0
    assert out.shape == hidden.shape and bool(torch.isfinite(out).all())
    return dict(step_s=step_s, tokens=N, conds=conds, layers=28, sample_s=step_s)


def config0(steps: int = 4, tr=None) -> dict:
    """BASELINE configs[0], timed for real through the reference's own pipeline call (loop alone, then loop + decode)."""
    _install()
    from ltx_video.models.autoencoders.causal_video_autoencoder import CausalVideoAutoencoder
    from ltx_video.models.transformers.symmetric_patchifier import SymmetricPatchifier
    from ltx_video.pipelines.pipeline_ltx_video import LTXVideoPipeline
    from ltx_video.schedulers.rf import RectifiedFlowScheduler
    from ltx_video.utils.diffusers_config_mapping import OURS_SCHEDULER_CONFIG, OURS_VAE_CONFIG
    from oracle import ltx_oracle as O
    if tr is None:
        tr = reference_2b_transformer()
    vsd = O.make_vae_decoder_state_dict(seed=1)
    vae = CausalVideoAutoencoder.from_config(dict(OURS_VAE_CONFIG))
    vae.decoder.load_state_dict({k[len("decoder."):]: v for k, v in vsd.items() if k.startswith("decoder.")}, strict=True)
    vae.register_buffer("std_of_means", vsd["std_of_means"])
    vae.register_buffer("mean_of_means", vsd["mean_of_means"])
    vae.eval()
    pipe = LTXVideoPipeline(tokenizer=None, text_encoder=None, vae=vae, transformer=tr,
                            scheduler=RectifiedFlowScheduler.from_config(dict(OURS_SCHEDULER_CONFIG)),
                            patchifier=SymmetricPatchifier(patch_size=1), prompt_enhancer_image_caption_model=None,
                            prompt_enhancer_image_caption_processor=None, prompt_enhancer_llm_model=None,
                            prompt_enhancer_llm_tokenizer=None)
    pe = torch.randn(1, 32, 4096, generator=torch.Generator().manual_seed(42))
    kw = dict(height=256, width=256, num_frames=9, frame_rate=30.0, prompt_embeds=pe, prompt_attention_mask=torch.ones(1, 32),
              negative_prompt_embeds=None, negative_prompt_attention_mask=None, num_inference_steps=steps, return_dict=False,
              joint_pass=True, ltxv_model=_NoInterrupt(), is_video=True, vae_per_channel_normalize=True, guidance_scale=1.0,
              stg_scale=0.0, rescaling_scale=1.0)
    cwd = os.getcwd()
    os.chdir("/tmp")                              # the reference writes a scratch file into the cwd (:1288)
    try:
        with _cuda_to_cpu():
            t0 = time.perf_counter()
            pipe(generator=torch.Generator("cpu").manual_seed(42), output_type="latent", **kw)
            loop_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            pipe(generator=torch.Generator("cpu").manual_seed(42), output_type="pt", **kw)
            video_s = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    return dict(steps=steps, loop_s=loop_s, video_s=video_s, steps_per_s=steps / loop_s)
