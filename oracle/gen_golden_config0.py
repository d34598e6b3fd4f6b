"""BASELINE.json configs[0] — "LTX-Video 2B random-init t2v 256x256x9 frames, 4 denoise steps + VAE decode, fp32 on
CPU (reference path, no GPU)" — run through the UNMODIFIED reference at FULL depth (28 layers, SURVEY.md §8d Config 1)
and recorded as tests/golden/ltx_config0.pt (TEST INFRASTRUCTURE ONLY).

Run in the build container only (needs /root/reference; ~16 GB of RAM for the two fp32 copies of the 1.9 B weights):
    python oracle/gen_golden_config0.py
It (1) builds the reference LTXVideoPipeline (pipeline_ltx_video.py:763) with the reference Transformer3DModel (28 layers)
and CausalVideoAutoencoder on the oracle's seeded weights, (2) runs the reference end to end on the CPU in fp32 — latents
drawn by the reference's own prepare_latents (:632-710) from Generator('cpu').manual_seed(42) — with output_type 'latent'
and 'pt', (3) asserts the oracle's denoise loop + decode reproduce both, (4) stores the reference's final latents, the
oracle's per-step latents and the reference's decoded frames (fp16), and the wall-clock of the reference run on this
container's cores (a reported number).
"""
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

from oracle import ltx_oracle as O  # noqa: E402
from oracle.gen_golden import GOLD, _NoInterrupt, _check, _cuda_to_cpu, build_ref_transformer, build_ref_vae  # noqa: E402

torch.set_grad_enabled(False)

H = W = 256
FRAMES, FPS, STEPS, LAYERS = 9, 30.0, 4, 28
LAT = (2, 8, 8)                                   # latent (1,128,2,8,8), N = 128 tokens
SEED_W, SEED_VAE, SEED_PROMPT, SEED_NOISE = 0, 1, 42, 42


def main():
    from ltx_video.pipelines.pipeline_ltx_video import LTXVideoPipeline
    from ltx_video.schedulers.rf import RectifiedFlowScheduler
    from ltx_video.models.transformers.symmetric_patchifier import SymmetricPatchifier
    from ltx_video.utils.diffusers_config_mapping import OURS_SCHEDULER_CONFIG
    cfg = O.LTX_2B
    t0 = time.perf_counter()
    sd = O.make_transformer_state_dict(cfg, seed=SEED_W, num_layers=LAYERS)
    print(f"  weights: {sum(v.numel() for v in sd.values()) / 1e9:.2f} B parameters in {time.perf_counter() - t0:.1f} s")
    tr = build_ref_transformer(LAYERS, sd)
    vsd = O.make_vae_decoder_state_dict(seed=SEED_VAE)
    vae = build_ref_vae(vsd)
    pipe = LTXVideoPipeline(tokenizer=None, text_encoder=None, vae=vae, transformer=tr,
                            scheduler=RectifiedFlowScheduler.from_config(dict(OURS_SCHEDULER_CONFIG)),
                            patchifier=SymmetricPatchifier(patch_size=1),
                            prompt_enhancer_image_caption_model=None, prompt_enhancer_image_caption_processor=None,
                            prompt_enhancer_llm_model=None, prompt_enhancer_llm_tokenizer=None)
    pe = torch.randn(1, 32, 4096, generator=torch.Generator().manual_seed(SEED_PROMPT))
    pm = torch.ones(1, 32)
    kw = dict(height=H, width=W, num_frames=FRAMES, frame_rate=FPS, prompt_embeds=pe, prompt_attention_mask=pm,
              negative_prompt_embeds=None, negative_prompt_attention_mask=None, num_inference_steps=STEPS,
              return_dict=False, joint_pass=True, ltxv_model=_NoInterrupt(), is_video=True, vae_per_channel_normalize=True,
              guidance_scale=1.0, stg_scale=0.0, rescaling_scale=1.0)
    cwd = os.getcwd()
    os.chdir("/tmp")                              # the reference writes a scratch file into the cwd (:1288)
    try:
        with _cuda_to_cpu():
            t0 = time.perf_counter()
            lat = pipe(generator=torch.Generator("cpu").manual_seed(SEED_NOISE), output_type="latent", **kw)[0]
            t_loop = time.perf_counter() - t0
            t0 = time.perf_counter()
            img = pipe(generator=torch.Generator("cpu").manual_seed(SEED_NOISE), output_type="pt", **kw)[0]
            t_video = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    del pipe, tr
    f, h, w = LAT
    assert tuple(lat.shape) == (1, 128, f, h, w) and tuple(img.shape) == (1, 3, FRAMES, H, W)

    # the oracle on the same noise the reference drew (prepare_latents :696-699: randn of the token-major shape)
    noise = torch.randn(1, f * h * w, 128, generator=torch.Generator("cpu").manual_seed(SEED_NOISE))
    per_step = []
    mine = O.denoise_loop(sd, cfg, noise, pe, pm, num_frames_lat=f, lat_h=h, lat_w=w, frame_rate=FPS, num_steps=STEPS,
                          per_step=per_step)
    _check("config0 final latents (28 layers)", O.unpatchify(mine, f, h, w), lat, tol=2e-4)
    mine_img = O.postprocess(O.vae_decode(vsd, O.unpatchify(mine, f, h, w)))
    ps = O.psnr(mine_img, img)
    print(f"  config0 decoded frames: PSNR(oracle, reference) = {ps:.1f} dB")
    assert ps > 70
    cores = torch.get_num_threads()
    print(f"  reference on this container's CPU ({cores} threads, fp32): denoise loop {t_loop:.2f} s "
          f"({STEPS / t_loop:.3f} steps/s), loop + decode {t_video:.2f} s/video")
    out = dict(meta=dict(H=H, W=W, F=FRAMES, fps=FPS, steps=STEPS, num_layers=LAYERS, latent=LAT, seed_weights=SEED_W,
                         seed_vae=SEED_VAE, seed_prompt=SEED_PROMPT, seed_noise=SEED_NOISE),
               latents=lat.clone(), per_step_oracle=[p.clone() for p in per_step], frames=img.to(torch.float16).clone(),
               reference_cpu=dict(cores=cores, loop_s=t_loop, video_s=t_video, steps_per_s=STEPS / t_loop))
    torch.save(out, os.path.join(GOLD, "ltx_config0.pt"))
    print("  wrote tests/golden/ltx_config0.pt", os.path.getsize(os.path.join(GOLD, "ltx_config0.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
