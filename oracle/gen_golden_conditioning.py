"""Keyframe / sequence conditioning: the product's `LTXVideoPipeline.prepare_conditioning` (host-side torch code, run here on the CPU with
pre-encoded conditioning latents) against the UNMODIFIED reference method (pipeline_ltx_video.py:1344-1548, 1614-1687) whose `vae_encode` is
patched to return the same latents — bit-exact for latents, pixel coordinates, mask and the extra-token count.
Build container only (needs /root/reference):  python oracle/gen_golden_conditioning.py"""
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
import load_reference  # noqa: E402

load_reference.install()


def main():
    import ltx_video.pipelines.pipeline_ltx_video as R
    from ltx_video.models.autoencoders.causal_video_autoencoder import CausalVideoAutoencoder as RefVAE
    from ltx_video.models.transformers.symmetric_patchifier import SymmetricPatchifier as RefPatchifier
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import ConditioningItem, LTXVideoPipeline
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier

    F_l, H_l, W_l = 5, 4, 6                      # 33 frames, 128 x 192
    num_frames, height, width = 33, 128, 192
    g = torch.Generator().manual_seed(0)
    enc = {1: torch.randn(1, 128, 1, H_l, W_l, generator=g), 9: torch.randn(1, 128, 2, H_l, W_l, generator=g),
           17: torch.randn(1, 128, 3, H_l, W_l, generator=g)}             # latents of 1-, 9- and 17-frame media

    # ---- the reference method, bound to a stand-in object; vae_encode answers with the prepared latents
    from ltx_video.utils.diffusers_config_mapping import OURS_VAE_CONFIG
    ref_vae = RefVAE.from_config(dict(OURS_VAE_CONFIG)).eval()      # only isinstance(), dtype / device and the scale factors are used
    calls = []

    def fake_encode(media, vae, vae_per_channel_normalize=False):
        calls.append(media.shape[2])
        return enc[media.shape[2]].clone()
    R.vae_encode = fake_encode
    fake_self = SimpleNamespace(vae_scale_factor=32, vae=ref_vae, patchifier=RefPatchifier(patch_size=1),
                                transformer=SimpleNamespace(config=SimpleNamespace(causal_temporal_positioning=True), use_tpu_flash_attention=False),
                                _resize_conditioning_item=R.LTXVideoPipeline._resize_conditioning_item,
                                _handle_non_first_conditioning_sequence=R.LTXVideoPipeline._handle_non_first_conditioning_sequence)
    import types
    fake_self._get_latent_spatial_position = types.MethodType(R.LTXVideoPipeline._get_latent_spatial_position, fake_self)

    ours = LTXVideoPipeline.__new__(LTXVideoPipeline)
    ours.vae = SimpleNamespace(spatial_downscale_factor=32, temporal_downscale_factor=8)
    ours.patchifier = SymmetricPatchifier(1)
    ours.transformer = SimpleNamespace(config=SimpleNamespace(causal_temporal_positioning=True))

    cases = {
        "first_frame": [(1, 0, 1.0)],
        "keyframe_16": [(1, 16, 0.8)],
        "first_and_last": [(1, 0, 1.0), (1, 32, 1.0)],
        "sequence_at_8": [(17, 8, 1.0)],
        "prefix_only_sequence": [(9, 16, 0.5)],
    }
    out = {}
    for name, items in cases.items():
        init = torch.randn(1, 128, F_l, H_l, W_l, generator=torch.Generator().manual_seed(1))
        ref_items = [R.ConditioningItem(media_item=torch.zeros(1, 3, n, height, width), media_frame_number=f, conditioning_strength=s)
                     for n, f, s in items]
        r = R.LTXVideoPipeline.prepare_conditioning(fake_self, ref_items, init.clone(), num_frames, height, width,
                                                    vae_per_channel_normalize=True, generator=torch.Generator().manual_seed(2))
        our_items = [ConditioningItem(latents=enc[n].clone(), media_frame_number=f, conditioning_strength=s) for n, f, s in items]
        o = ours.prepare_conditioning(our_items, init.clone(), num_frames, height, width, vae_per_channel_normalize=True,
                                      generator=torch.Generator().manual_seed(2))
        assert torch.equal(o[0], r[0]) and torch.equal(o[1].to(r[1].dtype), r[1]) and torch.equal(o[2], r[2]) and o[3] == r[3], name
        print(f"  prepare_conditioning[{name}]: tokens {tuple(r[0].shape)}, extra tokens {r[3]}: bit-exact")
        out[name] = dict(items=items, tokens=r[0], coords=r[1], mask=r[2], extra=r[3])
    torch.save(dict(enc=enc, cases=out, geom=(F_l, H_l, W_l, num_frames, height, width)), os.path.join(ROOT, "tests", "golden", "ltx_conditioning.pt"))
    print("written tests/golden/ltx_conditioning.pt")


def keyframe_loop():
    """The whole call with a keyframe in the middle of the video, through the reference's OWN __call__ (its vae_encode answering with fixed
    latents): extra conditioning tokens with their own pixel coordinates and per-token timesteps are prepended (:1449-1503), carried through
    the loop and dropped before unpatchify (:1258-1262).  The composition the GPU test checks the CUDA path against — product
    prepare_conditioning -> oracle denoise_loop(pixel_coords, conditioning_mask) -> drop the extra tokens — must give the same latents."""
    import ltx_video.pipelines.pipeline_ltx_video as R
    from ltx_video.models.transformers.symmetric_patchifier import SymmetricPatchifier as RefPatchifier
    from ltx_video.schedulers.rf import RectifiedFlowScheduler
    from ltx_video.utils.diffusers_config_mapping import OURS_SCHEDULER_CONFIG
    from ltx_video_gpupoor_b200.ltx.pipeline_ltx_video import ConditioningItem, LTXVideoPipeline
    from ltx_video_gpupoor_b200.ltx.symmetric_patchifier import SymmetricPatchifier
    from oracle import ltx_oracle as O
    from oracle.gen_golden import _NoInterrupt, _check, _cuda_to_cpu, build_ref_transformer, build_ref_vae
    torch.set_grad_enabled(False)
    L = 2
    sd = O.make_transformer_state_dict(O.LTX_2B, seed=0, num_layers=L)
    tr = build_ref_transformer(L, sd)
    vae = build_ref_vae(O.make_vae_decoder_state_dict(seed=1))
    g = torch.Generator().manual_seed(3)
    pe, pm = torch.randn(1, 16, 4096, generator=g), torch.ones(1, 16)
    key = torch.randn(1, 128, 1, 4, 6, generator=g)
    R.vae_encode = lambda media, vae_, vae_per_channel_normalize=False: key.clone()
    pipe = R.LTXVideoPipeline(tokenizer=None, text_encoder=None, vae=vae, transformer=tr,
                              scheduler=RectifiedFlowScheduler.from_config(dict(OURS_SCHEDULER_CONFIG)), patchifier=RefPatchifier(patch_size=1),
                              prompt_enhancer_image_caption_model=None, prompt_enhancer_image_caption_processor=None,
                              prompt_enhancer_llm_model=None, prompt_enhancer_llm_tokenizer=None)
    H, W, F_, fps, steps, fno = 128, 192, 33, 25.0, 3, 16            # latent (1,128,5,4,6) + 24 extra tokens
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        with _cuda_to_cpu():
            lat = pipe(height=H, width=W, num_frames=F_, frame_rate=fps, prompt_embeds=pe, prompt_attention_mask=pm,
                       negative_prompt_embeds=None, negative_prompt_attention_mask=None, num_inference_steps=steps,
                       generator=torch.Generator().manual_seed(5), output_type="latent", return_dict=False, joint_pass=True,
                       ltxv_model=_NoInterrupt(), is_video=True, vae_per_channel_normalize=True, guidance_scale=1.0, stg_scale=0.0,
                       rescaling_scale=1.0, image_cond_noise_scale=0.0,
                       conditioning_items=[R.ConditioningItem(media_item=torch.zeros(1, 3, 1, H, W), media_frame_number=fno,
                                                              conditioning_strength=1.0)])[0]
    finally:
        os.chdir(cwd)
    ours = LTXVideoPipeline.__new__(LTXVideoPipeline)
    ours.vae = SimpleNamespace(spatial_downscale_factor=32, temporal_downscale_factor=8)
    ours.patchifier = SymmetricPatchifier(1)
    ours.transformer = SimpleNamespace(config=SimpleNamespace(causal_temporal_positioning=bool(getattr(tr.config, "causal_temporal_positioning", False))))
    gen = torch.Generator().manual_seed(5)                           # one stream: initial noise, then the keyframe noise (:1466-1471)
    init = O.unpatchify(torch.randn(1, 120, 128, generator=gen), 5, 4, 6)
    tok, px, cm, extra = ours.prepare_conditioning([ConditioningItem(latents=key.clone(), media_frame_number=fno, conditioning_strength=1.0)],
                                                   init.clone(), F_, H, W, vae_per_channel_normalize=True, generator=gen)
    assert extra == 24
    mine = O.denoise_loop(sd, O.LTX_2B, tok.float(), pe, pm, num_frames_lat=5, lat_h=4, lat_w=6, frame_rate=fps, num_steps=steps,
                          conditioning_mask=cm, pixel_coords=px)
    _check("keyframe call (reference __call__ vs prepare_conditioning + oracle loop)", O.unpatchify(mine[:, extra:], 5, 4, 6), lat, tol=5e-5)
    torch.save(dict(meta=dict(H=H, W=W, F=F_, fps=fps, steps=steps, num_layers=L, frame=fno), pe=pe, pm=pm, key=key, noise_seed=5, latents=lat.clone()),
               os.path.join(ROOT, "tests", "golden", "ltx_keyframe_loop.pt"))
    print("written tests/golden/ltx_keyframe_loop.pt")


if __name__ == "__main__":
    main()
    keyframe_loop()
