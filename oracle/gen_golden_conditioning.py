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


if __name__ == "__main__":
    main()
