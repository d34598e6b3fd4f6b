"""Pin oracle/wan_vae_oracle.py against the UNMODIFIED reference WanVAE_ (wan/modules/vae.py) run the way the reference runs
it — one latent frame per decoder call with the CACHE_T feature cache — and write tests/golden/wan_vae_decode.pt.
Build container only (needs /root/reference):  python oracle/gen_golden_wan_vae.py"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
import load_reference  # noqa: E402

load_reference.install()
from oracle import wan_vae_oracle as V  # noqa: E402
from oracle.ltx_oracle import rel_l2  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
torch.set_grad_enabled(False)

TINY = dict(V.WAN_VAE, dim=32)          # channels 128 / 64 / 32: every code path of the 96-wide model, 27x fewer FLOPs


def build_ref(cfg, sd):
    from wan.modules.vae import WanVAE_
    m = WanVAE_(dim=cfg["dim"], z_dim=cfg["z_dim"], dim_mult=cfg["dim_mult"], num_res_blocks=cfg["num_res_blocks"], attn_scales=[],
                temperal_downsample=cfg["temperal_upsample"][::-1], dropout=0.0)
    own = m.state_dict()
    pre = ("decoder.", "conv2.") if "conv2.weight" in sd else ("encoder.", "conv1.")
    missing = [k for k in own if k.startswith(pre) and k not in sd]
    assert not missing, missing
    unexpected = [k for k in sd if k not in own]
    assert not unexpected, unexpected
    for k, v in sd.items():
        assert own[k].shape == v.shape, (k, own[k].shape, v.shape)
    m.load_state_dict(sd, strict=False)
    return m.eval()


def main():
    cfg = TINY
    sd = V.make_wan_vae_decoder_state_dict(cfg, seed=0)
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 2e-5)):
        sdd = {k: v.to(dtype) for k, v in sd.items()}
        ref = build_ref(cfg, sdd).to(dtype)
        g = torch.Generator().manual_seed(7)
        z = torch.randn(16, 4, 6, 10, generator=g).to(dtype)
        mean, std = torch.tensor(V.WAN_VAE_MEAN, dtype=dtype), torch.tensor(V.WAN_VAE_STD, dtype=dtype)
        y_ref = ref.decode(z.unsqueeze(0), [mean, 1.0 / std]).clamp_(-1, 1).float().squeeze(0)      # WanVAE.decode (:825-829), tile_size 0
        y = V.wan_vae_decode(sdd, z, cfg, mean, std)
        e = rel_l2(y, y_ref)
        print(f"  wan vae decode ({dtype}): out {tuple(y.shape)}, rel_l2(oracle batched, reference streaming) = {e:.3e}, clamp hits {(y_ref.abs() >= 1).float().mean():.3f}")
        assert y.shape == (3, 13, 48, 80) and e < tol
        # any_end_frame: the last latent frame is an independent image (no feature caches)
        y_ref_e = ref.decode(z.unsqueeze(0), [mean, 1.0 / std], any_end_frame=True).clamp_(-1, 1).float().squeeze(0)
        y_e = V.wan_vae_decode(sdd, z, cfg, mean, std, any_end_frame=True)
        ee = rel_l2(y_e, y_ref_e)
        print(f"  wan vae decode any_end_frame ({dtype}): out {tuple(y_e.shape)}, rel_l2 = {ee:.3e}")
        assert y_e.shape == (3, 10, 48, 80) and ee < tol
    torch.save(dict(cfg=cfg, seed_weights=0, z=z.float(), out=y_ref.half(), out_end_frame=y_ref_e.half()), os.path.join(GOLD, "wan_vae_decode.pt"))
    print("written", os.path.join(GOLD, "wan_vae_decode.pt"))

    # ---- encode: the reference streams chunks of 1, 4, 4, ... frames with its feature caches (WanVAE_.encode :536-575)
    sd = V.make_wan_vae_encoder_state_dict(cfg, seed=1)
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 2e-5)):
        sdd = {k: v.to(dtype) for k, v in sd.items()}
        ref = build_ref(cfg, sdd).to(dtype)
        g = torch.Generator().manual_seed(11)
        for shape in ((3, 9, 48, 80), (3, 1, 32, 48), (3, 13, 32, 32)):
            video = (torch.rand(*shape, generator=g) * 2 - 1).to(dtype)
            mean, std = torch.tensor(V.WAN_VAE_MEAN, dtype=dtype), torch.tensor(V.WAN_VAE_STD, dtype=dtype)
            mu_ref = ref.encode(video.unsqueeze(0), [mean, 1.0 / std]).float().squeeze(0)                # WanVAE.encode (:806-816), tile_size 0
            mu = V.wan_vae_encode(sdd, video, cfg, mean, std)
            e = rel_l2(mu, mu_ref)
            print(f"  wan vae encode ({dtype}) {shape}: out {tuple(mu.shape)}, rel_l2(oracle one-pass, reference streaming) = {e:.3e}")
            assert mu.shape == (16, 1 + (shape[1] - 1) // 4, shape[2] // 8, shape[3] // 8) and e < tol
            if shape[1] == 9:
                keep = dict(video=video.float(), mu=mu_ref.float())
                video_e = torch.cat([video, (torch.rand(3, 1, *shape[2:], generator=g) * 2 - 1).to(dtype)], dim=1)      # + an end frame
                mu_ref_e = ref.encode(video_e.unsqueeze(0), [mean, 1.0 / std], any_end_frame=True).float().squeeze(0)
                mu_e = V.wan_vae_encode(sdd, video_e, cfg, mean, std, any_end_frame=True)
                ee = rel_l2(mu_e, mu_ref_e)
                print(f"  wan vae encode any_end_frame ({dtype}): out {tuple(mu_e.shape)}, rel_l2 = {ee:.3e}")
                assert mu_e.shape == (16, 4, 6, 10) and ee < tol
                keep.update(video_end_frame=video_e.float(), mu_end_frame=mu_ref_e.float())
    torch.save(dict(cfg=cfg, seed_weights=1, **keep), os.path.join(GOLD, "wan_vae_encode.pt"))
    print("written", os.path.join(GOLD, "wan_vae_encode.pt"))


if __name__ == "__main__":
    main()
