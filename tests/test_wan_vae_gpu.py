"""GPU parity of the Wan VAE decode drop-in against the fixture recorded from the unmodified reference (streaming decode) and the
live oracle; kernels new to this path (tap-shaped causal conv with zero padding, RMS_norm+SiLU, x2 upsample, row softmax) vs torch."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

from ltx_video_gpupoor_b200 import ops  # noqa: E402
from ltx_video_gpupoor_b200.wan.vae import WanVAE  # noqa: E402
from oracle import wan_vae_oracle as V  # noqa: E402
from oracle.ltx_oracle import psnr, rel_l2  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


def rnd(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(BF).to(DEV)


@pytest.mark.parametrize("kt,khw,Cin,Cout", [(3, 3, 64, 128), (1, 3, 128, 64), (3, 1, 64, 128), (1, 1, 192, 64)])
def test_conv_taps_zero_causal(kt, khw, Cin, Cout):
    B, T, H, W = 1, 4, 9, 13
    x = rnd(B, T, H, W, Cin, seed=1)
    w5 = rnd(Cout, Cin, kt, khw, khw, seed=2, scale=(kt * khw * khw * Cin) ** -0.5)
    b = rnd(Cout, seed=3)
    xr = F.pad(x.float().permute(0, 4, 1, 2, 3), (khw // 2, khw // 2, khw // 2, khw // 2, kt - 1, 0))
    ref = F.conv3d(xr, w5.float(), b.float()).permute(0, 2, 3, 4, 1)
    wp = w5.permute(0, 2, 3, 4, 1).reshape(Cout, -1).contiguous()
    out = ops.conv_taps(x, wp, b, kt, khw, True)
    torch.cuda.synchronize()
    assert rel_l2(out.float().cpu(), ref.cpu()) < 6e-3
    res = rnd(B, T, H, W, Cout, seed=4)
    out = ops.conv_taps(x, wp, b, kt, khw, True, residual=res)
    assert rel_l2(out.float().cpu(), (ref + res.float()).cpu()) < 6e-3


@pytest.mark.parametrize("C,c_real", [(64, 64), (128, 96), (192, 192), (384, 384)])
def test_l2norm_silu_upsample_softmax(C, c_real):
    x = rnd(3, 5, 7, C, seed=1)
    x[..., c_real:] = 0
    g = rnd(C, seed=2, scale=0.1) + 1
    g[c_real:] = 0
    ref = F.normalize(x.float()[..., :c_real], dim=-1) * c_real ** 0.5 * g.float()[:c_real]
    y = ops.l2norm_silu(x, g, c_real, silu=True)
    assert rel_l2(y.float()[..., :c_real].cpu(), F.silu(ref).cpu()) < 6e-3
    assert float(y.float()[..., c_real:].abs().max() if c_real < C else 0.0) == 0.0
    up = ops.upsample2x(x)
    assert torch.equal(up, x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))
    s = torch.randn(37, 333, generator=torch.Generator().manual_seed(3)).to(DEV) * 4
    p = ops.softmax_rows(s, 0.3)
    assert rel_l2(p.float().cpu(), torch.softmax(s.cpu() * 0.3, dim=-1)) < 6e-3


def test_wan_vae_decode_vs_reference_fixture(golden_dir):
    g = torch.load(os.path.join(golden_dir, "wan_vae_decode.pt"), weights_only=False)
    cfg = g["cfg"]
    sd = V.make_wan_vae_decoder_state_dict(cfg, seed=g["seed_weights"])
    vae = WanVAE(dim=cfg["dim"], dim_mult=cfg["dim_mult"], num_res_blocks=cfg["num_res_blocks"],
                 temperal_downsample=cfg["temperal_upsample"][::-1])
    vae.load_state_dict(sd)
    y = vae.decode([g["z"].to(DEV)], tile_size=0)[0]
    torch.cuda.synchronize()
    ref = g["out"].float()
    assert tuple(y.shape) == tuple(ref.shape) == (3, 13, 48, 80)
    p = psnr(y.cpu() * 0.5 + 0.5, ref * 0.5 + 0.5)
    e = rel_l2(y.cpu(), ref)
    print(f"wan vae decode: PSNR vs reference = {p:.1f} dB, rel_l2 = {e:.3e}")
    assert p >= 40.0
    # and against the live oracle in fp32 on the same weights
    mean, std = torch.tensor(V.WAN_VAE_MEAN), torch.tensor(V.WAN_VAE_STD)
    yo = V.wan_vae_decode(sd, g["z"], cfg, mean, std)
    assert psnr(y.cpu() * 0.5 + 0.5, yo * 0.5 + 0.5) >= 40.0


@pytest.mark.parametrize("kt,khw,st,shw,off,H,W", [(1, 3, 1, 2, 1, 12, 20), (1, 3, 1, 2, 1, 9, 13), (3, 1, 2, 1, 0, 6, 10), (3, 3, 1, 1, 0, 5, 7)])
def test_conv_taps_strided(kt, khw, st, shw, off, H, W):
    """Wan encoder Resample pieces (vae.py:90-97,150-165): ZeroPad2d((0,1,0,1)) + Conv2d stride 2; time_conv stride (2,1,1)."""
    B, T, Cin, Cout = 1, 9, 64, 128
    x = rnd(B, T, H, W, Cin, seed=1)
    w5 = rnd(Cout, Cin, kt, khw, khw, seed=2, scale=(kt * khw * khw * Cin) ** -0.5)
    b = rnd(Cout, seed=3)
    xr = x.float().permute(0, 4, 1, 2, 3)
    lo = khw // 2 - off
    xr = F.pad(xr, (lo, khw // 2, lo, khw // 2, kt - 1, 0))
    ref = F.conv3d(xr, w5.float(), b.float(), stride=(st, shw, shw)).permute(0, 2, 3, 4, 1)
    out = ops.conv_taps_strided(x, w5.permute(0, 2, 3, 4, 1).reshape(Cout, -1).contiguous(), b, kt, khw, st, shw, off)
    torch.cuda.synchronize()
    assert tuple(out.shape) == tuple(ref.shape)
    assert rel_l2(out.float().cpu(), ref.cpu()) < 6e-3


def test_wan_vae_encode_vs_reference_fixture(golden_dir):
    g = torch.load(os.path.join(golden_dir, "wan_vae_encode.pt"), weights_only=False)
    cfg = g["cfg"]
    sd = V.make_wan_vae_encoder_state_dict(cfg, seed=g["seed_weights"])
    vae = WanVAE(dim=cfg["dim"], dim_mult=cfg["dim_mult"], num_res_blocks=cfg["num_res_blocks"],
                 temperal_downsample=cfg["temperal_upsample"][::-1])
    vae.load_state_dict(sd)
    mu = vae.encode([g["video"].to(DEV)], tile_size=0)[0]
    torch.cuda.synchronize()
    assert tuple(mu.shape) == tuple(g["mu"].shape) == (16, 3, 6, 10)
    e = rel_l2(mu.cpu(), g["mu"])
    print(f"wan vae encode: rel_l2 vs reference fixture = {e:.3e}")
    assert e < 2e-2
    # single image (the i2v conditioning frame) and the first-frame causality of the one-pass form
    mu1 = vae.encode([g["video"][:, :1].to(DEV)])[0]
    assert rel_l2(mu1.cpu(), g["mu"][:, :1]) < 2e-2
    with pytest.raises(ValueError):
        vae.encode([g["video"][:, :7].to(DEV)])
    with pytest.raises(RuntimeError):
        v2 = WanVAE(dim=cfg["dim"], dim_mult=cfg["dim_mult"], num_res_blocks=cfg["num_res_blocks"], temperal_downsample=cfg["temperal_upsample"][::-1])
        v2.load_state_dict(V.make_wan_vae_decoder_state_dict(cfg, seed=0))
        v2.encode([g["video"].to(DEV)])


def test_wan_vae_any_end_frame_vs_reference_fixture(golden_dir):
    """WanVAE.encode / decode with any_end_frame=True (last frame coded without the feature caches) against the reference's outputs."""
    g = torch.load(os.path.join(golden_dir, "wan_vae_encode.pt"), weights_only=False)
    cfg = g["cfg"]
    mk = lambda: WanVAE(dim=cfg["dim"], dim_mult=cfg["dim_mult"], num_res_blocks=cfg["num_res_blocks"], temperal_downsample=cfg["temperal_upsample"][::-1])
    vae = mk()
    vae.load_state_dict(V.make_wan_vae_encoder_state_dict(cfg, seed=g["seed_weights"]))
    mu = vae.encode([g["video_end_frame"].to(DEV)], tile_size=0, any_end_frame=True)[0]
    assert tuple(mu.shape) == (16, 4, 6, 10)
    e = rel_l2(mu.cpu(), g["mu_end_frame"])
    print(f"wan vae encode any_end_frame: rel_l2 vs reference = {e:.3e}")
    assert e < 2e-2
    d = torch.load(os.path.join(golden_dir, "wan_vae_decode.pt"), weights_only=False)
    vae2 = mk()
    vae2.load_state_dict(V.make_wan_vae_decoder_state_dict(d["cfg"], seed=d["seed_weights"]))
    y = vae2.decode([d["z"].to(DEV)], tile_size=0, any_end_frame=True)[0]
    torch.cuda.synchronize()
    assert tuple(y.shape) == (3, 10, 48, 80)
    p = psnr(y.cpu() * 0.5 + 0.5, d["out_end_frame"].float() * 0.5 + 0.5)
    print(f"wan vae decode any_end_frame: PSNR vs reference = {p:.1f} dB")
    assert p >= 40.0
