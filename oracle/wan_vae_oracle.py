"""CPU oracle for the Wan2.1 VAE decode and encode (TEST INFRASTRUCTURE ONLY — see ltx_oracle.py header).

Plain-PyTorch restatement of WanVAE.decode -> WanVAE_.decode -> Decoder3d (wan/modules/vae.py:386-493,578-609,
825-829) as ONE pass over the whole latent sequence.  The reference decodes one latent frame per call and threads a
feature cache through every causal convolution (CACHE_T = 2, :14, :210-227); that streaming is equivalent to:
  * every CausalConv3d = zero-padded (2 frames in front) causal convolution over the full sequence (:27-37);
  * Resample('upsample3d') (:107-143): the FIRST frame bypasses `time_conv` (the 'Rep' sentinel), the others go through
    `time_conv` as a causal conv over frames 1.. (zero front padding — frame 0 is NOT in their history), whose 2C output
    channels are interleaved into time (channels [0,C) -> frame 2k, [C,2C) -> frame 2k+1): T -> 1 + 2(T-1);
  * spatial part: nearest(-exact) x2 upsample in fp32 + Conv2d 3x3 per frame (:80-88);
  * RMS_norm = F.normalize(x, dim=channel) * sqrt(C) * gamma (:41-58);  AttentionBlock = per-frame single-head
    attention over h*w tokens with head dim C (:234-272).
Encode (WanVAE.encode -> WanVAE_.encode -> Encoder3d, :275-383, 536-575, 806-816): the reference feeds chunks of 1, 4, 4, ...
frames through the encoder with the same 2-frame caches; as ONE pass:
  * Resample('downsample2d'|'downsample3d') (:90-97): per frame ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2);
  * 'downsample3d' (:150-165): the FIRST frame bypasses `time_conv` (it only seeds the cache); every later chunk runs
    CausalConv3d((3,1,1), stride (2,1,1)) over [last frame of the previous chunk, chunk], so output frame m >= 1 reads
    input frames (2m-2, 2m-1, 2m): T -> 1 + (T-1)/2;
  * mu = first z_dim channels of conv1(encoder(x)); returned as (mu - mean) / std (:566-575).
Pinned by oracle/gen_golden_wan_vae.py against the unmodified reference run in its own streaming mode.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

WAN_VAE = dict(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, temperal_upsample=[True, True, False])   # vae.py:733-741
# WanVAE latent statistics (vae.py:766-776)
WAN_VAE_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632, -0.1922,
                -0.9497, 0.2503, -0.2921]
WAN_VAE_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253,
               2.8251, 1.9160]


def decoder_layout(cfg) -> List[tuple]:
    """The module list of Decoder3d.upsamples (vae.py:415-433) as ('res', cin, cout) / ('up3d'|'up2d', c) entries."""
    dm = cfg["dim_mult"]
    dims = [cfg["dim"] * u for u in [dm[-1]] + dm[::-1]]
    out = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(cfg["num_res_blocks"] + 1):
            out.append(("res", cin, cout))
            cin = cout
        if i != len(dm) - 1:
            out.append(("up3d" if cfg["temperal_upsample"][i] else "up2d", cout))
    return out


def make_wan_vae_decoder_state_dict(cfg=WAN_VAE, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded random init with the reference key names (WanVAE_: conv2.*, decoder.*).  AttentionBlock.proj is
    zero-initialised by the reference (:247), which would hide the attention path: overridden."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(name, cout, cin, k):
        fan = cin * math.prod(k)
        sd[name + ".weight"] = (torch.rand(cout, cin, *k, generator=g) * 2 - 1) / math.sqrt(fan)
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) / math.sqrt(fan)

    def gamma(name, c, images):
        sd[name + ".gamma"] = (1.0 + 0.1 * torch.randn(c, generator=g)).view(c, *([1, 1] if images else [1, 1, 1]))

    def res(p, cin, cout):
        gamma(p + "residual.0", cin, False); conv(p + "residual.2", cout, cin, (3, 3, 3))
        gamma(p + "residual.3", cout, False); conv(p + "residual.6", cout, cout, (3, 3, 3))
        if cin != cout:
            conv(p + "shortcut", cout, cin, (1, 1, 1))

    z = cfg["z_dim"]
    d0 = cfg["dim"] * cfg["dim_mult"][-1]
    conv("conv2", z, z, (1, 1, 1))
    conv("decoder.conv1", d0, z, (3, 3, 3))
    res("decoder.middle.0.", d0, d0)
    gamma("decoder.middle.1.norm", d0, True)
    conv("decoder.middle.1.to_qkv", 3 * d0, d0, (1, 1)); conv("decoder.middle.1.proj", d0, d0, (1, 1))
    res("decoder.middle.2.", d0, d0)
    c_last = d0
    for i, ent in enumerate(decoder_layout(cfg)):
        p = f"decoder.upsamples.{i}."
        if ent[0] == "res":
            res(p, ent[1], ent[2]); c_last = ent[2]
        else:
            c = ent[1]
            conv(p + "resample.1", c // 2, c, (3, 3))
            if ent[0] == "up3d":
                conv(p + "time_conv", 2 * c, c, (3, 1, 1))
            c_last = c // 2
    gamma("decoder.head.0", c_last, False)
    conv("decoder.head.2", 3, c_last, (3, 3, 3))
    return sd


def rms_norm(x: Tensor, gamma: Tensor) -> Tensor:
    """vae.py:52-58 (channel dim 1)"""
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma


def causal_conv3d(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """vae.py:17-37 over a whole sequence: zero padding, 2*pad_t frames in front, none behind."""
    kt, kh, kw = w.shape[2:]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0))
    return F.conv3d(x, w, b)


def res_block(sd, p, x: Tensor) -> Tensor:
    h = causal_conv3d(x, sd[p + "shortcut.weight"], sd[p + "shortcut.bias"]) if (p + "shortcut.weight") in sd else x
    y = causal_conv3d(F.silu(rms_norm(x, sd[p + "residual.0.gamma"])), sd[p + "residual.2.weight"], sd[p + "residual.2.bias"])
    y = causal_conv3d(F.silu(rms_norm(y, sd[p + "residual.3.gamma"])), sd[p + "residual.6.weight"], sd[p + "residual.6.bias"])
    return y + h


def attention_block(sd, p, x: Tensor) -> Tensor:
    """vae.py:249-272: per frame, one head of width C over h*w tokens."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = rms_norm(y, sd[p + "norm.gamma"])
    qkv = F.conv2d(y, sd[p + "to_qkv.weight"], sd[p + "to_qkv.bias"]).reshape(b * t, 3 * c, h * w).transpose(1, 2)   # [bt, hw, 3c]
    q, k, v = qkv.chunk(3, dim=-1)
    a = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1) @ v                                               # [bt, hw, c]
    y = F.conv2d(a.transpose(1, 2).reshape(b * t, c, h, w), sd[p + "proj.weight"], sd[p + "proj.bias"])
    return y.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4) + x


def resample(sd, p, x: Tensor, mode: str) -> Tensor:
    b, c, t, h, w = x.shape
    if mode == "up3d" and t > 1:
        y = causal_conv3d(x[:, :, 1:], sd[p + "time_conv.weight"], sd[p + "time_conv.bias"])        # [b, 2c, t-1, h, w]
        y = y.reshape(b, 2, c, t - 1, h, w)
        y = torch.stack((y[:, 0], y[:, 1]), 3).reshape(b, c, 2 * (t - 1), h, w)                        # :140-143
        x = torch.cat([x[:, :, :1], y], dim=2)
    t = x.shape[2]
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").to(x.dtype)
    y = F.conv2d(y, sd[p + "resample.1.weight"], sd[p + "resample.1.bias"], padding=1)
    return y.reshape(b, t, c // 2, 2 * h, 2 * w).permute(0, 2, 1, 3, 4)


def wan_vae_decode(sd: Dict[str, Tensor], z: Tensor, cfg=WAN_VAE, mean: Optional[Tensor] = None, std: Optional[Tensor] = None,
                   any_end_frame: bool = False) -> Tensor:
    """WanVAE.decode for one video (vae.py:578-609, 825-829): z [16, T, H, W] -> [3, 1 + 4(T-1), 8H, 8W] in [-1, 1] (float32).
    any_end_frame (:597-601): the LAST latent frame is decoded on its own, without the feature caches (one image), and appended:
    [16, T, H, W] -> [3, 1 + 4(T-2) + 1, 8H, 8W]."""
    x = z.unsqueeze(0)
    if mean is not None:
        x = x * std.view(1, -1, 1, 1, 1).to(x.dtype) + mean.view(1, -1, 1, 1, 1).to(x.dtype)      # z / (1/std) + mean (:581-586)
    x = causal_conv3d(x, sd["conv2.weight"], sd["conv2.bias"])
    if any_end_frame:
        parts = [_decoder(sd, x[:, :, :-1], cfg), _decoder(sd, x[:, :, -1:], cfg)]
        return torch.cat(parts, dim=2).clamp(-1, 1).float().squeeze(0)
    return _decoder(sd, x, cfg).clamp(-1, 1).float().squeeze(0)


def _decoder(sd: Dict[str, Tensor], x: Tensor, cfg) -> Tensor:
    """Decoder3d.forward over a whole (sub)sequence (vae.py:438-493)"""
    x = causal_conv3d(x, sd["decoder.conv1.weight"], sd["decoder.conv1.bias"])
    x = res_block(sd, "decoder.middle.0.", x)
    x = attention_block(sd, "decoder.middle.1.", x)
    x = res_block(sd, "decoder.middle.2.", x)
    for i, ent in enumerate(decoder_layout(cfg)):
        p = f"decoder.upsamples.{i}."
        x = res_block(sd, p, x) if ent[0] == "res" else resample(sd, p, x, ent[0])
    return causal_conv3d(F.silu(rms_norm(x, sd["decoder.head.0.gamma"])), sd["decoder.head.2.weight"], sd["decoder.head.2.bias"])


# ---------------------------------------------------------------------------------------------------------------------
# encode
# ---------------------------------------------------------------------------------------------------------------------
def encoder_layout(cfg) -> List[tuple]:
    """The module list of Encoder3d.downsamples (vae.py:302-318) as ('res', cin, cout) / ('down3d'|'down2d', c) entries."""
    dm = cfg["dim_mult"]
    dims = [cfg["dim"] * u for u in [1] + dm]
    tdown = cfg["temperal_upsample"][::-1]
    out = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(cfg["num_res_blocks"]):
            out.append(("res", cin, cout))
            cin = cout
        if i != len(dm) - 1:
            out.append(("down3d" if tdown[i] else "down2d", cout))
    return out


def make_wan_vae_encoder_state_dict(cfg=WAN_VAE, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded random init with the reference key names (WanVAE_: conv1.*, encoder.*); AttentionBlock.proj overridden as above."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(name, cout, cin, k):
        fan = cin * math.prod(k)
        sd[name + ".weight"] = (torch.rand(cout, cin, *k, generator=g) * 2 - 1) / math.sqrt(fan)
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) / math.sqrt(fan)

    def gamma(name, c, images):
        sd[name + ".gamma"] = (1.0 + 0.1 * torch.randn(c, generator=g)).view(c, *([1, 1] if images else [1, 1, 1]))

    def res(p, cin, cout):
        gamma(p + "residual.0", cin, False); conv(p + "residual.2", cout, cin, (3, 3, 3))
        gamma(p + "residual.3", cout, False); conv(p + "residual.6", cout, cout, (3, 3, 3))
        if cin != cout:
            conv(p + "shortcut", cout, cin, (1, 1, 1))

    z = cfg["z_dim"]
    conv("encoder.conv1", cfg["dim"], 3, (3, 3, 3))
    c = cfg["dim"]
    for i, ent in enumerate(encoder_layout(cfg)):
        p = f"encoder.downsamples.{i}."
        if ent[0] == "res":
            res(p, ent[1], ent[2]); c = ent[2]
        else:
            conv(p + "resample.1", c, c, (3, 3))
            if ent[0] == "down3d":
                conv(p + "time_conv", c, c, (3, 1, 1))
    res("encoder.middle.0.", c, c)
    gamma("encoder.middle.1.norm", c, True)
    conv("encoder.middle.1.to_qkv", 3 * c, c, (1, 1)); conv("encoder.middle.1.proj", c, c, (1, 1))
    res("encoder.middle.2.", c, c)
    gamma("encoder.head.0", c, False)
    conv("encoder.head.2", 2 * z, c, (3, 3, 3))
    conv("conv1", 2 * z, 2 * z, (1, 1, 1))
    return sd


def downsample(sd, p, x: Tensor, mode: str) -> Tensor:
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.conv2d(F.pad(y, (0, 1, 0, 1)), sd[p + "resample.1.weight"], sd[p + "resample.1.bias"], stride=2)      # :90-93
    y = y.reshape(b, t, c, y.shape[-2], y.shape[-1]).permute(0, 2, 1, 3, 4)
    if mode == "down3d" and t > 1:
        z = F.conv3d(y, sd[p + "time_conv.weight"], sd[p + "time_conv.bias"], stride=(2, 1, 1))                  # frames (2m-2, 2m-1, 2m)
        y = torch.cat([y[:, :, :1], z], dim=2)
    return y


def wan_vae_encode(sd: Dict[str, Tensor], video: Tensor, cfg=WAN_VAE, mean: Optional[Tensor] = None, std: Optional[Tensor] = None,
                   any_end_frame: bool = False) -> Tensor:
    """WanVAE.encode for one video (vae.py:536-575, 806-816, tile_size 0): video [3, 1+4k, H, W] in [-1, 1] ->
    mu [z_dim, 1+k, H/8, W/8] float32, normalised with (mu - mean) / std.
    any_end_frame (:541-542, 553-557): the LAST frame is encoded on its own, without the feature caches, and appended:
    video [3, 1+4k+1, H, W] -> mu [z_dim, 1+k+1, H/8, W/8]."""
    x = video.unsqueeze(0)
    if any_end_frame:
        assert (video.shape[1] - 2) % 4 == 0
        x = torch.cat([_encoder(sd, x[:, :, :-1], cfg), _encoder(sd, x[:, :, -1:], cfg)], dim=2)
    else:
        assert (video.shape[1] - 1) % 4 == 0, "the reference's 1,4,4,... chunking drops trailing frames otherwise"
        x = _encoder(sd, x, cfg)
    mu = causal_conv3d(x, sd["conv1.weight"], sd["conv1.bias"])[:, : cfg["z_dim"]]
    if mean is not None:
        mu = (mu - mean.view(1, -1, 1, 1, 1).to(mu.dtype)) * (1.0 / std).view(1, -1, 1, 1, 1).to(mu.dtype)
    return mu.float().squeeze(0)


def _encoder(sd: Dict[str, Tensor], x: Tensor, cfg) -> Tensor:
    """Encoder3d.forward over a whole (sub)sequence (vae.py:329-383)"""
    x = causal_conv3d(x, sd["encoder.conv1.weight"], sd["encoder.conv1.bias"])
    for i, ent in enumerate(encoder_layout(cfg)):
        p = f"encoder.downsamples.{i}."
        x = res_block(sd, p, x) if ent[0] == "res" else downsample(sd, p, x, ent[0])
    x = res_block(sd, "encoder.middle.0.", x)
    x = attention_block(sd, "encoder.middle.1.", x)
    x = res_block(sd, "encoder.middle.2.", x)
    return causal_conv3d(F.silu(rms_norm(x, sd["encoder.head.0.gamma"])), sd["encoder.head.2.weight"], sd["encoder.head.2.bias"])
