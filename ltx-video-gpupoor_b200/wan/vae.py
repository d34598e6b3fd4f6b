"""WanVAE — B200-native drop-in for wan/modules/vae.py: `WanVAE.decode` (:825-829 → `WanVAE_.decode` :578-609 →
`Decoder3d.forward` :438-493) and `WanVAE.encode` (:806-816 → `WanVAE_.encode` :536-575 → `Encoder3d.forward` :329-383).

The reference decodes one latent frame per call and carries a 2-frame feature cache through every causal convolution
(CACHE_T, :14).  That streaming is equivalent to one pass over the whole sequence (proved against the unmodified reference
by oracle/gen_golden_wan_vae.py, 0.0 difference in fp64), which is what runs here:
  * activations are NDHWC bf16 with channels padded to a multiple of 64 (96 → 128, 16 → 64, 3 → 8; pad channels carry
    zeros, their weights / gammas are zero);
  * every CausalConv3d is the TMA-tiled implicit-GEMM tcgen05 kernel with zero temporal padding (`ltxb200_conv_taps_bf16`);
    the Conv2d of Resample is the same kernel with one temporal tap, `time_conv` with one spatial tap; the residual add is
    fused in the epilogue;
  * Resample('upsample3d'): frame 0 bypasses `time_conv` (the 'Rep' branch, :107-113), frames 1.. go through it and their 2C
    output channels are interleaved into time (:140-143);
  * RMS_norm + SiLU: one memory-bound kernel; the single 384-wide attention head of the middle block: two GEMMs around a row
    softmax (per frame).
Encode: the reference streams chunks of 1, 4, 4, ... frames; as one pass (oracle/wan_vae_oracle.py, exact in fp64):
  * Resample('downsample2d'|'downsample3d'): ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2) = the same implicit-GEMM kernel with a
    striding TMA box and taps (h, h+1, h+2) (`ltxb200_conv_taps_strided_bf16`, off_hw = 1); `time_conv` (3,1,1) stride 2 =
    the temporal-stride variant, whose output frame 0 is replaced by the bypassed first frame (:150-165);
  * the 3 input channels are zero-padded to 64 (one k-block per tap).
Same state_dict keys as `WanVAE_` (conv1.*, conv2.*, encoder.*, decoder.*); either half may be absent.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from .. import ops
from ..module_like import ModuleLike

BF16 = torch.bfloat16
WAN_VAE_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632, -0.1922,
                -0.9497, 0.2503, -0.2921]                                                  # vae.py:766-769
WAN_VAE_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253,
               2.8251, 1.9160]                                                             # vae.py:770-773


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def _pad8(c: int) -> int:
    return (c + 7) // 8 * 8


class WanVAE(ModuleLike):
    def __init__(self, z_dim: int = 16, vae_pth: Optional[str] = None, dtype=torch.float, device="cuda", dim: int = 96,
                 dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2, temperal_downsample=(False, True, True)):
        self.z_dim, self.dim, self.dim_mult, self.num_res_blocks = z_dim, dim, list(dim_mult), num_res_blocks
        self.temperal_upsample = list(temperal_downsample)[::-1]
        self.dtype, self.device = dtype, torch.device(device)
        self.mean = torch.tensor(WAN_VAE_MEAN[:z_dim], dtype=torch.float32)
        self.std = torch.tensor(WAN_VAE_STD[:z_dim], dtype=torch.float32)
        self.scale = [self.mean, 1.0 / self.std]
        self.w: Dict[str, torch.Tensor] = {}
        if vae_pth is not None:
            raise NotImplementedError("checkpoint loading is out of scope: call load_state_dict(state_dict)")

    # ---------------------------------------------------------------------------------------------
    @staticmethod
    def get_VAE_tile_size(vae_config, device_mem_capacity, mixed_precision):
        """wan/modules/vae.py:789-811 answers 0 (no tiling) from 24 GB up; `tile_size` is accepted by encode / decode and selects nothing."""
        return 0

    def _layout(self):
        dm = self.dim_mult
        dims = [self.dim * u for u in [dm[-1]] + dm[::-1]]
        out = []
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            if i in (1, 2, 3):
                cin = cin // 2
            for _ in range(self.num_res_blocks + 1):
                out.append(("res", cin, cout))
                cin = cout
            if i != len(dm) - 1:
                out.append(("up3d" if self.temperal_upsample[i] else "up2d", cout))
        return out

    def _enc_layout(self):
        """Encoder3d.downsamples (vae.py:302-318)"""
        dm = self.dim_mult
        dims = [self.dim * u for u in [1] + dm]
        tdown = self.temperal_upsample[::-1]
        out = []
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(self.num_res_blocks):
                out.append(("res", cin, cout))
                cin = cout
            if i != len(dm) - 1:
                out.append(("down3d" if tdown[i] else "down2d", cout))
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor], device=None, strict: bool = True):
        dev = self.device = torch.device(device) if device is not None else self.device
        w: Dict[str, torch.Tensor] = {}

        def conv(name, pad_out=_pad64):
            """[Cout, Cin, (kt,) kh, kw] -> [Cout_p, taps*Cin_p] tap-major bf16 (+ bias [Cout_p])"""
            wt = sd[name + ".weight"].float()
            if wt.dim() == 4:
                wt = wt.unsqueeze(2)
            co, ci = wt.shape[:2]
            cop, cip = pad_out(co), _pad64(ci)
            full = torch.zeros(cop, *wt.shape[2:], cip)
            full[:co, ..., :ci] = wt.permute(0, 2, 3, 4, 1)
            b = torch.zeros(cop)
            b[:co] = sd[name + ".bias"].float()
            w[name + ".w"] = full.reshape(cop, -1).to(device=dev, dtype=BF16).contiguous()
            w[name + ".b"] = b.to(device=dev, dtype=BF16)
            w[name + ".taps"] = (wt.shape[2], wt.shape[3])

        def gamma(name):
            g = sd[name + ".gamma"].float().flatten()
            full = torch.zeros(_pad64(g.numel()))
            full[: g.numel()] = g
            w[name] = full.to(device=dev, dtype=BF16)

        def res(p, cin, cout):
            gamma(p + "residual.0"); conv(p + "residual.2"); gamma(p + "residual.3"); conv(p + "residual.6")
            if cin != cout:
                conv(p + "shortcut")

        d0 = self.dim * self.dim_mult[-1]
        if "encoder.conv1.weight" in sd:
            conv("encoder.conv1")
            for i, ent in enumerate(self._enc_layout()):
                p = f"encoder.downsamples.{i}."
                if ent[0] == "res":
                    res(p, ent[1], ent[2])
                else:
                    conv(p + "resample.1")
                    if ent[0] == "down3d":
                        conv(p + "time_conv")
            res("encoder.middle.0.", d0, d0)
            gamma("encoder.middle.1.norm"); conv("encoder.middle.1.to_qkv"); conv("encoder.middle.1.proj")
            res("encoder.middle.2.", d0, d0)
            gamma("encoder.head.0"); conv("encoder.head.2"); conv("conv1")
        if "decoder.conv1.weight" not in sd:
            if strict and "encoder.conv1.weight" not in sd:
                raise KeyError("state_dict holds neither encoder.* nor decoder.* keys")
            self.w = w
            return [], []
        conv("conv2"); conv("decoder.conv1")
        res("decoder.middle.0.", d0, d0)
        gamma("decoder.middle.1.norm"); conv("decoder.middle.1.to_qkv"); conv("decoder.middle.1.proj")
        res("decoder.middle.2.", d0, d0)
        for i, ent in enumerate(self._layout()):
            p = f"decoder.upsamples.{i}."
            if ent[0] == "res":
                res(p, ent[1], ent[2])
            else:
                conv(p + "resample.1")
                if ent[0] == "up3d":
                    conv(p + "time_conv")
        gamma("decoder.head.0"); conv("decoder.head.2", pad_out=_pad8)
        self.w = w
        return [], []

    # ---------------------------------------------------------------------------------------------
    def _conv(self, name, x, residual=None):
        kt, khw = self.w[name + ".taps"]
        return ops.conv_taps(x, self.w[name + ".w"], self.w[name + ".b"], kt, khw, True, residual)

    def _res(self, p, x, cin, cout):
        h = self._conv(p + "shortcut", x) if cin != cout else x
        y = self._conv(p + "residual.2", ops.l2norm_silu(x, self.w[p + "residual.0"], cin))
        return self._conv(p + "residual.6", ops.l2norm_silu(y, self.w[p + "residual.3"], cout), residual=h)

    def _attn(self, p, x, c):
        """AttentionBlock (vae.py:234-272) per frame: x [1, T, H, W, C]."""
        _, T, H, W, C = x.shape
        assert C == c, "the attention block sits at a 64-multiple width (384 in the reference config)"
        hw = H * W
        hw8 = (hw + 7) // 8 * 8                                          # GEMM N / K granularity; pad keys with zeros
        y = ops.l2norm_silu(x, self.w[p + "norm"], c, silu=False)
        qkv = self._conv(p + "to_qkv", y).view(T, hw, 3 * C)
        out = torch.empty(T, hw, C, device=x.device, dtype=BF16)
        kpad = torch.zeros(hw8, C, device=x.device, dtype=BF16)
        vt = torch.zeros(C, hw8, device=x.device, dtype=BF16)
        pr = torch.zeros(hw, hw8, device=x.device, dtype=BF16)
        for f in range(T):
            q = qkv[f, :, :C]
            kpad[:hw] = qkv[f, :, C:2 * C]
            vt[:, :hw] = qkv[f, :, 2 * C:].t()
            s = ops.gemm(q, kpad, None, out_f32=True)                    # [hw, hw8] = q k^T  (fp32)
            ops.softmax_rows(s[:, :hw], c ** -0.5, out=pr[:, :hw])       # pad columns of P stay zero
            ops.gemm(pr, vt, None, out=out[f])                           # [hw, C] = P v
        return self._conv(p + "proj", out.view(1, T, H, W, C), residual=x)

    def _resample(self, p, x, mode, c):
        _, T, H, W, C = x.shape
        if mode == "up3d" and T > 1:
            assert C == c, "time_conv sits at 64-multiple widths"
            y = self._conv(p + "time_conv", x[:, 1:].contiguous())                                # [1, T-1, H, W, 2C]
            y = y.view(1, T - 1, H, W, 2, C).permute(0, 1, 4, 2, 3, 5).reshape(1, 2 * (T - 1), H, W, C)   # :140-143
            x = torch.cat([x[:, :1], y], dim=1)
            T = x.shape[1]
        up = ops.upsample2x(x.view(T, H, W, C))
        return self._conv(p + "resample.1", up.view(1, T, 2 * H, 2 * W, C))

    @torch.no_grad()
    def decode_one(self, z: torch.Tensor) -> torch.Tensor:
        """z [16, T, H, W] -> [3, 1 + 4(T-1), 8H, 8W] float32 in [-1, 1]."""
        dev = self.device
        zc, T, H, W = z.shape
        x = z.to(dev, torch.float32) * self.std.to(dev).view(-1, 1, 1, 1) + self.mean.to(dev).view(-1, 1, 1, 1)   # :581-586
        xp = torch.zeros(1, T, H, W, _pad64(zc), device=dev, dtype=BF16)
        xp[0, ..., :zc] = x.permute(1, 2, 3, 0).to(BF16)
        x = self._conv("conv2", xp)
        x = self._conv("decoder.conv1", x)
        d0 = self.dim * self.dim_mult[-1]
        x = self._res("decoder.middle.0.", x, d0, d0)
        x = self._attn("decoder.middle.1.", x, d0)
        x = self._res("decoder.middle.2.", x, d0, d0)
        c = d0
        for i, ent in enumerate(self._layout()):
            p = f"decoder.upsamples.{i}."
            if ent[0] == "res":
                x = self._res(p, x, ent[1], ent[2]); c = ent[2]
            else:
                x = self._resample(p, x, ent[0], ent[1]); c = ent[1] // 2
        x = self._conv("decoder.head.2", ops.l2norm_silu(x, self.w["decoder.head.0"], c))        # [1, T', H', W', 8]
        return x[0, ..., :3].permute(3, 0, 1, 2).float().clamp_(-1, 1)

    def decode(self, zs: List[torch.Tensor], tile_size: int = 0, any_end_frame: bool = False) -> List[torch.Tensor]:
        """vae.py:825-829.  tile_size is accepted and ignored: 180 GB of HBM hold the whole video (the reference tiles to fit
        consumer cards, :92-115)."""
        if any_end_frame:
            # vae.py:597-601: the last latent frame is decoded on its own, without the feature caches (one image), and appended;
            # conv2 is 1x1x1, so splitting before it is the same computation (oracle: exact in fp64)
            return [torch.cat([self.decode_one(u[:, :-1]), self.decode_one(u[:, -1:])], dim=1) for u in zs]
        return [self.decode_one(u) for u in zs]

    def _downsample(self, p, x, mode):
        """Resample('downsample2d'|'downsample3d') (vae.py:90-97, 150-165) over the whole sequence."""
        y = ops.conv_taps_strided(x, self.w[p + "resample.1.w"], self.w[p + "resample.1.b"], 1, 3, 1, 2, off_hw=1)
        if mode == "down3d" and y.shape[1] > 1:
            z = ops.conv_taps_strided(y, self.w[p + "time_conv.w"], self.w[p + "time_conv.b"], 3, 1, 2, 1)
            z[:, 0] = y[:, 0]                      # the first chunk (one frame) bypasses time_conv and only seeds its cache
            y = z
        return y

    @torch.no_grad()
    def encode_one(self, video: torch.Tensor) -> torch.Tensor:
        """video [3, 1+4k, H, W] in [-1, 1] -> mu [z_dim, 1+k, H/8, W/8] float32, (mu - mean) / std (vae.py:566-575)."""
        if "encoder.conv1.w" not in self.w:
            raise RuntimeError("WanVAE.encode: no encoder weights were loaded (load_state_dict with encoder.* / conv1.* keys)")
        dev = self.device
        c3, T, H, W = video.shape
        if (T - 1) % 4:
            raise ValueError("WanVAE.encode: the reference's 1,4,4,... chunking needs 1 + 4k frames")
        xp = torch.zeros(1, T, H, W, 64, device=dev, dtype=BF16)
        xp[0, ..., :c3] = video.to(dev).permute(1, 2, 3, 0).to(BF16)
        x = self._conv("encoder.conv1", xp)
        for i, ent in enumerate(self._enc_layout()):
            p = f"encoder.downsamples.{i}."
            x = self._res(p, x, ent[1], ent[2]) if ent[0] == "res" else self._downsample(p, x, ent[0])
        d0 = self.dim * self.dim_mult[-1]
        x = self._res("encoder.middle.0.", x, d0, d0)
        x = self._attn("encoder.middle.1.", x, d0)
        x = self._res("encoder.middle.2.", x, d0, d0)
        x = self._conv("encoder.head.2", ops.l2norm_silu(x, self.w["encoder.head.0"], d0))
        x = self._conv("conv1", x)                                                                  # [1, T', H', W', 64]
        mu = x[0, ..., : self.z_dim].permute(3, 0, 1, 2).float()
        return (mu - self.mean.to(dev).view(-1, 1, 1, 1)) * (1.0 / self.std).to(dev).view(-1, 1, 1, 1)

    def encode(self, videos: List[torch.Tensor], tile_size: int = 0, any_end_frame: bool = False) -> List[torch.Tensor]:
        """vae.py:806-816.  tile_size is accepted and ignored (see decode)."""
        if any_end_frame:
            # vae.py:541-542, 553-557: the last frame is encoded on its own, without the feature caches, and appended (conv1 is 1x1x1)
            return [torch.cat([self.encode_one(u[:, :-1]), self.encode_one(u[:, -1:])], dim=1) for u in videos]
        return [self.encode_one(u) for u in videos]
