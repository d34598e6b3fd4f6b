"""LatentUpsampler — B200-native drop-in for ltx_video/models/autoencoders/latent_upsampler.py:42-149.

Same constructor keys, state_dict layout and `forward(latent [B,C,F,H,W]) -> [B,C,F,2H,2W]` for the released configuration
(dims=3, spatial_upsample=True, temporal_upsample=False: ltxv-spatial-upscaler-0.9.7).  Activations are NDHWC bf16; the centred,
zero-padded 3x3x3 / 3x3 convolutions run on the implicit-GEMM tcgen05 kernel (TMA out-of-bounds fill is the padding), GroupNorm(32)
+ SiLU (+ residual) is one stats pass and one apply pass, PixelShuffleND(2) is folded into the Conv2d's output-channel order.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import ops
from ..module_like import ModuleLike

BF16 = torch.bfloat16


def _pack(w: torch.Tensor, perm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[Cout, Cin, (kt,) kh, kw] -> [Cout, taps*Cin], k = tap-major then input channel."""
    if perm is not None:
        w = w[perm]
    order = (0, 2, 3, 4, 1) if w.dim() == 5 else (0, 2, 3, 1)
    return w.permute(*order).reshape(w.shape[0], -1).to(BF16).contiguous()


class LatentUpsampler(ModuleLike):
    def __init__(self, in_channels: int = 128, mid_channels: int = 512, num_blocks_per_stage: int = 4, dims: int = 3,
                 spatial_upsample: bool = True, temporal_upsample: bool = False):
        if dims != 3 or not spatial_upsample or temporal_upsample:
            raise NotImplementedError("only the released dims=3 spatial x2 upsampler (latent_upsampler.py:90-93) is implemented")
        if in_channels % 64 or mid_channels % 256:
            raise NotImplementedError("in_channels % 64 == 0 and mid_channels % 256 == 0 (GroupNorm kernel: 8-channel lanes per group)")
        self.in_channels, self.mid_channels, self.num_blocks_per_stage = in_channels, mid_channels, num_blocks_per_stage
        self.dims, self.spatial_upsample, self.temporal_upsample = dims, spatial_upsample, temporal_upsample
        self.device, self.dtype = torch.device("cuda"), BF16
        self.w: Dict[str, tuple] = {}

    @classmethod
    def from_config(cls, config):
        """latent_upsampler.py:151-160 (note the reference's defaults there differ from __init__'s)"""
        return cls(in_channels=config.get("in_channels", 4), mid_channels=config.get("mid_channels", 128),
                   num_blocks_per_stage=config.get("num_blocks_per_stage", 4), dims=config.get("dims", 2),
                   spatial_upsample=config.get("spatial_upsample", True), temporal_upsample=config.get("temporal_upsample", False))

    @classmethod
    def from_pretrained(cls, pretrained_model_path, *args, device="cuda", **kwargs):
        """latent_upsampler.py:177-199: .safetensors with the config in its metadata."""
        from .checkpoint_io import load_upsampler_checkpoint
        config, sd = load_upsampler_checkpoint(pretrained_model_path)
        up = cls.from_config(config)
        up.load_state_dict(sd, device=device)
        return up

    def config(self):
        return {"_class_name": "LatentUpsampler", "in_channels": self.in_channels, "mid_channels": self.mid_channels,
                "num_blocks_per_stage": self.num_blocks_per_stage, "dims": self.dims, "spatial_upsample": self.spatial_upsample,
                "temporal_upsample": self.temporal_upsample}

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True, device="cuda", **_):
        self.device = dev = torch.device(device)
        used = set()

        def get(name):
            used.add(name)
            return state_dict[name].to(dev)

        def conv(name, perm=None):
            w, b = get(name + ".weight"), get(name + ".bias")
            if perm is not None:
                b = b[perm]
            return _pack(w, perm), b.to(BF16).contiguous()

        def gn(name):
            return get(name + ".weight").to(BF16).contiguous(), get(name + ".bias").to(BF16).contiguous()

        w = {"initial_conv": conv("initial_conv"), "initial_norm": gn("initial_norm"), "final_conv": conv("final_conv")}
        for stage in ("res_blocks", "post_upsample_res_blocks"):
            for j in range(self.num_blocks_per_stage):
                p = f"{stage}.{j}."
                w[p + "conv1"], w[p + "norm1"], w[p + "conv2"], w[p + "norm2"] = conv(p + "conv1"), gn(p + "norm1"), conv(p + "conv2"), gn(p + "norm2")
        # PixelShuffleND(2) "b (c p1 p2) h w -> b c (h p1) (w p2)" (pixel_shuffle.py:14-20): reorder the Conv2d's output channels
        # to (p1, p2, c) so that every sub-pixel's channels are contiguous in the NHWC result
        m = self.mid_channels
        perm = (torch.arange(m).view(1, m) * 4 + torch.arange(4).view(4, 1)).reshape(-1).to(dev)
        w["upsampler"] = conv("upsampler.0", perm)
        extra = [k for k in state_dict if k not in used]
        if strict and extra:
            raise KeyError(f"unexpected keys in state_dict: {extra[:5]} ...")
        self.w = w
        return [], extra

    # ---------------------------------------------------------------------------------------------
    def _res(self, x, p):
        """ResBlock.forward (latent_upsampler.py:30-39)"""
        w = self.w
        h = ops.groupnorm_silu(ops.conv_taps(x, *w[p + "conv1"], 3, 3, centered=True), *w[p + "norm1"])
        h = ops.conv_taps(h, *w[p + "conv2"], 3, 3, centered=True)
        return ops.groupnorm_silu(h, *w[p + "norm2"], residual=x)

    def forward_ndhwc(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, F, H, W, C] bf16 -> [B, F, 2H, 2W, C] bf16"""
        w, m = self.w, self.mid_channels
        B, Fr, H, W, _ = x.shape
        x = ops.groupnorm_silu(ops.conv_taps(x, *w["initial_conv"], 3, 3, centered=True), *w["initial_norm"])
        for j in range(self.num_blocks_per_stage):
            x = self._res(x, f"res_blocks.{j}.")
        u = ops.conv_taps(x.view(1, B * Fr, H, W, m), *w["upsampler"], 1, 3, centered=True)          # Conv2d per frame, [.., (p1,p2,c)]
        x = u.view(B, Fr, H, W, 2, 2, m).permute(0, 1, 2, 4, 3, 5, 6).reshape(B, Fr, 2 * H, 2 * W, m)   # index permutation only
        for j in range(self.num_blocks_per_stage):
            x = self._res(x, f"post_upsample_res_blocks.{j}.")
        return ops.conv_taps(x, *w["final_conv"], 3, 3, centered=True)

    def forward(self, latent: torch.Tensor) -> torch.Tensor:
        """latent_upsampler.py:109-149; returns fp32 [B, C, F, 2H, 2W]."""
        z = latent.to(self.device)
        z = (z if z.dtype in (torch.float32, BF16) else z.float()).contiguous()
        return ops.latent_from_ndhwc(self.forward_ndhwc(ops.latent_to_ndhwc(z, None, None)), None, None)

    __call__ = forward


def adain_filter_latent(latents: torch.Tensor, reference_latents: torch.Tensor, factor: float = 1.0) -> torch.Tensor:
    """pipeline_ltx_video.py:1709-1737"""
    return ops.adain(latents.float().contiguous(), reference_latents.to(latents.device).float().contiguous(), factor)
