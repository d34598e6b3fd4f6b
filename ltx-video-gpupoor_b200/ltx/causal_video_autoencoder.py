"""CausalVideoAutoencoder (decode + encode) — B200-native drop-in for
ltx_video/models/autoencoders/causal_video_autoencoder.py:33-300,560-802,1023-1258,1282-1299 and
vae.py:343-413, vae_encode.py:94-165,239-247.

Activations are kept NDHWC bf16 so that (a) every 3x3x3 convolution is a TMA-tiled implicit GEMM on
tcgen05 (spatial zero padding = TMA out-of-bounds fill, temporal replicate padding = clamped frame
coordinate, no concat), (b) PixelNorm is a per-voxel reduction over the contiguous channel vector,
(c) depth-to-space, the first-frame drop, the residual add and the final 4x4 unpatchify are store
patterns of the conv epilogue instead of extra passes.
Encoder (causal_video_autoencoder.py:313-557, vae.py:265-306, vae_encode.py:22-91,228-237): the same
kernels, always causal; the stride-2 "compress_all" convolutions run on a TMA descriptor that strides over the input
(no im2col / space-to-depth copy).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch

from .. import ops
from ..module_like import ModuleLike

BF16 = torch.bfloat16

# ltx_video/utils/diffusers_config_mapping.py:106-130 (OURS_VAE_CONFIG)
LTX_VAE_CONFIG = dict(
    _class_name="CausalVideoAutoencoder", dims=3, in_channels=3, out_channels=3, latent_channels=128,
    blocks=[["res_x", 4], ["compress_all", 1], ["res_x_y", 1], ["res_x", 3], ["compress_all", 1], ["res_x_y", 1],
            ["res_x", 3], ["compress_all", 1], ["res_x", 3], ["res_x", 4]],
    scaling_factor=1.0, norm_layer="pixel_norm", patch_size=4, latent_log_var="uniform", use_quant_conv=False,
    causal_decoder=False,
)


@dataclass
class DecoderOutput:
    sample: torch.Tensor


def _pack_conv(w5: torch.Tensor, perm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[Cout, Cin, 3,3,3] -> [Cout, 27*Cin] with k = ((kt*3+kh)*3+kw)*Cin + ci; optional output-row permutation."""
    if perm is not None:
        w5 = w5[perm]
    return w5.permute(0, 2, 3, 4, 1).reshape(w5.shape[0], -1).to(BF16).contiguous()


class DiagonalGaussianDistribution:
    """diffusers.models.autoencoders.vae.DiagonalGaussianDistribution as used at vae.py:306 / vae_encode.py:77:
    logvar clamped to [-30, 20], sample = mean + exp(0.5 logvar) * randn."""

    def __init__(self, mean: torch.Tensor, logvar: torch.Tensor):
        self.mean = mean
        self.logvar = torch.clamp(logvar, -30.0, 20.0).expand_as(mean)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator=None, noise: Optional[torch.Tensor] = None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device if generator is None or
                                generator.device.type != "cpu" else "cpu", dtype=self.mean.dtype)
        return self.mean + self.std * noise.to(self.mean.device)

    def mode(self):
        return self.mean


class _Decoder:
    timestep_conditioning = False


class CausalVideoAutoencoder(ModuleLike):
    def __init__(self, **config):
        cfg = dict(LTX_VAE_CONFIG)
        cfg.update({k: v for k, v in config.items()})
        if cfg["norm_layer"] != "pixel_norm" or cfg["dims"] != 3:
            raise NotImplementedError("only the pixel_norm / 3D configuration of the named 2B VAE is implemented")
        self._cfg = cfg
        self.config = SimpleNamespace(**{k: v for k, v in cfg.items() if not k.startswith("_")})
        self.decoder = _Decoder()
        self.decoder.causal = cfg["causal_decoder"]
        self.decoder.patch_size = cfg["patch_size"]
        self.decoder.timestep_conditioning = bool(cfg.get("timestep_conditioning", False))     # read by the pipeline (:1271)
        self.dtype = BF16
        self.device = torch.device("cuda")
        self.use_z_tiling = False
        self.use_hw_tiling = False
        self.plan: List[Tuple] = []
        self.w: Dict[str, torch.Tensor] = {}
        self.std_of_means = None
        self.mean_of_means = None

    @staticmethod
    def from_config(config):
        assert config["_class_name"] == "CausalVideoAutoencoder"
        return CausalVideoAutoencoder(**config)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *args, device="cuda", **kwargs):
        """causal_video_autoencoder.py:34-120: legacy directory, diffusers directory or single .safetensors."""
        from .checkpoint_io import load_vae_checkpoint
        config, sd = load_vae_checkpoint(pretrained_model_name_or_path)
        vae = cls.from_config(config)
        vae.load_state_dict(sd, device=device)
        return vae

    # vae.py:92-115 — on a 180 GB part tiling is never needed
    @staticmethod
    def get_VAE_tile_size(vae_config, device_mem_capacity, mixed_precision):
        return (0, 0)

    # vae.py:117-154 — the tiling switches exist so reference callers keep working; they select nothing here (whole-video decode / encode)
    def set_tiling_params(self, sample_size: int = 512, overlap_factor: float = 0.25):
        self.tile_sample_min_size, self.tile_latent_min_size, self.tile_overlap_factor = sample_size, int(sample_size / 32), overlap_factor

    def enable_z_tiling(self, z_sample_size: int = 4):
        self.z_sample_size = z_sample_size

    def disable_z_tiling(self):
        pass

    def enable_hw_tiling(self):
        pass

    def disable_hw_tiling(self):
        pass

    @property
    def spatial_downscale_factor(self):
        n = len([b for b in self._cfg["blocks"] if b[0] in ("compress_space", "compress_all", "compress_all_res", "compress_space_res")])
        return 2 ** n * self._cfg["patch_size"]

    @property
    def temporal_downscale_factor(self):
        n = len([b for b in self._cfg["blocks"] if b[0] in ("compress_time", "compress_all", "compress_all_res", "compress_space_res")])
        return 2 ** n

    def _plan(self):
        ch = 128 * (2 ** len([b for b in self._cfg["blocks"] if b[0] == "res_x_y"]))
        plan = []
        for idx, (name, p) in enumerate(reversed(self._cfg["blocks"])):
            if name == "res_x":
                plan.append(("res_x", idx, ch, ch, int(p)))
            elif name == "res_x_y":
                plan.append(("res_x_y", idx, ch, ch // 2, 1))
                ch //= 2
            elif name == "compress_all":
                plan.append(("d2s", idx, ch, ch, 1))
            else:
                raise NotImplementedError(f"decoder block {name}")
        return plan, ch

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True, device="cuda", **_):
        """Reference key layout ('decoder.…', optional 'vae.' prefix, 'per_channel_statistics.*' or the
        registered std_of_means / mean_of_means buffers; causal_video_autoencoder.py:239-297)."""
        if any(k.startswith("vae.") for k in state_dict):
            state_dict = {k.replace("vae.", "", 1): v for k, v in state_dict.items() if k.startswith("vae.")}
        self.device = dev = torch.device(device)
        sd = state_dict
        w = {}

        def conv(name, perm=None):
            wt = sd[name + ".weight"].to(dev)
            b = sd[name + ".bias"].to(dev)
            if perm is not None:
                b = b[perm]
            return _pack_conv(wt, perm), b.to(BF16).contiguous()

        tc = self.decoder.timestep_conditioning

        def temb(name):
            lin = lambda n_: (sd[n_ + ".weight"].to(dev).to(BF16).contiguous(), sd[n_ + ".bias"].to(dev).to(BF16).contiguous())
            return lin(name + ".timestep_embedder.linear_1") + lin(name + ".timestep_embedder.linear_2")

        self.plan, last = self._plan()
        w["conv_in"] = conv("decoder.conv_in.conv")
        if tc:
            self.timestep_scale_multiplier = float(sd["decoder.timestep_scale_multiplier"])
            w["last_temb"] = temb("decoder.last_time_embedder")
            w["last_sst"] = sd["decoder.last_scale_shift_table"].to(dev).to(BF16).unsqueeze(0).contiguous()      # [1, 2, C]
        for kind, idx, cin, cout, n in self.plan:
            p = f"decoder.up_blocks.{idx}."
            if kind == "res_x":
                for j in range(n):
                    w[p + f"{j}.conv1"] = conv(p + f"res_blocks.{j}.conv1.conv")
                    w[p + f"{j}.conv2"] = conv(p + f"res_blocks.{j}.conv2.conv")
                if tc:                                   # UNetMidBlock3D.time_embedder + per-block tables (:850-853, 1207-1210)
                    w[p + "temb"] = temb(p + "time_embedder")
                    w[p + "sst"] = torch.stack([sd[p + f"res_blocks.{j}.scale_shift_table"].to(dev) for j in range(n)], 0) \
                        .to(BF16).contiguous()                                                    # [n, 4, C]
            elif kind == "res_x_y":
                w[p + "conv1"] = conv(p + "conv1.conv")
                w[p + "conv2"] = conv(p + "conv2.conv")
                w[p + "shortcut"] = (sd[p + "conv_shortcut.weight"].to(dev).reshape(cout, cin).to(BF16).contiguous(),
                                     sd[p + "conv_shortcut.bias"].to(dev).to(BF16).contiguous())
                w[p + "norm3"] = (sd[p + "norm3.norm.weight"].to(dev).to(BF16).contiguous(),
                                  sd[p + "norm3.norm.bias"].to(dev).to(BF16).contiguous())
            else:
                # pixel-shuffle channel order (c p1 p2 p3) -> (p1 p2 p3 c) so each 2x2x2 sub-voxel is contiguous
                perm = torch.arange(8 * cin, device=dev).reshape(cin, 8).t().reshape(-1)
                w[p + "conv"] = conv(p + "conv.conv", perm)
        ps = self._cfg["patch_size"]
        co = self._cfg["out_channels"]
        # unpatchify channel order (c r q) -> (c q r): r (width) innermost
        perm = torch.arange(co * ps * ps, device=dev).reshape(co, ps, ps).permute(0, 2, 1).reshape(-1)
        w["conv_out"] = conv("decoder.conv_out.conv", perm)
        self.enc_plan = self._enc_plan()
        self.has_encoder = "encoder.conv_in.conv.weight" in sd
        if self.has_encoder:
            cin0 = self._cfg["in_channels"] * ps * ps                         # 48 patchified channels, stored 64-padded
            wt = torch.zeros(sd["encoder.conv_in.conv.weight"].shape[0], 64, 3, 3, 3, device=dev)
            wt[:, :cin0] = sd["encoder.conv_in.conv.weight"].to(dev)
            w["enc.conv_in"] = (_pack_conv(wt), sd["encoder.conv_in.conv.bias"].to(dev).to(BF16).contiguous())
            for kind, idx, cin, cout, n in self.enc_plan:
                p = f"encoder.down_blocks.{idx}."
                if kind == "res_x":
                    for j in range(n):
                        w[p + f"{j}.conv1"] = conv(p + f"res_blocks.{j}.conv1.conv")
                        w[p + f"{j}.conv2"] = conv(p + f"res_blocks.{j}.conv2.conv")
                elif kind == "res_x_y":
                    w[p + "conv1"] = conv(p + "conv1.conv")
                    w[p + "conv2"] = conv(p + "conv2.conv")
                    w[p + "shortcut"] = (sd[p + "conv_shortcut.weight"].to(dev).reshape(cout, cin).to(BF16).contiguous(),
                                         sd[p + "conv_shortcut.bias"].to(dev).to(BF16).contiguous())
                    w[p + "norm3"] = (sd[p + "norm3.norm.weight"].to(dev).to(BF16).contiguous(),
                                      sd[p + "norm3.norm.bias"].to(dev).to(BF16).contiguous())
                else:
                    w[p + "conv"] = conv(p + "conv")      # the block IS a CausalConv3d: keys down_blocks.N.conv.{weight,bias}
            # conv_out: latent_channels + 1 ("uniform" log-variance) output channels, stored padded to a multiple of 8
            wo, bo = sd["encoder.conv_out.conv.weight"].to(dev), sd["encoder.conv_out.conv.bias"].to(dev)
            cop = (wo.shape[0] + 7) // 8 * 8
            wp = torch.zeros(cop, *wo.shape[1:], device=dev); wp[: wo.shape[0]] = wo
            bp = torch.zeros(cop, device=dev); bp[: bo.shape[0]] = bo
            w["enc.conv_out"] = (_pack_conv(wp), bp.to(BF16).contiguous())
        std = sd.get("std_of_means", sd.get("per_channel_statistics.std-of-means"))
        mean = sd.get("mean_of_means", sd.get("per_channel_statistics.mean-of-means"))
        if std is not None:
            self.std_of_means = std.to(dev)
            self.mean_of_means = (mean if mean is not None else torch.zeros_like(std)).to(dev)
        self.w = w
        return [], []

    def _enc_plan(self):
        """encoder.down_blocks (causal_video_autoencoder.py:373-478)"""
        ch, plan = 128, []
        for idx, (name, p) in enumerate(self._cfg["blocks"]):
            if name == "res_x":
                plan.append(("res_x", idx, ch, ch, int(p)))
            elif name == "res_x_y":
                plan.append(("res_x_y", idx, ch, ch * 2, 1)); ch *= 2
            elif name == "compress_all":
                plan.append(("down", idx, ch, ch, 1))
            else:
                raise NotImplementedError(f"encoder block {name}")
        return plan

    def _encode(self, x: torch.Tensor):
        """Encoder.forward (:510-557): x [B,3,F,H,W] in [-1,1] -> (mean [B,128,F',H',W'], logvar [B,1,F',H',W']) fp32."""
        if not self.has_encoder:
            raise RuntimeError("this CausalVideoAutoencoder was loaded without encoder weights")
        w, ps = self.w, self._cfg["patch_size"]
        B, C, Fr, H, W = x.shape
        assert C == 3 and H % ps == 0 and W % ps == 0
        h, wd = H // ps, W // ps
        # patchify 'b c (h q) (w r) -> b (c r q) f h w' (:1261-1279) straight into NDHWC, 48 -> 64 padded channels
        xp = torch.zeros(B, Fr, h, wd, 64, device=self.device, dtype=BF16)
        xp[..., : C * ps * ps] = (x.to(self.device, torch.float32).view(B, C, Fr, h, ps, wd, ps)
                                  .permute(0, 2, 3, 5, 1, 6, 4).reshape(B, Fr, h, wd, C * ps * ps).to(BF16))
        y = ops.conv3d(xp, *w["enc.conv_in"], causal=True)
        y, yn = self._run_blocks(y, self.enc_plan, "encoder.down_blocks.", True, None,
                                 lambda t, p: ops.conv3d_strided(t, *w[p + "conv"], stride_t=2, stride_hw=2))
        if yn is None:
            yn = ops.pixelnorm_silu(y)
        y = ops.conv3d(yn, *w["enc.conv_out"], causal=True)                               # [B, F', H', W', 136]
        lc = self._cfg["latent_channels"]
        mom = y[..., : lc + 1].permute(0, 4, 1, 2, 3).float()
        return mom[:, :lc].contiguous(), mom[:, lc:].contiguous()

    def encode(self, z: torch.Tensor, return_dict: bool = True):
        """vae.py:265-306: returns the DiagonalGaussianDistribution of the latent (mean, uniform log-variance)."""
        mean, logvar = self._encode(z)
        post = DiagonalGaussianDistribution(mean, logvar)
        if not return_dict:
            return (post,)
        return SimpleNamespace(latent_dist=post)

    # ---------------------------------------------------------------------------------------------
    def _time_embed(self, t: torch.Tensor, tw) -> torch.Tensor:
        """PixArtAlphaCombinedTimestepSizeEmbeddings(dim, 0): Timesteps(256) -> Linear -> SiLU -> Linear; t [1] fp32 -> [1, dim]"""
        return ops.gemm(ops.gemm(ops.timestep_embed(t, 256), tw[0], tw[1], act=ops.ACT_SILU), tw[2], tw[3])

    @staticmethod
    def _conv_want(h, c, causal, residual, want):
        """One convolution of a res block and the PixelNorm + SiLU of its OUTPUT (the next convolution's input), fused into the conv's
        epilogue when one N tile holds the whole channel vector (Cout <= 256).  want: "raw" -> (y, None); "both" -> (y, silu(pn(y)));
        "norm" -> (None, silu(pn(y))) (the raw row is never written)."""
        if want == "raw":
            return ops.conv3d(h, c[0], c[1], causal=causal, residual=residual), None
        if c[0].shape[0] > 256 or os.environ.get("LTXB200_VAE_FUSED_NORM") == "0":      # the switch exists for A/B runs and tests
            y = ops.conv3d(h, c[0], c[1], causal=causal, residual=residual)
            return (y if want == "both" else None), ops.pixelnorm_silu(y)
        return ops.conv3d_norm(h, c[0], c[1], causal=causal, residual=residual, keep_raw=(want == "both"))

    def _resnet(self, x, c1, c2, causal, shortcut=None, norm3=None, ada=None, xn=None, want="raw"):
        """ResnetBlock3D.forward (causal_video_autoencoder.py:1197-1258); `ada` [4, C] = scale_shift_table + timestep embedding
        (shift1, scale1, shift2, scale2) for timestep-conditioned decoders (:1212-1237).  `xn` = silu(pixelnorm(x)) when the kernel that
        produced x already wrote it; returns (y, silu(pixelnorm(y))) as `want` asks (see _conv_want)."""
        if ada is not None:
            h = ops.conv3d(ops.pixelnorm_silu(x, scale=ada[1], shift=ada[0]), c1[0], c1[1], causal=causal)
            h = ops.pixelnorm_silu(h, scale=ada[3], shift=ada[2])
            return ops.conv3d(h, c2[0], c2[1], causal=causal, residual=x), None
        if xn is None:
            xn = ops.pixelnorm_silu(x)
        _, hn = self._conv_want(xn, c1, causal, None, "norm")                      # norm2 + SiLU of conv1's output, in conv1's epilogue
        res = x
        if shortcut is not None:
            B, T, H, W, C = x.shape
            xs = ops.norm_mod(x.view(-1, C), weight=norm3[0], bias=norm3[1], eps=1e-6, layer_norm=True)
            res = ops.gemm(xs, shortcut[0], shortcut[1]).view(B, T, H, W, -1)
        return self._conv_want(hn, c2, causal, res, want)                           # + the NEXT block's norm1 + SiLU

    def _run_blocks(self, x, plan, prefix, causal, ada_fn, resample_fn, head_want="norm"):
        """The res-block / resample sequence of the encoder or decoder.  Every block is told what its consumer reads: the next res block
        needs the raw tensor (residual) and its PixelNorm + SiLU, a resampling convolution only the raw tensor, the output head only
        the normalised one.  -> (x | None, silu(pixelnorm(x)))"""
        w = self.w
        flat = []
        for kind, idx, cin, cout, n in plan:
            p = f"{prefix}{idx}."
            if kind == "res_x":
                flat += [("res", p, j, n, cin) for j in range(n)]
            elif kind == "res_x_y":
                flat.append(("resxy", p, 0, 1, cin))
            else:
                flat.append(("resample", p, 0, 1, cin))
        xn = None
        for i, (kind, p, j, n, cin) in enumerate(flat):
            nxt = flat[i + 1][0] if i + 1 < len(flat) else "head"
            want = "raw" if nxt == "resample" else (head_want if nxt == "head" else "both")
            if kind == "resample":
                x, xn = resample_fn(x, p), None
                continue
            ada = ada_fn(p, n, cin)[j] if (ada_fn is not None and kind == "res") else None
            if kind == "res":
                x, xn = self._resnet(x, w[p + f"{j}.conv1"], w[p + f"{j}.conv2"], causal, ada=ada, xn=xn, want=want)
            else:
                x, xn = self._resnet(x, w[p + "conv1"], w[p + "conv2"], causal, w[p + "shortcut"], w[p + "norm3"], xn=xn, want=want)
        return x, xn

    def _decode(self, z: torch.Tensor, target_shape=None, timestep=None, per_channel_normalize: bool = False,
                out_f32: bool = False) -> torch.Tensor:
        """vae.py:343-355 + Decoder.forward (causal_video_autoencoder.py:735-802)."""
        causal = self.decoder.causal
        w = self.w
        z = z.to(self.device)
        if z.dtype not in (torch.float32, BF16):
            z = z.float()
        z = z.contiguous()
        if per_channel_normalize:
            x = ops.latent_to_ndhwc(z, self.std_of_means.float().contiguous(), self.mean_of_means.float().contiguous())
        else:
            x = ops.latent_to_ndhwc(z, None, None)
        tc = self.decoder.timestep_conditioning
        if tc:
            assert timestep is not None, "should pass timestep with timestep_conditioning=True"          # :757-761
            tt = torch.as_tensor(timestep, dtype=torch.float32).flatten()
            if tt.numel() > 1 and not bool((tt == tt[0]).all()):
                # the reference passes one decode timestep per video (vae_encode.py:110-118) and its callers make them all equal
                # (pipeline_ltx_video.py:1271-1285); the modulation tables here are built once per call
                raise NotImplementedError("timestep-conditioned decode with DIFFERENT timesteps per video")
            ts = (tt[:1] * self.timestep_scale_multiplier).to(self.device)
        x = ops.conv3d(x, *w["conv_in"], causal=causal)
        ada_cache = {}

        def ada_fn(p, n, cin):          # UNetMidBlock3D: one timestep embedding per stage + per-block tables (:850-853, 1207-1210)
            if p not in ada_cache:
                ada_cache[p] = ops.ada_add(w[p + "sst"], self._time_embed(ts, w[p + "temb"])).view(n, 4, cin)
            return ada_cache[p]

        x, xn = self._run_blocks(x, self.plan, "decoder.up_blocks.", causal, ada_fn if tc else None,
                                 lambda t, p: ops.conv3d(t, *w[p + "conv"], causal=causal, store=ops.CONV_D2S),
                                 head_want="raw" if tc else "norm")      # the conditioned head modulates the norm: it reads the raw tensor
        if tc:                                                                         # :773-797
            ada = ops.ada_add(w["last_sst"], self._time_embed(ts, w["last_temb"])).view(2, -1)
            xn = ops.pixelnorm_silu(x, scale=ada[1], shift=ada[0])
        elif xn is None:
            xn = ops.pixelnorm_silu(x)
        return ops.conv3d(xn, *w["conv_out"], causal=causal, store=ops.CONV_UNPATCH, out_f32=out_f32)

    def decode(self, z: torch.Tensor, return_dict: bool = True, target_shape=None, timestep=None):
        """vae.py:357-413 (tiling branches are low-VRAM workarounds and are not needed on 180 GB)."""
        assert target_shape is not None, "target_shape must be provided for decoding"
        dec = self._decode(z, target_shape=target_shape, timestep=timestep)
        if not return_dict:
            return (dec,)
        return DecoderOutput(sample=dec)


def un_normalize_latents(latents, vae, vae_per_channel_normalize=False):
    """vae_encode.py:239-247 (kept for API parity; `vae_decode` below fuses it into the layout kernel)."""
    if vae_per_channel_normalize:
        return latents * vae.std_of_means.to(latents.dtype).view(1, -1, 1, 1, 1) + vae.mean_of_means.to(latents.dtype).view(1, -1, 1, 1, 1)
    return latents / vae.config.scaling_factor


def get_vae_size_scale_factor(vae) -> Tuple[int, int, int]:
    """vae_encode.py:168-187"""
    return (vae.temporal_downscale_factor, vae.spatial_downscale_factor, vae.spatial_downscale_factor)


def vae_decode(latents: torch.Tensor, vae: CausalVideoAutoencoder, is_video: bool = True, split_size: int = 1,
               vae_per_channel_normalize: bool = False, timestep=None) -> torch.Tensor:
    """vae_encode.py:94-165: latents [B,128,F,H,W] -> frames [B,3,8(F-1)+1,32H,32W] (bf16)."""
    if split_size != 1:
        raise NotImplementedError("split_size > 1 is a memory workaround that is not needed here")
    if not vae_per_channel_normalize and vae.config.scaling_factor != 1.0:
        latents = latents / vae.config.scaling_factor
    return vae._decode(latents.to(vae.dtype), per_channel_normalize=vae_per_channel_normalize, timestep=timestep)


def normalize_latents(latents, vae, vae_per_channel_normalize=False):
    """vae_encode.py:228-237"""
    if vae_per_channel_normalize:
        return (latents - vae.mean_of_means.to(latents.dtype).view(1, -1, 1, 1, 1)) / vae.std_of_means.to(latents.dtype).view(1, -1, 1, 1, 1)
    return latents * vae.config.scaling_factor


def vae_encode(media_items: torch.Tensor, vae: CausalVideoAutoencoder, split_size: int = 1, vae_per_channel_normalize: bool = False,
               generator=None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """vae_encode.py:22-91: media [B,3,F,H,W] (F = 8k+1) in [-1,1] -> normalised latents [B,128,F/8+1,H/32,W/32] (fp32).
    `latent_dist.sample()` is kept (the reference draws from the global RNG; pass `generator` or `noise` for reproducibility)."""
    if media_items.dim() != 5 or media_items.shape[1] != 3:
        raise ValueError(f"Expects [B, 3, F, H, W] tensors, got {tuple(media_items.shape)}")
    if split_size != 1:
        raise NotImplementedError("split_size > 1 is a memory workaround that is not needed here")
    latents = vae.encode(media_items).latent_dist.sample(generator=generator, noise=noise)
    return normalize_latents(latents, vae, vae_per_channel_normalize)
