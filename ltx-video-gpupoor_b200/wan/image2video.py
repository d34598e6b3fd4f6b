"""WanI2V — B200-native drop-in for the denoise loop of wan/image2video.py:124-414 (`WanI2V.generate`).

Scope: noise, UniPC schedule, RoPE tables, the per-step two-sequence forward with the i2v conditioning
(`y` = [mask(4) | image latent(16)] channels concatenated to the noisy latent, :232-244,279-280, and the 257 CLIP tokens
consumed by WanI2VCrossAttention), CFG with the optional CFG-Zero* projection, scheduler step.  `y` is either passed in or
built here the way the reference builds it (:232-244, 262-277): the start image (a [3, H, W] tensor in [-1, 1], already at the
output size — PIL / lanczos resizing is media I/O) padded with F-1 zero frames goes through `WanVAE.encode` (wan/vae.py) and
the 4-channel conditioning-frame mask is stacked on top; an end image (`image_end`, :191-199) adds one frame that is encoded without the
VAE's feature caches.  CLIP visual (:219-224) and T5 are out of scope: pass `clip_fea=`
[1, 257, 1280] and `context=` / `context_null=`.  Returns what the reference returns (:414-424): the video decoded by
`self.vae.decode(..., any_end_frame=...)` with the pixel frame of an added end image dropped, [3, F, H, W] fp32 in [-1, 1].  Without a `vae`
(or with `return_latents=True`, the parity tests' hook) the denoised latent [16, (F-1)/4+1, H/8, W/8] (fp32) comes back instead.
"""
from __future__ import annotations

import random
import sys
from typing import Optional

import torch

from .. import ops
from .fm_solvers import FlowDPMSolverMultistepScheduler, get_sampling_sigmas, retrieve_timesteps
from .fm_solvers_unipc import FlowUniPCMultistepScheduler
from .model import WanModel
from .posemb_layers import get_rotary_pos_embed
from .text2video import _interrupted


class WanI2V:
    def __init__(self, model: WanModel, device="cuda", num_train_timesteps: int = 1000, vae_stride=(4, 8, 8),
                 patch_size=(1, 2, 2), z_dim: int = 16, vae=None):
        assert model.model_type == "i2v"
        self.model = model
        self.vae = vae                                    # WanVAE with encoder weights, only needed when y= is not passed
        self.device = torch.device(device)
        self.num_train_timesteps = num_train_timesteps
        self.vae_stride, self.patch_size, self.z_dim = vae_stride, patch_size, z_dim
        self._interrupt = False

    def first_frame_mask(self, frame_num: int, lat_h: int, lat_w: int, any_end_frame: bool = False,
                         add_frames_for_end_image: bool = True) -> torch.Tensor:
        """image2video.py:232-244: ones on the conditioning frames (the first, and the last when there is an end image), the first (and an
        ADDED last) frame repeated 4x so that the frames fold into [4, latent frames, lat_h, lat_w]."""
        msk = torch.ones(1, frame_num, lat_h, lat_w, device=self.device)
        if any_end_frame:
            msk[:, 1:-1] = 0
            if add_frames_for_end_image:
                msk = torch.concat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:-1],
                                    torch.repeat_interleave(msk[:, -1:], repeats=4, dim=1)], dim=1)
            else:
                msk = torch.concat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:]], dim=1)
        else:
            msk[:, 1:] = 0
            msk = torch.concat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:]], dim=1)
        msk = msk.view(1, msk.shape[1] // 4, 4, lat_h, lat_w)
        return msk.transpose(1, 2)[0]

    @torch.no_grad()
    def encode_conditioning(self, image_start: torch.Tensor, frame_num: int, image_end: Optional[torch.Tensor] = None,
                            add_frames_for_end_image: bool = True) -> torch.Tensor:
        """image2video.py:262-277: y = [mask(4) | VAE latent(16) of (image, zeros ..., [end image])] -> [20, latent frames, H/8, W/8] fp32.
        `frame_num` is the frame count AFTER the reference's `frame_num += 1` for an added end frame (:196-199)."""
        if self.vae is None:
            raise RuntimeError("WanI2V(vae=WanVAE with encoder weights) is needed to build y from image_start")
        img = image_start.to(self.device, torch.float32)
        assert img.dim() == 3 and img.shape[0] == 3, "image_start: [3, H, W] tensor in [-1, 1] at the output size"
        h, w = img.shape[1:]
        assert h % (self.vae_stride[1] * self.patch_size[1]) == 0 and w % (self.vae_stride[2] * self.patch_size[2]) == 0
        any_end = image_end is not None
        if any_end:
            end = image_end.to(self.device, torch.float32)
            assert tuple(end.shape) == tuple(img.shape)
            enc = torch.concat([img[:, None], torch.zeros(3, frame_num - 2, h, w, device=self.device), end[:, None]], dim=1)
        else:
            enc = torch.concat([img[:, None], torch.zeros(3, frame_num - 1, h, w, device=self.device)], dim=1)
        lat_y = self.vae.encode([enc], 0, any_end_frame=any_end and add_frames_for_end_image)[0]
        msk = self.first_frame_mask(frame_num, h // self.vae_stride[1], w // self.vae_stride[2], any_end, add_frames_for_end_image)
        return torch.concat([msk, lat_y])

    @torch.no_grad()
    def generate(self, input_prompt=None, image_start=None, image_end=None, height=720, width=1280, fit_into_canvas=True,
                 frame_num=81, shift=5.0, sample_solver="unipc", sampling_steps=40, guide_scale=5.0, n_prompt="", seed=-1,
                 offload_model=True, callback=None, enable_RIFLEx=False, VAE_tile_size=0, joint_pass=False, slg_layers=None,
                 slg_start=0.0, slg_end=1.0, cfg_star_switch=True, cfg_zero_step=5, audio_scale=None, audio_cfg_scale=None,
                 audio_proj=None, audio_context_lens=None, model_filename=None,
                 context: Optional[torch.Tensor] = None, context_null: Optional[torch.Tensor] = None,
                 clip_fea: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None,
                 noise: Optional[torch.Tensor] = None, _per_step_latents=None, return_latents: bool = False, **bbargs):
        if audio_proj is not None or audio_scale is not None:
            raise NotImplementedError("fantasytalking audio conditioning is out of scope")
        any_end_frame = image_end is not None
        add_frames_for_end_image = model_filename is None or "image2video" in model_filename or "fantasy" in model_filename   # :191
        lat_frames = (frame_num - 1) // self.vae_stride[0] + 1
        if any_end_frame and add_frames_for_end_image:                                        # :194-199
            frame_num += 1
            lat_frames = (frame_num - 2) // self.vae_stride[0] + 2
        if clip_fea is None:
            raise NotImplementedError("CLIP visual is out of scope: pass clip_fea= [1,257,1280]")
        if y is None:
            if not torch.is_tensor(image_start):
                raise NotImplementedError("pass y= [20,T,H/8,W/8], or image_start= as a [3,H,W] tensor in [-1,1] (PIL resizing is media I/O)")
            if any_end_frame and not torch.is_tensor(image_end):
                raise NotImplementedError("image_end= must be a [3,H,W] tensor in [-1,1] (PIL resizing is media I/O)")
            y = self.encode_conditioning(image_start, frame_num, image_end, add_frames_for_end_image)
            height, width = image_start.shape[1:]
        if context is None or (guide_scale != 1 and context_null is None):
            raise NotImplementedError("the T5 text encoder is out of scope: pass context= / context_null= embeddings [L, 4096]")
        if sample_solver not in ("unipc", "dpm++"):
            raise NotImplementedError("Unsupported solver.")
        dev = self.device
        target_shape = (self.z_dim, lat_frames, height // self.vae_stride[1], width // self.vae_stride[2])
        assert tuple(y.shape) == (20,) + target_shape[1:], f"y must be [20, {target_shape[1:]}]"
        if noise is None:
            seed_g = torch.Generator(device=dev)
            seed_g.manual_seed(seed if seed >= 0 else random.randint(0, sys.maxsize))                    # :226
            noise = torch.randn(*target_shape, dtype=torch.float32, device=dev, generator=seed_g)     # :226-230
        latents = noise.to(device=dev, dtype=torch.float32).contiguous()
        assert tuple(latents.shape) == tuple(target_shape)
        if sample_solver == "unipc":
            sch = FlowUniPCMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1, use_dynamic_shifting=False)
            sch.set_timesteps(sampling_steps, device=dev, shift=shift)
        else:                                                                                             # 'dpm++'
            sch = FlowDPMSolverMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1, use_dynamic_shifting=False)
            retrieve_timesteps(sch, device=dev, sigmas=get_sampling_sigmas(sampling_steps, shift))                                        # :290-296
        freqs = get_rotary_pos_embed(latents.shape[1:], enable_RIFLEx=bool(enable_RIFLEx))
        freqs = (freqs[0].to(dev), freqs[1].to(dev))
        scratch = torch.empty(2 * 148, device=dev, dtype=torch.float32)
        ctx = context.to(dev)
        ctx0 = context_null.to(dev) if context_null is not None else None
        yd, clip = y.to(dev), clip_fea.to(dev)
        if self.model.enable_teacache:                                                                    # image2video.py:318-321
            self.model.previous_residual = [None] * 2
            if getattr(self.model, "teacache_multiplier", 0):
                self.model.compute_teacache_threshold(self.model.teacache_start_step, sch.timesteps_host, self.model.teacache_multiplier)
        if callback is not None:
            callback(-1, None, True)
        for i, t in enumerate(sch.timesteps_host):
            ts = torch.tensor([t], device=dev)
            slg = slg_layers if int(slg_start * sampling_steps) <= i < int(slg_end * sampling_steps) else None   # :333
            kw = dict(t=ts, clip_fea=clip, y=yd, freqs=freqs, pipeline=self, current_step=i, slg_layers=slg)
            if guide_scale == 1:
                pred = self.model([latents], context=[ctx], **kw)[0]                                     # :340-341
                if pred is None:
                    return None
            else:
                c, u = self.model([latents, latents], context=[ctx, ctx0], **kw)                        # joint pass :345-349
                if c is None:
                    return None
                pred = ops.cfg_combine(c.contiguous(), u.contiguous(), guide_scale,
                                       use_alpha=bool(cfg_star_switch and i > cfg_zero_step), scratch=scratch)   # :384-398
            latents = sch.step(pred.unsqueeze(0), t, latents.unsqueeze(0), return_dict=False)[0].squeeze(0)      # :403-409
            if _per_step_latents is not None:
                _per_step_latents.append(latents.clone())
            if callback is not None:
                callback(i, latents, False)
            if _interrupted(self, self.model.sp_group is not None, dev):
                return None
        if return_latents or self.vae is None:
            return latents
        # image2video.py:414-420: decode (the appended end frame without feature caches), drop the frame added for the end image
        video = self.vae.decode([latents], VAE_tile_size, any_end_frame=any_end_frame and add_frames_for_end_image)[0]
        if any_end_frame and add_frames_for_end_image:
            video = video[:, :-1]
        return video
