"""WanT2V — B200-native drop-in for the denoise loop of wan/text2video.py:281-607 (`WanT2V.generate`).

Scope: noise, UniPC schedule, RoPE tables, the per-step two-sequence forward (cond / uncond
batched), CFG with the optional CFG-Zero* projection, scheduler step, `self.vae.decode` of the result
(text2video.py:579-596).  Text encoding (T5), VACE / phantom / recam are out of scope: prompt embeddings are passed in
(`context=`, `context_null=`).  Returns what the reference returns — the decoded video [3, F, H, W] fp32 in [-1, 1] —
when the pipeline holds a `vae`; without one (or with `return_latents=True`, the parity tests' hook) the denoised
latent [16, (F-1)/4+1, H/8, W/8] (fp32).
"""
from __future__ import annotations

import math
import random
import sys
from typing import Optional

import torch

from .. import ops
from .fm_solvers import FlowDPMSolverMultistepScheduler, get_sampling_sigmas, retrieve_timesteps
from .fm_solvers_unipc import FlowUniPCMultistepScheduler
from .model import WanModel
from .posemb_layers import get_rotary_pos_embed


class WanT2V:
    def __init__(self, model: WanModel, device="cuda", num_train_timesteps: int = 1000, vae_stride=(4, 8, 8),
                 patch_size=(1, 2, 2), z_dim: int = 16, vae=None):
        self.model = model
        self.vae = vae                      # WanVAE (wan/vae.py) or None: generate() then returns the latents
        self.device = torch.device(device)
        self.num_train_timesteps = num_train_timesteps
        self.vae_stride, self.patch_size, self.z_dim = vae_stride, patch_size, z_dim
        self._interrupt = False

    @torch.no_grad()
    def generate(self, input_prompt=None, input_frames=None, input_masks=None, input_ref_images=None, input_video=None,
                 target_camera=None, context_scale=1.0, width=1280, height=720, fit_into_canvas=True, frame_num=81,
                 shift=5.0, sample_solver="unipc", sampling_steps=50, guide_scale=5.0, n_prompt="", seed=-1,
                 offload_model=True, callback=None, enable_RIFLEx=None, VAE_tile_size=0, joint_pass=False,
                 slg_layers=None, slg_start=0.0, slg_end=1.0, cfg_star_switch=True, cfg_zero_step=5,
                 overlapped_latents=None, return_latent_slice=None, overlap_noise=0, conditioning_latents_size=0,
                 model_filename=None, context: Optional[torch.Tensor] = None, context_null: Optional[torch.Tensor] = None,
                 noise: Optional[torch.Tensor] = None, _per_step_latents=None, cfg_parallel=None,
                 return_latents: bool = False, **bbargs):
        if input_frames is not None or input_ref_images is not None or target_camera is not None or overlapped_latents is not None:
            raise NotImplementedError("VACE / phantom / recam inputs are out of scope")
        if context is None or (guide_scale != 1 and context_null is None):
            raise NotImplementedError("the T5 text encoder is out of scope: pass context= / context_null= embeddings [L, 4096]")
        if sample_solver not in ("unipc", "dpm++"):
            raise NotImplementedError("Unsupported solver.")
        dev = self.device
        F = frame_num
        target_shape = (self.z_dim, (F - 1) // self.vae_stride[0] + 1, height // self.vae_stride[1], width // self.vae_stride[2])
        if noise is None:
            seed_g = torch.Generator(device=dev)
            seed_g.manual_seed(seed if seed >= 0 else random.randint(0, sys.maxsize))                    # :354
            noise = torch.randn(*target_shape, dtype=torch.float32, device=dev, generator=seed_g)     # :410
        latents = noise.to(device=dev, dtype=torch.float32).contiguous()
        assert tuple(latents.shape) == tuple(target_shape)
        if sample_solver == "unipc":
            sch = FlowUniPCMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1, use_dynamic_shifting=False)
            sch.set_timesteps(sampling_steps, device=dev, shift=shift)
        else:                                                                                             # 'dpm++'
            sch = FlowDPMSolverMultistepScheduler(num_train_timesteps=self.num_train_timesteps, shift=1, use_dynamic_shifting=False)
            retrieve_timesteps(sch, device=dev, sigmas=get_sampling_sigmas(sampling_steps, shift))                                        # :419-422
        freqs = get_rotary_pos_embed(latents.shape[1:], enable_RIFLEx=bool(enable_RIFLEx))
        freqs = (freqs[0].to(dev), freqs[1].to(dev))
        scratch = torch.empty(2 * 148, device=dev, dtype=torch.float32)
        ctx = context.to(dev)
        ctx0 = context_null.to(dev) if context_null is not None else None
        # under CFG-parallel the unconditional half never runs the x_id == 0 pass that takes the TeaCache decision (model.py:1029):
        # the decision only depends on the (replicated) time embedding, so every rank takes it itself
        self.model._teacache_every_rank_decides = cfg_parallel is not None
        collective = cfg_parallel is not None or self.model.sp_group is not None
        if self.model.enable_teacache:                                                                    # :461-464
            self.model.previous_residual = [None] * 2
            if getattr(self.model, "teacache_multiplier", 0):
                self.model.compute_teacache_threshold(self.model.teacache_start_step, sch.timesteps_host, self.model.teacache_multiplier)
        if callback is not None:
            callback(-1, None, True)
        for i, t in enumerate(sch.timesteps_host):
            ts = torch.tensor([t], device=dev)
            slg = slg_layers if int(slg_start * sampling_steps) <= i < int(slg_end * sampling_steps) else None   # :492
            if guide_scale == 1:
                pred = self.model([latents], t=ts, context=[ctx], freqs=freqs, pipeline=self, current_step=i, slg_layers=slg)[0]
                if pred is None:
                    return None
            elif cfg_parallel is not None:
                # cond on one half of the ranks, uncond on the other (distributed/cfg_parallel.py), one 2-rank all-gather
                mine = self.model([latents], t=ts, context=[cfg_parallel.select(ctx, ctx0)], freqs=freqs, pipeline=self,
                                  current_step=i, x_id=cfg_parallel.branch, slg_layers=slg)[0]     # x_id 1 = the unconditional pass (:1077)
                if mine is None:
                    return None
                c, u = cfg_parallel.exchange(mine)
                pred = ops.cfg_combine(c, u, guide_scale, use_alpha=bool(cfg_star_switch and i > cfg_zero_step), scratch=scratch)
            else:
                # cond and uncond sequences in one batched forward (the reference's joint_pass list call, :509)
                c, u = self.model([latents, latents], t=ts, context=[ctx, ctx0], freqs=freqs, pipeline=self, current_step=i,
                                  slg_layers=slg)
                if c is None:
                    return None
                pred = ops.cfg_combine(c.contiguous(), u.contiguous(), guide_scale,
                                       use_alpha=bool(cfg_star_switch and i > cfg_zero_step), scratch=scratch)   # :551-562
            latents = sch.step(pred.unsqueeze(0), t, latents.unsqueeze(0), return_dict=False)[0].squeeze(0)
            if _per_step_latents is not None:
                _per_step_latents.append(latents.clone())
            if callback is not None:
                callback(i, latents, False)
            if _interrupted(self, collective, dev):
                return None
        if return_latents or self.vae is None:
            return latents
        if return_latent_slice is not None:                                                               # :578-595
            return {"x": self.vae.decode([latents], VAE_tile_size)[0], "latent_slice": latents[:, return_latent_slice].clone()}
        return self.vae.decode([latents], VAE_tile_size)[0]                                               # :590,596


def _interrupted(pipe, collective: bool, dev) -> bool:
    """`_interrupt` is set asynchronously by the host application on ONE process.  With several ranks the peers' kernels wait for
    each other (Ulysses exchange flags, the CFG-parallel all-gather), so leaving the loop must be a collective decision: the flag
    is OR-reduced once per step (one tiny all-reduce + host read) and every rank leaves at the same step boundary."""
    if not collective:
        return bool(pipe._interrupt)
    import torch.distributed as dist
    f = torch.tensor([1 if pipe._interrupt else 0], device=dev, dtype=torch.int32)
    dist.all_reduce(f, op=dist.ReduceOp.MAX)
    return bool(f.item())
