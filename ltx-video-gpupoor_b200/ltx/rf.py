"""RectifiedFlowScheduler — drop-in for ltx_video/schedulers/rf.py:176-392 (SD3 resolution-dependent
shift + terminal stretch; Euler step with per-token timesteps).  The timestep table is host-side fp32
math identical to the reference (bit-exact); `step` runs the fused sm_100a guidance/step kernel."""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional, Tuple, Union

import torch

from .. import ops


def time_shift(mu: float, sigma: float, t: torch.Tensor):
    """rf.py:69-70"""
    return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)


def get_normal_shift(n_tokens: int, min_tokens: int = 1024, max_tokens: int = 4096, min_shift: float = 0.95,
                     max_shift: float = 2.05) -> float:
    """rf.py:73-82"""
    m = (max_shift - min_shift) / (max_tokens - min_tokens)
    b = min_shift - m * min_tokens
    return m * n_tokens + b


def strech_shifts_to_terminal(shifts: torch.Tensor, terminal: float = 0.1):
    """rf.py:85-109"""
    if shifts.numel() == 0:
        raise ValueError("The 'shifts' tensor must not be empty.")
    if terminal <= 0 or terminal >= 1:
        raise ValueError("The terminal value must be between 0 and 1 (exclusive).")
    one_minus_z = 1 - shifts
    scale_factor = one_minus_z[-1] / (1 - terminal)
    return 1 - (one_minus_z / scale_factor)


def sd3_resolution_dependent_timestep_shift(samples_shape, timesteps: torch.Tensor,
                                            target_shift_terminal: Optional[float] = None) -> torch.Tensor:
    """rf.py:112-149"""
    if len(samples_shape) == 3:
        _, m, _ = samples_shape
    elif len(samples_shape) in [4, 5]:
        m = math.prod(samples_shape[2:])
    else:
        raise ValueError("Samples must have shape (b, t, c), (b, c, h, w) or (b, c, f, h, w)")
    shifts = time_shift(get_normal_shift(m), 1, timesteps)
    if target_shift_terminal is not None:
        shifts = strech_shifts_to_terminal(shifts, target_shift_terminal)
    return shifts


@dataclass
class RectifiedFlowSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class RectifiedFlowScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, shifting: Optional[str] = "SD3", base_resolution=None,
                 target_shift_terminal: Optional[float] = 0.1, sampler: Optional[str] = "Uniform",
                 shift: Optional[float] = None, **_):
        if sampler != "Uniform":
            raise NotImplementedError("only the 'Uniform' sampler of OURS_SCHEDULER_CONFIG is implemented")
        if shifting not in (None, "SD3"):
            raise NotImplementedError("only SD3 shifting is implemented")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, shifting=shifting,
                                      base_resolution=base_resolution, target_shift_terminal=target_shift_terminal,
                                      sampler=sampler, shift=shift)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.shifting = shifting
        self.target_shift_terminal = target_shift_terminal
        self.timesteps = self.sigmas = torch.linspace(1, 1 / num_train_timesteps, num_train_timesteps)
        self._scratch = None

    @classmethod
    def from_config(cls, config: dict):
        return cls(**{k: v for k, v in config.items() if not k.startswith("_")})

    def shift_timesteps(self, samples_shape, timesteps):
        if self.shifting == "SD3":
            return sd3_resolution_dependent_timestep_shift(samples_shape, timesteps, self.target_shift_terminal)
        return timesteps

    def set_timesteps(self, num_inference_steps: Optional[int] = None, samples_shape=None, timesteps=None,
                      device: Union[str, torch.device] = None):
        """rf.py:227-257 — computed on the host in fp32 exactly as the reference, then moved."""
        if timesteps is not None and num_inference_steps is not None:
            raise ValueError("You cannot provide both `timesteps` and `num_inference_steps`.")
        if timesteps is None:
            num_inference_steps = min(self.config.num_train_timesteps, num_inference_steps)
            timesteps = torch.linspace(1, 1 / num_inference_steps, num_inference_steps)
            timesteps = self.shift_timesteps(samples_shape, timesteps)
        else:
            timesteps = torch.as_tensor(timesteps, dtype=torch.float32).cpu()
            num_inference_steps = len(timesteps)
        self.timesteps_host = timesteps.clone()
        self.timesteps = timesteps.to(device) if device is not None else timesteps
        self.num_inference_steps = num_inference_steps
        self.sigmas = self.timesteps

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True,
             stochastic_sampling: Optional[bool] = False, **kwargs) -> Union[RectifiedFlowSchedulerOutput, Tuple]:
        """rf.py:311-380.  `timestep`: 0-d (global) or [1,1]; per-token timesteps are expressed through the pipeline's conditioning
        mask path (`ops.guidance_step`).  stochastic_sampling (:370-373) re-noises the x0 estimate to the next timestep; the
        N(0,1) draw comes from `generator=` / `noise=` in kwargs (the reference uses the global RNG)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        t = float(torch.as_tensor(timestep).reshape(-1)[0])
        assert torch.as_tensor(timestep).numel() == 1, "use the pipeline path for per-token timesteps"
        dev = model_output.device
        lat = sample.to(torch.float32).contiguous().clone().view(-1)
        pred = model_output.to(torch.bfloat16).contiguous().view(1, -1)
        ts = self.timesteps.to(device=dev, dtype=torch.float32).contiguous()
        noise = None
        if stochastic_sampling:
            noise = kwargs.get("noise")
            if noise is None:
                noise = torch.randn(sample.shape, device=dev, dtype=torch.float32, generator=kwargs.get("generator"))
            noise = noise.to(device=dev, dtype=torch.float32).contiguous().view(-1)
        ops.guidance_step(pred, lat, ts, t, num_conds=1, has_cfg=False, has_stg=False, do_rescale=False,
                          guidance_scale=1.0, stg_scale=0.0, rescale=1.0, channels=sample.shape[-1], cond_mask=None,
                          scratch=None, noise=noise)
        prev = lat.view(sample.shape)
        if not return_dict:
            return (prev,)
        return RectifiedFlowSchedulerOutput(prev_sample=prev)

    def add_noise(self, original_samples, noise, timesteps):
        """rf.py:382-392"""
        sigmas = timesteps
        while sigmas.ndim < original_samples.ndim:
            sigmas = sigmas.unsqueeze(-1)
        return (1 - sigmas) * original_samples + sigmas * noise
