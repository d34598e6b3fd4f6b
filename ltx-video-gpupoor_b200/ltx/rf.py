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


def linear_quadratic_schedule(num_steps: int, threshold_noise: float = 0.025, linear_steps: Optional[int] = None) -> torch.Tensor:
    """rf.py:25-47 ("LinearQuadratic" sampler): noise level rises linearly to `threshold_noise` over the first half of the steps and
    quadratically to 1 over the rest; returned as timesteps 1 - level.  Python-float arithmetic in the reference's operation order
    (the table is compared bit for bit)."""
    if num_steps == 1:
        return torch.tensor([1.0])
    n_lin = num_steps // 2 if linear_steps is None else linear_steps
    n_quad = num_steps - n_lin
    levels = [i * threshold_noise / n_lin for i in range(n_lin)]
    gap = n_lin - threshold_noise * num_steps
    a = gap / (n_lin * n_quad ** 2)
    b = threshold_noise / n_lin - 2 * gap / (n_quad ** 2)
    c = a * (n_lin ** 2)
    levels += [a * (i ** 2) + b * i + c for i in range(n_lin, num_steps)]
    return torch.tensor([1.0 - x for x in levels])


def simple_diffusion_resolution_dependent_timestep_shift(samples_shape, timesteps: torch.Tensor, n: int = 32 * 32) -> torch.Tensor:
    """rf.py:50-66 ("SimpleDiffusion" shifting): log-SNR moved by 2*log(tokens / base resolution)."""
    if len(samples_shape) == 3:
        _, m, _ = samples_shape
    elif len(samples_shape) in [4, 5]:
        m = math.prod(samples_shape[2:])
    else:
        raise ValueError("Samples must have shape (b, t, c), (b, c, h, w) or (b, c, f, h, w)")
    snr = (timesteps / (1 - timesteps)) ** 2
    return torch.sigmoid(0.5 * (torch.log(snr) + 2 * math.log(m / n)))


@dataclass
class RectifiedFlowSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class RectifiedFlowScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, shifting: Optional[str] = "SD3", base_resolution=None,
                 target_shift_terminal: Optional[float] = 0.1, sampler: Optional[str] = "Uniform",
                 shift: Optional[float] = None, **_):
        """Defaults are OURS_SCHEDULER_CONFIG (diffusers_config_mapping.py:63-72), the configuration every released checkpoint carries
        (the reference class itself defaults to no shifting, rf.py:180-188; it is always built through from_config / from_pretrained)."""
        if sampler not in ("Uniform", "LinearQuadratic", "Constant"):
            raise ValueError(f"unknown sampler {sampler!r}")
        if shifting not in (None, "SD3", "SimpleDiffusion"):
            raise ValueError(f"unknown shifting {shifting!r}")
        if base_resolution is None:
            base_resolution = 32 ** 2                      # the reference constructor's default (rf.py:184)
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, shifting=shifting,
                                      base_resolution=base_resolution, target_shift_terminal=target_shift_terminal,
                                      sampler=sampler, shift=shift)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.sampler, self.shifting, self.shift = sampler, shifting, shift
        self.base_resolution, self.target_shift_terminal = base_resolution, target_shift_terminal
        self.timesteps = self.sigmas = self.get_initial_timesteps(num_train_timesteps, shift=shift)
        self._scratch = None

    def get_initial_timesteps(self, num_timesteps: int, shift: Optional[float] = None) -> torch.Tensor:
        """rf.py:201-214"""
        if self.sampler == "LinearQuadratic":
            return linear_quadratic_schedule(num_timesteps)
        uniform = torch.linspace(1, 1 / num_timesteps, num_timesteps)
        if self.sampler == "Constant":
            assert shift is not None, "Shift must be provided for constant time shift sampler."
            return time_shift(shift, 1, uniform)
        return uniform

    @staticmethod
    def from_pretrained(pretrained_model_path):
        """rf.py:262-268: a JSON scheduler config.  A single-file .safetensors checkpoint (config in its metadata, as the model classes read
        it) is accepted as well."""
        import json
        if str(pretrained_model_path).endswith(".safetensors"):
            from safetensors import safe_open
            with safe_open(str(pretrained_model_path), framework="pt", device="cpu") as f:
                return RectifiedFlowScheduler.from_config(json.loads(f.metadata()["config"])["scheduler"])
        with open(pretrained_model_path, "r", encoding="utf-8") as reader:
            return RectifiedFlowScheduler.from_config(json.loads(reader.read()))

    @classmethod
    def from_config(cls, config: dict):
        return cls(**{k: v for k, v in config.items() if not k.startswith("_")})

    def shift_timesteps(self, samples_shape, timesteps):
        if self.shifting == "SD3":
            return sd3_resolution_dependent_timestep_shift(samples_shape, timesteps, self.target_shift_terminal)
        if self.shifting == "SimpleDiffusion":
            return simple_diffusion_resolution_dependent_timestep_shift(samples_shape, timesteps, self.base_resolution)
        return timesteps

    def set_timesteps(self, num_inference_steps: Optional[int] = None, samples_shape=None, timesteps=None,
                      device: Union[str, torch.device] = None):
        """rf.py:227-257 — computed on the host in fp32 exactly as the reference, then moved."""
        if timesteps is not None and num_inference_steps is not None:
            raise ValueError("You cannot provide both `timesteps` and `num_inference_steps`.")
        if timesteps is None:
            num_inference_steps = min(self.config.num_train_timesteps, num_inference_steps)
            timesteps = self.shift_timesteps(samples_shape, self.get_initial_timesteps(num_inference_steps, shift=self.shift))
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
        """rf.py:311-380: prev = sample - dt * model_output, or with stochastic_sampling (:370-373)
        add_noise(sample - t * model_output, randn, t - dt); fp32 arithmetic (fp32 predictions are not truncated), result in the promoted
        dtype of `sample` and `model_output`, as the reference's tensor arithmetic gives.  The N(0,1) draw comes from `generator=` /
        `noise=` in kwargs (the reference uses the global RNG).  A GLOBAL timestep (0-d, or any one-element tensor) is a host scalar
        on the linear-combination kernel; PER-TOKEN timesteps [B, N] (:361-367; sample [B, N, C]) go through `ltxb200_rf_step_tokens_f32`,
        which repeats the reference expression's fp32 rounding points (bit-identical to the reference on fp32 inputs)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        tt = torch.as_tensor(timestep)
        if tt.numel() != 1:
            assert tt.ndim == 2, "timestep must be 0-d (global) or [B, N] (per token)"                   # rf.py:362
            assert sample.dim() == 3 and tuple(sample.shape[:2]) == tuple(tt.shape) and model_output.shape == sample.shape
            dev = model_output.device
            out_dtype = torch.promote_types(torch.promote_types(sample.dtype, model_output.dtype), tt.dtype if tt.is_floating_point() else torch.float32)
            C = sample.shape[-1]
            x = sample.to(device=dev, dtype=torch.float32).contiguous().view(-1, C)
            v = model_output.to(torch.float32).contiguous().view(-1, C)
            z = None
            if stochastic_sampling:
                z = kwargs.get("noise")
                if z is None:
                    z = torch.randn(sample.shape, device=dev, dtype=torch.float32, generator=kwargs.get("generator"))
                z = z.to(device=dev, dtype=torch.float32).contiguous().view(-1, C)
            sched = self.timesteps_host.to(device=dev, dtype=torch.float32).contiguous()
            prev = ops.rf_step_tokens(x, v, tt.to(device=dev, dtype=torch.float32).contiguous().view(-1), sched, noise=z)
            prev = prev.view(sample.shape).to(out_dtype)
            if not return_dict:
                return (prev,)
            return RectifiedFlowSchedulerOutput(prev_sample=prev)
        t = float(tt.reshape(-1)[0])
        dev = model_output.device
        out_dtype = torch.promote_types(sample.dtype, model_output.dtype)
        lower = 0.0
        for v in self.timesteps_host.tolist():          # descending: the first one strictly below t - eps is the closest
            if v < t - 1e-6:
                lower = float(v)
                break
        x = sample.to(device=dev, dtype=torch.float32).contiguous().view(-1)
        v = model_output.to(torch.float32).contiguous().view(-1)
        n = x.numel()
        pad = (-n) % 4                                   # the kernel works on float4
        if pad:
            x, v = torch.nn.functional.pad(x, (0, pad)), torch.nn.functional.pad(v, (0, pad))
        out = torch.empty_like(x)
        if stochastic_sampling:
            noise = kwargs.get("noise")
            if noise is None:
                noise = torch.randn(sample.shape, device=dev, dtype=torch.float32, generator=kwargs.get("generator"))
            z = noise.to(device=dev, dtype=torch.float32).contiguous().view(-1)
            if pad:
                z = torch.nn.functional.pad(z, (0, pad))
            ops.lincomb(out, [(1.0 - lower, x), (-(1.0 - lower) * t, v), (lower, z)])
        else:
            ops.lincomb(out, [(1.0, x), (-(t - lower), v)])
        prev = out[:n].view(sample.shape).to(out_dtype)
        if not return_dict:
            return (prev,)
        return RectifiedFlowSchedulerOutput(prev_sample=prev)

    def add_noise(self, original_samples, noise, timesteps):
        """rf.py:382-392"""
        sigmas = timesteps
        while sigmas.ndim < original_samples.ndim:
            sigmas = sigmas.unsqueeze(-1)
        return (1 - sigmas) * original_samples + sigmas * noise
