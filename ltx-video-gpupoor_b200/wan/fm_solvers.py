"""FlowDPMSolverMultistepScheduler / get_sampling_sigmas / retrieve_timesteps — drop-in for wan/utils/fm_solvers.py:22-859 as
`sample_solver='dpm++'` uses it (text2video.py:423-432): order-2 multistep DPM-Solver++ (midpoint) on the flow parameterisation,
final sigma zero.  Scalar coefficients (log-SNR steps, exp(-h) - 1, r0) are computed on the host with the reference's torch fp32
scalar ops; each step is two `ops.lincomb` launches (x0 conversion :378-380, first / second order update :415-594)."""
from __future__ import annotations

import inspect
from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from .. import ops


def get_sampling_sigmas(sampling_steps, shift):
    """fm_solvers.py:22-26"""
    sigma = np.linspace(1, 0, sampling_steps + 1)[:sampling_steps]
    return shift * sigma / (1 + (shift - 1) * sigma)


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None, **kwargs):
    """fm_solvers.py:29-66"""
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed. Please choose one to set custom values")
    if timesteps is not None:
        if "timesteps" not in set(inspect.signature(scheduler.set_timesteps).parameters.keys()):
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support custom"
                             f" timestep schedules. Please check whether you are using the correct scheduler.")
        scheduler.set_timesteps(timesteps=timesteps, device=device, **kwargs)
    elif sigmas is not None:
        if "sigmas" not in set(inspect.signature(scheduler.set_timesteps).parameters.keys()):
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support custom"
                             f" sigmas schedules. Please check whether you are using the correct scheduler.")
        scheduler.set_timesteps(sigmas=sigmas, device=device, **kwargs)
    else:
        scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
    return scheduler.timesteps, len(scheduler.timesteps)


class FlowDPMSolverMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, prediction_type: str = "flow_prediction",
                 shift: Optional[float] = 1.0, use_dynamic_shifting=False, thresholding: bool = False,
                 algorithm_type: str = "dpmsolver++", solver_type: str = "midpoint", lower_order_final: bool = True,
                 euler_at_final: bool = False, final_sigmas_type: Optional[str] = "zero", **_):
        if (prediction_type != "flow_prediction" or algorithm_type != "dpmsolver++" or solver_type != "midpoint" or thresholding
                or use_dynamic_shifting or final_sigmas_type != "zero" or solver_order not in (1, 2)):
            raise NotImplementedError("only the configuration WanT2V constructs for 'dpm++' is implemented (text2video.py:424-427)")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, solver_order=solver_order, shift=shift,
                                      lower_order_final=lower_order_final, euler_at_final=euler_at_final,
                                      final_sigmas_type=final_sigmas_type, algorithm_type=algorithm_type, solver_type=solver_type,
                                      prediction_type=prediction_type)
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sigmas = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        self.sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = self.sigmas * num_train_timesteps
        self.sigma_min, self.sigma_max = self.sigmas[-1].item(), self.sigmas[0].item()
        self.num_inference_steps = None
        self._step_index = None

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas: Optional[List[float]] = None, mu=None, shift=None):
        """:226-290"""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        sigmas = np.asarray(sigmas)
        if shift is None:
            shift = self.config.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.config.num_train_timesteps
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0]]).astype(np.float32))      # host scalars
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.timesteps_host = [int(t) for t in torch.from_numpy(timesteps).to(torch.int64)]
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self._step_index = None
        self._bufs = {}

    def _buf(self, name, like):
        b = self._bufs.get(name)
        if b is None or b.shape != like.shape or b.device != like.device:
            b = torch.empty_like(like)
            self._bufs[name] = b
        return b

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None, variance_noise=None,
             return_dict: bool = True):
        """:706-798.  model_output / sample: fp32 CUDA tensors of identical shape."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self._step_index is None:
            t = int(timestep)
            idx = [i for i, v in enumerate(self.timesteps_host) if v == t]
            self._step_index = idx[1] if len(idx) > 1 else idx[0]                    # index_for_timestep :679-691
        i, n = self._step_index, len(self.timesteps_host)
        cfg = self.config
        lower_order_final = (i == n - 1) and (cfg.euler_at_final or (cfg.lower_order_final and n < 15) or cfg.final_sigmas_type == "zero")
        v = model_output.to(torch.float32).contiguous()
        x = sample.to(torch.float32).contiguous()
        sig = self.sigmas
        m0 = self._buf(f"m{i % 3}", x)
        ops.lincomb(m0, [(1.0, x), (-float(sig[i]), v)])                              # x0 = x - sigma*v (:378-380)
        for k in range(cfg.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = m0
        s_t, s_s0 = sig[i + 1], sig[i]
        a_t, a_s0 = 1 - s_t, 1 - s_s0
        lam_t, lam_s0 = torch.log(a_t) - torch.log(s_t), torch.log(a_s0) - torch.log(s_s0)
        h = lam_t - lam_s0
        c0 = a_t * (torch.exp(-h) - 1.0)
        prev = self._buf(f"xp{i % 2}", x)
        if cfg.solver_order == 1 or self.lower_order_nums < 1 or lower_order_final:
            ops.lincomb(prev, [(float(s_t / s_s0), x), (-float(c0), m0)])            # :462-465
        else:
            s_s1 = sig[i - 1]
            lam_s1 = torch.log(1 - s_s1) - torch.log(s_s1)
            r0 = (lam_s0 - lam_s1) / h
            # x_t = s_t/s_s0 x - c0 D0 - 0.5 c0 D1, D1 = (m0 - m1)/r0   (:571-579)
            c1 = float(0.5 * c0 * (1.0 / r0))
            ops.lincomb(prev, [(float(s_t / s_s0), x), (-float(c0) - c1, m0), (c1, self.model_outputs[-2])])
        if self.lower_order_nums < cfg.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev)

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        """:815-854 for schedule timesteps: alpha_t * x + sigma_t * noise"""
        out = []
        for t in torch.as_tensor(timesteps).flatten().tolist():
            idx = [i for i, v in enumerate(self.timesteps_host) if v == int(t)]
            out.append(float(self.sigmas[idx[1] if len(idx) > 1 else idx[0]]))
        s = torch.tensor(out, device=original_samples.device, dtype=original_samples.dtype)
        while s.ndim < original_samples.ndim:
            s = s.unsqueeze(-1)
        return (1 - s) * original_samples + s * noise

    def __len__(self):
        return self.config.num_train_timesteps
