"""FlowUniPCMultistepScheduler — drop-in for wan/utils/fm_solvers_unipc.py:20-800 (order-2 UniPC, bh2,
predict_x0, flow_prediction; Wan's default sampler).  All scalar coefficients (log-SNR steps, expm1 terms,
rho's, the 2x2 solve) are computed on the host with the same torch fp32 scalar ops as the reference; every
tensor update is a linear combination executed by one fp32 kernel (`ops.lincomb`): x0 conversion (:321),
UniC corrector (:590-626) and UniP predictor (:458-484) are three launches instead of ~25."""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from .. import ops


class FlowUniPCMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, prediction_type: str = "flow_prediction",
                 shift: Optional[float] = 1.0, use_dynamic_shifting=False, thresholding: bool = False,
                 predict_x0: bool = True, solver_type: str = "bh2", lower_order_final: bool = True,
                 disable_corrector: List[int] = [], final_sigmas_type: Optional[str] = "zero", **_):
        if prediction_type != "flow_prediction" or not predict_x0 or solver_type != "bh2" or thresholding or use_dynamic_shifting:
            raise NotImplementedError("only the configuration WanT2V/WanI2V construct is implemented (text2video.py:419-422)")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, solver_order=solver_order, shift=shift,
                                      lower_order_final=lower_order_final, final_sigmas_type=final_sigmas_type,
                                      solver_type=solver_type, prediction_type=prediction_type)
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sigmas = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.sigmas = sigmas
        self.sigma_min, self.sigma_max = self.sigmas[-1].item(), self.sigmas[0].item()
        self.disable_corrector = disable_corrector
        self.num_inference_steps = None
        self._step_index = None

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, shift=None):
        """:160-227"""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.config.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.config.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)                       # host: scalars stay on the CPU (:226-227)
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.timesteps_host = [int(t) for t in torch.from_numpy(timesteps).to(torch.int64)]
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self._step_index = None
        self.this_order = 1
        self._bufs = {}

    # ---- host scalar helpers (identical op sequence to :386-452 / :523-588) ----
    def _coeffs(self, s_t, s_s0):
        a_t, a_s0 = 1 - s_t, 1 - s_s0
        lam_t = torch.log(a_t) - torch.log(s_t)
        lam_s0 = torch.log(a_s0) - torch.log(s_s0)
        h = lam_t - lam_s0
        hh = -h
        h_phi_1 = torch.expm1(hh)
        return a_t, lam_s0, h, hh, h_phi_1, torch.expm1(hh)

    @staticmethod
    def _rb(rks, hh, h_phi_1, B_h, order):
        R, b = [], []
        h_phi_k = h_phi_1 / hh - 1
        fact = 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return torch.stack(R), torch.tensor(b)

    def _buf(self, name, like):
        b = self._bufs.get(name)
        if b is None or b.shape != like.shape or b.device != like.device:
            b = torch.empty_like(like)
            self._bufs[name] = b
        return b

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True, generator=None):
        """:655-739.  model_output / sample: fp32 CUDA tensors of identical shape."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self._step_index is None:
            t = int(timestep)
            idx = [i for i, v in enumerate(self.timesteps_host) if v == t]
            self._step_index = idx[1] if len(idx) > 1 else idx[0]                    # index_for_timestep :626-640
        i = self._step_index
        v = model_output.to(torch.float32).contiguous()
        x = sample.to(torch.float32).contiguous()
        sig = self.sigmas
        # ring of three x0 buffers so history is never overwritten while still referenced
        m_t = self._buf(f"m{i % 3}", x)
        ops.lincomb(m_t, [(1.0, x), (-float(sig[i]), v)])                             # x0 = x - sigma*v (:321)
        use_corrector = i > 0 and (i - 1) not in self.disable_corrector and self.last_sample is not None
        if use_corrector:
            order = self.this_order
            s_t, s_s0 = sig[i], sig[i - 1]
            a_t, lam_s0, h, hh, h_phi_1, B_h = self._coeffs(s_t, s_s0)
            m0 = self.model_outputs[-1]
            rks, hist = [], []
            for k in range(1, order):
                si = i - (k + 1)
                lam_si = torch.log(1 - sig[si]) - torch.log(sig[si])
                rk = (lam_si - lam_s0) / h
                rks.append(rk)
                hist.append((self.model_outputs[-(k + 1)], rk))
            rks.append(1.0)
            R, b = self._rb(torch.tensor(rks), hh, h_phi_1, B_h, order)
            rhos_c = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
            # x = s_t/s_s0*last - a_t*h_phi_1*m0 - a_t*B_h*(sum_k rho_k (m_k - m0)/rk + rho_last (m_t - m0))
            c_m0 = -float(a_t * h_phi_1) + float(a_t * B_h * rhos_c[-1])
            terms = [(float(s_t / s_s0), self.last_sample), (-float(a_t * B_h * rhos_c[-1]), m_t)]
            for k, (mk, rk) in enumerate(hist):
                ck = float(a_t * B_h * rhos_c[k] / rk)
                terms.append((-ck, mk))
                c_m0 += ck
            terms.append((c_m0, m0))
            xc = self._buf(f"xc{i % 2}", x)
            ops.lincomb(xc, terms)
            x = xc
        for k in range(self.config.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = m_t
        this_order = min(self.config.solver_order, len(self.timesteps_host) - i) if self.config.lower_order_final else self.config.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = x
        order = self.this_order
        s_t, s_s0 = sig[i + 1], sig[i]
        a_t, lam_s0, h, hh, h_phi_1, B_h = self._coeffs(s_t, s_s0)
        terms = [(float(s_t / s_s0), x)]
        c_m0 = -float(a_t * h_phi_1)
        if order == 2:
            si = i - 1
            lam_si = torch.log(1 - sig[si]) - torch.log(sig[si])
            rk = (lam_si - lam_s0) / h
            c1 = float(a_t * B_h * 0.5 / rk)                                          # rhos_p = 0.5 (:458-459)
            terms.append((-c1, self.model_outputs[-2]))
            c_m0 += c1
        elif order > 2:
            raise NotImplementedError("solver_order > 2")
        terms.append((c_m0, m_t))
        prev = self._buf(f"xp{i % 2}", x)
        ops.lincomb(prev, terms)
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev)

    def scale_model_input(self, sample, *args, **kwargs):
        return sample
