"""B200-native flow-matching UniPC sampler update (drop-in for the part of reference
omnivideo/utils/fm_solvers_unipc.py that the sampling loop uses, x2x_gen_unified.py:409-416,
:486-491): `FlowUniPCMultistepScheduler(num_train_timesteps, shift, use_dynamic_shifting=False)`,
`.set_timesteps(n, device, shift)`, `.timesteps`, `.step(model_output, t, sample, return_dict=False)`.

Every tensor expression of a step is a linear combination of latent-sized tensors with scalar
coefficients: x0 = sample - sigma*v (:319-323), UniC corrector (:610-623), UniP predictor (:467-475).
The scalars follow the reference's own fp32 arithmetic (0-dim CPU tensors); each combination is
ONE launch of `ovb_lincomb` instead of ~6 ATen kernels.  `step_cfg` additionally folds classifier-free
guidance `u + g (c - u)` (x2x_gen_unified.py:484) into the x0 prediction.

Supported: the configuration the pipeline instantiates — flow_prediction, predict_x0, bh2 (or bh1),
solver_order <= 3, lower_order_final, final sigma 0, no thresholding / dynamic shifting / solver_p.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .. import ops

__all__ = ["FlowUniPCMultistepScheduler"]


class _Cfg:
    pass


class FlowUniPCMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2,
                 prediction_type: str = "flow_prediction", shift: Optional[float] = 1.0,
                 use_dynamic_shifting: bool = False, thresholding: bool = False,
                 predict_x0: bool = True, solver_type: str = "bh2", lower_order_final: bool = True,
                 disable_corrector: List[int] = [], solver_p=None, final_sigmas_type: str = "zero", **unused):
        if prediction_type != "flow_prediction" or not predict_x0 or thresholding or use_dynamic_shifting \
                or solver_p is not None or final_sigmas_type != "zero":
            raise NotImplementedError("only the flow_prediction / predict_x0 / final sigma 0 configuration "
                                      "used by the Omni-Video pipeline is implemented")
        if solver_type in ("midpoint", "heun", "logrho"):
            solver_type = "bh2"
        if solver_type not in ("bh1", "bh2"):
            raise NotImplementedError(f"{solver_type} is not implemented for {self.__class__}")
        if not 1 <= solver_order <= 3:
            raise NotImplementedError("solver_order must be 1, 2 or 3")
        self.config = _Cfg()
        self.config.num_train_timesteps = num_train_timesteps
        self.config.solver_order = solver_order
        self.config.solver_type = solver_type
        self.config.lower_order_final = lower_order_final
        self.config.shift = shift
        self.predict_x0 = True
        self.disable_corrector = list(disable_corrector)
        # reference :107-133
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sigmas = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.sigmas = sigmas
        self.timesteps = sigmas * num_train_timesteps
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()
        self.num_inference_steps = None
        self._reset()

    def _reset(self):
        self.model_outputs = [None] * self.config.solver_order
        self.timestep_list = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self._step_index = None
        self._begin_index = None
        self.this_order = None

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, shift=None):
        """reference :161-228"""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.config.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.config.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)            # CPU, fp32 (reference :227)
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        self._reset()

    # ------------------------------------------------------------------ scalar side (fp32, as :404-456)
    def _lambda(self, sigma):
        return torch.log(1 - sigma) - torch.log(sigma)

    def _bh_coefficients(self, sigma_t, sigma_s0, rk_sigmas, order):
        """h_phi_1, B_h, rks, R, b of :404-456 / :549-600 as python floats / fp32 tensors."""
        lambda_t, lambda_s0 = self._lambda(sigma_t), self._lambda(sigma_s0)
        h = lambda_t - lambda_s0
        rks = [((self._lambda(s) - lambda_s0) / h) for s in rk_sigmas]
        rks.append(1.0)
        rks_t = torch.tensor(rks)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = hh if self.config.solver_type == "bh1" else torch.expm1(hh)
        R, b = [], []
        factorial_i = 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks_t, i - 1))
            b.append(h_phi_k * factorial_i / B_h)
            factorial_i *= i + 1
            h_phi_k = h_phi_k / hh - 1 / factorial_i
        return h_phi_1, B_h, rks, torch.stack(R), torch.tensor(b)

    # ------------------------------------------------------------------ tensor side (one launch each)
    def _combine(self, terms, like):
        """sum of (coefficient, tensor) with the reference's dtype behaviour: a `0-dim fp32 * bf16
        tensor` product is bf16 (type promotion), everything else fp32; result in `like.dtype`."""
        terms = [(float(c), x) for c, x in terms if x is not None]
        rnd = [k for k, (_, x) in enumerate(terms) if k == 0 and x.dtype == torch.bfloat16]
        return ops.lincomb(terms, out_dtype=like.dtype, round_bf16=rnd)

    def _predictor(self, sample, order):
        """multistep_uni_p_bh_update, reference :351-485"""
        m0 = self.model_outputs[-1]
        si = self.step_index
        sigma_t, sigma_s0 = self.sigmas[si + 1], self.sigmas[si]
        alpha_t = 1 - sigma_t
        rk_sig = [self.sigmas[si - i] for i in range(1, order)]
        h_phi_1, B_h, rks, R, b = self._bh_coefficients(sigma_t, sigma_s0, rk_sig, order)
        c_m = [0.0] * order                       # coefficients of m0, m_{-2}, m_{-3}
        c_m[0] = float(-alpha_t * h_phi_1)
        if order > 1:
            rhos_p = torch.tensor([0.5]) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
            for k in range(order - 1):            # pred_res = sum_k rhos_p[k] * (m_k - m0) / rk_k
                w = float(-alpha_t * B_h * rhos_p[k] / rks[k])
                c_m[k + 1] += w
                c_m[0] -= w
        terms = [(sigma_t / sigma_s0, sample)] + [(c_m[k], self.model_outputs[-(k + 1)]) for k in range(order)]
        return self._combine(terms, sample)

    def _corrector(self, model_t, last_sample, this_sample, order):
        """multistep_uni_c_bh_update, reference :487-627"""
        m0 = self.model_outputs[-1]
        si = self.step_index
        sigma_t, sigma_s0 = self.sigmas[si], self.sigmas[si - 1]
        alpha_t = 1 - sigma_t
        rk_sig = [self.sigmas[si - (i + 1)] for i in range(1, order)]
        h_phi_1, B_h, rks, R, b = self._bh_coefficients(sigma_t, sigma_s0, rk_sig, order)
        rhos_c = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
        c_m = [0.0] * order
        c_m[0] = float(-alpha_t * h_phi_1)
        for k in range(order - 1):
            w = float(-alpha_t * B_h * rhos_c[k] / rks[k])
            c_m[k + 1] += w
            c_m[0] -= w
        c_t = float(-alpha_t * B_h * rhos_c[-1])  # * (model_t - m0)
        c_m[0] -= c_t
        terms = [(sigma_t / sigma_s0, last_sample)] + \
                [(c_m[k], self.model_outputs[-(k + 1)]) for k in range(order)] + [(c_t, model_t)]
        return self._combine(terms, this_sample)

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        st = self.timesteps if schedule_timesteps is None else schedule_timesteps
        indices = (st == timestep).nonzero()
        pos = 1 if len(indices) > 1 else 0
        return indices[pos].item()

    def _init_step_index(self, timestep):
        if self.begin_index is None:
            if isinstance(timestep, torch.Tensor):
                timestep = timestep.to(self.timesteps.device)
            self._step_index = self.index_for_timestep(timestep)
        else:
            self._step_index = self._begin_index

    # ------------------------------------------------------------------ public step
    def step(self, model_output, timestep, sample, return_dict: bool = True, generator=None):
        """reference :656-740.  Returns a tuple `(prev_sample,)` (the pipeline passes
        return_dict=False); with return_dict=True an object with `.prev_sample`."""
        return self._step([(1.0, model_output)], timestep, sample, return_dict)

    def step_cfg(self, cond, uncond, guide_scale: float, timestep, sample, return_dict: bool = False):
        """step() on `uncond + guide_scale * (cond - uncond)` without materialising it."""
        return self._step([(1.0 - guide_scale, uncond), (guide_scale, cond)], timestep, sample, return_dict)

    def _step(self, v_terms, timestep, sample, return_dict):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after "
                             "creating the scheduler")
        v_terms = [(c, x.reshape(sample.shape).contiguous()) for c, x in v_terms]
        sample = sample.contiguous()
        if self.step_index is None:
            self._init_step_index(timestep)
        use_corrector = (self.step_index > 0 and self.step_index - 1 not in self.disable_corrector
                         and self.last_sample is not None)
        sigma = float(self.sigmas[self.step_index])
        # convert_model_output: x0 = sample - sigma * v   (fp32, :319-323)
        x0 = ops.lincomb([(1.0, sample)] + [(-sigma * c, x) for c, x in v_terms], out_dtype=torch.float32)
        if use_corrector:
            sample = self._corrector(x0, self.last_sample, sample, self.this_order)
        for i in range(self.config.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
            self.timestep_list[i] = self.timestep_list[i + 1]
        self.model_outputs[-1] = x0
        self.timestep_list[-1] = timestep
        if self.config.lower_order_final:
            this_order = min(self.config.solver_order, len(self.timesteps) - self.step_index)
        else:
            this_order = self.config.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        assert self.this_order > 0
        self.last_sample = sample
        prev_sample = self._predictor(sample, self.this_order)
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        if not return_dict:
            return (prev_sample,)
        out = _Cfg()
        out.prev_sample = prev_sample
        return out

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def __len__(self):
        return self.config.num_train_timesteps
