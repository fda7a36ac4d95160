"""B200-native flow-matching DPM-Solver++ sampler update (drop-in for the part of reference
omnivideo/utils/fm_solvers.py that the sampling loop uses when `sample_solver == 'dpm++'`,
x2x_gen_unified.py:417-427, :486-491): `get_sampling_sigmas(steps, shift)`, `retrieve_timesteps(scheduler,
device=..., sigmas=...)`, `FlowDPMSolverMultistepScheduler(num_train_timesteps, shift, use_dynamic_shifting=False)`
with `.set_timesteps`, `.timesteps`, `.step(model_output, t, sample, return_dict=False)`.

As in fm_solvers_unipc.py every tensor expression of a step is a linear combination of latent-sized tensors with
scalar coefficients — x0 = sample - sigma*v (:381-383), the first-order update (:463-469), the second-order
multistep update in its midpoint and Heun forms (:548-559) and the third-order one (:659-667) — so each is ONE
launch of `ovb_lincomb`; the scalars follow the reference's fp32 arithmetic on 0-dim CPU tensors.  `step_cfg` folds
classifier-free guidance (x2x_gen_unified.py:484) into the x0 prediction.

Supported: what the pipeline instantiates and its deterministic neighbours — algorithm_type "dpmsolver++",
flow_prediction, solver_order <= 3, solver_type midpoint / heun, lower_order_final, euler_at_final, final sigma 0;
no SDE variants (they draw noise), thresholding or dynamic shifting.
"""
from __future__ import annotations

import inspect
from typing import Optional

import numpy as np
import torch

from .. import ops

__all__ = ["FlowDPMSolverMultistepScheduler", "get_sampling_sigmas", "retrieve_timesteps"]


def get_sampling_sigmas(sampling_steps, shift):
    """reference :23-27"""
    sigma = np.linspace(1, 0, sampling_steps + 1)[:sampling_steps]
    return shift * sigma / (1 + (shift - 1) * sigma)


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None, **kwargs):
    """reference :30-67: `scheduler.set_timesteps` with custom timesteps or sigmas -> (timesteps, count)."""
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed. Please choose one to set custom values")
    accepted = set(inspect.signature(scheduler.set_timesteps).parameters.keys())
    if timesteps is not None:
        if "timesteps" not in accepted:
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support "
                             "custom timestep schedules. Please check whether you are using the correct scheduler.")
        scheduler.set_timesteps(timesteps=timesteps, device=device, **kwargs)
    elif sigmas is not None:
        if "sigmas" not in accepted:
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support "
                             "custom sigmas schedules. Please check whether you are using the correct scheduler.")
        scheduler.set_timesteps(sigmas=sigmas, device=device, **kwargs)
    else:
        scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
        return scheduler.timesteps, num_inference_steps
    return scheduler.timesteps, len(scheduler.timesteps)


class _Cfg:
    pass


class FlowDPMSolverMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2,
                 prediction_type: str = "flow_prediction", shift: Optional[float] = 1.0,
                 use_dynamic_shifting: bool = False, thresholding: bool = False,
                 algorithm_type: str = "dpmsolver++", solver_type: str = "midpoint", lower_order_final: bool = True,
                 euler_at_final: bool = False, final_sigmas_type: Optional[str] = "zero", **unused):
        if algorithm_type == "deis":                                   # reference :155-157
            algorithm_type = "dpmsolver++"
        if prediction_type != "flow_prediction" or algorithm_type != "dpmsolver++" or thresholding \
                or use_dynamic_shifting or final_sigmas_type != "zero":
            raise NotImplementedError("only the deterministic dpmsolver++ / flow_prediction / final sigma 0 "
                                      "configuration used by the Omni-Video pipeline is implemented")
        if solver_type in ("logrho", "bh1", "bh2"):                    # reference :162-164
            solver_type = "midpoint"
        if solver_type not in ("midpoint", "heun"):
            raise NotImplementedError(f"{solver_type} is not implemented for {self.__class__}")
        if not 1 <= solver_order <= 3:
            raise NotImplementedError("solver_order must be 1, 2 or 3")
        self.config = _Cfg()
        self.config.num_train_timesteps = num_train_timesteps
        self.config.solver_order = solver_order
        self.config.solver_type = solver_type
        self.config.algorithm_type = algorithm_type
        self.config.lower_order_final = lower_order_final
        self.config.euler_at_final = euler_at_final
        self.config.final_sigmas_type = final_sigmas_type
        self.config.shift = shift
        # reference :176-200
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
        self.lower_order_nums = 0
        self._step_index = None
        self._begin_index = None

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, shift=None):
        """reference :227-290 (the pipeline passes the already shifted `get_sampling_sigmas` and shift = 1)"""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        sigmas = np.asarray(sigmas)
        if shift is None:
            shift = self.config.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.config.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)      # final_sigmas_type == "zero"
        self.sigmas = torch.from_numpy(sigmas)                         # CPU, fp32
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        self._reset()

    # ------------------------------------------------------------------ scalar side (0-dim fp32, as the reference)
    @staticmethod
    def _lambda(sigma):
        return torch.log(1 - sigma) - torch.log(sigma)

    def _update_terms(self, sample, order):
        """[(coefficient, tensor)] of x_t for the update of this order; the model-output history holds x0
        predictions, newest last."""
        si = self.step_index
        sigma_t, sigma_s0 = self.sigmas[si + 1], self.sigmas[si]
        alpha_t = 1 - sigma_t
        lambda_t, lambda_s0 = self._lambda(sigma_t), self._lambda(sigma_s0)
        h = lambda_t - lambda_s0
        e = torch.exp(-h) - 1.0
        a = alpha_t * e                                                # x_t = (sigma_t/sigma_s0) x - a D0 ...
        m = self.model_outputs
        head = (float(sigma_t / sigma_s0), sample)
        if order == 1:                                                 # :463-469
            return [head, (float(-a), m[-1])]
        lambda_s1 = self._lambda(self.sigmas[si - 1])
        r0 = (lambda_s0 - lambda_s1) / h
        if order == 2:                                                 # :548-559, D1 = (m0 - m1) / r0
            w = 0.5 * a if self.config.solver_type == "midpoint" else -(alpha_t * (e / h + 1.0))
            return [head, (float(-a - w / r0), m[-1]), (float(w / r0), m[-2])]
        lambda_s2 = self._lambda(self.sigmas[si - 2])                  # :640-667
        r1 = (lambda_s1 - lambda_s2) / h
        b = alpha_t * (e / h + 1.0)                                    # + b D1
        c2 = alpha_t * ((e + h) / h ** 2 - 0.5)                        # - c2 D2
        k = r0 / (r0 + r1)                                             # D1 = D1_0 + k (D1_0 - D1_1), D2 = (D1_0 - D1_1)/(r0+r1)
        p = b * (1 + k) - c2 / (r0 + r1)                               # coefficient of D1_0 = (m0 - m1) / r0
        q = -b * k + c2 / (r0 + r1)                                    # coefficient of D1_1 = (m1 - m2) / r1
        return [head, (float(-a + p / r0), m[-1]), (float(-p / r0 + q / r1), m[-2]), (float(-q / r1), m[-3])]

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
    def step(self, model_output, timestep, sample, generator=None, variance_noise=None, return_dict: bool = True):
        """reference :707-799.  Returns `(prev_sample,)` with return_dict=False (what the pipeline passes)."""
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
        n, si, cfg = len(self.timesteps), self.step_index, self.config
        lower_order_final = si == n - 1 and (cfg.euler_at_final or (cfg.lower_order_final and n < 15)
                                             or cfg.final_sigmas_type == "zero")
        lower_order_second = si == n - 2 and cfg.lower_order_final and n < 15
        # convert_model_output: x0 = sample - sigma * v in the promoted dtype (:381-383).  The DiT returns fp32, so x0
        # and with it the new sample are fp32 whatever the latent's dtype (:791 casts to x0's dtype); with a bf16
        # model output the reference would also round the products of the update to bf16, here they stay fp32
        dtype = sample.dtype
        for _, x in v_terms:
            dtype = torch.promote_types(dtype, x.dtype)
        sigma = float(self.sigmas[si])
        x0 = ops.lincomb([(1.0, sample)] + [(-sigma * c, x) for c, x in v_terms], out_dtype=dtype)
        for i in range(cfg.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = x0
        if cfg.solver_order == 1 or self.lower_order_nums < 1 or lower_order_final:
            order = 1
        elif cfg.solver_order == 2 or self.lower_order_nums < 2 or lower_order_second:
            order = 2
        else:
            order = 3
        # the reference upcasts the sample to fp32, evaluates the update there and casts to x0's dtype (:762, :791)
        prev_sample = ops.lincomb(self._update_terms(sample, order), out_dtype=dtype)
        if self.lower_order_nums < cfg.solver_order:
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
