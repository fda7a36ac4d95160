"""ORACLE — test infrastructure only.  Not imported by the product (omni_video_b200/*).

CPU restatement of the reference's flow-matching DPM-Solver++ multistep scheduler as the sampling loop uses it
with `sample_solver='dpm++'` (SAIS-FUXI/Omni-Video omnivideo/utils/fm_solvers.py; call sites
x2x_gen_unified.py:417-427, :486-491): algorithm dpmsolver++, flow prediction, solver_order 2 (1 and 3 supported),
midpoint / heun, lower_order_final, final sigma 0.  Written with explicit difference tensors D0, D1, D2 like the
reference (not as pre-multiplied coefficients like the CUDA path), in torch CPU fp32 so that scalar rounding
matches the reference's 0-dim tensor arithmetic.

Pinning: tests/golden/dpm_*.npz hold trajectories produced by the reference class itself
(oracle/gen_golden_dpm.py, diffusers shimmed); tests/test_oracle_dpm_golden.py.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch


def sampling_sigmas(steps: int, shift: float) -> np.ndarray:
    """get_sampling_sigmas, :23-27"""
    s = np.linspace(1, 0, steps + 1)[:steps]
    return shift * s / (1 + (shift - 1) * s)


class DPMOracle:
    def __init__(self, num_train_timesteps=1000, solver_order=2, solver_type="midpoint", lower_order_final=True,
                 euler_at_final=False):
        self.T, self.order_max, self.solver_type = num_train_timesteps, solver_order, solver_type
        self.lower_order_final, self.euler_at_final = lower_order_final, euler_at_final

    def set_sigmas(self, sigmas: np.ndarray):
        """set_timesteps(sigmas=...) with shift = 1 and final sigma 0, :227-290"""
        self.timesteps = torch.from_numpy(sigmas * self.T).to(torch.int64)
        self.sigmas = torch.from_numpy(np.concatenate([sigmas, [0]]).astype(np.float32))
        self.m: List = [None] * self.order_max        # x0 predictions, newest last
        self.lower, self.i = 0, 0

    @staticmethod
    def _lam(s):
        return torch.log(1 - s) - torch.log(s)

    def step(self, v: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        """:707-799 for the next timestep in order; v = model output (flow), returns prev_sample."""
        i, n = self.i, len(self.timesteps)
        final = i == n - 1                               # final_sigmas_type == "zero": always first order there
        second = i == n - 2 and self.lower_order_final and n < 15
        x0 = sample - self.sigmas[i] * v                 # convert_model_output, :381-383
        self.m = self.m[1:] + [x0]
        x = sample.to(torch.float32)
        s_t, s_0 = self.sigmas[i + 1], self.sigmas[i]
        a_t = 1 - s_t
        h = self._lam(s_t) - self._lam(s_0)
        if self.order_max == 1 or self.lower < 1 or final:                     # :463-469
            out = (s_t / s_0) * x - (a_t * (torch.exp(-h) - 1.0)) * x0
        elif self.order_max == 2 or self.lower < 2 or second:                  # :527-559
            m0, m1 = self.m[-1], self.m[-2]
            r0 = (self._lam(s_0) - self._lam(self.sigmas[i - 1])) / h
            D0, D1 = m0, (1.0 / r0) * (m0 - m1)
            if self.solver_type == "midpoint":
                out = (s_t / s_0) * x - (a_t * (torch.exp(-h) - 1.0)) * D0 - 0.5 * (a_t * (torch.exp(-h) - 1.0)) * D1
            else:
                out = (s_t / s_0) * x - (a_t * (torch.exp(-h) - 1.0)) * D0 + \
                      (a_t * ((torch.exp(-h) - 1.0) / h + 1.0)) * D1
        else:                                                                  # :640-667
            m0, m1, m2 = self.m[-1], self.m[-2], self.m[-3]
            l0, l1, l2 = self._lam(s_0), self._lam(self.sigmas[i - 1]), self._lam(self.sigmas[i - 2])
            r0, r1 = (l0 - l1) / h, (l1 - l2) / h
            D1_0, D1_1 = (1.0 / r0) * (m0 - m1), (1.0 / r1) * (m1 - m2)
            D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
            D2 = (1.0 / (r0 + r1)) * (D1_0 - D1_1)
            out = (s_t / s_0) * x - (a_t * (torch.exp(-h) - 1.0)) * m0 + \
                  (a_t * ((torch.exp(-h) - 1.0) / h + 1.0)) * D1 - \
                  (a_t * ((torch.exp(-h) - 1.0 + h) / h ** 2 - 0.5)) * D2
        if self.lower < self.order_max:
            self.lower += 1
        self.i += 1
        return out.to(x0.dtype)                          # :791
