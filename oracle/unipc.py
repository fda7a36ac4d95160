"""ORACLE — test infrastructure only.  Not imported by the product (omni_video_b200/*).

CPU restatement of the reference's flow-matching UniPC multistep scheduler as the sampling loop
uses it (SAIS-FUXI/Omni-Video omnivideo/utils/fm_solvers_unipc.py; call sites
x2x_gen_unified.py:409-416, :486-491): bh2, solver_order 2 (general order supported),
predict_x0 with flow prediction, lower_order_final, final sigma 0.  Written with explicit
difference tensors like the reference (not as pre-multiplied coefficients like the CUDA path), in
torch CPU fp32 so that scalar rounding matches the reference's 0-dim tensor arithmetic.

Pinning: tests/golden/unipc_*.npz hold trajectories produced by the reference class itself
(oracle/gen_golden_unipc.py, diffusers shimmed); tests/test_oracle_unipc_golden.py.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch


class UniPCOracle:
    def __init__(self, num_train_timesteps=1000, solver_order=2, solver_type="bh2", lower_order_final=True):
        self.T, self.order_max, self.solver_type = num_train_timesteps, solver_order, solver_type
        self.lower_order_final = lower_order_final
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        s = torch.from_numpy(1.0 - alphas).to(torch.float32)       # :107-116 with shift = 1
        self.sigma_min, self.sigma_max = s[-1].item(), s[0].item()

    def set_timesteps(self, n, shift):
        """:161-228"""
        sig = np.linspace(self.sigma_max, self.sigma_min, n + 1).copy()[:-1]
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.timesteps = torch.from_numpy(sig * self.T).to(torch.int64)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0]]).astype(np.float32))
        self.m: List = [None] * self.order_max       # x0 predictions, newest last
        self.lower, self.last_sample, self.i, self.this_order = 0, None, 0, None

    @staticmethod
    def _lam(s):
        return torch.log(1 - s) - torch.log(s)

    def _coeffs(self, s_t, s_0, s_prev, order):
        """shared scalar part of UniP (:404-456) and UniC (:549-600)"""
        h = self._lam(s_t) - self._lam(s_0)
        rks = torch.tensor([(self._lam(s) - self._lam(s_0)) / h for s in s_prev] + [1.0])
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = hh if self.solver_type == "bh1" else torch.expm1(hh)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return h_phi_1, B_h, rks, torch.stack(R), torch.tensor(b)

    def step(self, v: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        """:656-740 for the next timestep in order; v = model output (flow), returns prev_sample."""
        i = self.i
        x0 = sample - self.sigmas[i] * v                                   # :319-323
        if i > 0 and self.last_sample is not None:                         # UniC, :487-627
            order, m0 = self.this_order, self.m[-1]
            s_t, s_0 = self.sigmas[i], self.sigmas[i - 1]
            prev = [self.sigmas[i - (k + 1)] for k in range(1, order)]
            h_phi_1, B_h, rks, R, b = self._coeffs(s_t, s_0, prev, order)
            D1s = [(self.m[-(k + 1)] - m0) / rks[k - 1] for k in range(1, order)]
            rhos = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
            x_ = s_t / s_0 * self.last_sample - (1 - s_t) * h_phi_1 * m0
            corr = sum(rhos[k] * D1s[k] for k in range(len(D1s))) if D1s else 0
            sample = (x_ - (1 - s_t) * B_h * (corr + rhos[-1] * (x0 - m0))).to(sample.dtype)
        self.m = self.m[1:] + [x0]
        n = len(self.timesteps)
        order = min(self.order_max, n - i) if self.lower_order_final else self.order_max
        self.this_order = order = min(order, self.lower + 1)
        self.last_sample = sample
        m0 = self.m[-1]                                                    # UniP, :351-485
        s_t, s_0 = self.sigmas[i + 1], self.sigmas[i]
        prev = [self.sigmas[i - k] for k in range(1, order)]
        h_phi_1, B_h, rks, R, b = self._coeffs(s_t, s_0, prev, order)
        D1s = [(self.m[-(k + 1)] - m0) / rks[k - 1] for k in range(1, order)]
        x_ = s_t / s_0 * sample - (1 - s_t) * h_phi_1 * m0
        if D1s:
            rhos = torch.tensor([0.5]) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
            x_ = x_ - (1 - s_t) * B_h * sum(rhos[k] * D1s[k] for k in range(len(D1s)))
        if self.lower < self.order_max:
            self.lower += 1
        self.i += 1
        return x_.to(sample.dtype)


def synthetic_flow(step: int, shape, seed: int = 11) -> torch.Tensor:
    """Deterministic stand-in for the DiT output at `step` (PCG64: identical on every machine)."""
    rng = np.random.default_rng(seed * 1000 + step)
    return torch.from_numpy(rng.standard_normal(shape, dtype=np.float32))
