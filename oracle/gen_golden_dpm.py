"""ORACLE support — DPM-Solver++ goldens from the UNMODIFIED reference scheduler
(/root/reference/omnivideo/utils/fm_solvers.py) with `diffusers` shimmed, driven exactly as the pipeline drives it
(x2x_gen_unified.py:417-427: `get_sampling_sigmas` -> `retrieve_timesteps(sigmas=...)`).
Run in the build container:   python -m oracle.gen_golden_dpm
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden_unipc import GOLDEN_DIR, load_reference_scheduler  # noqa: E402
from oracle.unipc import synthetic_flow  # noqa: E402

CASES = [
    dict(name="dpm_8steps_shift5", steps=8, shift=5.0, shape=(1, 16, 2, 4, 4), dtype="float32"),
    dict(name="dpm_50steps_shift5_first12", steps=50, shift=5.0, shape=(1, 16, 2, 4, 4), dtype="float32", run=12),
    dict(name="dpm_20steps_shift3_order3_bf16latent", steps=20, shift=3.0, shape=(1, 16, 1, 4, 8), dtype="bfloat16",
         solver_order=3),
    dict(name="dpm_6steps_shift3_heun", steps=6, shift=3.0, shape=(1, 16, 1, 4, 8), dtype="float32", solver_type="heun"),
]


def main():
    cls, mod = load_reference_scheduler("fm_solvers.py", "FlowDPMSolverMultistepScheduler", module=True)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for case in CASES:
        kw = {k: case[k] for k in ("solver_order", "solver_type") if k in case}
        sch = cls(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False, **kw)
        timesteps, _ = mod.retrieve_timesteps(sch, device="cpu", sigmas=mod.get_sampling_sigmas(case["steps"], case["shift"]))
        dt = getattr(torch, case["dtype"])
        x = synthetic_flow(-1, case["shape"]).to(dt)
        traj = []
        n_run = case.get("run", case["steps"])
        for i, t in enumerate(timesteps[:n_run]):
            v = synthetic_flow(i, case["shape"])
            x = sch.step(v, t, x, return_dict=False)[0]
            traj.append(x.float().numpy().copy())
        meta = dict(case, kind="dpm", timesteps=[int(t) for t in sch.timesteps], sigmas=[float(s) for s in sch.sigmas],
                    out_dtype=str(x.dtype).replace("torch.", ""),
                    reference="omnivideo/utils/fm_solvers.py FlowDPMSolverMultistepScheduler.step, CPU")
        np.savez_compressed(os.path.join(GOLDEN_DIR, case["name"] + ".npz"), meta=json.dumps(meta),
                            **{f"out{i}": a for i, a in enumerate(traj)})
        print(f"wrote {case['name']}: {len(traj)} steps, out dtype {x.dtype}, final absmax {np.abs(traj[-1]).max():.4f}")


if __name__ == "__main__":
    main()
