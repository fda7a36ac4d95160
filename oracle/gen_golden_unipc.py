"""ORACLE support — UniPC goldens from the UNMODIFIED reference scheduler
(/root/reference/omnivideo/utils/fm_solvers_unipc.py) with `diffusers` shimmed.
Run in the build container:   python -m oracle.gen_golden_unipc
"""
from __future__ import annotations

import importlib.util
import inspect
import json
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import REF_ROOT  # noqa: E402
from oracle.unipc import synthetic_flow  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference_scheduler(filename="fm_solvers_unipc.py", cls_name="FlowUniPCMultistepScheduler", module=False):
    """diffusers is absent: provide the three names the file imports (ConfigMixin with a working
    `register_to_config`, SchedulerMixin/SchedulerOutput, deprecate).  Also serves oracle/gen_golden_dpm.py
    (fm_solvers.py; `module=True` returns the module next to the class)."""
    class ConfigMixin:
        def register_to_config(self, **kw):
            for k, v in kw.items():
                setattr(self.config, k, v)

    def register_to_config(init):
        sig = inspect.signature(init)

        def wrapped(self, *a, **kw):
            bound = sig.bind(self, *a, **kw)
            bound.apply_defaults()
            self.config = types.SimpleNamespace(**{k: v for k, v in bound.arguments.items() if k != "self"})
            return init(self, *a, **kw)
        return wrapped

    class SchedulerMixin:
        pass

    class SchedulerOutput:
        def __init__(self, prev_sample):
            self.prev_sample = prev_sample

    mods = {
        "diffusers": types.ModuleType("diffusers"),
        "diffusers.configuration_utils": types.ModuleType("diffusers.configuration_utils"),
        "diffusers.schedulers": types.ModuleType("diffusers.schedulers"),
        "diffusers.schedulers.scheduling_utils": types.ModuleType("diffusers.schedulers.scheduling_utils"),
        "diffusers.utils": types.ModuleType("diffusers.utils"),
        "diffusers.utils.torch_utils": types.ModuleType("diffusers.utils.torch_utils"),
    }
    mods["diffusers.utils.torch_utils"].randn_tensor = None        # only the SDE variants draw noise
    mods["diffusers.configuration_utils"].ConfigMixin = ConfigMixin
    mods["diffusers.configuration_utils"].register_to_config = register_to_config
    su = mods["diffusers.schedulers.scheduling_utils"]
    su.KarrasDiffusionSchedulers, su.SchedulerMixin, su.SchedulerOutput = (), SchedulerMixin, SchedulerOutput
    mods["diffusers.utils"].deprecate = lambda *a, **k: None
    mods["diffusers.utils"].is_scipy_available = lambda: False
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        spec = importlib.util.spec_from_file_location(
            "ref_" + filename[:-3], os.path.join(REF_ROOT, "omnivideo", "utils", filename))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    # the reference applies @register_to_config?  it calls self.config.* — make sure config exists
    cls = getattr(mod, cls_name)
    if not hasattr(cls.__init__, "__wrapped__") and "config" not in cls.__init__.__code__.co_names:
        cls.__init__ = register_to_config(cls.__init__)
    return (cls, mod) if module else cls


CASES = [
    dict(name="unipc_8steps_shift5", steps=8, shift=5.0, shape=(1, 16, 2, 4, 4), dtype="float32"),
    dict(name="unipc_50steps_shift5_first12", steps=50, shift=5.0, shape=(1, 16, 2, 4, 4), dtype="float32", run=12),
    dict(name="unipc_6steps_shift3_bf16", steps=6, shift=3.0, shape=(1, 16, 1, 4, 8), dtype="bfloat16"),
]


def main():
    cls = load_reference_scheduler()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for case in CASES:
        sch = cls(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        sch.set_timesteps(case["steps"], device="cpu", shift=case["shift"])
        dt = getattr(torch, case["dtype"])
        x = synthetic_flow(-1, case["shape"]).to(dt)
        traj = []
        n_run = case.get("run", case["steps"])
        for i, t in enumerate(sch.timesteps[:n_run]):
            v = synthetic_flow(i, case["shape"])
            x = sch.step(v, t, x, return_dict=False)[0]
            traj.append(x.float().numpy().copy())
        meta = dict(case, kind="unipc", timesteps=[int(t) for t in sch.timesteps],
                    sigmas=[float(s) for s in sch.sigmas],
                    reference="omnivideo/utils/fm_solvers_unipc.py FlowUniPCMultistepScheduler.step, CPU")
        np.savez_compressed(os.path.join(GOLDEN_DIR, case["name"] + ".npz"), meta=json.dumps(meta),
                            **{f"out{i}": a for i, a in enumerate(traj)})
        print(f"wrote {case['name']}: {len(traj)} steps, final absmax {np.abs(traj[-1]).max():.4f}")


if __name__ == "__main__":
    main()
