"""oracle/dpm.py is held to trajectories produced by the reference's DPM-Solver++ scheduler itself, driven the way
the pipeline drives it (tests/golden/dpm_*.npz, oracle/gen_golden_dpm.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle.dpm import DPMOracle, sampling_sigmas
from oracle.unipc import synthetic_flow
from tests.golden_util import golden_names
from tests.test_oracle_unipc_golden import load


@pytest.mark.parametrize("name", golden_names("dpm"))
def test_oracle_trajectory_matches_reference_scheduler(name):
    meta, traj = load(name)
    o = DPMOracle(solver_order=meta.get("solver_order", 2), solver_type=meta.get("solver_type", "midpoint"))
    o.set_sigmas(sampling_sigmas(meta["steps"], meta["shift"]))
    assert [int(t) for t in o.timesteps] == meta["timesteps"]
    assert np.array_equal(o.sigmas.numpy(), np.array(meta["sigmas"], dtype=np.float32))
    x = synthetic_flow(-1, tuple(meta["shape"])).to(getattr(torch, meta["dtype"]))
    for i, want in enumerate(traj):
        x = o.step(synthetic_flow(i, tuple(meta["shape"])), x)
        assert x.dtype == getattr(torch, meta["out_dtype"])        # fp32 model output -> fp32 sample (:791)
        err = (x.float() - want).abs().max().item()
        assert err <= 2e-6 * max(1.0, want.abs().max().item()), f"step {i}: {err}"
