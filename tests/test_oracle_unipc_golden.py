"""oracle/unipc.py is held to trajectories produced by the reference scheduler itself
(tests/golden/unipc_*.npz, oracle/gen_golden_unipc.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.unipc import UniPCOracle, synthetic_flow
from tests.golden_util import GOLDEN_DIR, golden_names


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    n = len([k for k in z.files if k.startswith("out")])
    return meta, [torch.from_numpy(z[f"out{i}"]) for i in range(n)]


@pytest.mark.parametrize("name", golden_names("unipc"))
def test_oracle_trajectory_matches_reference_scheduler(name):
    meta, traj = load(name)
    o = UniPCOracle()
    o.set_timesteps(meta["steps"], meta["shift"])
    assert [int(t) for t in o.timesteps] == meta["timesteps"]
    assert np.allclose(o.sigmas.numpy(), np.array(meta["sigmas"], dtype=np.float32), rtol=0, atol=0)
    dt = getattr(torch, meta["dtype"])
    x = synthetic_flow(-1, tuple(meta["shape"])).to(dt)
    tol = 2e-6 if dt == torch.float32 else 2 ** -7       # bf16 trajectory: one ulp of the rounded state
    for i, want in enumerate(traj):
        x = o.step(synthetic_flow(i, tuple(meta["shape"])), x)
        assert x.dtype == dt
        err = (x.float() - want).abs().max().item()
        assert err <= tol * max(1.0, want.abs().max().item()), f"step {i}: {err}"
