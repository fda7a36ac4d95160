"""Host logic of the product's flow-matching UniPC scheduler (omni_video_b200/utils/fm_solvers_unipc.py: the
sigma/timestep tables, orders, the scalar coefficients of UniP/UniC, the model-output history) on CPU, with its
single kernel `ovb_lincomb` replaced by the same linear combination in torch — against trajectories of the
reference scheduler (tests/golden/unipc_*.npz).  The kernel itself: tests/test_gpu_unipc.py."""
import pytest
import torch

from oracle.unipc import synthetic_flow
from tests.golden_util import golden_names
from tests.test_oracle_unipc_golden import load

bf16, f32 = torch.bfloat16, torch.float32


@pytest.fixture
def cpu_lincomb(monkeypatch):
    from omni_video_b200 import ops

    def lincomb(terms, out=None, out_dtype=None, round_bf16=()):
        acc = None
        for k, (c, x) in enumerate(terms):
            v = x.float()
            if k in round_bf16:
                v = v.to(bf16).float()
            acc = c * v if acc is None else acc + c * v
        res = acc.to(out_dtype or terms[0][1].dtype)
        if out is not None:
            out.copy_(res)
            return out
        return res

    monkeypatch.setattr(ops, "lincomb", lincomb)


@pytest.mark.parametrize("name", golden_names("unipc"))
def test_scheduler_host_logic_vs_reference_trajectory(cpu_lincomb, name):
    from omni_video_b200.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler
    meta, traj = load(name)
    sch = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    sch.set_timesteps(meta["steps"], device="cpu", shift=meta["shift"])
    assert [int(t) for t in sch.timesteps] == meta["timesteps"]
    dt = getattr(torch, meta["dtype"])
    x = synthetic_flow(-1, tuple(meta["shape"])).to(dt)
    tol = 2e-5 if dt == f32 else 2 ** -6
    for i, (t, want) in enumerate(zip(sch.timesteps, traj)):
        x = sch.step(synthetic_flow(i, tuple(meta["shape"])), t, x, return_dict=False)[0]
        assert x.dtype == dt
        err = (x.float() - want).abs().max().item()
        assert err <= tol * max(1.0, want.abs().max().item()), f"step {i}: {err}"


def test_unsupported_configurations_are_refused():
    from omni_video_b200.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler
    for kw in (dict(prediction_type="epsilon"), dict(predict_x0=False), dict(thresholding=True),
               dict(use_dynamic_shifting=True), dict(final_sigmas_type="sigma_min")):
        with pytest.raises(NotImplementedError):
            FlowUniPCMultistepScheduler(**kw)
