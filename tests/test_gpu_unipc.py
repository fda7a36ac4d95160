"""The CUDA sampler update (omni_video_b200/utils/fm_solvers_unipc.py: one ovb_lincomb launch per
tensor expression) against trajectories of the reference scheduler (goldens) and the oracle."""
import pytest
import torch

from oracle.unipc import UniPCOracle, synthetic_flow
from tests.golden_util import golden_names
from tests.test_oracle_unipc_golden import load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("unipc"))
def test_scheduler_trajectory_vs_reference_golden(name):
    from omni_video_b200.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler
    meta, traj = load(name)
    sch = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    sch.set_timesteps(meta["steps"], device="cuda", shift=meta["shift"])
    assert [int(t) for t in sch.timesteps] == meta["timesteps"]
    dt = getattr(torch, meta["dtype"])
    x = synthetic_flow(-1, tuple(meta["shape"])).to(dt).cuda()
    tol = 1e-5 if dt == torch.float32 else 2 ** -6
    for i, (t, want) in enumerate(zip(sch.timesteps, traj)):
        v = synthetic_flow(i, tuple(meta["shape"])).cuda()
        x = sch.step(v, t, x, return_dict=False)[0]
        assert x.dtype == dt and x.shape == tuple(meta["shape"])
        err = (x.float().cpu() - want).abs().max().item()
        assert err <= tol * max(1.0, want.abs().max().item()), f"step {i}: {err}"


def test_step_cfg_equals_step_on_guided_prediction():
    from omni_video_b200.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler
    shape, g = (1, 16, 3, 8, 8), 3.0
    a = FlowUniPCMultistepScheduler(shift=1)
    b = FlowUniPCMultistepScheduler(shift=1)
    o = UniPCOracle()
    for s in (a, b):
        s.set_timesteps(10, device="cuda", shift=5.0)
    o.set_timesteps(10, 5.0)
    xa = xb = synthetic_flow(-1, shape).cuda()
    xo = synthetic_flow(-1, shape)
    for i, t in enumerate(a.timesteps):
        c, u = synthetic_flow(2 * i, shape).cuda(), synthetic_flow(2 * i + 1, shape).cuda()
        xa = a.step(u + g * (c - u), t, xa, return_dict=False)[0]
        xb = b.step_cfg(c, u, g, t, xb)[0]
        xo = o.step((u + g * (c - u)).cpu(), xo)
        assert (xa - xb).abs().max().item() <= 2e-5 * max(1.0, xa.abs().max().item())
        assert (xb.cpu() - xo).abs().max().item() <= 2e-5 * max(1.0, xo.abs().max().item())


def test_lincomb_kernel(  ):
    from omni_video_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = [torch.randn(1000 * 4, generator=g, device="cuda") for _ in range(5)]
    xb = x[1].bfloat16()
    c = [0.5, -1.25, 3.0, 1e-3, -7.0]
    out = ops.lincomb([(c[0], x[0]), (c[1], xb), (c[2], x[2]), (c[3], x[3]), (c[4], x[4])])
    ref = c[0] * x[0] + c[1] * xb.float() + c[2] * x[2] + c[3] * x[3] + c[4] * x[4]
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    outb = ops.lincomb([(2.0, xb), (1.0, x[0])], out_dtype=torch.bfloat16, round_bf16=[0])
    refb = ((2.0 * xb.float()).bfloat16().float() + x[0]).bfloat16()
    assert (outb.float() - refb.float()).abs().max().item() <= 2 ** -7 * refb.float().abs().max().item()


@pytest.mark.parametrize("name", golden_names("dpm"))
def test_dpm_scheduler_trajectory_vs_reference_golden(name):
    """`sample_solver='dpm++'` (reference utils/fm_solvers.py, x2x_gen_unified.py:417-427): the DPM-Solver++ update
    as ovb_lincomb launches against trajectories of the reference scheduler — second order midpoint (the default),
    third order on a bf16 latent, heun."""
    from omni_video_b200.utils.fm_solvers import (FlowDPMSolverMultistepScheduler, get_sampling_sigmas,
                                                  retrieve_timesteps)
    meta, traj = load(name)
    kw = {k: meta[k] for k in ("solver_order", "solver_type") if k in meta}
    sch = FlowDPMSolverMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False, **kw)
    timesteps, _ = retrieve_timesteps(sch, device="cuda", sigmas=get_sampling_sigmas(meta["steps"], meta["shift"]))
    assert timesteps.is_cuda and [int(t) for t in timesteps] == meta["timesteps"]
    x = synthetic_flow(-1, tuple(meta["shape"])).to(getattr(torch, meta["dtype"])).cuda()
    for i, (t, want) in enumerate(zip(timesteps, traj)):
        x = sch.step(synthetic_flow(i, tuple(meta["shape"])).cuda(), t, x, return_dict=False)[0]
        assert x.dtype == getattr(torch, meta["out_dtype"]) and x.shape == tuple(meta["shape"])
        err = (x.float().cpu() - want).abs().max().item()
        assert err <= 2e-5 * max(1.0, want.abs().max().item()), f"step {i}: {err}"
