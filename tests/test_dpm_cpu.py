"""Host logic of the product's flow-matching DPM-Solver++ scheduler (omni_video_b200/utils/fm_solvers.py: sigma and
timestep tables, order selection, the pre-multiplied scalar coefficients of the first / second (midpoint, heun) /
third order updates, the x0 history) on CPU, with its single kernel `ovb_lincomb` replaced by the same linear
combination in torch — against trajectories of the reference scheduler (tests/golden/dpm_*.npz) — and the
`sample_solver='dpm++'` branch of the denoising loop.  The kernel itself: tests/test_gpu_unipc.py."""
import numpy as np
import pytest
import torch

from oracle.dpm import DPMOracle, sampling_sigmas
from oracle.unipc import synthetic_flow
from tests.golden_util import golden_names
from tests.test_oracle_unipc_golden import load
from tests.test_unipc_cpu import cpu_lincomb  # noqa: F401  (fixture)

f32 = torch.float32


@pytest.mark.parametrize("name", golden_names("dpm"))
def test_scheduler_host_logic_vs_reference_trajectory(cpu_lincomb, name):  # noqa: F811
    from omni_video_b200.utils.fm_solvers import (FlowDPMSolverMultistepScheduler, get_sampling_sigmas,
                                                  retrieve_timesteps)
    meta, traj = load(name)
    kw = {k: meta[k] for k in ("solver_order", "solver_type") if k in meta}
    sch = FlowDPMSolverMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False, **kw)
    timesteps, n = retrieve_timesteps(sch, device="cpu", sigmas=get_sampling_sigmas(meta["steps"], meta["shift"]))
    assert n == meta["steps"] and [int(t) for t in timesteps] == meta["timesteps"]
    assert np.array_equal(sch.sigmas.numpy(), np.array(meta["sigmas"], dtype=np.float32))
    x = synthetic_flow(-1, tuple(meta["shape"])).to(getattr(torch, meta["dtype"]))
    for i, (t, want) in enumerate(zip(timesteps, traj)):
        x = sch.step(synthetic_flow(i, tuple(meta["shape"])), t, x, return_dict=False)[0]
        assert x.dtype == getattr(torch, meta["out_dtype"])
        err = (x.float() - want).abs().max().item()
        assert err <= 2e-5 * max(1.0, want.abs().max().item()), f"step {i}: {err}"


def test_step_cfg_equals_step_on_the_guided_prediction(cpu_lincomb):  # noqa: F811
    from omni_video_b200.utils.fm_solvers import FlowDPMSolverMultistepScheduler, get_sampling_sigmas
    shape, g = (1, 16, 1, 4, 8), 5.0
    a = FlowDPMSolverMultistepScheduler()
    b = FlowDPMSolverMultistepScheduler()
    for s in (a, b):
        s.set_timesteps(sigmas=get_sampling_sigmas(6, 3.0), device="cpu")
    xa = xb = synthetic_flow(-1, shape)
    for i, t in enumerate(a.timesteps):
        cond, uncond = synthetic_flow(i, shape), synthetic_flow(100 + i, shape)
        xa = a.step(uncond + g * (cond - uncond), t, xa, return_dict=False)[0]
        xb = b.step_cfg(cond, uncond, g, t, xb)[0]
        assert (xa - xb).abs().max().item() <= 2e-5 * max(1.0, xa.abs().max().item())


def test_unsupported_configurations_are_refused():
    from omni_video_b200.utils.fm_solvers import FlowDPMSolverMultistepScheduler, retrieve_timesteps
    for kw in (dict(prediction_type="epsilon"), dict(algorithm_type="sde-dpmsolver++"), dict(thresholding=True),
               dict(use_dynamic_shifting=True), dict(final_sigmas_type="sigma_min"), dict(solver_type="euler"),
               dict(solver_order=4)):
        with pytest.raises(NotImplementedError):
            FlowDPMSolverMultistepScheduler(**kw)
    with pytest.raises(ValueError):
        retrieve_timesteps(FlowDPMSolverMultistepScheduler(), timesteps=[1], sigmas=[0.5])
    with pytest.raises(ValueError):                                     # set_timesteps takes sigmas, not timesteps
        retrieve_timesteps(FlowDPMSolverMultistepScheduler(), timesteps=[999, 500])
    with pytest.raises(ValueError):
        FlowDPMSolverMultistepScheduler().step(torch.zeros(4), 0, torch.zeros(4))      # set_timesteps not called


def test_denoise_loop_with_dpm_solver_matches_the_reference_loop(cpu_lincomb):  # noqa: F811
    """`DenoiseLoop.sample(sample_solver='dpm++')` (x2x_gen_unified.py:417-427, :474-492) against the loop written
    out with the oracle scheduler: timesteps from the shifted sigmas, expert switch at the boundary, CFG."""
    from omni_video_b200.sampling import DenoiseLoop

    class Expert:
        def __init__(self, a):
            self.a = a

        def __call__(self, xs, t=None, context=None, seq_len=None):
            return [self.a * x.float() + c.float().mean() + 1e-4 * float(tt) for x, c, tt in zip(xs, context, t)]

    hi, lo = Expert(0.3), Expert(-0.2)
    noise = synthetic_flow(-1, (16, 1, 4, 8))
    ctx_c, ctx_n = torch.full((3, 8), 0.5), torch.zeros(3, 8)
    steps, shift, g, boundary = 8, 5.0, 4.0, 0.875
    loop = DenoiseLoop(hi, lo, boundary=boundary)
    got = loop.sample(noise, dict(context=[ctx_c], seq_len=32), dict(context=[ctx_n], seq_len=32),
                      sampling_steps=steps, shift=shift, guide_scale=g, sample_solver="dpm++")[0]
    o = DPMOracle()
    o.set_sigmas(sampling_sigmas(steps, shift))
    x = noise
    for t in o.timesteps:
        model = hi if float(t) >= boundary * 1000 else lo
        cond = model([x], t=t.unsqueeze(0), context=[ctx_c])[0]
        uncond = model([x], t=t.unsqueeze(0), context=[ctx_n])[0]
        x = o.step((uncond + g * (cond - uncond)).unsqueeze(0), x.unsqueeze(0)).squeeze(0)
    assert loop.calls["high_noise_model"] > 0 and loop.calls["low_noise_model"] > 0
    assert got.shape == noise.shape and (got - x).abs().max().item() <= 2e-5 * max(1.0, x.abs().max().item())
    with pytest.raises(NotImplementedError):
        loop.sample(noise, {}, {}, sample_solver="euler")
