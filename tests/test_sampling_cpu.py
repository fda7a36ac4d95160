"""Control flow of the denoising loop (omni_video_b200/sampling.py) on CPU: expert selection at the
boundary, CFG batching vs two calls, argument merging — against a direct restatement of the reference loop
(x2x_gen_unified.py:474-492) with the oracle scheduler.  The scheduler's CUDA launches are replaced by the
oracle's arithmetic here; the kernels themselves are covered by tests/test_gpu_unipc.py and, end to end, by
tests/test_gpu_sampling.py."""
import pytest
import torch

from oracle.unipc import UniPCOracle


class _OracleScheduler:
    """FlowUniPCMultistepScheduler's interface on top of oracle/unipc.py."""

    def __init__(self, num_train_timesteps=1000, shift=1, use_dynamic_shifting=False):
        self.o = UniPCOracle(num_train_timesteps)

    def set_timesteps(self, n, device=None, shift=None):
        self.o.set_timesteps(n, shift)
        self.timesteps = self.o.timesteps

    def step_cfg(self, cond, uncond, g, t, sample, return_dict=False):
        return (self.o.step(uncond + g * (cond - uncond), sample),)


class _Expert:
    """Deterministic stand-in for a DiT expert: out = a * x + b * mean(context) + c * t."""

    def __init__(self, a, b):
        self.a, self.b, self.batches = a, b, []

    def __call__(self, xs, t=None, context=None, seq_len=None, **kw):
        self.batches.append(len(xs))
        assert len(context) == len(xs) and t.shape == (len(xs),)
        return [self.a * x.float() + self.b * c.float().mean() + 1e-4 * float(tt) for x, c, tt in zip(xs, context, t)]


@pytest.fixture
def loop_mod(monkeypatch):
    from omni_video_b200 import sampling
    monkeypatch.setattr(sampling, "FlowUniPCMultistepScheduler", _OracleScheduler)
    return sampling


def _reference_loop(hi, lo, noise, ctx_c, ctx_u, steps, shift, g, boundary):
    o = UniPCOracle()
    o.set_timesteps(steps, shift)
    lat, used = noise, []
    for t in o.timesteps:
        m = hi if t.item() >= boundary * 1000 else lo
        used.append("hi" if m is hi else "lo")
        ts = t.unsqueeze(0)
        c = m([lat], t=ts, context=[ctx_c], seq_len=8)[0]
        u = m([lat], t=ts, context=[ctx_u], seq_len=8)[0]
        lat = o.step(u + g * (c - u), lat.unsqueeze(0)).squeeze(0)
    return lat, used


@pytest.mark.parametrize("batch_cfg", [True, False])
def test_loop_matches_reference_restatement(loop_mod, batch_cfg):
    torch.manual_seed(0)
    noise = torch.randn(16, 2, 4, 4)
    ctx_c, ctx_u = torch.randn(5, 8), torch.randn(3, 8)
    hi, lo = _Expert(0.9, 0.3), _Expert(-0.5, 0.7)
    want, used = _reference_loop(_Expert(0.9, 0.3), _Expert(-0.5, 0.7), noise, ctx_c, ctx_u, 12, 5.0, 4.0, 0.875)
    loop = loop_mod.DenoiseLoop(hi, lo, boundary=0.875, batch_cfg=batch_cfg)
    seen = []
    got = loop.sample(noise, dict(context=[ctx_c], seq_len=8), dict(context=[ctx_u], seq_len=8), sampling_steps=12,
                      shift=5.0, guide_scale=4.0, callback=lambda i, t, x: seen.append(int(t)))[0]
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    n_hi = used.count("hi")
    assert 0 < n_hi < 12, "the schedule must cross the expert boundary for this test to mean anything"
    per = 1 if batch_cfg else 2
    assert loop.calls == {"high_noise_model": n_hi * per, "low_noise_model": (12 - n_hi) * per}
    assert set(hi.batches) == ({2} if batch_cfg else {1}) and len(seen) == 12 and seen == sorted(seen, reverse=True)


def test_single_expert_and_unsupported_solver(loop_mod):
    e = _Expert(1.0, 0.0)
    loop = loop_mod.DenoiseLoop(e)
    assert loop.low_noise_model is e
    with pytest.raises(NotImplementedError):           # x2x_gen_unified.py:428: anything but unipc / dpm++
        loop.sample(torch.zeros(16, 1, 2, 2), {}, {}, sample_solver="euler")


def test_merge_cfg_args():
    from omni_video_b200.sampling import merge_cfg_args
    st = {"<prp_st>": torch.zeros(1, 8)}
    c, u, a, an, v = torch.randn(5, 8), torch.randn(3, 8), torch.randn(6, 4), torch.zeros(2, 4), torch.randn(16, 1, 8, 8)
    base = dict(seq_len=32, special_token_dict=st, classifier_free_ratio=0.0, unconditioned_context=None,
                condition_mode="full")
    m = merge_cfg_args(dict(context=[c], ar_vision_input=[a], visual_emb=[v], **base),
                       dict(context=[u], ar_vision_input=[an], visual_emb=[v], **base))
    assert m["context"][0] is c and m["context"][1] is u and m["ar_vision_input"][1] is an
    assert m["visual_emb"][0] is v and m["visual_emb"][1] is v and m["special_token_dict"] is st and m["seq_len"] == 32
    # tensors with / without a leading batch dimension (unified_model.py:226-233)
    m = merge_cfg_args(dict(context=c[None], visual_emb=v, seq_len=1), dict(context=u[None], visual_emb=v[None], seq_len=1))
    assert torch.equal(m["context"][0], c) and m["visual_emb"][0].shape == v.shape == m["visual_emb"][1].shape
    # absent on both sides stays absent; present on one side only, or different scalars, cannot share a forward
    assert merge_cfg_args(dict(context=[c], ar_vision_input=None), dict(context=[u], ar_vision_input=None))["ar_vision_input"] is None
    assert merge_cfg_args(dict(context=[c], ar_vision_input=[a]), dict(context=[u], ar_vision_input=None)) is None
    assert merge_cfg_args(dict(context=[c], seq_len=32), dict(context=[u], seq_len=64)) is None
    assert merge_cfg_args(dict(context=[c, c]), dict(context=[u])) is None
