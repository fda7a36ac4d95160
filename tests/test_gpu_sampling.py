"""The denoising loop (omni_video_b200/sampling.py) end to end on the GPU: two resident experts switched at
the boundary, CFG-batched forwards, fused UniPC update — against a step-by-step restatement of the reference
loop (x2x_gen_unified.py:474-492) that calls the same experts one sample at a time and does guidance and the
scheduler update with the fp32 CPU oracle."""
import random

import pytest
import torch

from oracle import dit as O
from oracle import synth
from oracle.unipc import UniPCOracle
from tests.golden_util import load_golden
from tests.test_oracle_unified_golden import unified_case

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _expert(seed):
    from omni_video_b200.modules.model import WanModel
    cfg = dict(synth.TINY_DIT)
    with torch.device("cuda"):
        m = WanModel(**cfg)
    m.load_state_dict({k: v.cuda() for k, v in synth.dit_state_dict(cfg, seed=seed).items()})
    return m.to(bf16).eval().requires_grad_(False)


def test_denoise_loop_two_experts_matches_stepwise_restatement():
    from omni_video_b200.sampling import DenoiseLoop
    hi, lo = _expert(11), _expert(12)
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(16, 3, 8, 12, generator=g)
    ctx_c = torch.randn(20, 64, generator=g).to(bf16).cuda()
    ctx_u = torch.randn(9, 64, generator=g).to(bf16).cuda()
    L, steps, shift, gs, boundary = 3 * 4 * 6, 6, 5.0, 3.0, 0.9

    o = UniPCOracle()
    o.set_timesteps(steps, shift)
    want, n_hi = noise.clone(), 0
    for t in o.timesteps:
        m = hi if t.item() >= boundary * 1000 else lo
        n_hi += m is hi
        ts = t.unsqueeze(0).cuda()
        c = m([want.cuda()], t=ts, context=[ctx_c], seq_len=L)[0].cpu()
        u = m([want.cuda()], t=ts, context=[ctx_u], seq_len=L)[0].cpu()
        want = o.step(u + gs * (c - u), want.unsqueeze(0)).squeeze(0)
    assert 0 < n_hi < steps

    loop = DenoiseLoop(hi, lo, boundary=boundary)
    got = loop.sample(noise.cuda(), dict(context=[ctx_c], seq_len=L), dict(context=[ctx_u], seq_len=L),
                      sampling_steps=steps, shift=shift, guide_scale=gs)[0]
    assert got.shape == noise.shape and bool(torch.isfinite(got).all())
    assert loop.calls == {"high_noise_model": n_hi, "low_noise_model": steps - n_hi}     # one B=2 forward per step
    assert O.rel_l2(got.float().cpu(), want) <= 2e-3


def test_denoise_loop_vs_cpu_oracle_end_to_end():
    """The WHOLE loop against the CPU oracle, nothing shared with the GPU path: every forward of the
    restatement is `oracle.dit.dit_forward` (bf16-emulating, on bf16-rounded weights, fp32 CPU), guidance and
    the UniPC update are `oracle.unipc`.  Four steps of a two-expert schedule on the tiny DiT; the latent after
    the last step must agree within the north-star per-step tolerance, although errors compound over steps."""
    from omni_video_b200.sampling import DenoiseLoop
    cfg = dict(synth.TINY_DIT)
    sd_hi, sd_lo = synth.dit_state_dict(cfg, seed=11), synth.dit_state_dict(cfg, seed=12)
    hi, lo = _expert(11), _expert(12)
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(16, 2, 8, 8, generator=g).to(bf16).float()
    ctx_c = torch.randn(12, 64, generator=g).to(bf16)
    ctx_u = torch.randn(5, 64, generator=g).to(bf16)
    L, steps, shift, gs, boundary = 2 * 4 * 4, 4, 5.0, 3.0, 0.9

    o = UniPCOracle()
    o.set_timesteps(steps, shift)
    want = noise.clone()
    for t in o.timesteps:
        sd = synth.round_bf16(sd_hi if t.item() >= boundary * 1000 else sd_lo)
        xin = [want.to(bf16).float()]                     # the loop keeps bf16 latents (x2x_gen_unified.py:382-391)
        c = O.dit_forward(sd, cfg, xin, t.reshape(1), [ctx_c.float()], L, emulate_bf16=True)[0]
        u = O.dit_forward(sd, cfg, xin, t.reshape(1), [ctx_u.float()], L, emulate_bf16=True)[0]
        want = o.step(u + gs * (c - u), want.unsqueeze(0)).squeeze(0)

    loop = DenoiseLoop(hi, lo, boundary=boundary)
    got = loop.sample(noise.cuda(), dict(context=[ctx_c.cuda()], seq_len=L), dict(context=[ctx_u.cuda()], seq_len=L),
                      sampling_steps=steps, shift=shift, guide_scale=gs)[0]
    err = O.rel_l2(got.float().cpu(), want)
    print(f"denoise loop, {steps} steps, two experts: rel-L2 vs the CPU oracle loop {err:.3e}")
    assert bool(torch.isfinite(got).all()) and err <= 1e-2


def test_denoise_loop_unified_model_batched_equals_unbatched_and_keeps_rng_stream():
    from omni_video_b200.sampling import DenoiseLoop
    from tests.test_gpu_unified import build, to_cuda
    meta, _ = load_golden("unified_full")
    cfg, sd, ar, ctx, vis, st = unified_case(meta)
    model = build(meta, cfg, sd)
    model.wan_model.load_state_dict({k: v.cuda() for k, v in synth.dit_state_dict(dict(cfg), seed=3).items()})
    model = model.to(bf16)
    ar, ctx, vis, st = to_cuda(ar, ctx, vis, st)
    null_ctx = [torch.zeros(4, cfg["text_dim"], device="cuda", dtype=bf16)]
    null_ar = [torch.zeros(2, meta["vlm_in_dim"], device="cuda", dtype=bf16) + 1e-6]
    base = dict(seq_len=32, visual_emb=vis, special_token_dict=st, classifier_free_ratio=0.0,
                unconditioned_context=None, condition_mode="full")
    arg_c = dict(context=ctx, ar_vision_input=ar, **base)
    arg_null = dict(context=null_ctx, ar_vision_input=null_ar, **base)
    noise = torch.randn(16, 2, 8, 8, generator=torch.Generator().manual_seed(1)).cuda()
    outs, after = [], []
    for batch_cfg in (True, False):
        random.seed(77)
        loop = DenoiseLoop(model, batch_cfg=batch_cfg)
        outs.append(loop.sample(noise, arg_c, arg_null, sampling_steps=4, guide_scale=2.0)[0])
        after.append(random.random())
        assert sum(loop.calls.values()) == (4 if batch_cfg else 8)
    assert after[0] == after[1]                      # 4 draws per step either way (unified_model.py:160-162)
    assert O.rel_l2(outs[0].float().cpu(), outs[1].float().cpu()) <= 1e-3
    assert model.context_cache_misses <= 3           # the front-end ran per distinct condition set, not per step
