"""Per-token timesteps, `t` of shape [B, seq_len] (reference model.py:466-476): oracle and — with emulated
kernels — the engine's host path against the reference's own output (tests/golden/dit_tiny_pertoken.npz)."""
import torch

from oracle import dit as O
from oracle import synth
from tests import cpu_emulation
from tests.golden_util import load_golden

bf16 = torch.bfloat16


def pertoken_case():
    meta, gold = load_golden("dit_tiny_pertoken")
    cfg = dict(meta["cfg"])
    cfg["patch_size"], cfg["window_size"] = tuple(cfg["patch_size"]), tuple(cfg["window_size"])
    sd = synth.dit_state_dict(cfg, seed=meta["weight_seed"])
    xs, _, ctx = synth.dit_inputs(cfg, tuple(meta["grid"]), meta["ctx_len"], seed=meta["input_seed"],
                                  batch=meta["batch"], timestep=0)
    L = meta["seq_len"]
    t = torch.stack([(torch.arange(L) * 37 + 11 + 400 * b) % 1000 for b in range(meta["batch"])]).to(torch.int64)
    return cfg, sd, xs, t, ctx, L, gold


def test_oracle_per_token_timesteps_match_reference_golden():
    cfg, sd, xs, t, ctx, seq_len, gold = pertoken_case()
    outs = O.dit_forward(sd, cfg, xs, t, ctx, seq_len, emulate_bf16=False)
    for o, g in zip(outs, gold):
        assert o.shape == g.shape and (o - g).abs().max().item() <= 2e-4 * max(1.0, g.abs().max().item())


def test_engine_per_token_timesteps_with_emulated_kernels(monkeypatch):
    from omni_video_b200.modules.model import WanModel
    cpu_emulation.install(monkeypatch)
    cfg, sd, xs, t, ctx, seq_len, gold = pertoken_case()
    m = WanModel(**cfg, skip_init=True)
    m.load_state_dict(sd, strict=True)
    m = m.to(bf16).eval().requires_grad_(False)
    outs = m(xs, t=t, context=[c.to(bf16) for c in ctx], seq_len=seq_len)
    for o, g in zip(outs, gold):
        assert o.shape == g.shape and O.rel_l2(o, g) <= 1e-2
