"""A batch of samples with DIFFERENT latent grids in one `WanModel.forward` call (reference model.py:455-464:
every sample zero-padded to seq_len, attending to its own tokens, rotated with its own (f, h, w)): oracle and — with
emulated kernels — the engine's host path against the reference's own output (tests/golden/dit_tiny_ragged.npz)."""
import torch

from oracle import dit as O
from oracle import synth
from tests import cpu_emulation
from tests.golden_util import load_golden

bf16 = torch.bfloat16


def ragged_case():
    meta, gold = load_golden("dit_tiny_ragged")
    cfg = dict(meta["cfg"])
    cfg["patch_size"], cfg["window_size"] = tuple(cfg["patch_size"]), tuple(cfg["window_size"])
    sd = synth.dit_state_dict(cfg, seed=meta["weight_seed"])
    xs, ctx = [], []
    for grid, n, seed in zip(meta["grids"], meta["ctx_lens"], meta["input_seeds"]):
        x, t, c = synth.dit_inputs(cfg, tuple(grid), n, seed=seed, batch=1, timestep=meta["t"])
        xs += x
        ctx += c
    return cfg, sd, xs, torch.cat([t, t]), ctx, meta["seq_len"], gold


def test_oracle_ragged_batch_matches_reference_golden():
    cfg, sd, xs, t, ctx, seq_len, gold = ragged_case()
    outs = O.dit_forward(sd, cfg, xs, t, ctx, seq_len, emulate_bf16=False)
    for o, g in zip(outs, gold):
        assert o.shape == g.shape and (o - g).abs().max().item() <= 2e-4 * max(1.0, g.abs().max().item())


def test_engine_ragged_batch_with_emulated_kernels(monkeypatch):
    from omni_video_b200.modules.model import WanModel
    cpu_emulation.install(monkeypatch)
    cfg, sd, xs, t, ctx, seq_len, gold = ragged_case()
    m = WanModel(**cfg, skip_init=True)
    m.load_state_dict(sd, strict=True)
    m = m.to(bf16).eval().requires_grad_(False)
    outs = m(xs, t=t, context=[c.to(bf16) for c in ctx], seq_len=seq_len)
    assert [tuple(o.shape) for o in outs] == [tuple(g.shape) for g in gold]
    for o, g in zip(outs, gold):
        assert O.rel_l2(o, g) <= 1e-2
