"""The image-to-video flavour of WanModel.forward (`y` concatenated to `x` along channels, reference model.py
:417-424, :446-447; `model_type='i2v'` makes `y` mandatory): oracle and — with emulated kernels — the engine's
host path against the reference's own output (tests/golden/dit_tiny_i2v.npz).  CPU only."""
import pytest
import torch

from oracle import dit as O
from tests import cpu_emulation
from tests.golden_util import dit_case, load_golden

bf16 = torch.bfloat16


def i2v_case():
    meta, gold = load_golden("dit_tiny_i2v")
    cfg, sd, xs, t, ctx, seq_len = dit_case(meta)
    nx = cfg["in_dim"] - meta["y_channels"]
    return cfg, sd, [u[:nx] for u in xs], [u[nx:] for u in xs], xs, t, ctx, seq_len, gold[0]


def test_oracle_i2v_matches_reference_golden():
    cfg, sd, x, y, xcat, t, ctx, seq_len, gold = i2v_case()
    out = O.dit_forward(sd, cfg, xcat, t, ctx, seq_len, emulate_bf16=False)[0]        # cat([x, y]) is all `y` does
    assert (out - gold).abs().max().item() <= 2e-4 * max(1.0, gold.abs().max().item())


def test_engine_i2v_path_with_emulated_kernels(monkeypatch):
    from omni_video_b200.modules.model import WanModel
    cpu_emulation.install(monkeypatch)
    cfg, sd, x, y, xcat, t, ctx, seq_len, gold = i2v_case()
    m = WanModel(**cfg, skip_init=True)
    m.load_state_dict(sd, strict=True)
    m = m.to(bf16).eval().requires_grad_(False)
    out = m(x, t=t, context=[c.to(bf16) for c in ctx], seq_len=seq_len, y=y)[0]
    assert out.shape == gold.shape and O.rel_l2(out, gold) <= 1e-2
    with pytest.raises(AssertionError):
        m(x, t=t, context=[c.to(bf16) for c in ctx], seq_len=seq_len)               # i2v without y (model.py:446-447)
