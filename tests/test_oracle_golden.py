"""The oracle restatement is held to outputs of the reference itself (tests/golden/*.npz, produced
by oracle/gen_golden.py from /root/reference).  CPU only."""
import pytest
import torch

from oracle import dit as O
from tests.golden_util import dit_case, golden_names, load_golden


@pytest.mark.parametrize("name", golden_names("dit_forward"))
def test_oracle_fp32_matches_reference_golden(name):
    meta, gold = load_golden(name)
    cfg, sd, xs, t, ctx, seq_len = dit_case(meta)
    outs = O.dit_forward(sd, cfg, xs, t, ctx, seq_len, emulate_bf16=False)
    assert len(outs) == len(gold)
    for o, g in zip(outs, gold):
        assert o.shape == g.shape and o.dtype == torch.float32
        err = (o - g).abs().max().item()
        assert err <= 2e-4 * max(1.0, g.abs().max().item()), f"{name}: max-abs err {err}"


@pytest.mark.parametrize("name", golden_names("dit_forward")[:2])
def test_oracle_bf16_emulation_within_north_star_tolerance(name):
    """bf16-emulating mode (what the GPU path is compared to) stays within the north-star
    tolerance of the fp32 reference output: rel-L2 <= 1e-2."""
    meta, gold = load_golden(name)
    cfg, sd, xs, t, ctx, seq_len = dit_case(meta)
    from oracle.synth import round_bf16
    outs = O.dit_forward(round_bf16(sd), cfg, [x.bfloat16().float() for x in xs], t,
                         [c.bfloat16().float() for c in ctx], seq_len, emulate_bf16=True)
    for o, g in zip(outs, gold):
        assert O.rel_l2(o, g) <= 1e-2


@pytest.mark.parametrize("name", golden_names("dit_forward_full"))
def test_oracle_fp32_matches_reference_at_baseline_configs(name):
    """Full-depth OmniVideo2-1.3B at BASELINE.json configs[0] (30 layers, 1280 tokens, T5 context
    [77, 4096]) and the A14B width (dim 5120, 40 heads, 2 layers, 2160 tokens): the oracle is pinned
    to the reference's own output where rounding has 30 layers to accumulate."""
    meta, gold = load_golden(name)
    cfg, sd, xs, t, ctx, seq_len = dit_case(meta)
    out = O.dit_forward(sd, cfg, xs, t, ctx, seq_len, emulate_bf16=False)[0]
    assert out.shape == gold[0].shape
    assert O.rel_l2(out, gold[0]) <= 1e-4, f"{name}: rel-L2 {O.rel_l2(out, gold[0])}"
    assert (out - gold[0]).abs().max().item() <= 2e-3 * gold[0].abs().max().item()
