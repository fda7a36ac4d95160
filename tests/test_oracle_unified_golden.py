"""The conditioning front-end restatement (oracle/unified.py) is held to what the reference's
UnifiedWanWithMixedConditionModel.forward itself hands to WanModel (tests/golden/unified_*.npz,
produced by oracle/gen_golden_unified.py from /root/reference).  CPU only."""
import pytest
import torch

from oracle import synth
from oracle import unified as U
from tests.golden_util import golden_names, load_golden


def unified_case(meta):
    cfg = dict(meta["cfg"])
    patch = tuple(meta["patch"]) if meta["patch"] else (1, 4, 4)
    sd = synth.unified_state_dict(cfg, meta["vlm_in_dim"], patch, seed=meta["weight_seed"],
                                  with_adapter=meta["adapter"])
    visual = tuple(meta["visual"]) if meta["visual"] else None
    ar, ctx, vis, st = synth.unified_inputs(cfg, meta["vlm_in_dim"], meta["n_vlm"], meta["n_text"], visual,
                                            seed=meta["input_seed"], special_tokens=meta["special"])
    return cfg, sd, ar, ctx, vis, st


@pytest.mark.parametrize("name", golden_names("unified_context"))
def test_oracle_mixed_context_matches_reference_golden(name):
    meta, gold = load_golden(name)
    cfg, sd, ar, ctx, vis, st = unified_case(meta)
    outs = U.mixed_context(sd, len(meta["n_text"]), ctx, ar, vis, st, meta["max_len"], cfg["text_dim"], eps=1e-6)
    assert len(outs) == len(gold)
    for o, g in zip(outs, gold):
        assert o.shape == g.shape
        assert (o - g).abs().max().item() <= 1e-4 * max(1.0, g.abs().max().item())


def test_golden_set_covers_truncation_and_absent_conditions():
    metas = [load_golden(n)[0] for n in golden_names("unified_context")]
    assert any(m["max_len"] < sum(x or 0 for x in m["n_text"]) + sum(x or 0 for x in m["n_vlm"]) + 40 for m in metas)
    assert any(not m["adapter"] for m in metas) and any(None in m["n_vlm"] for m in metas)
