"""The umT5-encoder restatement (oracle/t5.py) is held to outputs of the reference's own
T5Encoder (tests/golden/t5_*.npz, produced by oracle/gen_golden_t5.py from /root/reference).  CPU only.
The CUDA side of this row is not built yet: this pins the checker it will be built against."""
import pytest
import torch

from oracle import synth
from oracle import t5 as T
from tests.golden_util import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names("t5_encode"))
def test_oracle_t5_matches_reference_golden(name):
    meta, gold = load_golden(name)
    cfg = dict(meta["cfg"])
    sd = synth.t5_state_dict(cfg, seed=meta["weight_seed"])
    ids, mask = synth.t5_inputs(cfg, meta["lengths"], meta["seq_len"], seed=meta["input_seed"])
    outs = T.t5_encode(sd, cfg, ids, mask)
    assert len(outs) == len(gold)
    for o, g in zip(outs, gold):
        assert o.shape == g.shape
        assert (o - g).abs().max().item() <= 2e-5 * max(1.0, g.abs().max().item())


def test_relative_position_buckets():
    """bidirectional, 32 buckets: 16 per sign; exact up to distance 7, log-spaced to 128, saturating beyond"""
    b = T.relative_position_bucket(300, 300)
    assert b.shape == (300, 300) and int(b.min()) == 0 and int(b.max()) == 31
    assert b[5, 5] == 0 and b[5, 6] == 17 and b[6, 5] == 1 and b[0, 7] == 16 + 7 and b[7, 0] == 7
    assert b[0, 128] == 31 and b[0, 299] == 31 and b[299, 0] == 15 and b[128, 0] == 15
    assert torch.equal(b[10:20, 10:20], b[110:120, 110:120])            # depends on the distance only


def test_padding_keys_do_not_influence_valid_rows():
    cfg = dict(synth.TINY_T5)
    sd = synth.t5_state_dict(cfg, seed=5)
    ids, mask = synth.t5_inputs(cfg, [6], 16, seed=1)
    a = T.t5_encode(sd, cfg, ids, mask)[0]
    ids2 = ids.clone()
    ids2[0, 6:] = 3                                                     # different padding tokens
    b = T.t5_encode(sd, cfg, ids2, mask)[0]
    assert a.shape == (6, cfg["dim"]) and torch.allclose(a, b, atol=1e-6)
