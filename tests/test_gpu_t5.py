"""umT5 encoder (omni_video_b200/modules/t5.py) on the GPU against the fp32 oracle that is
pinned to the reference's T5Encoder (tests/golden/t5_*.npz): ragged batch with padding keys, and a sequence
whose distances exceed the 128-position bucket range."""
import pytest
import torch

from oracle import dit as O
from oracle import synth
from oracle import t5 as T

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


@pytest.mark.parametrize("shared_pos,lengths,seq_len", [(False, [9, 64], 64), (False, [130], 256)])
def test_t5_encoder_vs_oracle(shared_pos, lengths, seq_len):
    from omni_video_b200.modules.t5 import T5Encoder
    cfg = dict(synth.TINY_T5, shared_pos=shared_pos)
    sd = synth.t5_state_dict(cfg, seed=1818)
    with torch.device("cuda"):
        m = T5Encoder(**cfg)
    m.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=True)
    m = m.to(torch.bfloat16).eval().requires_grad_(False)
    ids, mask = synth.t5_inputs(cfg, lengths, seq_len, seed=7)
    want = T.t5_encode(synth.round_bf16(sd), cfg, ids, mask)
    out = m(ids.cuda(), mask.cuda())
    assert out.shape == (len(lengths), seq_len, cfg["dim"]) and out.dtype == torch.bfloat16
    for b, n in enumerate(lengths):
        assert bool(torch.isfinite(out[b, :n].float()).all())
        assert O.rel_l2(out[b, :n].float().cpu(), want[b]) <= 2.5e-2
