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


def test_t5_encoder_at_umt5_xxl_width_vs_oracle():
    """One block at the REAL umT5-XXL width (reference t5.py:467-476: dim 4096, 64 heads of 64, ffn 10240, 32
    buckets, per-layer position bias) on a 512-token prompt batch (77 and 512 valid tokens): every GEMM shape, the
    64-head score/softmax/PV loop and the padding-key mask at the size the pipeline runs.  The depth (24 blocks)
    repeats this block; the vocabulary only sizes an embedding lookup and is cut to 512 rows."""
    from omni_video_b200.modules.t5 import T5Encoder
    cfg = dict(vocab=512, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=1, num_buckets=32,
               shared_pos=False)
    sd = synth.t5_state_dict(cfg, seed=1818)
    with torch.device("cuda"):
        m = T5Encoder(**cfg)
    m.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=True)
    m = m.to(torch.bfloat16).eval().requires_grad_(False)
    lengths, seq_len = [77, 512], 512
    ids, mask = synth.t5_inputs(cfg, lengths, seq_len, seed=7)
    want = T.t5_encode(synth.round_bf16(sd), cfg, ids, mask)
    out = m(ids.cuda(), mask.cuda())
    assert out.shape == (2, seq_len, 4096) and out.dtype == torch.bfloat16
    for b, n in enumerate(lengths):
        err = O.rel_l2(out[b, :n].float().cpu(), want[b])
        print(f"umT5-XXL width, {n} tokens: rel-L2 vs fp32 oracle {err:.3e}")
        # at the real width the averaging over 4096 channels brings the bf16 error under the north-star bound
        # (measured 8.4e-3 / 9.0e-3); the toy cases above keep 2.5e-2 (dim 128, sharp unscaled softmaxes)
        assert bool(torch.isfinite(out[b, :n].float()).all()) and err <= 1e-2
