"""Host logic of the umT5 encoder composition (omni_video_b200/modules/t5.py) on CPU: the kernels it calls are
replaced by torch emulations WITH THEIR ROUNDING POINTS (bf16 GEMM results, fp32 score buffer, bf16 P ...), and
the result is compared with the fp32 oracle that is pinned to the reference's T5Encoder.  This checks indexing,
head views of the fused QKV buffer, the relative-bias table and key counts, the gated FFN plumbing and the residual flow; the kernels themselves are
covered by tests/test_gpu_kernels.py / test_gpu_vae.py, the module on the GPU by tests/test_gpu_t5.py."""
import pytest
import torch
import torch.nn.functional as F

from oracle import dit as O
from oracle import synth
from oracle import t5 as T

bf16, f32 = torch.bfloat16, torch.float32


def _emulate(monkeypatch):
    from omni_video_b200 import ops

    def gemm(a, w, bias, *, epilogue=ops.EPI_BF16, out=None, resid=None, gate=None, rows_per_sample=0, **kw):
        assert a.dtype == bf16 and w.dtype == bf16 and a.shape[1] == w.shape[1] and a.stride(1) == 1 and w.stride(1) == 1
        acc = a.float() @ w.float().T + (0 if bias is None else bias.float())
        accb = acc.to(bf16).float()
        if epilogue == ops.EPI_BF16:
            res = accb.to(bf16)
        elif epilogue == ops.EPI_GELU_BF16:
            res = F.gelu(acc, approximate="tanh").to(bf16)
        elif epilogue == ops.EPI_RESID_F32:
            res = resid + accb
        elif epilogue == ops.EPI_GATE_RESID_F32:
            assert gate.shape == acc.shape and gate.dtype == f32
            res = resid + accb * gate
        elif epilogue == ops.EPI_RESID_BF16:
            res = (resid.float() + accb).to(bf16)
        else:
            raise AssertionError(epilogue)
        if out is None:
            return res
        assert out.shape == res.shape and out.dtype == res.dtype
        out.copy_(res)
        return out

    def rmsnorm_rope_(x, weight, eps, out=None, **kw):
        assert not kw and x.dtype == bf16
        xf = x.float()
        dst = x if out is None else out                     # out: the input stays untouched (the residual stream)
        dst.copy_(((xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(bf16).float() * weight.float()).to(bf16))
        return dst

    def attention_relbias(q, k, v, rel_bias, k_lens=None, softmax_scale=1.0, out=None):
        """ovb_attn_relbias_fwd (include/ovb200.h): fp32 scores + bias[h, (j - i) + L - 1], keys >= k_lens[b]
        masked, fp32 softmax, P rounded to bf16, fp32 accumulation of P V."""
        B, L, H, D = q.shape
        assert q.dtype == bf16 and k.shape == q.shape and v.shape == q.shape and out.shape == q.shape
        assert rel_bias.dtype == f32 and rel_bias.shape == (H, 2 * L - 1)
        assert k_lens is None or (k_lens.dtype == torch.int32 and k_lens.shape == (B,))
        delta = torch.arange(L)[None, :] - torch.arange(L)[:, None] + L - 1            # [i, j]
        S = torch.einsum("bihc,bjhc->bhij", q.float(), k.float()) * softmax_scale + rel_bias[:, delta][None]
        if k_lens is not None:
            S = S.masked_fill(torch.arange(L)[None, None, None, :] >= k_lens.view(B, 1, 1, 1), float("-inf"))
        P = torch.softmax(S, dim=-1).to(bf16).float()
        out.copy_(torch.einsum("bhij,bjhc->bihc", P, v.float()).to(bf16))
        return out

    monkeypatch.setattr(ops, "gemm", gemm)
    monkeypatch.setattr(ops, "rmsnorm_rope_", rmsnorm_rope_)
    monkeypatch.setattr(ops, "attention_relbias", attention_relbias)
    monkeypatch.setattr(ops, "cast_bf16_to_f32", lambda x: x.float())
    monkeypatch.setattr(ops, "cast_f32_to_bf16", lambda x: x.to(bf16))


@pytest.mark.parametrize("shared_pos,lengths,seq_len", [(False, [9, 64], 64), (False, [130], 256), (True, [70], 128)])
def test_t5_encoder_composition_matches_oracle(monkeypatch, shared_pos, lengths, seq_len):
    from omni_video_b200.modules.t5 import T5Encoder
    _emulate(monkeypatch)
    cfg = dict(synth.TINY_T5, shared_pos=shared_pos)
    sd = synth.t5_state_dict(cfg, seed=1818)
    m = T5Encoder(**cfg).eval().requires_grad_(False)
    m.load_state_dict(sd, strict=True)
    ids, mask = synth.t5_inputs(cfg, lengths, seq_len, seed=7)
    want = T.t5_encode(synth.round_bf16(sd), cfg, ids, mask)
    out = m._encode(ids, mask.gt(0).sum(1).to(torch.int32))          # the whole ragged batch in one pass
    assert out.shape == (len(lengths), seq_len, cfg["dim"]) and out.dtype == bf16
    for b, n in enumerate(lengths):
        got = out[b, :n]
        assert got.shape == want[b].shape
        # bf16 rounding of every GEMM result against an fp32 oracle, through unscaled (sharp) softmaxes: ~1.4e-2
        # here; a slicing / masking / residual mistake gives O(1)
        assert O.rel_l2(got.float(), want[b]) <= 2.5e-2


def test_t5_state_dict_names_are_the_references():
    from omni_video_b200.modules.t5 import T5Encoder
    cfg = dict(synth.TINY_T5)
    ours = set(T5Encoder(**cfg).state_dict().keys())
    assert ours == set(synth.t5_state_dict(cfg).keys())     # the synthetic dict loads strict=True into the reference too
    with pytest.raises(RuntimeError):
        T5Encoder(**cfg)(torch.zeros(1, 64, dtype=torch.int64))          # no CPU fallback
    with pytest.raises(NotImplementedError):                             # 32-wide heads: no such kernel
        T5Encoder(**dict(cfg, num_heads=4))._encode(torch.zeros(1, 24, dtype=torch.int64), None)


def test_relative_bucket_vector_is_the_diagonal_form_of_the_reference_matrix():
    """`relative_buckets(L)[(j - i) + L - 1] == buckets(L, L)[i, j]` (reference t5.py:241-263): the [heads, 2L-1]
    bias table ovb_attn_relbias_fwd reads holds exactly the reference's [heads, L, L] bias."""
    from omni_video_b200.modules.t5 import T5RelativeEmbedding
    for L in (1, 7, 64, 300):
        e = T5RelativeEmbedding(32, 2, bidirectional=True)
        full, vec = e.buckets(L, L), e.relative_buckets(L)
        assert vec.shape == (2 * L - 1,)
        delta = torch.arange(L)[None, :] - torch.arange(L)[:, None] + L - 1
        assert torch.equal(vec[delta], full)


def test_prompt_cleaning_matches_reference_rules():
    from omni_video_b200.modules.t5 import clean_whitespace
    assert clean_whitespace("  a  cat\n\tsits &amp;amp; purrs  ") == "a cat sits & purrs"
    assert clean_whitespace("plain") == "plain"
