"""Parity of every libovb200 kernel against the oracle (oracle/dit.py, CPU) on seeded inputs, and
against a plain fp32 torch restatement at sizes the CPU oracle would take too long for.  All calls
go through the C ABI (omni_video_b200.ops -> ctypes -> libovb200.so)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dit as O

pytestmark = pytest.mark.gpu

bf16, f32 = torch.bfloat16, torch.float32


def _rng(seed):
    return np.random.default_rng(seed)


def _randn(rng, *shape, scale=1.0):
    return torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(scale))


@pytest.fixture(scope="module")
def ops():
    from omni_video_b200 import _lib, ops as _ops
    lib = _lib.load()
    assert lib.ovb_check_device() == 0, lib.ovb_last_error()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _ops


# ------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    (128, 256, 64), (128, 256, 256), (256, 512, 128), (200, 264, 192), (1000, 1536, 1536),
    (77, 8, 64), (4095, 4608, 1536), (512, 1536, 4096), (300, 8960, 1536), (300, 1536, 8960),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bias_bf16(ops, M, N, K):
    rng = _rng(M * 7 + N * 3 + K)
    a = _randn(rng, M, K).to(bf16).cuda()
    w = _randn(rng, N, K, scale=1 / math.sqrt(K)).to(bf16).cuda()
    b = _randn(rng, N, scale=0.1).to(bf16).cuda()
    out = ops.gemm(a, w, b, epilogue=ops.EPI_BF16)
    ref = (a.float() @ w.float().T + b.float())
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    # bf16 output rounding: half an ulp of the largest magnitude
    tol = ref.abs().max().item() * 2 ** -8 + 1e-3
    assert err <= tol, f"max err {err} > {tol}"
    assert O.rel_l2(out.float().cpu(), ref.cpu()) < 3e-3


def test_gemm_epilogues(ops):
    rng = _rng(11)
    M, N, K, rps = 384, 512, 320, 192
    a = _randn(rng, M, K).to(bf16).cuda()
    w = _randn(rng, N, K, scale=1 / math.sqrt(K)).to(bf16).cuda()
    b = _randn(rng, N, scale=0.1).to(bf16).cuda()
    resid = _randn(rng, M, N).cuda()
    gate_s = _randn(rng, M // rps, N).cuda()
    gate_t = _randn(rng, M, N).cuda()
    acc = a.float() @ w.float().T + b.float()
    accb = acc.to(bf16).float()
    # fp32
    o = ops.gemm(a, w, b, epilogue=ops.EPI_F32)
    assert (o - acc).abs().max().item() < 2e-3
    # gelu
    o = ops.gemm(a, w, b, epilogue=ops.EPI_GELU_BF16)
    ref = F.gelu(accb, approximate="tanh")
    assert (o.float() - ref).abs().max().item() <= ref.abs().max().item() * 2 ** -7 + 1e-3
    # residual
    o = ops.gemm(a, w, b, epilogue=ops.EPI_RESID_F32, resid=resid)
    assert (o - (resid + accb)).abs().max().item() <= accb.abs().max().item() * 2 ** -7 + 1e-3
    # gated residual, per-sample gate
    o = ops.gemm(a, w, b, epilogue=ops.EPI_GATE_RESID_F32, resid=resid, gate=gate_s, rows_per_sample=rps)
    ref = resid + accb * gate_s.repeat_interleave(rps, 0)
    assert O.rel_l2(o.cpu(), ref.cpu()) < 3e-3
    # gated residual, per-token gate, in place
    x = resid.clone()
    ops.gemm(a, w, b, epilogue=ops.EPI_GATE_RESID_F32, resid=x, out=x, gate=gate_t)
    ref = resid + accb * gate_t
    assert O.rel_l2(x.cpu(), ref.cpu()) < 3e-3


def test_gemm_split_layouts(ops):
    """n_split writes the Ulysses send layout, k_split reads the receive layout."""
    rng = _rng(12)
    M, N, K, p = 300, 768, 512, 4
    a = _randn(rng, M, K).to(bf16).cuda()
    w = _randn(rng, N, K, scale=1 / math.sqrt(K)).to(bf16).cuda()
    b = _randn(rng, N, scale=0.1).to(bf16).cuda()
    ref = ops.gemm(a, w, b)
    o = ops.gemm(a, w, b, n_split=p)
    assert o.shape == (p, M, N // p)
    assert torch.equal(o.permute(1, 0, 2).reshape(M, N), ref)
    a_split = a.view(M, p, K // p).permute(1, 0, 2).contiguous()
    o2 = ops.gemm(a_split, w, b, k_split=p)
    assert torch.equal(o2, ref)


def test_gemm_rejects_bad_arguments(ops):
    from omni_video_b200._lib import OvbError
    a = torch.zeros(16, 60, dtype=bf16, device="cuda")
    w = torch.zeros(8, 60, dtype=bf16, device="cuda")
    with pytest.raises((OvbError, AssertionError)):
        ops.gemm(a, w, None)  # K not a multiple of 8
    with pytest.raises(OvbError):
        ops.gemm(torch.zeros(4, 64, dtype=bf16), torch.zeros(8, 64, dtype=bf16), None)  # CPU tensors


# ------------------------------------------------------------------------------------- attention
def _attn_ref_gpu(q, k, v, k_lens):
    """fp32 torch restatement on the GPU (same math as oracle.dit.attention without rounding)."""
    B, Lq, H, D = q.shape
    out = torch.empty((B, Lq, H, D), device=q.device, dtype=f32)
    for b in range(B):
        kl = k.shape[1] if k_lens is None else k_lens[b]
        qb = q[b].float().permute(1, 0, 2)
        kb = k[b, :kl].float().permute(1, 0, 2)
        vb = v[b, :kl].float().permute(1, 0, 2)
        for h in range(H):
            s = (qb[h] @ kb[h].T) / math.sqrt(D)
            out[b, :, h] = torch.softmax(s, dim=-1) @ vb[h]
    return out


ATTN_CASES = [
    # B, H, Lq, Lk, k_lens
    (1, 1, 128, 128, None), (1, 2, 256, 256, None), (1, 2, 300, 300, None), (1, 2, 1280, 1280, None),
    (1, 3, 515, 77, None), (2, 2, 384, 384, [384, 200]), (1, 2, 1000, 1024, [1000]),
    (1, 1, 100, 512, [1]), (1, 12, 2048, 2048, [2040]),
]


@pytest.mark.parametrize("B,H,Lq,Lk,k_lens", ATTN_CASES)
def test_attention_vs_oracle(ops, B, H, Lq, Lk, k_lens):
    rng = _rng(B * 1000 + H * 100 + Lq + Lk)
    q = _randn(rng, B, Lq, H, 128).to(bf16)
    k = _randn(rng, B, Lk, H, 128).to(bf16)
    v = _randn(rng, B, Lk, H, 128).to(bf16)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), k_lens=k_lens)
    torch.cuda.synchronize()
    if Lq * Lk * H <= 1280 * 1280 * 2:
        ref = torch.stack([O.attention(q[b].float(), k[b].float(), v[b].float(),
                                       None if k_lens is None else k_lens[b], emulate=True)
                           for b in range(B)])
        refs = [("oracle-bf16", ref, 6e-3)]
    else:
        refs = []
    refs.append(("torch-fp32", _attn_ref_gpu(q.cuda(), k.cuda(), v.cuda(), k_lens).cpu(), 8e-3))
    for name, ref, tol in refs:
        err = O.rel_l2(out.float().cpu(), ref)
        assert err < tol, f"{name}: rel-L2 {err}"
        assert (out.float().cpu() - ref).abs().max().item() < 0.05, name


SPLIT_CASES = [
    # B, H, Lq, Lk, k_lens: launches whose last wave is cut along the keys (ovb_attn_fwd, AttnSplit)
    (1, 1, 2560, 8192, None),            # 10 items on 148 SMs -> 7-way key split of every item
    (1, 1, 2500, 8192, [8100]),          # ragged rows at the end of Lq + a masked last key tile
    (2, 1, 1280, 4096, [4096, 2100]),    # slices of different length per sample
    (1, 3, 13056, 4096, None),           # 153 items: one whole wave + 5 tail items
]


@pytest.mark.parametrize("B,H,Lq,Lk,k_lens", SPLIT_CASES)
def test_attention_key_split_last_wave(ops, B, H, Lq, Lk, k_lens):
    rng = _rng(B * 7 + H * 5 + Lq + Lk)
    q = _randn(rng, B, Lq, H, 128).to(bf16).cuda()
    k = _randn(rng, B, Lk, H, 128).to(bf16).cuda()
    v = _randn(rng, B, Lk, H, 128).to(bf16).cuda()
    from omni_video_b200 import _lib as L
    before = L.launch_count()
    out = ops.attention(q, k, v, k_lens=k_lens)
    assert L.launch_count() - before == 2, "expected the key-split form (attention + combine launch)"
    whole = ops.attention(q, k, v, k_lens=k_lens, split_keys=False)
    ref = _attn_ref_gpu(q, k, v, k_lens)
    assert O.rel_l2(out.float().cpu(), ref.cpu()) < 8e-3
    assert (out.float() - ref).abs().max().item() < 0.05
    # the two decompositions agree to bf16 rounding of the output
    assert (out.float() - whole.float()).abs().max().item() <= 2 ** -7 * max(1.0, whole.float().abs().max().item())


def test_attention_strided_qkv_buffer(ops):
    """q, k, v living inside one fused [L, 3*H*128] projection buffer (how DiTEngine calls it)."""
    rng = _rng(5)
    L, H = 700, 3
    qkv = _randn(rng, 1, L, 3, H, 128).to(bf16).cuda()
    out = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], k_lens=[650])
    ref = _attn_ref_gpu(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], [650])
    assert O.rel_l2(out.float().cpu(), ref.cpu()) < 8e-3


def test_attention_large_scores(ops):
    """exercise the lazy-rescale path: score scale large enough that the running max keeps growing"""
    rng = _rng(6)
    L, H = 1024, 2
    q = (_randn(rng, 1, L, H, 128) * 4).to(bf16).cuda()
    k = (_randn(rng, 1, L, H, 128) * 4).to(bf16).cuda()
    # keys sorted by norm so later tiles tend to have larger logits
    v = _randn(rng, 1, L, H, 128).to(bf16).cuda()
    out = ops.attention(q, k, v)
    ref = _attn_ref_gpu(q, k, v, None)
    assert torch.isfinite(out.float()).all()
    assert O.rel_l2(out.float().cpu(), ref.cpu()) < 1e-2


def _relbias_ref(q, k, v, rel_bias, k_lens, scale):
    """fp32 torch restatement of reference t5.py:101-112 with the bias given in (key - query) form; P is rounded to
    bf16 before P V like the kernel (and like the reference's bf16 execution, t5.py:111)."""
    B, L, H, D = q.shape
    delta = (torch.arange(L)[None, :] - torch.arange(L)[:, None] + L - 1).to(q.device)
    S = torch.einsum("bihc,bjhc->bhij", q.float(), k.float()) * scale + rel_bias[:, delta][None]
    if k_lens is not None:
        S = S.masked_fill(torch.arange(L, device=q.device)[None, None, None, :] >= k_lens.view(B, 1, 1, 1).long(), float("-inf"))
    P = torch.softmax(S, dim=-1).to(bf16).float()
    return torch.einsum("bhij,bjhc->bihc", P, v.float())


@pytest.mark.parametrize("B,L,H,D,k_lens,scale", [
    (1, 64, 2, 64, None, 1.0),              # one tile
    (2, 512, 64, 64, [77, 512], 1.0),       # the umT5-XXL launch: 64 heads of 64, ragged batch
    (3, 200, 3, 64, [1, 130, 200], 1.0),    # L not a multiple of the 64-row tiles, a single valid key
    (2, 130, 2, 128, [130, 65], 0.5),       # 128-wide heads, scaled scores
])
def test_attention_relbias_vs_fp32(ops, B, L, H, D, k_lens, scale):
    """ovb_attn_relbias_fwd (umT5 self-attention, all heads in one launch) on column blocks of a fused QKV buffer."""
    rng = _rng(11 + L)
    qkv = _randn(rng, B, L, 3, H, D).to(bf16).cuda()
    qkv[:, :, 0] *= 0.35                                           # unscaled T5 scores: keep the softmax from saturating
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    rel_bias = _randn(rng, H, 2 * L - 1, scale=0.7).cuda()
    kl = None if k_lens is None else torch.tensor(k_lens, dtype=torch.int32, device="cuda")
    out = ops.attention_relbias(q, k, v, rel_bias, kl, scale)
    ref = _relbias_ref(q, k, v, rel_bias, kl, scale)
    assert out.shape == (B, L, H, D) and out.dtype == bf16 and bool(torch.isfinite(out.float()).all())
    for b in range(B):                                             # per sample, so a wrong key count cannot hide
        err = O.rel_l2(out[b].float().cpu(), ref[b].cpu())
        assert err < 6e-3, (b, err)
    # rows are independent of the launch shape: one sample alone gives the same bits as inside the batch
    if B > 1:
        alone = ops.attention_relbias(q[1:2], k[1:2], v[1:2], rel_bias, None if kl is None else kl[1:2].contiguous(), scale)
        assert torch.equal(alone[0], out[1])


def test_attention_relbias_rejects_bad_arguments(ops):
    q = torch.zeros(1, 64, 2, 32, dtype=bf16, device="cuda")
    with pytest.raises(ValueError):
        ops.attention_relbias(q, q, q, torch.zeros(2, 127, device="cuda"))
    from omni_video_b200 import _lib
    import ctypes as C
    a = _lib.AttnRelBiasArgs()
    assert _lib.load().ovb_attn_relbias_fwd(C.byref(a), None) == -1          # null pointers
    assert b"null pointer" in _lib.load().ovb_last_error()


# ----------------------------------------------------------------------------------- elementwise
@pytest.mark.parametrize("rows,d,rps", [(64, 256, 32), (100, 1536, 50), (33, 5120, 33)])
def test_layernorm_modulate(ops, rows, d, rps):
    rng = _rng(rows + d)
    x = _randn(rng, rows, d) * 3 + 0.5
    ns = rows // rps
    tbl = _randn(rng, ns, 6, d, scale=0.3)
    ref = O.layer_norm(x, 1e-6) * (1 + tbl[:, 1].repeat_interleave(rps, 0)) + tbl[:, 0].repeat_interleave(rps, 0)
    tc = tbl.cuda()
    out = ops.layernorm_modulate(x.cuda(), tc[:, 1], tc[:, 0], 1e-6, rows_per_sample=rps)
    assert (out.float().cpu() - ref).abs().max().item() <= ref.abs().max().item() * 2 ** -8 + 1e-5
    out32 = ops.layernorm_modulate(x.cuda(), tc[:, 1], tc[:, 0], 1e-6, rows_per_sample=rps, out_dtype=f32)
    assert (out32.cpu() - ref).abs().max().item() < 1e-4
    # per-token table
    tt = _randn(rng, rows, 2, d, scale=0.3)
    ref = O.layer_norm(x, 1e-6) * (1 + tt[:, 1]) + tt[:, 0]
    ttc = tt.cuda()
    out32 = ops.layernorm_modulate(x.cuda(), ttc[:, 1], ttc[:, 0], 1e-6, out_dtype=f32)
    assert (out32.cpu() - ref).abs().max().item() < 1e-4


def test_layernorm_affine_and_bf16_input(ops):
    rng = _rng(3)
    rows, d = 70, 1536
    x = _randn(rng, rows, d) * 2
    w = (1 + _randn(rng, d, scale=0.1)).to(bf16)
    b = _randn(rng, d, scale=0.1).to(bf16)
    ref = O.layer_norm(x, 1e-6, w.float(), b.float())
    out = ops.layernorm_affine(x.cuda(), w.cuda(), b.cuda(), 1e-6, out_dtype=f32)
    assert (out.cpu() - ref).abs().max().item() < 1e-4
    xb = x.to(bf16)
    z = torch.zeros(1, d)
    ref = O.layer_norm(xb.float(), 1e-6).to(bf16).float()
    out = ops.layernorm_modulate(xb.cuda(), z.cuda(), z.cuda(), 1e-6, out_dtype=f32, round_ln_bf16=True)
    assert (out.cpu() - ref).abs().max().item() <= 2 ** -7


@pytest.mark.parametrize("grid,H,pad,offset", [((3, 4, 5), 2, 4, 0), ((5, 16, 16), 12, 0, 0),
                                               ((4, 6, 8), 2, 0, 96)])
def test_rmsnorm_rope(ops, grid, H, pad, offset):
    rng = _rng(sum(grid) + H)
    Fg, Hg, Wg = grid
    L = Fg * Hg * Wg + pad - offset
    d = H * 128
    x = _randn(rng, L, d).to(bf16)
    w = (1 + _randn(rng, d, scale=0.1)).to(bf16)
    ang = O.rope_angles(1024, 128)
    ref = O.rms_norm(x.float(), w.float(), 1e-6, emulate=True)
    ref = O.rope_apply(ref.view(L, H, 128), grid, ang, token_offset=offset).to(bf16).float().view(L, d)
    cs = ops.rope_cos_sin_table(1024, 128).cuda()
    y = x.clone().cuda()
    ops.rmsnorm_rope_(y, w.cuda(), 1e-6, cs_table=cs, grid=grid, rows_per_sample=L, token_offset=offset)
    diff = (y.float().cpu() - ref).abs()
    # identical up to 1 bf16 ulp at rounding boundaries (fp32 vs fp64 rotation)
    assert diff.max().item() <= ref.abs().max().item() * 2 ** -7
    assert (diff > 0).float().mean().item() < 0.02
    # RMSNorm only, split-column layout
    p = 2
    xs = x.view(L, p, d // p).permute(1, 0, 2).contiguous().cuda()
    ops.rmsnorm_rope_(xs, w.cuda(), 1e-6)
    ref2 = O.rms_norm(x.float(), w.float(), 1e-6, emulate=True)
    got2 = xs.permute(1, 0, 2).reshape(L, d).float().cpu()
    # the row sum-of-squares is reduced in a different order than torch's CPU mean: a value sitting
    # exactly on a bf16 rounding boundary may land one ulp away, nothing more
    d2 = (got2 - ref2).abs()
    assert d2.max().item() <= ref2.abs().max().item() * 2 ** -7
    assert (d2 > 0).float().mean().item() < 1e-3


def test_scatter_forms_on_one_gpu_match_the_plain_forms(ops):
    """The sequence-parallel scatter forms (include/ovb200.h: `dst_ptrs` of ovb_rmsnorm_rope, `out_ptrs` of
    ovb_attn_fwd) write through raw destination pointers — peer-mapped memory under Ulysses.  With local buffers
    as destinations they must give the bits of the plain forms: this keeps the scatter instantiations covered on
    a one-GPU box (the multi-GPU checks are scripts/sp_check.py and the `parity` object of bench.py)."""
    rng = _rng(21)
    rows, groups, dc = 700, 2, 256                                   # d = 512: 4 heads, 2 per column group
    x = _randn(rng, groups, rows, dc).to(bf16).cuda()
    w = (1 + 0.1 * _randn(rng, groups * dc)).to(bf16).cuda()
    cs = ops.rope_cos_sin_table().cuda()
    kw = dict(cs_table=cs, grid=(7, 10, 10), rows_per_sample=rows)
    want = ops.rmsnorm_rope_(x.clone(), w, 1e-6, **kw)
    ld = dc + 8
    dst = [torch.full((rows + 3, ld), 7.0, dtype=bf16, device="cuda") for _ in range(groups)]
    src = x.clone()
    ops.rmsnorm_rope_(src, w, 1e-6, dst_ptrs=[t.data_ptr() for t in dst], dst_ld=ld, **kw)
    assert torch.equal(src, x)                                       # the scatter form leaves its input alone
    for g in range(groups):
        assert torch.equal(dst[g][:rows, :dc], want[g])
        assert bool((dst[g][rows:] == 7.0).all()) and bool((dst[g][:, dc:] == 7.0).all())      # nothing else touched
    ops.rmsnorm_rope_(src, w, 1e-6, dst_ptrs=[t.data_ptr() for t in dst], dst_ld=ld, copy_only=True)
    for g in range(groups):
        assert torch.equal(dst[g][:rows, :dc], x[g])

    # `out=`: the same form with one local destination (how the umT5 encoder keeps its residual stream untouched)
    x2 = _randn(rng, 300, 4096).to(bf16).cuda()
    w2 = (1 + 0.1 * _randn(rng, 4096)).to(bf16).cuda()
    keep, dst2 = x2.clone(), torch.empty_like(x2)
    assert ops.rmsnorm_rope_(x2, w2, 1e-6, out=dst2) is dst2
    assert torch.equal(x2, keep) and torch.equal(dst2, ops.rmsnorm_rope_(keep.clone(), w2, 1e-6))

    L, H, per = 600, 2, 300
    q, k, v = (_randn(rng, 1, L, H, 128).to(bf16).cuda() for _ in range(3))
    whole = ops.attention(q, k, v, k_lens=[555])
    parts = [torch.zeros(per, H * 128 + 64, dtype=bf16, device="cuda") for _ in range(L // per)]
    assert ops.attention(q, k, v, k_lens=[555], out_ptrs=[t.data_ptr() for t in parts], rows_per_out=per,
                         out_row_stride=parts[0].stride(0)) is None
    for i, t in enumerate(parts):
        assert torch.equal(t[:, :H * 128].view(per, H, 128), whole[0, i * per:(i + 1) * per])


@pytest.mark.parametrize("M,N,K,act_in,act_out", [(5000, 64, 1536, 0, 0), (4100, 40, 256, 1, 1), (65520, 64, 1536, 0, 0)])
def test_linear_f32_tall_form_matches_the_general_kernel_bit_for_bit(ops, monkeypatch, M, N, K, act_in, act_out):
    """The DiT head's shape (many rows, <= 64 outputs) takes the register-blocked kernel; each output is the same
    fp32 FMA chain over ascending k as in the general kernel, so the bits agree — and both agree with fp64."""
    rng = _rng(41 + N)
    x = _randn(rng, M, K).cuda()
    w = _randn(rng, N, K, scale=K ** -0.5).to(bf16).cuda()
    b = _randn(rng, N).to(bf16).cuda()
    tall = ops.linear_f32(x, w, b, act_in=act_in, act_out=act_out)
    monkeypatch.setenv("OVB_LINEAR_F32_GENERAL", "1")
    general = ops.linear_f32(x, w, b, act_in=act_in, act_out=act_out)
    monkeypatch.delenv("OVB_LINEAR_F32_GENERAL")
    assert tall.shape == (M, N) and torch.equal(tall, general)
    if not act_in and not act_out:
        want = x[:512].double() @ w.double().T + b.double()
        assert (tall[:512].double() - want).abs().max().item() <= 3e-6 * want.abs().max().item()


@pytest.mark.parametrize("M,N,K,act_in,act_out", [(2, 9216, 1536, 1, 0), (1, 1536, 256, 0, 1), (8, 100, 264, 0, 0)])
def test_linear_f32_few_rows_form(ops, monkeypatch, M, N, K, act_in, act_out):
    """The time-embedding MLPs (one row per sample, reference model.py:390-392) take the warp-per-output-column
    kernel; its lane-parallel fp32 sums agree with the general kernel's chains and with fp64 to fp32 accuracy."""
    rng = _rng(51 + M)
    x = _randn(rng, M, K).cuda()
    w = _randn(rng, N, K, scale=K ** -0.5).to(bf16).cuda()
    b = _randn(rng, N).to(bf16).cuda()
    few = ops.linear_f32(x, w, b, act_in=act_in, act_out=act_out)
    monkeypatch.setenv("OVB_LINEAR_F32_GENERAL", "1")
    general = ops.linear_f32(x, w, b, act_in=act_in, act_out=act_out)
    monkeypatch.delenv("OVB_LINEAR_F32_GENERAL")
    xin = F.silu(x.double()) if act_in else x.double()
    want = xin @ w.double().T + b.double()
    want = F.silu(want) if act_out else want
    scale = want.abs().max().item()
    assert few.shape == (M, N)
    assert (few.double() - want).abs().max().item() <= 3e-6 * scale
    assert (few - general).abs().max().item() <= 3e-6 * scale


def test_patchify_unpatchify(ops):
    rng = _rng(9)
    u = _randn(rng, 16, 3, 8, 12)
    tok = ops.patchify(u.cuda())
    assert torch.equal(tok.float().cpu(), O.patchify(u).to(bf16).float())
    y = _randn(rng, 3 * 4 * 6 + 5, 64)
    out = ops.unpatchify(y.cuda(), 16, (3, 4, 6))
    assert torch.equal(out.cpu(), O.unpatchify(y, (3, 4, 6), 16))


def test_time_embedding_path(ops):
    rng = _rng(10)
    d = 256
    t = torch.tensor([999.0, 17.0])
    sin = ops.sinusoidal_embedding(t.cuda(), 256)
    ref = O.sinusoidal_embedding_1d(256, t).float()
    assert (sin.cpu() - ref).abs().max().item() < 1e-6
    w = _randn(rng, 6 * d, d, scale=0.05).to(bf16)
    b = _randn(rng, 6 * d, scale=0.05).to(bf16)
    x = _randn(rng, 2, d)
    out = ops.linear_f32(x.cuda(), w.cuda(), b.cuda(), act_in=1)
    ref = F.linear(F.silu(x), w.float(), b.float())
    assert (out.cpu() - ref).abs().max().item() < 1e-4
    out = ops.linear_f32(x.cuda(), w.cuda(), b.cuda(), act_out=1)
    ref = F.silu(F.linear(x, w.float(), b.float()))
    assert (out.cpu() - ref).abs().max().item() < 1e-4
    mod = _randn(rng, 3, 6, d, scale=0.1).to(bf16)
    e0 = _randn(rng, 2, 6 * d)
    tbl = ops.modulation_table(mod.cuda(), e0.cuda(), True)
    ref = mod.float()[:, None] + e0.view(1, 2, 6, d)
    assert torch.equal(tbl.cpu(), ref)
