"""Test infrastructure: torch (CPU) stand-ins for the `omni_video_b200.ops` entry points, with the kernels'
rounding points, so that the HOST orchestration of the engines (buffer views and strides, epilogue choices,
sequence-parallel layouts, token offsets, gathers) can be exercised without a GPU — including on two gloo ranks.
Each stand-in restates the contract documented in include/ovb200.h / ops.py; where the oracle has the arithmetic
(attention, RoPE, norms, patchify) it is reused.  The kernels themselves are checked on the GPU
(tests/test_gpu_*.py); nothing here is reachable from the product."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle import dit as O

bf16, f32, f64 = torch.bfloat16, torch.float32, torch.float64


def _rb(x):
    return x.to(bf16).float()


def gemm(a, w, bias, *, epilogue=0, out=None, resid=None, gate=None, rows_per_sample=0, n_split=1, k_split=1):
    from omni_video_b200 import ops
    if k_split > 1:
        assert a.dim() == 3 and a.shape[0] == k_split
        a = torch.cat([a[g] for g in range(k_split)], dim=1)
    assert a.dtype == bf16 and w.dtype == bf16 and a.shape[1] == w.shape[1]
    acc = a.float() @ w.float().T + (0 if bias is None else bias.float())
    accb = _rb(acc)
    M, N = acc.shape
    if epilogue == ops.EPI_BF16:
        res = accb.to(bf16)
    elif epilogue == ops.EPI_GELU_BF16:
        res = F.gelu(acc, approximate="tanh").to(bf16)
    elif epilogue == ops.EPI_F32:
        res = acc
    elif epilogue == ops.EPI_RESID_F32:
        res = resid + accb
    elif epilogue == ops.EPI_GATE_RESID_F32:
        rps = rows_per_sample or M
        g = gate if gate.shape[0] == M and M != (M + rps - 1) // rps else gate.repeat_interleave(rps, 0)[:M]
        res = resid + accb * g
    elif epilogue == ops.EPI_RESID_BF16:
        res = (resid.float() + accb).to(bf16)
    else:
        raise AssertionError(epilogue)
    if n_split > 1:
        ns = N // n_split
        if out is None:
            out = torch.empty((n_split, M, ns), dtype=res.dtype)
        for g in range(n_split):
            out[g].copy_(res[:, g * ns:(g + 1) * ns])
        return out
    if out is None:
        return res
    assert out.shape == res.shape and out.dtype == res.dtype, (out.shape, res.shape, out.dtype, res.dtype)
    out.copy_(res)
    return out


def _ln(x, eps):
    return F.layer_norm(x.float(), (x.shape[-1],), None, None, eps)


def layernorm_modulate(x, scale, shift, eps, *, rows_per_sample=0, out_dtype=bf16, round_ln_bf16=False, out=None):
    rows = x.shape[0]
    y = _ln(x, eps)
    if round_ln_bf16:
        y = _rb(y)
    rps = rows_per_sample or rows
    if scale.shape[0] != rows or rows == (rows + rps - 1) // rps:
        scale, shift = scale.repeat_interleave(rps, 0)[:rows], shift.repeat_interleave(rps, 0)[:rows]
    y = y * (1.0 + scale.float()) + shift.float()
    if out is None:
        return y.to(out_dtype)
    out.copy_(y.to(out.dtype))
    return out


def layernorm_affine(x, weight, bias, eps, *, out_dtype=bf16, out=None):
    y = _ln(x, eps) * weight.float() + bias.float()
    if out is None:
        return y.to(out_dtype)
    out.copy_(y.to(out.dtype))
    return out


def rmsnorm_rope_(x, weight, eps, *, cs_table=None, grid=None, rows_per_sample=0, token_offset=0, dst_ptrs=None,
                  dst_ld=0, copy_only=False):
    assert dst_ptrs is None, "the peer-memory scatter form needs a GPU"
    full = torch.cat([x[g] for g in range(x.shape[0])], dim=1) if x.dim() == 3 else x      # [rows, d]
    rows, d = full.shape
    xf = full.float()
    y = _rb(_rb(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)) * weight.float())
    if cs_table is not None:
        rps = rows_per_sample or rows
        ang = torch.atan2(cs_table[..., 1].double(), cs_table[..., 0].double())            # [T, 64] angles back from (cos, sin)
        pieces = []
        for r0 in range(0, rows, rps):
            blk = y[r0:r0 + rps].view(-1, d // 128, 128)
            pieces.append(O.rope_apply(blk, grid, ang, token_offset=token_offset).reshape(-1, d))
        y = _rb(torch.cat(pieces, dim=0))
    yb = y.to(bf16)
    if x.dim() == 3:
        dc = x.shape[2]
        for g in range(x.shape[0]):
            x[g].copy_(yb[:, g * dc:(g + 1) * dc])
    else:
        x.copy_(yb)
    return x


def attention(q, k, v, k_lens=None, softmax_scale=None, out=None, split_keys=True, out_ptrs=None, rows_per_out=0,
              out_row_stride=0):
    assert out_ptrs is None, "the peer-memory scatter form needs a GPU"
    B = q.shape[0]
    res = torch.stack([O.attention(q[b].float(), k[b].float(), v[b].float(), None if k_lens is None else int(k_lens[b]),
                                   emulate=True, softmax_scale=softmax_scale) for b in range(B)]).to(bf16)
    if out is None:
        return res
    out.copy_(res)
    return out


def patchify(x, out=None):
    res = O.patchify(x.float()).to(bf16)
    if out is None:
        return res
    out[:res.shape[0]].copy_(res)
    return out


def unpatchify(y, C_out, grid):
    return O.unpatchify(y.float(), tuple(int(g) for g in grid), C_out).contiguous()


def linear_f32(x, w, bias, *, act_in=0, act_out=0):
    h = F.silu(x) if act_in else x
    y = h @ w.float().T + (0 if bias is None else bias.float())
    return F.silu(y) if act_out else y


def sinusoidal_embedding(t, dim):
    return O.sinusoidal_embedding_1d(dim, t.to(f32).reshape(-1)).to(f32)


def modulation_table(mod, e, e_has_k):
    n_layers, K, d = mod.shape
    B = e.shape[0]
    ev = e.view(B, K, d) if e_has_k else e.view(B, 1, d).expand(B, K, d)
    return (mod.float()[:, None] + ev[None]).contiguous()


def install(monkeypatch):
    """Replace the CUDA entry points of `omni_video_b200.ops` by the stand-ins above (and lift the engine's
    "parameters must be on a CUDA device" check, which is there precisely because the product has no CPU path)."""
    from omni_video_b200 import ops
    from omni_video_b200.modules.model import DiTEngine
    monkeypatch.setattr(DiTEngine, "_check_device", staticmethod(lambda dev: None))
    for name, fn in dict(gemm=gemm, layernorm_modulate=layernorm_modulate, layernorm_affine=layernorm_affine,
                         rmsnorm_rope_=rmsnorm_rope_, attention=attention, patchify=patchify, unpatchify=unpatchify,
                         linear_f32=linear_f32, sinusoidal_embedding=sinusoidal_embedding,
                         modulation_table=modulation_table, cast_bf16_to_f32=lambda x: x.float(),
                         cast_f32_to_bf16=lambda x: x.to(bf16)).items():
        monkeypatch.setattr(ops, name, fn)
