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


# ------------------------------------------------------------------------------------------------- VAE
def conv3d(x, pc, out_extent, *, out=None, out_strides=None, out_base=0, n_group=0, group_stride=0, resid=None,
           out_mode=0, plane_stride=0, dh_shift=0, norm_gamma=None, norm_out=None, norm_silu=True):
    """ovb_conv3d (include/ovb200.h): taps with zero fill outside x, bias, optional bf16 residual (the conv result
    is rounded to bf16 before the add), output scattered through (os_t, os_h, os_w, out_base, n_group)."""
    from omni_video_b200 import vae_ops
    Ti, Hi, Wi, Cin = x.shape
    To, Ho, Wo = out_extent
    st, sh, sw = pc.stride
    xf = x.float()
    t = torch.arange(To).view(To, 1, 1)
    h = torch.arange(Ho).view(1, Ho, 1)
    w = torch.arange(Wo).view(1, 1, Wo)
    acc = torch.zeros((To, Ho, Wo, pc.cout), dtype=f32)
    for k, (dt, dh, dw) in enumerate(pc.taps):
        ti, hi, wi = t * st + dt, h * sh + dh + dh_shift, w * sw + dw
        ok = ((ti >= 0) & (ti < Ti)) & ((hi >= 0) & (hi < Hi)) & ((wi >= 0) & (wi < Wi))
        ti, hi, wi = (ti.clamp(0, Ti - 1).expand(To, Ho, Wo), hi.clamp(0, Hi - 1).expand(To, Ho, Wo),
                      wi.clamp(0, Wi - 1).expand(To, Ho, Wo))
        xs = xf[ti, hi, wi] * ok.unsqueeze(-1)
        acc += xs @ pc.w[:pc.cout, k * Cin:(k + 1) * Cin].float().T
    if pc.bias is not None:
        acc += pc.bias.float()
    n = torch.arange(pc.cout)
    if out_strides is None:
        out_strides = (Ho * Wo * pc.cout, Wo * pc.cout, pc.cout)
    os_t, os_h, os_w = out_strides
    base = out_base + t * os_t + h * os_h + w * os_w                                   # [To, Ho, Wo]
    if out_mode == vae_ops.CONV_OUT_BF16:
        if out is None:
            out = torch.empty((To, Ho, Wo, pc.cout), dtype=bf16)
        nmap = (n // n_group) * group_stride + n % n_group if n_group else n
        idx = (base.unsqueeze(-1) + nmap.view(1, 1, 1, -1)).reshape(-1)
        y = acc
        if resid is not None:
            y = _rb(y) + resid.reshape(-1)[idx].float().view_as(y)
        if norm_gamma is not None:
            # fused RMS_norm (+SiLU) of the bf16 value the unfused formulation would have stored
            yn = rmsnorm_silu(y.to(bf16), norm_gamma, silu=norm_silu)
            if norm_out is None:
                out.view(-1)[idx] = yn.reshape(-1)
                return out
            norm_out.view(-1)[idx] = yn.reshape(-1)
        out.view(-1)[idx] = y.to(bf16).reshape(-1)
        return out
    idx = (base.unsqueeze(-1) + (n * plane_stride).view(1, 1, 1, -1)).reshape(-1)
    y = acc.clamp(-1.0, 1.0) if out_mode == vae_ops.CONV_OUT_F32_PLANAR_CLAMP else acc
    out.view(-1)[idx] = y.reshape(-1)
    return out


def tap_sum(y, taps, bias, channels, out_extent, *, dh_shift=0, clamp=False, out=None):
    """ovb_tap_sum (include/ovb200.h): out[c,t,h,w] = bias[c] + sum_k y[k, t+dt, h+dh+dh_shift, w+dw, c], zero outside."""
    n, Ty, Hy, Wy, _ = y.shape
    To, Ho, Wo = out_extent
    t = torch.arange(To).view(To, 1, 1)
    h = torch.arange(Ho).view(1, Ho, 1)
    w = torch.arange(Wo).view(1, 1, Wo)
    acc = torch.zeros((To, Ho, Wo, 4), dtype=f32)
    for k, (dt, dh, dw) in enumerate(taps):
        ti, hi, wi = t + dt, h + dh + dh_shift, w + dw
        ok = ((ti >= 0) & (ti < Ty)) & ((hi >= 0) & (hi < Hy)) & ((wi >= 0) & (wi < Wy))
        ti, hi, wi = (ti.clamp(0, Ty - 1).expand(To, Ho, Wo), hi.clamp(0, Hy - 1).expand(To, Ho, Wo),
                      wi.clamp(0, Wy - 1).expand(To, Ho, Wo))
        acc += y[k].float()[ti, hi, wi] * ok.unsqueeze(-1)
    r = acc[..., :channels]
    if bias is not None:
        r = r + bias.float()
    if clamp:
        r = r.clamp(-1.0, 1.0)
    r = r.permute(3, 0, 1, 2).contiguous()
    if out is None:
        return r
    out.copy_(r)
    return out


def rmsnorm_silu(x, gamma, silu=True, out=None):
    xf = x.float()
    y = xf / xf.norm(dim=-1, keepdim=True).clamp_min(1e-12) * (x.shape[-1] ** 0.5) * gamma.float()
    if silu:
        y = F.silu(y)
    if out is None:
        return y.to(bf16)
    out.copy_(y.to(bf16))
    return out


def planar_to_cl(src, scale, shift, cpad):
    C = src.shape[0]
    y = (src * scale.view(-1, 1, 1, 1) + shift.view(-1, 1, 1, 1)).permute(1, 2, 3, 0)
    out = torch.zeros(y.shape[:3] + (cpad,), dtype=bf16)
    out[..., :C] = y.to(bf16)
    return out


def cl_to_planar(src, scale, shift, c):
    return (src[..., :c].float() * scale.view(1, 1, 1, -1) + shift.view(1, 1, 1, -1)).permute(3, 0, 1, 2).contiguous()


def softmax_rows(S, scale, out=None):
    p = torch.softmax(S * scale, dim=-1).to(bf16)
    if out is None:
        return p
    out.copy_(p)
    return out


def install_vae(monkeypatch):
    """VAE counterpart of install(): `omni_video_b200.vae_ops` entry points + the engine's device check."""
    from omni_video_b200 import vae_ops
    from omni_video_b200.modules.vae2_1 import VAEEngine
    install(monkeypatch)
    monkeypatch.setattr(VAEEngine, "_check_device", staticmethod(lambda dev: None))
    for name, fn in dict(conv3d=conv3d, tap_sum=tap_sum, rmsnorm_silu=rmsnorm_silu, planar_to_cl=planar_to_cl, cl_to_planar=cl_to_planar,
                         softmax_rows=softmax_rows).items():
        monkeypatch.setattr(vae_ops, name, fn)
