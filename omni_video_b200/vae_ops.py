"""Tensor-level wrappers over the VAE part of the C ABI (include/ovb200.h, "Wan2.1 causal-3D VAE").

Activations are channels-last bf16 CUDA tensors `[T, H, W, C]`.  As in ops.py, PyTorch only owns
memory and the stream; there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from . import ops as ops_mod
from ._lib import CONV_OUT_BF16, CONV_OUT_F32_PLANAR, CONV_OUT_F32_PLANAR_CLAMP  # noqa: F401
from .ops import _need_cuda, _ptr, _stream

bf16 = torch.bfloat16
f32 = torch.float32


class PackedConv:
    """Weights of one convolution in the layout ovb_conv3d consumes: bf16 [rows, n_taps*Cin] with
    K index = tap*Cin + c, the taps' input offsets, and an fp32 bias."""

    def __init__(self, w2: torch.Tensor, bias: Optional[torch.Tensor], taps: Sequence[Tuple[int, int, int]],
                 cin: int, cout: int, stride=(1, 1, 1)):
        assert w2.dtype == bf16 and w2.dim() == 2 and w2.is_contiguous()
        assert w2.shape[1] == len(taps) * cin and w2.shape[0] % 16 == 0 and w2.shape[0] >= cout
        self.w, self.bias = w2, None if bias is None else bias.to(f32).contiguous()
        self.cin, self.cout, self.stride = cin, cout, tuple(stride)
        n = len(taps)
        self.n_taps = n
        self.taps = [tuple(t) for t in taps]
        self._dt = (C.c_int8 * n)(*[t[0] for t in taps])
        self._dh = (C.c_int8 * n)(*[t[1] for t in taps])
        self._dw = (C.c_int8 * n)(*[t[2] for t in taps])
        self._dh_shifted = {0: self._dh}

    @property
    def spatial_rows(self):
        """(lowest, highest) input-row offset any tap reads: which halo rows a strip needs."""
        return min(t[1] for t in self.taps), max(t[1] for t in self.taps)

    def dh_array(self, shift: int):
        if shift not in self._dh_shifted:
            self._dh_shifted[shift] = (C.c_int8 * self.n_taps)(*[t[1] + shift for t in self.taps])
        return self._dh_shifted[shift]


def _pad_rows(w2: torch.Tensor) -> torch.Tensor:
    rows = (w2.shape[0] + 15) // 16 * 16
    if rows == w2.shape[0]:
        return w2.contiguous()
    out = torch.zeros((rows, w2.shape[1]), device=w2.device, dtype=w2.dtype)
    out[:w2.shape[0]] = w2
    return out


def pack_conv3d(weight: torch.Tensor, bias: Optional[torch.Tensor], *, causal_pad: Optional[int] = None,
                stride=(1, 1, 1), spatial_pad: Optional[Tuple[int, int]] = None) -> PackedConv:
    """weight [Co, Ci, kt, kh, kw] (or [Co, Ci, kh, kw] for nn.Conv2d) -> PackedConv.
    Tap (it, ih, iw) reads input voxel (t*st + it - causal_pad, h*sh + ih - ph, w*sw + iw - pw):
    causal_pad defaults to kt-1 (CausalConv3d, reference vae2_1.py:23-24), ph/pw to (k-1)/2."""
    if weight.dim() == 4:
        weight = weight.unsqueeze(2)
    co, ci, kt, kh, kw = weight.shape
    cp = kt - 1 if causal_pad is None else causal_pad
    ph, pw = (kh // 2, kw // 2) if spatial_pad is None else spatial_pad
    cin = (ci + 15) // 16 * 16
    w = weight.detach().to(f32).permute(0, 2, 3, 4, 1)                  # [Co, kt, kh, kw, Ci]
    if cin != ci:
        w = torch.nn.functional.pad(w, (0, cin - ci))
    taps = [(it - cp, ih - ph, iw - pw) for it in range(kt) for ih in range(kh) for iw in range(kw)]
    w2 = _pad_rows(w.reshape(co, kt * kh * kw * cin).to(bf16))
    return PackedConv(w2, bias, taps, cin, co, stride)


def pack_tap_planes(weight: torch.Tensor, bias: Optional[torch.Tensor], *, causal_pad: Optional[int] = None):
    """A conv with at most 4 output channels (the decoder head, 96 -> 3) as per-tap partial products:
    returns (PackedConv of ONE 1x1x1 tap with rows k*4 + c = weight[c, :, tap k] — 4 columns per tap, zero rows up
    to a multiple of 32 —, the taps' input offsets, the fp32 bias).  `conv3d(x, pc, n_group=4, ...)` then writes the
    planes y[k][t][h][w][c] and `tap_sum` shifts and adds them (ovb200.h: ovb_tap_sum)."""
    if weight.dim() == 4:
        weight = weight.unsqueeze(2)
    co, ci, kt, kh, kw = weight.shape
    assert co <= 4 and ci % 16 == 0
    cp = kt - 1 if causal_pad is None else causal_pad
    taps = [(it - cp, ih - kh // 2, iw - kw // 2) for it in range(kt) for ih in range(kh) for iw in range(kw)]
    n = len(taps)
    rows = (4 * n + 31) // 32 * 32
    w2 = torch.zeros((rows, ci), device=weight.device, dtype=bf16)
    w = weight.detach().to(f32).reshape(co, ci, n).permute(2, 0, 1)           # [tap, Co, Ci]
    w2[:4 * n].view(n, 4, ci)[:, :co] = w.to(bf16)
    pc = PackedConv(w2, None, [(0, 0, 0)], ci, 4 * n)
    return pc, taps, None if bias is None else bias.detach().to(f32).contiguous()


def tap_sum(y: torch.Tensor, taps: Sequence[Tuple[int, int, int]], bias: Optional[torch.Tensor], channels: int,
            out_extent: Tuple[int, int, int], *, dh_shift: int = 0, clamp: bool = False,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y: bf16 [n_taps, Ty, Hy, Wy, 4] (per-tap planes).  out[c, t, h, w] = bias[c] + sum_k y[k, t+dt, h+dh+dh_shift,
    w+dw, c] with zero outside y's extents; fp32 [channels, To, Ho, Wo]."""
    _need_cuda(y, bias, out)
    lib = _lib.load()
    n = len(taps)
    assert y.dtype == bf16 and y.dim() == 5 and y.is_contiguous() and y.shape[0] == n and y.shape[4] == 4
    To, Ho, Wo = out_extent
    if out is None:
        out = torch.empty((channels, To, Ho, Wo), device=y.device, dtype=f32)
    assert out.dtype == f32 and out.is_contiguous() and out.numel() == channels * To * Ho * Wo
    a = _lib.TapSumArgs()
    a.y, a.n_taps = y.data_ptr(), n
    dt = (C.c_int8 * n)(*[t[0] for t in taps])
    dh = (C.c_int8 * n)(*[t[1] + dh_shift for t in taps])
    dw = (C.c_int8 * n)(*[t[2] for t in taps])
    a.tap_dt, a.tap_dh, a.tap_dw = (C.cast(v, C.POINTER(C.c_int8)) for v in (dt, dh, dw))
    a.Ty, a.Hy, a.Wy = y.shape[1], y.shape[2], y.shape[3]
    if bias is not None:
        assert bias.dtype == f32 and bias.numel() == channels
        a.bias = bias.data_ptr()
    a.out, a.C = out.data_ptr(), channels
    a.To, a.Ho, a.Wo = To, Ho, Wo
    a.clamp = int(clamp)
    ev = ops_mod._tl_begin()
    _lib.check(lib.ovb_tap_sum(C.byref(a), _stream()), "ovb_tap_sum")
    ops_mod._tl_end(ev, "tap_sum", float(To * Ho * Wo) * (8 * n + 4 * channels), (To, Ho, Wo, n, channels))
    return out


def pack_upsample_conv2d(weight: torch.Tensor, bias: Optional[torch.Tensor]):
    """nearest-2x upsample followed by Conv2d(3x3, padding=1) (reference vae2_1.py:75-84) as FOUR
    2x2 convolutions on the low-resolution input, one per output phase (a, b) in {0,1}^2:
        out[2i+a, 2j+b] = sum_{u,v in {0,1}} Wab[u, v] * in[i + u - (1-a), j + v - (1-b)]
    where the 3 kernel rows collapse onto 2 input rows: a=0 -> (W0, W1+W2), a=1 -> (W0+W1, W2),
    and likewise for columns.  2.25x fewer FLOPs than convolving the upsampled tensor."""
    co, ci, kh, kw = weight.shape
    assert (kh, kw) == (3, 3)
    w = weight.detach().to(f32)
    rows = {0: [w[:, :, 0], w[:, :, 1] + w[:, :, 2]], 1: [w[:, :, 0] + w[:, :, 1], w[:, :, 2]]}  # [Co,Ci,3(kw)]
    phases = {}
    for a in (0, 1):
        for b in (0, 1):
            taps, mats = [], []
            for u in (0, 1):
                r = rows[a][u]                                             # [Co, Ci, 3]
                cols = [r[:, :, 0], r[:, :, 1] + r[:, :, 2]] if b == 0 else [r[:, :, 0] + r[:, :, 1], r[:, :, 2]]
                for v in (0, 1):
                    taps.append((0, u - (1 - a), v - (1 - b)))
                    mats.append(cols[v])                                   # [Co, Ci]
            w2 = _pad_rows(torch.cat(mats, dim=1).to(bf16))               # [Co, 4*Ci]
            phases[(a, b)] = PackedConv(w2, bias, taps, ci, co)
    return phases


def conv3d(x: torch.Tensor, pc: PackedConv, out_extent: Tuple[int, int, int], *,
           out: Optional[torch.Tensor] = None, out_strides: Optional[Tuple[int, int, int]] = None,
           out_base: int = 0, n_group: int = 0, group_stride: int = 0,
           resid: Optional[torch.Tensor] = None, out_mode: int = CONV_OUT_BF16,
           plane_stride: int = 0, dh_shift: int = 0, norm_gamma: Optional[torch.Tensor] = None,
           norm_out: Optional[torch.Tensor] = None, norm_silu: bool = True) -> torch.Tensor:
    """x: bf16 [Ti, Hi, Wi, Cin] contiguous.  Default output: a new bf16 [To, Ho, Wo, Cout].
    dh_shift is added to every tap's row offset (a strip whose input carries `dh_shift` halo rows
    on top).
    norm_gamma (fp32 [Cout]): fuse the RMS_norm (+SiLU) that consumes this conv's output into the
    epilogue (`can_fuse_norm(pc)` must hold).  Without norm_out the returned tensor IS the normalised
    one; with norm_out (bf16, shaped like the output) the raw output is returned and norm_out filled."""
    _need_cuda(x, pc.w, out, resid)
    lib = _lib.load()
    assert x.dtype == bf16 and x.dim() == 4 and x.is_contiguous(), "x must be contiguous bf16 [T,H,W,C]"
    Ti, Hi, Wi, Cin = x.shape
    assert Cin == pc.cin, f"Cin mismatch: tensor {Cin}, weights {pc.cin}"
    To, Ho, Wo = out_extent
    if out is None:
        assert out_mode == CONV_OUT_BF16 and out_strides is None
        out = torch.empty((To, Ho, Wo, pc.cout), device=x.device, dtype=bf16)
    if out_strides is None:
        out_strides = (Ho * Wo * pc.cout, Wo * pc.cout, pc.cout)
    a = _lib.ConvArgs()
    a.x, a.Ti, a.Hi, a.Wi, a.Cin = x.data_ptr(), Ti, Hi, Wi, Cin
    a.w, a.w_rows = pc.w.data_ptr(), pc.w.shape[0]
    a.bias = None if pc.bias is None else pc.bias.data_ptr()
    a.Cout, a.n_taps = pc.cout, pc.n_taps
    a.tap_dt = C.cast(pc._dt, C.POINTER(C.c_int8))
    a.tap_dh = C.cast(pc.dh_array(dh_shift), C.POINTER(C.c_int8))
    a.tap_dw = C.cast(pc._dw, C.POINTER(C.c_int8))
    a.st, a.sh, a.sw = pc.stride
    a.To, a.Ho, a.Wo = To, Ho, Wo
    a.out = out.data_ptr()
    a.os_t, a.os_h, a.os_w = out_strides
    a.out_base = out_base
    a.n_group, a.group_stride = n_group, group_stride
    if resid is not None:
        assert resid.dtype == bf16 and resid.is_contiguous() and resid.numel() == out.numel()
        a.resid = resid.data_ptr()
    a.out_mode, a.plane_stride = out_mode, plane_stride
    if norm_gamma is not None:
        assert can_fuse_norm(pc), f"RMS_norm cannot be fused into a conv with {pc.cout} output channels"
        assert norm_gamma.dtype == f32 and norm_gamma.is_contiguous() and norm_gamma.numel() == pc.cout
        _need_cuda(norm_gamma, norm_out)
        a.norm_gamma, a.norm_silu = norm_gamma.data_ptr(), int(norm_silu)
        if norm_out is not None:
            assert norm_out.dtype == bf16 and norm_out.is_contiguous() and norm_out.numel() == out.numel()
            a.norm_out = norm_out.data_ptr()
    else:
        assert norm_out is None
    ev = ops_mod._tl_begin()
    _lib.check(lib.ovb_conv3d(C.byref(a), _stream()), "ovb_conv3d")
    ops_mod._tl_end(ev, "conv", 2.0 * To * Ho * Wo * pc.cout * pc.n_taps * Cin, (To, Ho, Wo, Cin, pc.cout, pc.n_taps))
    return out


def can_fuse_norm(pc: PackedConv) -> bool:
    """The conv epilogue can apply the following RMS_norm when one N tile holds every output channel
    (ovb_conv3d picks the largest multiple of 16 <= 256 dividing the weight rows as tile width)."""
    rows = pc.w.shape[0]
    return rows == pc.cout and rows <= 256 and rows % 32 == 0


def rmsnorm_silu(x: torch.Tensor, gamma: torch.Tensor, silu: bool = True,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: bf16 [..., C] contiguous (channels-last); gamma fp32 [C]."""
    _need_cuda(x, gamma, out)
    lib = _lib.load()
    assert x.dtype == bf16 and x.is_contiguous() and gamma.dtype == f32 and gamma.is_contiguous()
    Cc = x.shape[-1]
    assert gamma.numel() == Cc
    if out is None:
        out = torch.empty_like(x)
    assert out.dtype == bf16 and out.is_contiguous() and out.shape == x.shape
    ev = ops_mod._tl_begin()
    _lib.check(lib.ovb_rmsnorm_silu_cl(_ptr(x), _ptr(gamma), _ptr(out), x.numel() // Cc, Cc, int(silu),
                                       _stream()), "ovb_rmsnorm_silu_cl")
    ops_mod._tl_end(ev, "norm", 4.0 * x.numel(), (x.numel() // Cc, Cc))   # "flops" slot = algorithmic bytes
    return out


def planar_to_cl(src: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, cpad: int) -> torch.Tensor:
    """src fp32 [C, T, H, W] -> bf16 [T, H, W, cpad] = src*scale[c] + shift[c] (zeros for c >= C)."""
    _need_cuda(src, scale, shift)
    lib = _lib.load()
    assert src.dtype == f32 and src.dim() == 4 and src.is_contiguous()
    Cc, T, H, W = src.shape
    out = torch.empty((T, H, W, cpad), device=src.device, dtype=bf16)
    _lib.check(lib.ovb_planar_to_cl(_ptr(src), _ptr(scale), _ptr(shift), _ptr(out), Cc, cpad, T * H * W,
                                    _stream()), "ovb_planar_to_cl")
    return out


def cl_to_planar(src: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, c: int) -> torch.Tensor:
    """src bf16 [T, H, W, Cs] -> fp32 [c, T, H, W] = float(src[..., :c])*scale + shift."""
    _need_cuda(src, scale, shift)
    lib = _lib.load()
    assert src.dtype == bf16 and src.dim() == 4 and src.is_contiguous()
    T, H, W, Cs = src.shape
    out = torch.empty((c, T, H, W), device=src.device, dtype=f32)
    _lib.check(lib.ovb_cl_to_planar(_ptr(src), Cs, _ptr(scale), _ptr(shift), _ptr(out), c, T * H * W,
                                    _stream()), "ovb_cl_to_planar")
    return out


def softmax_rows(S: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """S fp32 [rows, cols] -> bf16 softmax(scale*S) row-wise."""
    _need_cuda(S, out)
    lib = _lib.load()
    assert S.dtype == f32 and S.dim() == 2 and S.stride(1) == 1
    rows, cols = S.shape
    if out is None:
        out = torch.empty((rows, cols), device=S.device, dtype=bf16)
    assert out.dtype == bf16 and out.shape == (rows, cols) and out.stride(1) == 1
    _lib.check(lib.ovb_softmax_rows(_ptr(S), S.stride(0), _ptr(out), out.stride(0), rows, cols, float(scale),
                                    _stream()), "ovb_softmax_rows")
    return out
