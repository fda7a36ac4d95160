"""Tensor-level wrappers over the C ABI: each takes CUDA torch tensors, validates the layout the
kernel needs, and launches on torch's current stream.  PyTorch is only the allocator and stream
owner here; no op has a PyTorch/CPU fallback."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (EPI_BF16, EPI_F32, EPI_GATE_RESID_F32, EPI_GELU_BF16, EPI_RESID_BF16,  # noqa: F401
                   EPI_RESID_F32, LN_AFFINE, LN_MODULATE)

bf16 = torch.bfloat16
f32 = torch.float32


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# Optional per-launch timing (bench.py's roofline leg): when `timeline` is a list, gemm() and
# attention() bracket their launch with CUDA events on the launching stream and append
# (kind, algorithmic FLOPs, shape, start_event, end_event).  None (the default) costs nothing.
timeline = None


def timeline_active() -> bool:
    return timeline is not None


def add_launch_count(n: int):
    """Account for kernels of this library launched by a CUDA-graph replay (they do not pass through
    the C entry points that ovb_launch_count() counts)."""
    _lib.load().ovb_add_launch_count(int(n))


def _tl_begin():
    if timeline is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _tl_end(ev, kind, flops, shape):
    if ev is None:
        return
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    timeline.append((kind, flops, shape, ev, end))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.OvbError("omni_video_b200 ops need CUDA tensors (there is no CPU fallback)")


def _row_major_2d(t: torch.Tensor, name: str):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got "
                         f"shape {tuple(t.shape)} strides {t.stride()}")


# ------------------------------------------------------------------------------------------ GEMM
def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], *, epilogue: int = EPI_BF16,
         out: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
         gate: Optional[torch.Tensor] = None, rows_per_sample: int = 0,
         n_split: int = 1, k_split: int = 1) -> torch.Tensor:
    """out = epilogue(a @ w.T + bias).

    a: bf16 [M, K] (or [k_split, M, K/k_split] contiguous chunks when k_split > 1)
    w: bf16 [N, K] (nn.Linear layout);  bias: bf16 [N] or None
    gate: fp32 [n_samples, N] (per-sample vector) or fp32 [M, N] (per-token)
    n_split > 1: out is [n_split, M, N/n_split].
    """
    _need_cuda(a, w, bias, out, resid, gate)
    lib = _lib.load()
    if k_split > 1:
        assert a.dim() == 3 and a.shape[0] == k_split and a.stride(2) == 1
        M, K = a.shape[1], a.shape[2] * k_split
        lda, a_split_stride = a.stride(1), a.stride(0)
    else:
        _row_major_2d(a, "a")
        M, K = a.shape
        lda, a_split_stride = a.stride(0), 0
    _row_major_2d(w, "w")
    N = w.shape[0]
    assert w.shape[1] == K, f"K mismatch: a has {K}, w has {w.shape[1]}"
    assert a.dtype == bf16 and w.dtype == bf16
    out_f32 = epilogue in (EPI_RESID_F32, EPI_GATE_RESID_F32, EPI_F32)
    if out is None:
        shape = (n_split, M, N // n_split) if n_split > 1 else (M, N)
        out = torch.empty(shape, device=a.device, dtype=f32 if out_f32 else bf16)
    assert out.dtype == (f32 if out_f32 else bf16)
    if n_split > 1:
        assert out.dim() == 3 and out.shape == (n_split, M, N // n_split) and out.stride(2) == 1
        ldo, out_split_stride = out.stride(1), out.stride(0)
    else:
        _row_major_2d(out, "out")
        assert out.shape == (M, N)
        ldo, out_split_stride = out.stride(0), 0
    args = _lib.GemmArgs()
    args.A, args.lda = a.data_ptr(), lda
    args.W, args.ldw = w.data_ptr(), w.stride(0)
    args.bias = None if bias is None else bias.data_ptr()
    args.out, args.ldo = out.data_ptr(), ldo
    args.M, args.N, args.K = M, N, K
    args.epilogue = epilogue
    if epilogue == EPI_RESID_BF16:
        assert resid is not None and resid.dtype == bf16 and resid.shape == (M, N) and n_split == 1
        _row_major_2d(resid, "resid")
        args.resid, args.ldr = resid.data_ptr(), resid.stride(0)
    if epilogue in (EPI_RESID_F32, EPI_GATE_RESID_F32):
        assert resid is not None and resid.dtype == f32 and resid.shape == (M, N)
        _row_major_2d(resid, "resid")
        args.resid, args.ldr = resid.data_ptr(), resid.stride(0)
    if epilogue == EPI_GATE_RESID_F32:
        assert gate is not None and gate.dtype == f32 and gate.dim() == 2 and gate.stride(1) == 1
        rps = rows_per_sample or M
        if gate.shape[0] == M and M != (M + rps - 1) // rps:  # per-token table
            args.gate_row_stride, args.gate_sample_stride = gate.stride(0), gate.stride(0) * rps
        else:  # one vector per sample
            assert gate.shape[0] == (M + rps - 1) // rps, "gate rows must equal #samples or M"
            args.gate_row_stride, args.gate_sample_stride = 0, gate.stride(0)
        args.gate = gate.data_ptr()
        args.rows_per_sample = rps
    args.n_split, args.out_split_stride = n_split, out_split_stride
    args.k_split, args.a_split_stride = k_split, a_split_stride
    ev = _tl_begin()
    _lib.check(lib.ovb_gemm_bf16(C.byref(args), _stream()), "ovb_gemm_bf16")
    _tl_end(ev, "gemm", 2.0 * M * N * K, (M, N, K))
    return out


# ------------------------------------------------------------------------------------- attention
_attn_ws = {}  # (device index, stream) -> uint8 scratch for the key-split last wave of ovb_attn_fwd


def _attn_workspace(lib, args, device):
    need = int(lib.ovb_attn_workspace_size(C.byref(args)))
    if need <= 0:
        return None
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _attn_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, device=device, dtype=torch.uint8)
        _attn_ws[key] = ws
    return ws


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
              k_lens: Optional[Sequence[int]] = None, softmax_scale: Optional[float] = None,
              out: Optional[torch.Tensor] = None, split_keys: bool = True,
              out_ptrs: Optional[Sequence[int]] = None, rows_per_out: int = 0,
              out_row_stride: int = 0) -> Optional[torch.Tensor]:
    """q: [B, Lq, H, 128], k/v: [B, Lk, H, 128] bf16 views (inner two dims contiguous).
    Returns bf16 [B, Lq, H, 128].  split_keys=False forces one CTA per (q block, head, sample).
    Scatter form (B == 1): out_ptrs = raw device pointers (possibly peer-mapped); query row r goes to
    out_ptrs[r // rows_per_out] + (r % rows_per_out) * out_row_stride + head * 128; returns None."""
    _need_cuda(q, k, v, out)
    lib = _lib.load()
    assert q.dim() == 4 and k.dim() == 4 and v.dim() == 4
    B, Lq, H, D = q.shape
    Lk = k.shape[1]
    if D != 128:
        raise ValueError(f"ovb_attn_fwd supports head_dim 128 only (got {D})")
    assert k.shape == (B, Lk, H, D) and v.shape == (B, Lk, H, D)
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        assert t.dtype == bf16, f"{n} must be bf16"
        assert t.stride(3) == 1 and t.stride(2) == D, f"{n}: heads must be packed [H,128] per row"
    args = _lib.AttnArgs()
    if out_ptrs is not None:
        assert B == 1 and rows_per_out > 0 and out_row_stride >= H * D and out is None
        keep = _ptr_array(out_ptrs)
        args.out_ptrs, args.n_out_ptrs, args.rows_per_out = keep, len(out_ptrs), rows_per_out
    else:
        if out is None:
            out = torch.empty((B, Lq, H, D), device=q.device, dtype=bf16)
        assert out.dtype == bf16 and out.shape == (B, Lq, H, D) and out.stride(3) == 1 and out.stride(2) == D
    args.q, args.q_row_stride, args.q_batch_stride = q.data_ptr(), q.stride(1), q.stride(0)
    args.k, args.k_row_stride, args.k_batch_stride = k.data_ptr(), k.stride(1), k.stride(0)
    args.v, args.v_row_stride, args.v_batch_stride = v.data_ptr(), v.stride(1), v.stride(0)
    if out_ptrs is None:
        args.out, args.o_row_stride, args.o_batch_stride = out.data_ptr(), out.stride(1), out.stride(0)
    else:
        args.out, args.o_row_stride, args.o_batch_stride = None, out_row_stride, 0
    args.B, args.H, args.Lq, args.Lk = B, H, Lq, Lk
    if k_lens is not None:
        kl = [int(x) for x in k_lens]
        assert len(kl) == B
        arr = (C.c_int32 * B)(*kl)
        args.k_lens = C.cast(arr, C.POINTER(C.c_int32))
    args.softmax_scale = float(softmax_scale) if softmax_scale is not None else 0.0
    args.no_split = 0 if split_keys else 1
    ws = _attn_workspace(lib, args, q.device) if split_keys else None
    if ws is not None:
        args.workspace, args.workspace_bytes = ws.data_ptr(), ws.numel()
    ev = _tl_begin()
    _lib.check(lib.ovb_attn_fwd(C.byref(args), _stream()), "ovb_attn_fwd")
    if ev is not None:
        keys = sum(int(x) for x in k_lens) if k_lens is not None else B * Lk
        _tl_end(ev, "attn", 4.0 * Lq * keys * H * D, (B, H, Lq, Lk))
    return out


def attention_relbias(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, rel_bias: torch.Tensor,
                      k_lens: Optional[torch.Tensor] = None, softmax_scale: float = 1.0,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """umT5 self-attention, all heads and samples in one launch (reference t5.py:96-117).
    q/k/v: bf16 [B, L, H, D] views (D = 64 or 128, heads packed per row); rel_bias: fp32 [H, 2L-1] with the bias of
    key j for query i at [h, (j - i) + L - 1]; k_lens: int32 DEVICE tensor [B] (valid keys per sample) or None.
    Returns bf16 [B, L, H, D]."""
    _need_cuda(q, k, v, rel_bias, k_lens, out)
    lib = _lib.load()
    assert q.dim() == 4 and q.shape == k.shape == v.shape
    B, L, H, D = q.shape
    if D not in (64, 128):
        raise ValueError(f"ovb_attn_relbias_fwd supports head_dim 64 or 128 (got {D})")
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        assert t.dtype == bf16, f"{n} must be bf16"
        assert t.stride(3) == 1 and t.stride(2) == D, f"{n}: heads must be packed [H,{D}] per row"
    assert rel_bias.dtype == f32 and rel_bias.shape == (H, 2 * L - 1) and rel_bias.stride(1) == 1
    if k_lens is not None:
        assert k_lens.dtype == torch.int32 and k_lens.shape == (B,) and k_lens.is_contiguous()
    if out is None:
        out = torch.empty((B, L, H, D), device=q.device, dtype=bf16)
    assert out.dtype == bf16 and out.shape == (B, L, H, D) and out.stride(3) == 1 and out.stride(2) == D
    args = _lib.AttnRelBiasArgs()
    args.q, args.q_row_stride, args.q_batch_stride = q.data_ptr(), q.stride(1), q.stride(0)
    args.k, args.k_row_stride, args.k_batch_stride = k.data_ptr(), k.stride(1), k.stride(0)
    args.v, args.v_row_stride, args.v_batch_stride = v.data_ptr(), v.stride(1), v.stride(0)
    args.out, args.o_row_stride, args.o_batch_stride = out.data_ptr(), out.stride(1), out.stride(0)
    args.rel_bias, args.bias_head_stride = rel_bias.data_ptr(), rel_bias.stride(0)
    args.k_lens = _ptr(k_lens)
    args.B, args.H, args.L, args.head_dim = B, H, L, D
    args.softmax_scale = float(softmax_scale)
    _lib.check(lib.ovb_attn_relbias_fwd(C.byref(args), _stream()), "ovb_attn_relbias_fwd")
    return out


# ------------------------------------------------------------------------------------- layernorm
def layernorm_modulate(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, eps: float, *,
                       rows_per_sample: int = 0, out_dtype=bf16, round_ln_bf16: bool = False,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LN(x) * (1 + scale) + shift.  x: [rows, d] fp32/bf16; scale/shift fp32 [n_samples, d] or
    [rows, d] (row stride may be larger than d: e.g. slices of the [.., 6, d] modulation table)."""
    return _layernorm(x, LN_MODULATE, scale, shift, eps, rows_per_sample, out_dtype, round_ln_bf16, out)


def layernorm_affine(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float, *,
                     out_dtype=bf16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _layernorm(x, LN_AFFINE, weight, bias, eps, 0, out_dtype, False, out)


def _layernorm(x, mode, a, b, eps, rows_per_sample, out_dtype, round_ln_bf16, out):
    _need_cuda(x, a, b, out)
    lib = _lib.load()
    _row_major_2d(x, "x")
    rows, d = x.shape
    assert x.dtype in (f32, bf16)
    if out is None:
        out = torch.empty((rows, d), device=x.device, dtype=out_dtype)
    _row_major_2d(out, "out")
    args = _lib.LnArgs()
    args.x, args.ldx, args.x_is_bf16 = x.data_ptr(), x.stride(0), int(x.dtype == bf16)
    args.out, args.ldo, args.out_is_f32 = out.data_ptr(), out.stride(0), int(out.dtype == f32)
    args.rows, args.d, args.eps, args.mode = rows, d, eps, mode
    if mode == LN_MODULATE:
        assert a.dtype == f32 and b.dtype == f32 and a.dim() == 2 and b.dim() == 2
        assert a.stride(1) == 1 and b.stride(1) == 1 and a.stride(0) == b.stride(0)
        assert a.shape == b.shape and a.shape[1] == d
        rps = rows_per_sample or rows
        n_samples = (rows + rps - 1) // rps
        if a.shape[0] == rows and rows != n_samples:
            args.mod_row_stride, args.mod_sample_stride = a.stride(0), a.stride(0) * rps
        else:
            assert a.shape[0] == n_samples, "modulation rows must equal #samples or #rows"
            args.mod_row_stride, args.mod_sample_stride = 0, a.stride(0)
        args.rows_per_sample = rps
    else:
        assert a.dtype == bf16 and b.dtype == bf16 and a.numel() == d and b.numel() == d
        assert a.is_contiguous() and b.is_contiguous()
    args.a, args.b = a.data_ptr(), b.data_ptr()
    args.round_ln_bf16 = int(round_ln_bf16)
    _lib.check(lib.ovb_layernorm(C.byref(args), _stream()), "ovb_layernorm")
    return out


# ------------------------------------------------------------------------------- rmsnorm (+rope)
def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[int(v) for v in ptrs])


def rmsnorm_rope_(x: torch.Tensor, weight: torch.Tensor, eps: float, *,
                  cs_table: Optional[torch.Tensor] = None, grid=None, rows_per_sample: int = 0,
                  token_offset: int = 0, dst_ptrs: Optional[Sequence[int]] = None, dst_ld: int = 0,
                  copy_only: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In place on x: bf16 [rows, d] (ld arbitrary) or [col_split, rows, d/col_split].
    out (bf16 [rows, d], 2-D x only): write the result there and leave x alone — the scatter form with one local
    destination (d <= 5120); returns out.
    With cs_table (fp32 [T, 64, 2]) and grid=(F, H, W) also applies the 3-D RoPE.
    Scatter form: dst_ptrs = col_split raw device pointers (possibly peer-mapped memory of other
    GPUs); column group g of row r is written to dst_ptrs[g] + r * dst_ld instead of in place;
    copy_only skips the normalisation (plain scatter)."""
    _need_cuda(x, weight, cs_table)
    lib = _lib.load()
    assert x.dtype == bf16 and weight.dtype == bf16 and weight.is_contiguous()
    args = _lib.RmsRopeArgs()
    if x.dim() == 3:
        col_split, rows, dc = x.shape
        assert x.stride(2) == 1
        d = dc * col_split
        args.ldx, args.col_split, args.col_split_stride = x.stride(1), col_split, x.stride(0)
    else:
        _row_major_2d(x, "x")
        rows, d = x.shape
        args.ldx, args.col_split, args.col_split_stride = x.stride(0), 1, 0
    assert weight.numel() == d
    args.x, args.rows, args.d = x.data_ptr(), rows, d
    args.weight, args.eps = weight.data_ptr(), eps
    if cs_table is not None:
        assert cs_table.dtype == f32 and cs_table.is_contiguous() and cs_table.shape[1:] == (64, 2)
        F, H, W = (int(g) for g in grid)
        args.rope, args.cs_table, args.table_len = 1, cs_table.data_ptr(), cs_table.shape[0]
        args.F, args.H, args.W = F, H, W
        args.rows_per_sample = rows_per_sample or rows
        args.token_offset = token_offset
    if out is not None:
        assert dst_ptrs is None and x.dim() == 2 and out.dtype == bf16 and out.shape == x.shape and out.stride(1) == 1
        _need_cuda(out)
        dst_ptrs, dst_ld = [out.data_ptr()], out.stride(0)
    if dst_ptrs is not None:
        assert len(dst_ptrs) == args.col_split
        keep = _ptr_array(dst_ptrs)
        args.dst_ptrs, args.dst_ld, args.copy_only = keep, dst_ld, int(copy_only)
    _lib.check(lib.ovb_rmsnorm_rope(C.byref(args), _stream()), "ovb_rmsnorm_rope")
    return x if out is None else out


def peer_signal_wait(pad_ptrs: Sequence[int], rank: int, state: torch.Tensor):
    """Cross-GPU barrier on peer-mapped flag pads (include/ovb200.h "peer transport"): makes this rank's
    earlier peer stores visible, signals every peer with the next epoch, waits for all peers' signal.
    state: int32 [2] on this device = (epoch counter, time-out status), both advanced by the kernel."""
    lib = _lib.load()
    assert state.dtype == torch.int32 and state.numel() == 2 and state.is_cuda
    _lib.check(lib.ovb_peer_signal_wait(_ptr_array(pad_ptrs), rank, len(pad_ptrs), 0,
                                        C.c_void_p(state.data_ptr()), C.c_void_p(state.data_ptr() + 4),
                                        _stream()),
               "ovb_peer_signal_wait")


# ----------------------------------------------------------------------------- patchify & friends
def patchify(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [C, F, H, W] fp32/bf16 contiguous -> bf16 [F*(H/2)*(W/2), C*4]."""
    _need_cuda(x, out)
    lib = _lib.load()
    assert x.dim() == 4 and x.is_contiguous() and x.dtype in (f32, bf16)
    Cc, F, H, W = x.shape
    L = F * (H // 2) * (W // 2)
    if out is None:
        out = torch.empty((L, Cc * 4), device=x.device, dtype=bf16)
    _row_major_2d(out, "out")
    assert out.dtype == bf16 and out.shape[0] >= L and out.shape[1] == Cc * 4
    _lib.check(lib.ovb_patchify(_ptr(x), int(x.dtype == bf16), _ptr(out), out.stride(0), Cc, F, H, W,
                                _stream()), "ovb_patchify")
    return out


def patchify3d(x: torch.Tensor, patch) -> torch.Tensor:
    """x: [C, F, H, W] fp32/bf16 contiguous -> bf16 [(F//pt)*(H//ph)*(W//pw), C*pt*ph*pw], columns in the
    order of Conv3d.weight.flatten(1) (a stride=kernel Conv3d is this gather followed by a GEMM)."""
    _need_cuda(x)
    lib = _lib.load()
    assert x.dim() == 4 and x.is_contiguous() and x.dtype in (f32, bf16)
    pt, ph, pw = (int(v) for v in patch)
    Cc, F, H, W = x.shape
    out = torch.empty(((F // pt) * (H // ph) * (W // pw), Cc * pt * ph * pw), device=x.device, dtype=bf16)
    _lib.check(lib.ovb_patchify3d(_ptr(x), int(x.dtype == bf16), _ptr(out), out.stride(0), Cc, F, H, W,
                                  pt, ph, pw, _stream()), "ovb_patchify3d")
    return out


def unpatchify(y: torch.Tensor, C_out: int, grid) -> torch.Tensor:
    """y: fp32 [>=F*Hp*Wp, 4*C_out] -> fp32 [C_out, F, 2*Hp, 2*Wp]."""
    _need_cuda(y)
    lib = _lib.load()
    _row_major_2d(y, "y")
    F, Hp, Wp = (int(g) for g in grid)
    assert y.dtype == f32 and y.shape[1] == 4 * C_out and y.shape[0] >= F * Hp * Wp
    out = torch.empty((C_out, F, Hp * 2, Wp * 2), device=y.device, dtype=f32)
    _lib.check(lib.ovb_unpatchify(_ptr(y), y.stride(0), _ptr(out), C_out, F, Hp, Wp, _stream()),
               "ovb_unpatchify")
    return out


def linear_f32(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], *, act_in: int = 0,
               act_out: int = 0) -> torch.Tensor:
    """fp32 x [M, K] times bf16 W [N, K]^T in fp32 (+bias), optional SiLU on the input/output."""
    _need_cuda(x, w, bias)
    lib = _lib.load()
    _row_major_2d(x, "x")
    _row_major_2d(w, "w")
    assert x.dtype == f32 and w.dtype == bf16 and x.shape[1] == w.shape[1]
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), device=x.device, dtype=f32)
    _lib.check(lib.ovb_linear_f32(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(out),
                                  out.stride(0), M, N, K, act_in, act_out, _stream()),
               "ovb_linear_f32")
    return out


def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """t: [n] (any real dtype) -> fp32 [n, dim]; computed in fp64 on device."""
    _need_cuda(t)
    lib = _lib.load()
    tf = t.to(f32).contiguous()
    out = torch.empty((tf.numel(), dim), device=t.device, dtype=f32)
    _lib.check(lib.ovb_sinusoidal_embedding(_ptr(tf), _ptr(out), tf.numel(), dim, _stream()),
               "ovb_sinusoidal_embedding")
    return out


def modulation_table(mod: torch.Tensor, e: torch.Tensor, e_has_k: bool) -> torch.Tensor:
    """mod: bf16 [n_layers, K, d]; e: fp32 [B, K*d] (e_has_k) or [B, d] -> fp32 [n_layers, B, K, d]."""
    _need_cuda(mod, e)
    lib = _lib.load()
    assert mod.dtype == bf16 and mod.is_contiguous() and mod.dim() == 3
    assert e.dtype == f32 and e.is_contiguous() and e.dim() == 2
    n_layers, K, d = mod.shape
    B = e.shape[0]
    assert e.shape[1] == (K * d if e_has_k else d)
    out = torch.empty((n_layers, B, K, d), device=e.device, dtype=f32)
    _lib.check(lib.ovb_modulation_table(_ptr(mod), _ptr(e), _ptr(out), n_layers, B, K, d,
                                        int(e_has_k), _stream()), "ovb_modulation_table")
    return out


def cast_bf16_to_f32(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    lib = _lib.load()
    assert x.dtype == bf16 and x.is_contiguous() and x.numel() % 8 == 0
    out = torch.empty(x.shape, device=x.device, dtype=f32)
    _lib.check(lib.ovb_cast_bf16_to_f32(_ptr(x), _ptr(out), x.numel(), _stream()),
               "ovb_cast_bf16_to_f32")
    return out


def cast_f32_to_bf16(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    lib = _lib.load()
    assert x.dtype == f32 and x.is_contiguous() and x.numel() % 8 == 0
    out = torch.empty(x.shape, device=x.device, dtype=bf16)
    _lib.check(lib.ovb_cast_f32_to_bf16(_ptr(x), _ptr(out), x.numel(), _stream()),
               "ovb_cast_f32_to_bf16")
    return out


def lincomb(terms, out: Optional[torch.Tensor] = None, out_dtype=None, round_bf16=()) -> torch.Tensor:
    """out = sum_k c_k * x_k for `terms` = [(c_k, x_k), ...] (<= 5 tensors of one shape, fp32 or bf16).
    `round_bf16`: indices of terms rounded to bf16 before the sum."""
    lib = _lib.load()
    xs = [x for _, x in terms]
    _need_cuda(*xs, out)
    assert 1 <= len(terms) <= 5
    ref = xs[0]
    for x in xs:
        assert x.shape == ref.shape and x.is_contiguous() and x.dtype in (f32, bf16)
    if out is None:
        out = torch.empty(ref.shape, device=ref.device, dtype=out_dtype or ref.dtype)
    assert out.shape == ref.shape and out.is_contiguous() and out.dtype in (f32, bf16)
    a = _lib.LinCombArgs()
    for k, (c, x) in enumerate(terms):
        a.x[k] = x.data_ptr()
        a.c[k] = float(c)
        a.is_bf16[k] = int(x.dtype == bf16)
    a.n_terms = len(terms)
    a.round_mask = sum(1 << k for k in round_bf16)
    a.out, a.out_is_bf16, a.n = out.data_ptr(), int(out.dtype == bf16), ref.numel()
    _lib.check(lib.ovb_lincomb(C.byref(a), _stream()), "ovb_lincomb")
    return out


# --------------------------------------------------------------------------------- host helpers
def rope_cos_sin_table(max_len: int = 1024, head_dim: int = 128, theta: float = 10000.0) -> torch.Tensor:
    """fp32 [max_len, head_dim/2, 2] (cos, sin) of the reference's complex `freqs` table
    (model.py:26-34, :405-411): three spectra of 22/21/21 pairs, computed in fp64 like the
    reference and rounded once to fp32.  Host (CPU) tensor; move to the device once."""
    c = head_dim // 2
    dims = [head_dim - 4 * (head_dim // 6), 2 * (head_dim // 6), 2 * (head_dim // 6)]
    assert sum(dims) // 2 == c
    pos = torch.arange(max_len, dtype=torch.float64)
    parts = []
    for dsub in dims:
        inv = 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float64),
                              torch.arange(0, dsub, 2, dtype=torch.float64) / dsub)
        parts.append(torch.outer(pos, inv))
    ang = torch.cat(parts, dim=1)  # [max_len, 64]
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).to(f32).contiguous()


def default_softmax_scale(head_dim: int) -> float:
    return 1.0 / math.sqrt(head_dim)
