"""Drop-in for the reference attention entry points (omnivideo/modules/attention.py:24-182).

Same names, argument meaning, return dtype and error behaviour; the arithmetic is one launch of the
tcgen05/TMEM kernel `ovb_attn_fwd` instead of flash_attn's varlen kernels.  Features the reference
accepts but no DiT/VAE call site uses (dropout, causal, sliding window, q_lens) raise
NotImplementedError rather than being silently ignored.
"""
from __future__ import annotations

import torch

from .. import ops

__all__ = ["flash_attention", "attention"]

_HALF = (torch.float16, torch.bfloat16)


def flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0.0, softmax_scale=None,
                    q_scale=None, causal=False, window_size=(-1, -1), deterministic=False,
                    dtype=torch.bfloat16, version=None):
    """q: [B, Lq, Nq, C], k: [B, Lk, Nk, C], v: [B, Lk, Nk, C]; k_lens: [B] valid keys per sample.
    Returns [B, Lq, Nq, C] in q's input dtype (reference attention.py:57,133)."""
    assert dtype in _HALF
    assert q.device.type == "cuda" and q.size(-1) <= 256  # same guard as the reference (:54)
    if dropout_p != 0.0 or causal or tuple(window_size) != (-1, -1):
        raise NotImplementedError("ovb flash_attention: dropout / causal / window are not on the "
                                  "Omni-Video hot path and are not implemented")
    if q.size(-1) != 128 or v.size(-1) != 128:
        raise NotImplementedError("ovb flash_attention: head_dim must be 128")
    if k.size(2) != q.size(2):
        raise NotImplementedError("ovb flash_attention: grouped-query attention is not implemented")
    if dtype != torch.bfloat16 or any(t.dtype == torch.float16 for t in (q, k, v)):
        raise NotImplementedError("ovb flash_attention computes in bf16 only")
    b, lq, out_dtype = q.size(0), q.size(1), q.dtype

    def half(x):  # reference :59-60 (cast to the compute dtype when not already half)
        x = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
        return x if (x.stride(3) == 1 and x.stride(2) == x.size(3)) else x.contiguous()

    q, k, v = half(q), half(k), half(v)
    if q_scale is not None:
        q = q * q_scale
    kl = None
    if k_lens is not None:
        kl = [int(x) for x in (k_lens.tolist() if torch.is_tensor(k_lens) else k_lens)]
    out = ops.attention(q, k, v, k_lens=kl, softmax_scale=softmax_scale)
    if q_lens is not None:
        # the reference packs only the first q_lens[b] queries; rows beyond are undefined there and
        # zero here
        ql = [int(x) for x in (q_lens.tolist() if torch.is_tensor(q_lens) else q_lens)]
        for i, n in enumerate(ql):
            out[i, n:].zero_()
    assert out.shape[:2] == (b, lq)
    return out.type(out_dtype)


def attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0.0, softmax_scale=None, q_scale=None,
              causal=False, window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16,
              fa_version=None):
    """Reference attention.py:136-182.  There the SDPA branch is a fallback for machines without
    flash_attn; here there is exactly one implementation."""
    return flash_attention(q=q, k=k, v=v, q_lens=q_lens, k_lens=k_lens, dropout_p=dropout_p,
                           softmax_scale=softmax_scale, q_scale=q_scale, causal=causal,
                           window_size=window_size, deterministic=deterministic, dtype=dtype,
                           version=fa_version)
