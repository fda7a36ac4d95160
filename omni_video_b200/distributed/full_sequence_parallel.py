"""Sequence-parallel hooks with the reference's names and signatures
(omnivideo/distributed/full_sequence_parallel.py:63,170,204; v1 names `usp_*` from
previous_version/Omni-Video1/nets/third_party/wan/distributed/xdit_context_parallel.py:66,149).

Two ways to use them, both supported:
  * the reference's own way — `types.MethodType` rebinding on a model (x2x_gen_unified.py:171-182).
    `sp_dit_forward` bound on a B200 `WanModel` simply sets `sp_group` and runs the fused DiTEngine
    (GEMM epilogues write the all-to-all layouts, cross-attention needs no collective);
  * per-module — `sp_attn_forward` / `sp_cross_attn_forward` bound on `block.self_attn` /
    `block.cross_attn` give the reference's exact dataflow (context sequence-sharded, four
    all-to-alls per attention) on the B200 kernels.
"""
from __future__ import annotations

import torch

from ..modules import blocks_api
from .ulysses import distributed_attention
from .util import get_rank, get_world_size


def sp_dit_forward(self, x, t, context, seq_len, y=None):
    """reference full_sequence_parallel.py:63-167.  `self` is a WanModel with `.sp_group` set."""
    if any("forward" in vars(b.self_attn) or "forward" in vars(b.cross_attn) for b in self.blocks):
        # the caller rebound the attention modules too (instance-level `forward`, as
        # x2x_gen_unified.py:180-181 does): honour the per-module path
        return blocks_api.dit_forward_modular(self, x, t, context, seq_len, y,
                                              sp_group=getattr(self, "sp_group", None))
    return self._engine.forward(x, t, context, seq_len, y)


def sp_attn_forward(self, x, seq_lens, grid_sizes, freqs, dtype=torch.bfloat16):
    """reference :170-201.  x: this rank's rows [B, L/p, C]."""
    b, s = x.shape[:2]
    g = getattr(self, "sp_group", None)
    q, k, v = blocks_api.qkv_project(self, x, grid_sizes, token_offset=get_rank(g) * s)
    a = distributed_attention(q, k, v, seq_lens, window_size=self.window_size, sp_group=g)
    return blocks_api._linear(self.o, a.reshape(b * s, -1)).view(b, s, -1)


def sp_cross_attn_forward(self, x, context, context_lens, dtype=torch.bfloat16):
    """reference :204-232: q from this rank's rows, k/v from this rank's slice of the context."""
    b, s = x.shape[:2]
    g = getattr(self, "sp_group", None)
    q, k, v = blocks_api.cross_qkv(self, x, context)
    a = distributed_attention(q, k, v, context_lens, sp_group=g)
    return blocks_api._linear(self.o, a.reshape(b * s, -1)).view(b, s, -1)


# Omni-Video 1 names
usp_dit_forward = sp_dit_forward
usp_attn_forward = sp_attn_forward

__all__ = ["sp_dit_forward", "sp_attn_forward", "sp_cross_attn_forward", "usp_dit_forward",
           "usp_attn_forward", "get_rank", "get_world_size"]
