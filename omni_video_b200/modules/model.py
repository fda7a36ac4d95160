"""B200-native WanModel (drop-in for reference omnivideo/modules/model.py).

The module tree (names, shapes, `[out, in]` Linear layout) is the reference's, so its checkpoints
load unchanged (`blocks.{i}.self_attn.{q,k,v,o}`, `norm_q`, `modulation`, `ffn.{0,2}`, `head.head`,
`patch_embedding`, `text_embedding.{0,2}`, `time_embedding.{0,2}`, `time_projection.1`).  None of
those submodules' own `forward` runs on the hot path: `WanModel.forward` hands the whole step to
`DiTEngine`, which drives the sm_100a kernels of libovb200 over flat row-major `[tokens, channels]`
buffers:

    per block (reference model.py:218-258)                         kernel
    -------------------------------------------------------------  ----------------------------
    LN(x)*(1+scale)+shift                     fp32 -> bf16          ovb_layernorm (MODULATE)
    [q|k|v] = h Wqkv^T + b                    one GEMM, N = 3d      ovb_gemm_bf16 (BF16)
    q,k <- RoPE(RMSNorm_d(.) * w)             in place              ovb_rmsnorm_rope  x2
    a = softmax(q k^T / sqrt(128)) v          k_lens = seq_lens     ovb_attn_fwd
    x += (a Wo^T + b) * gate                  fp32 residual         ovb_gemm_bf16 (GATE_RESID_F32)
    LN_affine(x)                              fp32 -> bf16          ovb_layernorm (AFFINE)
    qc = RMSNorm(h Wq^T + b) ; [kc|vc] from the text context        ovb_gemm_bf16, ovb_rmsnorm_rope
    x += softmax(qc kc^T) vc Wo^T + b                               ovb_attn_fwd, ovb_gemm_bf16 (RESID_F32)
    x += (GELU(LN(x)*(1+s)+t W1^T + b1) W2^T + b2) * gate           ovb_layernorm, ovb_gemm_bf16 x2

Rounding points follow the reference's bf16-autocast execution (SURVEY.md §5 "Precision regime"):
GEMM outputs are rounded to bf16 before gate/residual math, the residual stream and all
LayerNorm/modulation math is fp32, the time-embedding MLP and the head run in fp32 on bf16-stored
weights.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..ops import (EPI_BF16, EPI_GATE_RESID_F32, EPI_GELU_BF16, EPI_RESID_F32)

__all__ = ["WanModel", "DiTEngine"]

bf16 = torch.bfloat16
f32 = torch.float32


# ------------------------------------------------------------------------------------------------
# Parameter containers (same attribute names as the reference; forward() of the leaves is only
# used by callers that patch at module granularity, see distributed/full_sequence_parallel.py)
# ------------------------------------------------------------------------------------------------
class WanRMSNorm(nn.Module):
    """reference model.py:68-84"""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        shp = x.shape
        y = x.reshape(-1, shp[-1]).to(bf16).contiguous()
        ops.rmsnorm_rope_(y, self.weight.data.to(bf16), self.eps)
        return y.view(shp).type_as(x)


class WanLayerNorm(nn.LayerNorm):
    """reference model.py:87-97"""

    def __init__(self, dim, eps=1e-6, elementwise_affine=False):
        super().__init__(dim, elementwise_affine=elementwise_affine, eps=eps)

    def forward(self, x):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if x2.dtype not in (f32, bf16):
            x2 = x2.float()
        x2 = x2.contiguous()
        if self.elementwise_affine:
            y = ops.layernorm_affine(x2, self.weight.data.to(bf16), self.bias.data.to(bf16),
                                     self.eps, out_dtype=f32)
        else:
            zero = torch.zeros((1, shp[-1]), device=x.device, dtype=f32)
            y = ops.layernorm_modulate(x2, zero, zero, self.eps, out_dtype=f32)
        return y.view(shp).type_as(x)


class WanSelfAttention(nn.Module):
    """reference model.py:100-154 (parameters); the math lives in DiTEngine."""

    def __init__(self, dim, num_heads, window_size=(-1, -1), qk_norm=True, eps=1e-6):
        assert dim % num_heads == 0
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.window_size, self.qk_norm, self.eps = window_size, qk_norm, eps
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        self.o = nn.Linear(dim, dim)
        self.norm_q = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()
        self.norm_k = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()

    def forward(self, x, seq_lens, grid_sizes, freqs):
        """x: [B, L, C] -> [B, L, C] (bf16).  Module-granular entry (reference :125-154)."""
        from . import blocks_api
        return blocks_api.self_attention(self, x, seq_lens, grid_sizes)


class WanCrossAttention(WanSelfAttention):
    """reference model.py:157-179"""

    def forward(self, x, context, context_lens):
        from . import blocks_api
        return blocks_api.cross_attention(self, x, context, context_lens)


class WanAttentionBlock(nn.Module):
    """reference model.py:182-258"""

    def __init__(self, dim, ffn_dim, num_heads, window_size=(-1, -1), qk_norm=True,
                 cross_attn_norm=False, eps=1e-6):
        super().__init__()
        self.dim, self.ffn_dim, self.num_heads = dim, ffn_dim, num_heads
        self.window_size, self.qk_norm, self.cross_attn_norm, self.eps = (
            window_size, qk_norm, cross_attn_norm, eps)
        self.norm1 = WanLayerNorm(dim, eps)
        self.self_attn = WanSelfAttention(dim, num_heads, window_size, qk_norm, eps)
        self.norm3 = WanLayerNorm(dim, eps, elementwise_affine=True) if cross_attn_norm \
            else nn.Identity()
        self.cross_attn = WanCrossAttention(dim, num_heads, (-1, -1), qk_norm, eps)
        self.norm2 = WanLayerNorm(dim, eps)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"),
                                 nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim**0.5)

    def forward(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens):
        from . import blocks_api
        return blocks_api.attention_block(self, x, e, seq_lens, grid_sizes, context, context_lens)


class Head(nn.Module):
    """reference model.py:261-290"""

    def __init__(self, dim, out_dim, patch_size, eps=1e-6):
        super().__init__()
        self.dim, self.out_dim, self.patch_size, self.eps = dim, out_dim, patch_size, eps
        self.norm = WanLayerNorm(dim, eps)
        self.head = nn.Linear(dim, math.prod(patch_size) * out_dim)
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim**0.5)

    def forward(self, x, e):
        from . import blocks_api
        return blocks_api.head(self, x, e)


# ------------------------------------------------------------------------------------------------
# Engine
# ------------------------------------------------------------------------------------------------
class _Packed:
    """bf16 weight views / fused copies derived from the module parameters (owned by the engine,
    rebuilt when the parameters move or change version)."""
    pass


def _bf(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    return (t if t.dtype == bf16 else t.to(bf16)).contiguous()


class DiTEngine:
    """Runs WanModel.forward on the B200 kernels.  One instance per WanModel."""

    def __init__(self, model: "WanModel"):
        self.m = model
        self._packed = None
        self._packed_key = None
        self._cs_table = None
        self._graphs = {}           # forward signature -> _GraphedForward
        self.graph_replays = 0

    # -------------------------------------------------------------------------------- weights
    def _key(self):
        p = self.m.patch_embedding.weight
        sp = self.sp_size()
        return (p.device, p.dtype, sp, tuple(q._version for q in self.m.parameters()))

    def sp_group(self):
        return getattr(self.m, "sp_group", None)

    def sp_size(self) -> int:
        g = self.sp_group()
        if g is None:
            return 1
        import torch.distributed as dist
        return dist.get_world_size(group=g)

    def sp_rank(self) -> int:
        g = self.sp_group()
        if g is None:
            return 0
        import torch.distributed as dist
        return dist.get_rank(group=g)

    @staticmethod
    def _check_device(dev):
        """The product has no CPU path: parameters must be on the GPU the kernels run on."""
        if dev.type != "cuda":
            raise ops._lib.OvbError("WanModel parameters must live on a CUDA device: the B200 path "
                                    "has no CPU fallback")

    def packed(self) -> _Packed:
        key = self._key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        m = self.m
        dev = m.patch_embedding.weight.device
        self._check_device(dev)
        P = _Packed()
        sp = self.sp_size()
        d, h = m.dim, m.num_heads
        if h % sp != 0:
            # same condition and wording as reference distributed/ulysses.py:33-38
            raise ValueError(f"q.num_heads={h} must be divisible by sp_size={sp}")
        P.patch_w = _bf(m.patch_embedding.weight.flatten(1))
        P.patch_b = _bf(m.patch_embedding.bias)
        P.text_w0, P.text_b0 = _bf(m.text_embedding[0].weight), _bf(m.text_embedding[0].bias)
        P.text_w2, P.text_b2 = _bf(m.text_embedding[2].weight), _bf(m.text_embedding[2].bias)
        P.time_w0, P.time_b0 = _bf(m.time_embedding[0].weight), _bf(m.time_embedding[0].bias)
        P.time_w2, P.time_b2 = _bf(m.time_embedding[2].weight), _bf(m.time_embedding[2].bias)
        P.tproj_w, P.tproj_b = _bf(m.time_projection[1].weight), _bf(m.time_projection[1].bias)
        P.head_w, P.head_b = _bf(m.head.head.weight), _bf(m.head.head.bias)
        P.head_mod = _bf(m.head.modulation.reshape(1, 2, d))
        P.block_mod = _bf(torch.stack([b.modulation.detach().reshape(6, d) for b in m.blocks]))
        P.blocks = []
        dg = d // sp  # columns per Ulysses rank
        for blk in m.blocks:
            B = _Packed()
            sa, ca = blk.self_attn, blk.cross_attn
            wq, wk, wv = sa.q.weight.detach(), sa.k.weight.detach(), sa.v.weight.detach()
            bq, bk, bv = sa.q.bias.detach(), sa.k.bias.detach(), sa.v.bias.detach()
            if sp == 1:
                B.qkv_w = _bf(torch.cat([wq, wk, wv], 0))
                B.qkv_b = _bf(torch.cat([bq, bk, bv], 0))
            else:
                # rank-major row order [g: q_g | k_g | v_g] so that the GEMM's n_split epilogue
                # writes the Ulysses all-to-all send layout directly
                B.qkv_w = _bf(torch.cat([torch.cat([w[g * dg:(g + 1) * dg] for w in (wq, wk, wv)], 0)
                                         for g in range(sp)], 0))
                B.qkv_b = _bf(torch.cat([torch.cat([b_[g * dg:(g + 1) * dg] for b_ in (bq, bk, bv)], 0)
                                         for g in range(sp)], 0))
            B.o_w, B.o_b = _bf(sa.o.weight), _bf(sa.o.bias)
            B.nq_w = _bf(sa.norm_q.weight) if sa.qk_norm else None
            B.nk_w = _bf(sa.norm_k.weight) if sa.qk_norm else None
            B.cq_w, B.cq_b = _bf(ca.q.weight), _bf(ca.q.bias)
            B.ckv_w = _bf(torch.cat([ca.k.weight.detach(), ca.v.weight.detach()], 0))
            B.ckv_b = _bf(torch.cat([ca.k.bias.detach(), ca.v.bias.detach()], 0))
            B.co_w, B.co_b = _bf(ca.o.weight), _bf(ca.o.bias)
            B.cnq_w = _bf(ca.norm_q.weight) if ca.qk_norm else None
            B.cnk_w = _bf(ca.norm_k.weight) if ca.qk_norm else None
            if blk.cross_attn_norm:
                B.n3_w, B.n3_b = _bf(blk.norm3.weight), _bf(blk.norm3.bias)
            else:
                B.n3_w = B.n3_b = None
            B.f1_w, B.f1_b = _bf(blk.ffn[0].weight), _bf(blk.ffn[0].bias)
            B.f2_w, B.f2_b = _bf(blk.ffn[2].weight), _bf(blk.ffn[2].bias)
            P.blocks.append(B)
        self._packed, self._packed_key = P, key
        if self._cs_table is None or self._cs_table.device != dev:
            self._cs_table = ops.rope_cos_sin_table(1024, m.dim // m.num_heads).to(dev)
        return P

    # -------------------------------------------------------------------------------- forward
    @torch.no_grad()
    def embed(self, x, t, context, seq_len, sp: int = 1, rank: int = 0):
        """Embedding stage of reference model.py:444-485 for the rows owned by `rank` of `sp`.
        Returns (xres fp32 [B*s_loc, d], e fp32 [B, d], e0 fp32 [B, 6d], ctx bf16 [B*text_len, d],
        grids, seq_lens)."""
        m = self.m
        P = self.packed()
        dev = P.patch_w.device
        d = m.dim
        B = len(x)
        assert tuple(m.patch_size) == (1, 2, 2), "patch_size (1,2,2) is the only one on the hot path"
        grids = [(u.shape[1], u.shape[2] // 2, u.shape[3] // 2) for u in x]
        seq_lens = [g[0] * g[1] * g[2] for g in grids]
        assert max(seq_lens) <= seq_len  # reference model.py:460
        assert seq_len % sp == 0, "seq_len must be a multiple of the sequence-parallel size"
        s_loc = seq_len // sp           # rows per sample held by this rank
        row0 = rank * s_loc             # first token of this rank inside each sample

        # ---- patch embedding -> fp32 residual stream (padded rows are exactly zero, :461-464)
        xres = torch.zeros((B * s_loc, d), device=dev, dtype=f32)
        for b, u in enumerate(x):
            u = u.to(dev)
            if u.dtype not in (f32, bf16):
                u = u.float()
            tok = ops.patchify(u.contiguous())                      # [L_b, 64] bf16
            lo, hi = row0, min(row0 + s_loc, seq_lens[b])            # this rank's valid tokens
            if hi > lo:
                # out = 0 + bf16(acc + bias): the bf16-rounded embedding, widened to fp32 in place
                rows = xres[b * s_loc:b * s_loc + (hi - lo)]
                ops.gemm(tok[lo:hi], P.patch_w, P.patch_b, epilogue=EPI_RESID_F32, out=rows,
                         resid=rows)

        # ---- time embedding (fp32 math on bf16-stored weights, :466-476), one row per sample
        t = t.to(dev)
        if t.dim() == 2:
            # per-token timesteps t [B, seq_len] (reference model.py:466-476 embeds every token's own
            # t).  Uniform inside every sample -> the one-row-per-sample fast path (identical
            # numerics); otherwise one embedding row per token of this rank: e [B*s_loc, d],
            # e0 [B*s_loc, 6d], and the blocks read their modulation per ROW.
            assert t.shape == (B, seq_len), f"t must be [B] or [B, seq_len]; got {tuple(t.shape)}"
            if bool((t == t[:, :1]).all()):
                t = t[:, 0]
            else:
                t = t[:, row0:row0 + s_loc].contiguous()
        assert t.numel() in (B, B * s_loc)
        sin = ops.sinusoidal_embedding(t.reshape(-1), m.freq_dim)                   # [B | B*s_loc, 256]
        e = ops.linear_f32(ops.linear_f32(sin, P.time_w0, P.time_b0, act_out=1),
                           P.time_w2, P.time_b2)                                    # [B, d]
        e0 = ops.linear_f32(e, P.tproj_w, P.tproj_b, act_in=1)                      # [B, 6d]

        # ---- text context (:478-485): zero-pad to text_len, 2-layer GELU MLP; context_lens=None
        Lc = m.text_len
        ctx_in = torch.zeros((B * Lc, m.text_dim), device=dev, dtype=bf16)
        for b, c in enumerate(context):
            n = c.shape[0]
            assert n <= Lc, f"context length {n} exceeds text_len {Lc}"
            ctx_in[b * Lc:b * Lc + n] = c.to(device=dev, dtype=bf16)
        ctx = ops.gemm(ops.gemm(ctx_in, P.text_w0, P.text_b0, epilogue=EPI_GELU_BF16),
                       P.text_w2, P.text_b2, epilogue=EPI_BF16)                     # [B*Lc, d]
        return xres, e, e0, ctx, grids, seq_lens

    # ------------------------------------------------------------------------------ CUDA graph
    def use_graph(self) -> bool:
        v = getattr(self.m, "use_cuda_graph", None)
        return bool(v) if v is not None else os.environ.get("OVB_CUDA_GRAPH", "0") == "1"

    @torch.no_grad()
    def forward(self, x: List[torch.Tensor], t: torch.Tensor, context: List[torch.Tensor],
                seq_len: int, y: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
        """Eager launch sequence, or — `model.use_cuda_graph = True` / OVB_CUDA_GRAPH=1 — one CUDA graph
        per forward signature (shapes, context lengths, seq_len): the ~900 launches of a 30-block
        forward are captured once (all buffers come from the graph's private pool, the peer barriers
        keep their epoch in device memory) and every later call is input copies + one graph launch.
        What that buys is the launch gaps: nothing at 32760 tokens on one GPU, several per cent when
        sequence parallelism cuts every kernel to a fraction of a millisecond."""
        if (not self.use_graph() or not x[0].is_cuda or t.dim() != 1 or ops.timeline_active()
                or torch.cuda.is_current_stream_capturing() or (self.sp_size() > 1 and len(x) != 1)):
            return self._forward(x, t, context, seq_len, y)
        sig = (tuple((tuple(u.shape), u.dtype) for u in x), tuple(t.shape), t.dtype,
               tuple((tuple(c.shape), c.dtype) for c in context), seq_len,
               None if y is None else tuple((tuple(u.shape), u.dtype) for u in y), self._key())
        g = self._graphs.get(sig)
        if g is None:
            if len(self._graphs) >= 4:
                self._graphs.pop(next(iter(self._graphs)))
            g = self._graphs[sig] = _GraphedForward(self, x, t, context, seq_len, y)
        self.graph_replays += 1
        return g(x, t, context, y)

    @torch.no_grad()
    def _forward(self, x: List[torch.Tensor], t: torch.Tensor, context: List[torch.Tensor],
                 seq_len: int, y: Optional[List[torch.Tensor]] = None, finish: bool = True):
        """finish=False stops after the head projection and returns (head output of this rank's rows,
        grids): the part a CUDA graph captures.  What follows (`_finish`) holds the forward's only NCCL
        collective and is always run eagerly."""
        m = self.m
        if m.model_type == "i2v":
            assert y is not None
        P = self.packed()
        dev = P.patch_w.device
        d, h, eps = m.dim, m.num_heads, m.eps
        sp, rank = self.sp_size(), self.sp_rank()
        if y is not None:
            x = [torch.cat([u, v], dim=0) for u, v in zip(x, y)]
        B = len(x)
        if sp > 1 and B != 1:
            # the reference pipeline always calls with one sample per forward (x2x_gen_unified.py
            # :480-481); run samples one after another under sequence parallelism
            assert finish, "batched sequence-parallel forwards are not captured"
            out = []
            for i in range(B):
                out += self._forward([x[i]], t[i:i + 1], [context[i]], seq_len, None)
            return out
        xres, e, e0, ctx, grids, seq_lens = self.embed(x, t, context, seq_len, sp, rank)
        s_loc = seq_len // sp
        row0 = rank * s_loc
        M = B * s_loc
        Lc = m.text_len
        per_token = e.shape[0] != B                             # t was [B, seq_len] with distinct values
        # modulation = block parameter + time embedding: one row per sample [N, B, 6, d] fp32, or — with
        # per-token timesteps — one row per token, built layer by layer ([M, 6, d] is 1.2 GB at 32760 tokens)
        mod = None if per_token else ops.modulation_table(P.block_mod, e0, True)
        hmod = ops.modulation_table(P.head_mod, e, False)       # [1, B | M, 2, d] fp32

        # ---- work buffers
        hbuf = torch.empty((M, d), device=dev, dtype=bf16)
        attn = torch.empty((M, d), device=dev, dtype=bf16)
        qc = torch.empty((M, d), device=dev, dtype=bf16)
        ffn_mid = torch.empty((M, m.ffn_dim), device=dev, dtype=bf16)
        ckv = torch.empty((B * Lc, 2 * d), device=dev, dtype=bf16)
        if sp == 1:
            qkv = torch.empty((M, 3 * d), device=dev, dtype=bf16)
        else:
            from ..distributed import peer
            dg = d // sp
            qkv_send = torch.empty((sp, M, 3 * dg), device=dev, dtype=bf16)
            # peer transport (NVLink stores from the kernels + flag barriers) when the group's memory
            # can be mapped; otherwise NCCL all_to_all_single on send/receive buffers
            tr = peer.get_transport(self.sp_group(), dev, M, d)
            if tr is None:
                qkv_recv = torch.empty((sp, M, 3 * dg), device=dev, dtype=bf16)
                att_send = torch.empty((sp, M, dg), device=dev, dtype=bf16)
                att_recv = torch.empty((sp, M, dg), device=dev, dtype=bf16)
            else:
                qkv_recv, att_recv = tr.qkv_recv, tr.att_recv
        grid0 = grids[0]
        same_grid = all(g == grid0 for g in grids)      # a ragged batch (reference model.py:455-464 pads every
        #                                                 sample to seq_len) rotates each sample with its own grid

        for li, W in enumerate(P.blocks):
            # [B | M, 6, d]: shift_sa, scale_sa, gate_sa, shift_ffn, scale_ffn, gate_ffn
            mb = ops.modulation_table(P.block_mod[li:li + 1], e0, True)[0] if per_token else mod[li]
            # ---------------- self attention
            ops.layernorm_modulate(xres, mb[:, 1], mb[:, 0], eps, rows_per_sample=s_loc,
                                   round_ln_bf16=(li == 0), out=hbuf)
            if sp == 1:
                ops.gemm(hbuf, W.qkv_w, W.qkv_b, epilogue=EPI_BF16, out=qkv)
                q2, k2 = qkv[:, :d], qkv[:, d:2 * d]
                if W.nq_w is not None and same_grid:
                    ops.rmsnorm_rope_(q2, W.nq_w, eps, cs_table=self._cs_table, grid=grid0,
                                      rows_per_sample=s_loc)
                    ops.rmsnorm_rope_(k2, W.nk_w, eps, cs_table=self._cs_table, grid=grid0,
                                      rows_per_sample=s_loc)
                elif W.nq_w is not None:
                    for b in range(B):
                        rows = slice(b * s_loc, (b + 1) * s_loc)
                        ops.rmsnorm_rope_(q2[rows], W.nq_w, eps, cs_table=self._cs_table, grid=grids[b],
                                          rows_per_sample=s_loc)
                        ops.rmsnorm_rope_(k2[rows], W.nk_w, eps, cs_table=self._cs_table, grid=grids[b],
                                          rows_per_sample=s_loc)
                else:
                    raise NotImplementedError("qk_norm=False is not on the Omni-Video hot path")
                q4 = qkv.view(B, s_loc, 3, h, 128)
                ops.attention(q4[:, :, 0], q4[:, :, 1], q4[:, :, 2], k_lens=seq_lens,
                              out=attn.view(B, s_loc, h, 128))
                ops.gemm(attn, W.o_w, W.o_b, epilogue=EPI_GATE_RESID_F32, out=xres, resid=xres,
                         gate=mb[:, 2], rows_per_sample=s_loc)
            else:
                import torch.distributed as dist
                dg, hg = d // sp, h // sp
                ops.gemm(hbuf, W.qkv_w, W.qkv_b, epilogue=EPI_BF16, out=qkv_send, n_split=sp)
                s4 = qkv_send.view(sp, M, 3, dg)
                rope = dict(cs_table=self._cs_table, grid=grid0, rows_per_sample=s_loc, token_offset=row0)
                if tr is None:
                    ops.rmsnorm_rope_(s4[:, :, 0], W.nq_w, eps, **rope)
                    ops.rmsnorm_rope_(s4[:, :, 1], W.nk_w, eps, **rope)
                    dist.all_to_all_single(qkv_recv, qkv_send, group=self.sp_group())
                else:
                    # each rank's head group goes from the normalising kernel straight into that
                    # rank's receive buffer; V is a plain scatter
                    ops.rmsnorm_rope_(s4[:, :, 0], W.nq_w, eps, dst_ptrs=tr.q_k_v_dst(0), dst_ld=3 * dg, **rope)
                    ops.rmsnorm_rope_(s4[:, :, 1], W.nk_w, eps, dst_ptrs=tr.q_k_v_dst(1), dst_ld=3 * dg, **rope)
                    ops.rmsnorm_rope_(s4[:, :, 2], W.nq_w, eps, dst_ptrs=tr.q_k_v_dst(2), dst_ld=3 * dg,
                                      copy_only=True)
                    tr.barrier()
                # [sp(src rank = token chunk), s_loc, 3, hg, 128] == [seq_len, 3, hg, 128]
                r5 = qkv_recv.view(1, sp * s_loc, 3, hg, 128)
                if tr is None:
                    ops.attention(r5[:, :, 0], r5[:, :, 1], r5[:, :, 2], k_lens=seq_lens,
                                  out=att_send.view(1, sp * s_loc, hg, 128))
                    dist.all_to_all_single(att_recv, att_send, group=self.sp_group())
                else:
                    # the epilogue stores every query row into the buffer of the rank that owns it
                    ops.attention(r5[:, :, 0], r5[:, :, 1], r5[:, :, 2], k_lens=seq_lens,
                                  out_ptrs=tr.att_dst, rows_per_out=s_loc, out_row_stride=dg)
                    tr.barrier()
                ops.gemm(att_recv, W.o_w, W.o_b, epilogue=EPI_GATE_RESID_F32, out=xres, resid=xres,
                         gate=mb[:, 2], rows_per_sample=s_loc, k_split=sp)
            # ---------------- cross attention (context K/V replicated on every rank: no collective)
            if W.n3_w is not None:
                ops.layernorm_affine(xres, W.n3_w, W.n3_b, eps, out=hbuf)
            else:
                hbuf.copy_(xres)
            ops.gemm(hbuf, W.cq_w, W.cq_b, epilogue=EPI_BF16, out=qc)
            ops.gemm(ctx, W.ckv_w, W.ckv_b, epilogue=EPI_BF16, out=ckv)
            ops.rmsnorm_rope_(qc, W.cnq_w, eps)
            ops.rmsnorm_rope_(ckv[:, :d], W.cnk_w, eps)
            c4 = ckv.view(B, Lc, 2, h, 128)
            ops.attention(qc.view(B, s_loc, h, 128), c4[:, :, 0], c4[:, :, 1],
                          out=attn.view(B, s_loc, h, 128))
            ops.gemm(attn, W.co_w, W.co_b, epilogue=EPI_RESID_F32, out=xres, resid=xres)
            # ---------------- feed forward
            ops.layernorm_modulate(xres, mb[:, 4], mb[:, 3], eps, rows_per_sample=s_loc, out=hbuf)
            ops.gemm(hbuf, W.f1_w, W.f1_b, epilogue=EPI_GELU_BF16, out=ffn_mid)
            ops.gemm(ffn_mid, W.f2_w, W.f2_b, epilogue=EPI_GATE_RESID_F32, out=xres, resid=xres,
                     gate=mb[:, 5], rows_per_sample=s_loc)

        # ---- head (fp32, :278-290) and unpatchify (:526-549)
        hm = hmod[0]  # [B, 2, d]: shift, scale
        hn = ops.layernorm_modulate(xres, hm[:, 1], hm[:, 0], eps, rows_per_sample=s_loc,
                                    out_dtype=f32)
        yout = ops.linear_f32(hn, P.head_w, P.head_b)            # [M, 4*out_dim] fp32
        if not finish:
            return yout, grids
        return self._finish(yout, grids, seq_len)

    @torch.no_grad()
    def _finish(self, yout: torch.Tensor, grids, seq_len: int) -> List[torch.Tensor]:
        """Gather the head output of all sequence-parallel ranks (the forward's one NCCL collective) and
        unpatchify (reference model.py:526-549)."""
        sp = self.sp_size()
        if sp > 1:
            import torch.distributed as dist
            M = yout.shape[0]
            full = torch.empty((sp, M, yout.shape[1]), device=yout.device, dtype=f32)
            dist.all_gather_into_tensor(full, yout.contiguous(), group=self.sp_group())
            yout = full.view(sp * M, -1)                         # B == 1: rows are in token order
        return [ops.unpatchify(yout[b * seq_len:(b + 1) * seq_len], self.m.out_dim, grids[b])
                for b in range(len(grids))]


class _GraphedForward:
    """One captured DiT forward: static input buffers, a CUDA graph of everything up to the head projection
    (no NCCL inside: under sequence parallelism the exchanges are peer stores + flag barriers), and the eager
    tail (`DiTEngine._finish`: gather of the head output, unpatchify)."""

    def __init__(self, eng: DiTEngine, x, t, context, seq_len, y):
        dev = x[0].device
        self.x = [u.detach().clone() for u in x]
        self.t = t.detach().to(dev).clone()
        self.ctx = [c.detach().to(dev).clone() for c in context]
        self.y = None if y is None else [u.detach().clone() for u in y]
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up: weight packing, workspaces, peer rendezvous
            for _ in range(2):
                eng._forward(self.x, self.t, self.ctx, seq_len, self.y)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        n0 = _lib_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.yout, self.grids = eng._forward(self.x, self.t, self.ctx, seq_len, self.y, finish=False)
        self.launches = _lib_launch_count() - n0   # kernels of this library inside one replay
        self.eng, self.seq_len = eng, seq_len

    def __call__(self, x, t, context, y):
        for dst, src in zip(self.x, x):
            dst.copy_(src, non_blocking=True)
        self.t.copy_(t, non_blocking=True)
        for dst, src in zip(self.ctx, context):
            dst.copy_(src, non_blocking=True)
        if y is not None:
            for dst, src in zip(self.y, y):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        ops.add_launch_count(self.launches)
        return self.eng._finish(self.yout, self.grids, self.seq_len)


def _lib_launch_count() -> int:
    from .. import _lib
    return _lib.launch_count()


# ------------------------------------------------------------------------------------------------
class WanModel(nn.Module):
    """Wan DiT backbone; constructor and `forward` signature as reference model.py:293-524."""

    ignore_for_config = ["patch_size", "cross_attn_norm", "qk_norm", "text_dim", "window_size"]
    _no_split_modules = ["WanAttentionBlock"]

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048,
                 ffn_dim=8192, freq_dim=256, text_dim=4096, out_dim=16, num_heads=16,
                 num_layers=32, window_size=(-1, -1), qk_norm=True, cross_attn_norm=True, eps=1e-6,
                 use_gradient_checkpointing=False, skip_init=False):
        super().__init__()
        assert model_type in ["t2v", "i2v", "ti2v"]
        self.config = dict(model_type=model_type, patch_size=tuple(patch_size), text_len=text_len,
                           in_dim=in_dim, dim=dim, ffn_dim=ffn_dim, freq_dim=freq_dim,
                           text_dim=text_dim, out_dim=out_dim, num_heads=num_heads,
                           num_layers=num_layers, window_size=tuple(window_size), qk_norm=qk_norm,
                           cross_attn_norm=cross_attn_norm, eps=eps)
        self.model_type = model_type
        self.use_gradient_checkpointing = use_gradient_checkpointing
        self.patch_size, self.text_len, self.in_dim, self.dim = tuple(patch_size), text_len, in_dim, dim
        self.ffn_dim, self.freq_dim, self.text_dim, self.out_dim = ffn_dim, freq_dim, text_dim, out_dim
        self.num_heads, self.num_layers, self.window_size = num_heads, num_layers, tuple(window_size)
        self.qk_norm, self.cross_attn_norm, self.eps = qk_norm, cross_attn_norm, eps
        assert dim % num_heads == 0
        if dim // num_heads != 128:
            raise NotImplementedError("the B200 attention kernel is built for head_dim 128 "
                                      "(both OmniVideo2 configurations)")
        if tuple(window_size) != (-1, -1):
            raise NotImplementedError("windowed attention is not on the Omni-Video hot path")

        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=patch_size, stride=patch_size)
        self.text_embedding = nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate="tanh"),
                                            nn.Linear(dim, dim))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
        self.blocks = nn.ModuleList([
            WanAttentionBlock(dim, ffn_dim, num_heads, window_size, qk_norm, cross_attn_norm, eps)
            for _ in range(num_layers)])
        self.head = Head(dim, out_dim, patch_size, eps)
        # the reference keeps a complex128 `freqs` attribute (model.py:403-411); ours is the same
        # table as fp32 (cos, sin), created lazily on the parameters' device by the engine
        self.freqs = None
        if not skip_init:
            self.init_weights()
        self._engine = DiTEngine(self)

    # same initialisation scheme as reference model.py:551-573
    def init_weights(self):
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.xavier_uniform_(mod.weight)
                if mod.bias is not None:
                    nn.init.zeros_(mod.bias)
        nn.init.xavier_uniform_(self.patch_embedding.weight.flatten(1))
        for seq in (self.text_embedding, self.time_embedding):
            for mod in seq.modules():
                if isinstance(mod, nn.Linear):
                    nn.init.normal_(mod.weight, std=0.02)
        nn.init.zeros_(self.head.head.weight)

    def forward(self, x, t, context, seq_len, y=None):
        """x: list of [C_in, F, H, W]; t: [B]; context: list of [L_i, text_dim]; returns a list of
        fp32 [C_out, F, H, W] (reference model.py:417-524)."""
        if any("forward" in vars(b.self_attn) or "forward" in vars(b.cross_attn) or
               "forward" in vars(b) for b in self.blocks):
            # someone rebound block-level forwards on this instance (the reference's patching
            # idiom, x2x_gen_unified.py:176-182): run the module-granular assembly so they apply
            from . import blocks_api
            return blocks_api.dit_forward_modular(self, x, t, context, seq_len, y,
                                                  sp_group=getattr(self, "sp_group", None))
        return self._engine.forward(x, t, context, seq_len, y)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=None, **kwargs):
        """diffusers `ModelMixin.from_pretrained` for a local directory (config.json + safetensors / .bin),
        the call the reference makes at unified_model.py:43."""
        from ..utils.checkpoint import wan_from_pretrained
        return wan_from_pretrained(pretrained_model_name_or_path, subfolder=subfolder, torch_dtype=torch_dtype,
                                   **kwargs)

    def save_pretrained(self, save_directory, max_shard_size=5 << 30, safe_serialization=True):
        from ..utils.checkpoint import save_wan_pretrained
        save_wan_pretrained(self, save_directory, max_shard_bytes=max_shard_size,
                            safe_serialization=safe_serialization)

    def unpatchify(self, x, grid_sizes):
        """reference model.py:526-549; x: list/batch of [L, out_dim*4] fp32."""
        grids = grid_sizes.tolist() if torch.is_tensor(grid_sizes) else grid_sizes
        return [ops.unpatchify(u.float().contiguous(), self.out_dim, g) for u, g in zip(x, grids)]

    def enable_gradient_checkpointing(self):
        self.use_gradient_checkpointing = True

    def disable_gradient_checkpointing(self):
        self.use_gradient_checkpointing = False
