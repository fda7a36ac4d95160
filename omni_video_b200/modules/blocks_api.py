"""Module-granular entry points with the reference's signatures, for callers that rebind
`block.self_attn.forward` / `block.cross_attn.forward` themselves the way the reference pipeline
does (x2x_gen_unified.py:176-182).  They run the same kernels as DiTEngine, but on `[B, L, C]`
tensors handed over by PyTorch code, so they cannot keep the fused buffer layouts of the engine;
`WanModel.forward` never goes through here.
"""
from __future__ import annotations

import torch

from .. import ops
from ..ops import EPI_BF16, EPI_GELU_BF16

bf16 = torch.bfloat16
f32 = torch.float32

_cs_cache = {}


def _cs_table(device):
    key = str(device)
    if key not in _cs_cache:
        _cs_cache[key] = ops.rope_cos_sin_table(1024, 128).to(device)
    return _cs_cache[key]


def _w(p):
    p = p.detach()
    return (p if p.dtype == bf16 else p.to(bf16)).contiguous()


def _rows_bf16(x):
    x2 = x.reshape(-1, x.shape[-1])
    return (x2 if x2.dtype == bf16 else x2.to(bf16)).contiguous()


def _linear(lin, x2, epilogue=EPI_BF16):
    return ops.gemm(x2, _w(lin.weight), None if lin.bias is None else _w(lin.bias), epilogue=epilogue)


def _grid_list(grid_sizes):
    g = grid_sizes.tolist() if torch.is_tensor(grid_sizes) else list(grid_sizes)
    assert all(tuple(u) == tuple(g[0]) for u in g), "batched samples must share one latent grid"
    return tuple(int(v) for v in g[0])


def qkv_project(sa, x, grid_sizes, token_offset=0):
    """q, k (RMSNorm + RoPE applied) and v as bf16 [B, L, H, 128] (reference model.py:136-147)."""
    b, s = x.shape[:2]
    x2 = _rows_bf16(x)
    q, k, v = _linear(sa.q, x2), _linear(sa.k, x2), _linear(sa.v, x2)
    cs = _cs_table(x.device)
    grid = _grid_list(grid_sizes)
    if sa.qk_norm:
        ops.rmsnorm_rope_(q, _w(sa.norm_q.weight), sa.eps, cs_table=cs, grid=grid,
                          rows_per_sample=s, token_offset=token_offset)
        ops.rmsnorm_rope_(k, _w(sa.norm_k.weight), sa.eps, cs_table=cs, grid=grid,
                          rows_per_sample=s, token_offset=token_offset)
    else:
        raise NotImplementedError("qk_norm=False is not on the Omni-Video hot path")
    n = sa.num_heads
    return q.view(b, s, n, 128), k.view(b, s, n, 128), v.view(b, s, n, 128)


def self_attention(sa, x, seq_lens, grid_sizes):
    b, s = x.shape[:2]
    q, k, v = qkv_project(sa, x, grid_sizes)
    kl = seq_lens.tolist() if torch.is_tensor(seq_lens) else seq_lens
    a = ops.attention(q, k, v, k_lens=kl)
    return _linear(sa.o, a.view(b * s, -1)).view(b, s, -1)


def cross_qkv(ca, x, context):
    b, s = x.shape[:2]
    n = ca.num_heads
    q = _linear(ca.q, _rows_bf16(x))
    c2 = _rows_bf16(context)
    k, v = _linear(ca.k, c2), _linear(ca.v, c2)
    if ca.qk_norm:
        ops.rmsnorm_rope_(q, _w(ca.norm_q.weight), ca.eps)
        ops.rmsnorm_rope_(k, _w(ca.norm_k.weight), ca.eps)
    return q.view(b, s, n, 128), k.view(b, -1, n, 128), v.view(b, -1, n, 128)


def cross_attention(ca, x, context, context_lens):
    b, s = x.shape[:2]
    q, k, v = cross_qkv(ca, x, context)
    kl = None
    if context_lens is not None:
        kl = context_lens.tolist() if torch.is_tensor(context_lens) else context_lens
    a = ops.attention(q, k, v, k_lens=kl)
    return _linear(ca.o, a.view(b * s, -1)).view(b, s, -1)


def _mod6(blk, e):
    """e: fp32 [B, L1, 6, C] (L1 = 1 or L) -> 6 fp32 tensors [B*L1, C] (reference :236-239)."""
    assert e.dtype == f32
    b, l1 = e.shape[:2]
    tbl = (blk.modulation.detach().float().unsqueeze(0) + e).reshape(b * l1, 6, -1)
    return [tbl[:, i] for i in range(6)], l1


def attention_block(blk, x, e, seq_lens, grid_sizes, context, context_lens):
    """reference model.py:218-258 with [B, L, C] in / fp32 [B, L, C] out."""
    b, s, c = x.shape
    mods, l1 = _mod6(blk, e)
    rps = s if l1 == 1 else 0  # per-sample vectors vs per-token table
    x2 = x.reshape(b * s, c)
    first_bf16 = x2.dtype == bf16
    xf = x2.float().contiguous() if x2.dtype != f32 else x2.contiguous()
    h = ops.layernorm_modulate(xf, mods[1], mods[0], blk.eps, rows_per_sample=rps,
                               round_ln_bf16=first_bf16)
    y = blk.self_attn(h.view(b, s, c), seq_lens, grid_sizes, None).reshape(b * s, c)
    xf = _gated_add(xf, y, mods[2], rps)
    h = blk.norm3(xf.view(b, s, c)) if blk.cross_attn_norm else xf.view(b, s, c)
    y = blk.cross_attn(h, context, context_lens).reshape(b * s, c)
    xf = xf + y.float()
    h = ops.layernorm_modulate(xf, mods[4], mods[3], blk.eps, rows_per_sample=rps)
    mid = _linear(blk.ffn[0], h, EPI_GELU_BF16)
    y = _linear(blk.ffn[2], mid)
    xf = _gated_add(xf, y, mods[5], rps)
    return xf.view(b, s, c)


def _gated_add(xf, y, gate, rps):
    if rps:
        g = gate.repeat_interleave(rps, dim=0) if gate.shape[0] != xf.shape[0] else gate
    else:
        g = gate
    return xf + y.float() * g


def head(hd, x, e):
    """reference model.py:278-290: x [B, L, C], e fp32 [B, L1, C] -> fp32 [B, L, out*4]."""
    assert e.dtype == f32
    b, s, c = x.shape
    l1 = e.shape[1]
    tbl = (hd.modulation.detach().float().unsqueeze(0) + e.unsqueeze(2)).reshape(b * l1, 2, c)
    xf = x.reshape(b * s, c)
    xf = xf.float().contiguous() if xf.dtype != f32 else xf.contiguous()
    hn = ops.layernorm_modulate(xf, tbl[:, 1], tbl[:, 0], hd.eps,
                                rows_per_sample=s if l1 == 1 else 0, out_dtype=f32)
    y = ops.linear_f32(hn, _w(hd.head.weight), _w(hd.head.bias))
    return y.view(b, s, -1)


def dit_forward_modular(model, x, t, context, seq_len, y=None, sp_group=None):
    """WanModel.forward assembled the way the reference does it — embed, loop over
    `block(x, **kwargs)`, `head`, unpatchify (model.py:444-524) — so that per-module rebinding
    (`block.self_attn.forward = MethodType(...)`, x2x_gen_unified.py:176-182) takes effect.
    With `sp_group` it follows sp_dit_forward (full_sequence_parallel.py:63-167): rows and the
    context are chunked across the group, the head output is all-gathered."""
    from ..distributed import util as dutil
    if y is not None:
        x = [torch.cat([u, v], dim=0) for u, v in zip(x, y)]
    eng = model._engine
    B, d = len(x), model.dim
    xres, e, e0, ctx, grids, seq_lens = eng.embed(x, t, context, seq_len, 1, 0)
    # the embedding is bf16 in the reference (autocast), and block 0's LayerNorm result is rounded back to
    # it (`.type_as(x)`, model.py:97): hand block 0 a bf16 tensor (exact: the values are bf16 numbers) so
    # that attention_block rounds like DiTEngine and the reference do; later blocks get fp32 (model.py:245)
    xs = xres.view(B, seq_len, d).to(bf16)
    l1 = e.shape[0] // B                       # 1, or seq_len with per-token timesteps
    ev = e.view(B, l1, d)
    e0v = e0.view(B, l1, 6, d)
    ctxv = ctx.view(B, model.text_len, d)
    sp = dutil.get_world_size(sp_group) if sp_group is not None else 1
    if sp > 1:
        # rows and context are sharded below: BOTH attention forwards of every block must be the
        # sequence-parallel variants, or the one left local would silently attend within a 1/sp shard
        from ..distributed.full_sequence_parallel import sp_attn_forward, sp_cross_attn_forward
        for i, blk in enumerate(model.blocks):
            for mod, want in ((blk.self_attn, sp_attn_forward), (blk.cross_attn, sp_cross_attn_forward)):
                if getattr(mod.forward, "__func__", None) is not want:
                    raise RuntimeError(
                        f"dit_forward_modular with a sequence-parallel group: blocks[{i}].{type(mod).__name__}."
                        f"forward is not {want.__name__}; rebind both attention forwards "
                        "(x2x_gen_unified.py:176-182) or use the engine path (model.sp_group)")
        r = dutil.get_rank(sp_group)
        xs = torch.chunk(xs, sp, dim=1)[r]
        ctxv = torch.chunk(ctxv, sp, dim=1)[r]
        if l1 > 1:
            ev, e0v = torch.chunk(ev, sp, dim=1)[r], torch.chunk(e0v, sp, dim=1)[r]
    kwargs = dict(e=e0v, seq_lens=torch.tensor(seq_lens, dtype=torch.long),
                  grid_sizes=torch.tensor(grids, dtype=torch.long), freqs=model.freqs,
                  context=ctxv, context_lens=None)
    for block in model.blocks:
        xs = block(xs, **kwargs)
    out = model.head(xs, ev)
    if sp > 1:
        out = dutil.gather_forward(out, dim=1, group=sp_group)
    return [ops.unpatchify(out[b].contiguous(), model.out_dim, grids[b]) for b in range(B)]
