"""ORACLE — test infrastructure only.  Not imported by the product (omni_video_b200/*).

CPU restatement of the reference's DiT hot path (SAIS-FUXI/Omni-Video, omnivideo/modules/model.py and
omnivideo/modules/attention.py) as plain functions over a state dict with the reference's
parameter names.  Every function cites the reference lines it follows.

Pinning: the reference ships no tests, golden vectors or fixtures for this path,
so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container from
/root/reference by `oracle/gen_golden.py`; the resulting vectors are committed under tests/golden/
and `tests/test_oracle_golden.py` holds the oracle to them (fp32, max-abs <= 2e-4 relative to the
output scale).

Two arithmetic modes:
  emulate_bf16=False  everything fp32 (fp64 RoPE / sinusoid) — the reference's CPU plumbing run
                      (BASELINE.json configs[0]).
  emulate_bf16=True   rounds to bf16 exactly where the reference's CUDA bf16-autocast execution
                      produces bf16 tensors (SURVEY.md §5 "Precision regime"); fp32 in between.  This is
                      what the sm_100a path is compared against on the GPU.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

f32 = torch.float32
f64 = torch.float64


def _r(x: torch.Tensor, emulate: bool) -> torch.Tensor:
    """round-trip through bf16 where the reference holds a bf16 tensor"""
    return x.to(torch.bfloat16).to(f32) if emulate else x


# ---------------------------------------------------------------- embeddings / RoPE
def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    """model.py:13-23 — fp64, cos first then sin, f_i = 10000^(-i/half)."""
    half = dim // 2
    pos = position.to(f64)
    freqs = torch.pow(torch.tensor(10000.0, dtype=f64), -torch.arange(half, dtype=f64) / half)
    ang = torch.outer(pos, freqs)
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)


def rope_angles(max_len: int, head_dim: int, theta: float = 10000.0) -> torch.Tensor:
    """model.py:26-34 + :405-411 — angles [max_len, head_dim/2] of the three concatenated spectra
    (sub-dims d-4*(d//6), 2*(d//6), 2*(d//6))."""
    d = head_dim
    parts = []
    for dsub in (d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)):
        inv = 1.0 / torch.pow(torch.tensor(theta, dtype=f64), torch.arange(0, dsub, 2, dtype=f64) / dsub)
        parts.append(torch.outer(torch.arange(max_len, dtype=f64), inv))
    return torch.cat(parts, dim=1)


def rope_apply(x: torch.Tensor, grid: Sequence[int], angles: torch.Tensor, token_offset: int = 0) -> torch.Tensor:
    """model.py:37-65 (and the sequence-parallel variant full_sequence_parallel.py:22-60 through
    `token_offset`): x [L, H, D]; token n -> (f, h, w) in f-major order; pairs (x[2j], x[2j+1]);
    first c-2*(c//3) pairs rotate with f, next c//3 with h, last c//3 with w; tokens >= F*H*W pass
    through.  fp64 math, fp32 result."""
    L, H, D = x.shape
    c = D // 2
    Fg, Hg, Wg = grid
    n_f, n_h = c - 2 * (c // 3), c // 3
    tok = torch.arange(L) + token_offset
    valid = tok < Fg * Hg * Wg
    tc = torch.where(valid, tok, torch.zeros_like(tok))
    fi, hi, wi = tc // (Hg * Wg), (tc // Wg) % Hg, tc % Wg
    ang = torch.cat([angles[fi, :n_f], angles[hi, n_f:n_f + n_h], angles[wi, n_f + n_h:]], dim=1)  # [L, c]
    ang = torch.where(valid[:, None], ang, torch.zeros_like(ang))
    cos, sin = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
    xe, xo = x[..., 0::2].to(f64), x[..., 1::2].to(f64)
    out = torch.stack([xe * cos - xo * sin, xe * sin + xo * cos], dim=-1).flatten(-2)
    return out.to(f32)


# ---------------------------------------------------------------- norms
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float, emulate: bool) -> torch.Tensor:
    """model.py:68-84 — over the FULL last dim; cast back to the input dtype BEFORE the weight."""
    y = x.float() * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    return _r(_r(y, emulate) * weight, emulate)


def layer_norm(x: torch.Tensor, eps: float, weight=None, bias=None) -> torch.Tensor:
    """model.py:87-97 — fp32 LayerNorm, optional affine."""
    return F.layer_norm(x.float(), (x.shape[-1],), weight, bias, eps)


def linear(x, w, b, emulate: bool):
    """nn.Linear under bf16 autocast: bf16 operands, fp32 accumulate, bf16 result."""
    return _r(F.linear(_r(x, emulate), w, b), emulate)


# ---------------------------------------------------------------- attention
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, k_len: Optional[int],
              emulate: bool, softmax_scale: Optional[float] = None) -> torch.Tensor:
    """attention.py:24-133 for one sample: q [Lq, H, D], k/v [Lk, H, D]; keys >= k_len are dropped
    (:75-77), queries never are; scale 1/sqrt(D); non-causal.  In emulate mode the (unnormalised)
    probabilities are rounded to bf16 before the PV product and the output to bf16, as flash-attn
    does; statistics stay fp32."""
    Lq, H, D = q.shape
    if k_len is not None:
        k, v = k[:k_len], v[:k_len]
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    qh, kh, vh = (_r(t, emulate).permute(1, 0, 2) for t in (q, k, v))  # [H, L, D]
    out = torch.empty((H, Lq, D), dtype=f32)
    step = max(1, (1 << 26) // max(1, kh.shape[1]))  # bound the score matrix to 256 MB
    for h in range(H):
        for i0 in range(0, Lq, step):
            s = (qh[h, i0:i0 + step] @ kh[h].T) * scale
            m = s.max(dim=-1, keepdim=True).values
            p = torch.exp(s - m)
            l = p.sum(dim=-1, keepdim=True)
            out[h, i0:i0 + step] = (_r(p, emulate) @ vh[h]) / l
    return _r(out.permute(1, 0, 2).contiguous(), emulate)


# ---------------------------------------------------------------- blocks
def block_forward(sd: Dict[str, torch.Tensor], pre: str, x: torch.Tensor, e0: torch.Tensor,
                  seq_len_valid: int, grid, angles, context: torch.Tensor, num_heads: int,
                  eps: float, emulate: bool, first_block: bool, token_offset: int = 0,
                  kv_gather=None) -> torch.Tensor:
    """model.py:218-258 for one sample.  x [L, d]; e0 [6, d] fp32 (or [L, 6, d]); context [Lc, d].
    kv_gather: optional hook (q, k, v) -> (q, k, v, finish) used by the Ulysses tests."""
    L, d = x.shape
    H = num_heads
    mod = sd[pre + "modulation"].float().reshape(6, d) + e0          # :238 (fp32)
    m = [mod[..., i, :] for i in range(6)]

    def g(name):
        return sd[pre + name]

    # ---- self attention :241-246
    ln = layer_norm(x, eps)
    if first_block:
        ln = _r(ln, emulate)                                          # .type_as(x) with bf16 x (:97)
    h = ln * (1 + m[1]) + m[0]
    q = rms_norm(linear(h, g("self_attn.q.weight"), g("self_attn.q.bias"), emulate),
                 g("self_attn.norm_q.weight"), eps, emulate)
    k = rms_norm(linear(h, g("self_attn.k.weight"), g("self_attn.k.bias"), emulate),
                 g("self_attn.norm_k.weight"), eps, emulate)
    v = linear(h, g("self_attn.v.weight"), g("self_attn.v.bias"), emulate)
    q = rope_apply(q.view(L, H, -1), grid, angles, token_offset)
    k = rope_apply(k.view(L, H, -1), grid, angles, token_offset)
    v = v.view(L, H, -1)
    if kv_gather is None:
        a = attention(q, k, v, seq_len_valid, emulate)
    else:
        a = kv_gather(q, k, v, seq_len_valid)
    y = linear(a.reshape(L, d), g("self_attn.o.weight"), g("self_attn.o.bias"), emulate)
    x = x.float() + y * m[2]

    # ---- cross attention :250 (context_lens=None -> all padded rows attended, :479)
    h = layer_norm(x, eps, g("norm3.weight").float(), g("norm3.bias").float()) \
        if (pre + "norm3.weight") in sd else x
    qc = rms_norm(linear(h, g("cross_attn.q.weight"), g("cross_attn.q.bias"), emulate),
                  g("cross_attn.norm_q.weight"), eps, emulate)
    kc = rms_norm(linear(context, g("cross_attn.k.weight"), g("cross_attn.k.bias"), emulate),
                  g("cross_attn.norm_k.weight"), eps, emulate)
    vc = linear(context, g("cross_attn.v.weight"), g("cross_attn.v.bias"), emulate)
    a = attention(qc.view(L, H, -1), kc.view(-1, H, d // H), vc.view(-1, H, d // H), None, emulate)
    x = x + linear(a.reshape(L, d), g("cross_attn.o.weight"), g("cross_attn.o.bias"), emulate)

    # ---- ffn :251-254
    h = layer_norm(x, eps) * (1 + m[4]) + m[3]
    mid = linear(h, g("ffn.0.weight"), g("ffn.0.bias"), emulate)
    mid = _r(F.gelu(mid, approximate="tanh"), emulate)
    y = linear(mid, g("ffn.2.weight"), g("ffn.2.bias"), emulate)
    return x + y * m[5]


def patchify(u: torch.Tensor) -> torch.Tensor:
    """[C, F, H, W] -> [F*(H/2)*(W/2), C*4] with column order of patch_embedding.weight.flatten(1)
    (Conv3d k=s=(1,2,2) as a GEMM; model.py:384, :455-458)."""
    C, Fr, H, W = u.shape
    t = u.reshape(C, Fr, H // 2, 2, W // 2, 2).permute(1, 2, 4, 0, 3, 5)  # f hp wp c ph pw
    return t.reshape(Fr * (H // 2) * (W // 2), C * 4)


def unpatchify(y: torch.Tensor, grid, out_dim: int) -> torch.Tensor:
    """model.py:526-549."""
    Fg, Hg, Wg = grid
    u = y[:Fg * Hg * Wg].view(Fg, Hg, Wg, 1, 2, 2, out_dim)
    u = torch.einsum("fhwpqrc->cfphqwr", u)
    return u.reshape(out_dim, Fg, Hg * 2, Wg * 2)


def dit_forward(sd: Dict[str, torch.Tensor], cfg: dict, x: List[torch.Tensor], t: torch.Tensor,
                context: List[torch.Tensor], seq_len: int, emulate_bf16: bool = False,
                num_layers: Optional[int] = None) -> List[torch.Tensor]:
    """model.py:417-524.  sd: state dict (reference names; in emulate mode pass bf16-rounded
    values); returns a list of fp32 [C_out, F, H, W]."""
    em = emulate_bf16
    sd = {k: v.float() for k, v in sd.items()}
    d, H, eps = cfg["dim"], cfg["num_heads"], cfg.get("eps", 1e-6)
    nl = cfg["num_layers"] if num_layers is None else num_layers
    angles = rope_angles(1024, d // H)
    outs = []
    for b, u in enumerate(x):
        grid = (u.shape[1], u.shape[2] // 2, u.shape[3] // 2)
        L = grid[0] * grid[1] * grid[2]
        assert L <= seq_len
        tok = patchify(u.float())
        emb = linear(tok, sd["patch_embedding.weight"].flatten(1), sd["patch_embedding.bias"], em)
        xs = torch.cat([emb, emb.new_zeros(seq_len - L, d)])                      # :461-464
        # time :466-476 (fp32 math on the stored weights)
        per_token = t.dim() == 2                       # t [B, seq_len]: every token its own timestep (:466-467)
        tb = t[b].reshape(-1) if per_token else t[b].reshape(1)
        sin = sinusoidal_embedding_1d(cfg["freq_dim"], tb).float()
        e = F.linear(F.silu(F.linear(sin, sd["time_embedding.0.weight"], sd["time_embedding.0.bias"])),
                     sd["time_embedding.2.weight"], sd["time_embedding.2.bias"])   # [1 | seq_len, d]
        e0 = F.linear(F.silu(e), sd["time_projection.1.weight"], sd["time_projection.1.bias"])
        e0 = e0.view(-1, 6, d) if per_token else e0.view(6, d)
        # context :478-485
        c = context[b].float()
        c = torch.cat([c, c.new_zeros(cfg["text_len"] - c.shape[0], c.shape[1])])
        c = linear(c, sd["text_embedding.0.weight"], sd["text_embedding.0.bias"], em)
        c = _r(F.gelu(c, approximate="tanh"), em)
        c = linear(c, sd["text_embedding.2.weight"], sd["text_embedding.2.bias"], em)
        for li in range(nl):
            xs = block_forward(sd, f"blocks.{li}.", xs, e0, L, grid, angles, c, H, eps, em, li == 0)
        # head :278-290 (fp32)
        hm = sd["head.modulation"].reshape(2, d) + e.unsqueeze(-2)                 # [1 | seq_len, 2, d]: shift, scale
        hn = layer_norm(xs, eps) * (1 + hm[:, 1]) + hm[:, 0]
        y = F.linear(hn, sd["head.head.weight"], sd["head.head.bias"])
        outs.append(unpatchify(y, grid, cfg["out_dim"]).float())
    return outs


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
