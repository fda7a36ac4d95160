"""ORACLE (test infrastructure only — never imported by the product path).

CPU fp32 restatement of the umT5 ENCODER the reference uses for text conditioning
(omnivideo/modules/t5.py: T5Encoder :266-311 with T5SelfAttention :141-171, T5Attention :67-117,
T5FeedForward :120-138, T5LayerNorm :47-64, T5RelativeEmbedding :223-263; `T5EncoderModel.__call__`
:505-512 trims each sample to its number of valid tokens).  SURVEY.md §8(f) #4.  Pinned to outputs of the
reference itself: tests/golden/t5_*.npz (oracle/gen_golden_t5.py).

umT5-XXL: vocab 256384, dim 4096, dim_attn 4096, dim_ffn 10240, 64 heads of 64, 24 layers, 32 buckets,
`shared_pos=False` (every layer owns its relative-position table), no attention scaling, gated-GELU FFN,
no biases anywhere, dropout inactive in eval.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch


def t5_layer_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """:55-64 — RMS norm without mean subtraction."""
    x = x.float()
    return weight * (x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps))


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    """:38-44"""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def relative_position_bucket(lq: int, lk: int, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """:241-263, bidirectional: [lq, lk] int64 bucket of (key position - query position)."""
    rel = torch.arange(lk).unsqueeze(0) - torch.arange(lq).unsqueeze(1)
    half = num_buckets // 2
    bucket = (rel > 0).long() * half
    rel = rel.abs()
    max_exact = half // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_dist / max_exact) *
                         (half - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, half - 1))
    return bucket + torch.where(rel < max_exact, rel, large)


def t5_attention(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, mask: Optional[torch.Tensor],
                 pos_bias: torch.Tensor, num_heads: int) -> torch.Tensor:
    """:83-117 for one sample.  x [L, C]; mask [L] (0 = padding key) or None; pos_bias [N, L, L]."""
    L = x.shape[0]
    q = (x @ sd[p + "q.weight"].T).view(L, num_heads, -1)
    k = (x @ sd[p + "k.weight"].T).view(L, num_heads, -1)
    v = (x @ sd[p + "v.weight"].T).view(L, num_heads, -1)
    bias = pos_bias.clone()
    if mask is not None:
        bias = bias.masked_fill(mask.view(1, 1, L) == 0, torch.finfo(torch.float32).min)
    s = torch.einsum("inc,jnc->nij", q, k) + bias          # T5 does not scale the scores
    a = torch.softmax(s, dim=-1)
    out = torch.einsum("nij,jnc->inc", a, v).reshape(L, -1)
    return out @ sd[p + "o.weight"].T


def t5_ffn(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor) -> torch.Tensor:
    """:133-138: fc2( fc1(x) * GELU(gate(x)) )"""
    return ((x @ sd[p + "fc1.weight"].T) * gelu_tanh(x @ sd[p + "gate.0.weight"].T)) @ sd[p + "fc2.weight"].T


def t5_encode(sd: Dict[str, torch.Tensor], cfg: dict, ids: torch.Tensor, mask: Optional[torch.Tensor]) -> List[torch.Tensor]:
    """T5Encoder.forward :299-311 + the trimming of T5EncoderModel.__call__ :510-512.
    ids [B, L] int64, mask [B, L] (1 = token) or None -> list of fp32 [n_valid_b, dim]."""
    sd = {k: v.float() for k, v in sd.items()}
    B, L = ids.shape
    buckets = relative_position_bucket(L, L, cfg["num_buckets"])
    outs = []
    for b in range(B):
        x = sd["token_embedding.weight"][ids[b]]
        m = None if mask is None else mask[b]
        for i in range(cfg["num_layers"]):
            p = f"blocks.{i}."
            table = sd["pos_embedding.embedding.weight"] if cfg.get("shared_pos") else sd[p + "pos_embedding.embedding.weight"]
            pos_bias = table[buckets].permute(2, 0, 1)                         # [N, L, L]
            x = x + t5_attention(sd, p + "attn.", t5_layer_norm(x, sd[p + "norm1.weight"]), m, pos_bias,
                                 cfg["num_heads"])
            x = x + t5_ffn(sd, p + "ffn.", t5_layer_norm(x, sd[p + "norm2.weight"]))
        x = t5_layer_norm(x, sd["norm.weight"])
        n = L if m is None else int((m > 0).sum())
        outs.append(x[:n])
    return outs
