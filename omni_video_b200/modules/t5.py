"""umT5 encoder on the B200 kernels (drop-in for the ENCODER half of reference omnivideo/modules/t5.py;
SURVEY.md §8(f) #4).  Composed from the kernels of the DiT/VAE path; parity on the GPU against the oracle pinned
to the reference's T5Encoder (tests/test_gpu_t5.py), host logic also on CPU with emulated kernels
(tests/test_t5_cpu.py).

The text encoder is 4.7 TFLOP per 512-token prompt and runs twice per sample — under 0.02 % of a 50-step
sample — so nothing here is tuned; the point is that the reference's `T5EncoderModel(...)(texts, device)` keeps
working without PyTorch arithmetic on the device.  Parameter names are the reference's (`token_embedding`,
`blocks.{i}.{norm1,attn.{q,k,v,o},norm2,ffn.{gate.0,fc1,fc2},pos_embedding.embedding}`, `norm`), so
`models_t5_umt5-xxl-enc-bf16.pth` loads unchanged.

    reference (t5.py)                                            kernel
    -----------------------------------------------------------  ------------------------------------------
    T5LayerNorm :55-64 (RMS, cast to bf16 before the weight)     ovb_rmsnorm_rope (rope = 0)
    q | k | v = x W^T (no bias) :96-98                           ovb_gemm_bf16 (fused [Wq; Wk; Wv], whole batch)
    scores + position bias + key mask, fp32 softmax, P V         ovb_attn_relbias_fwd: all heads and samples in ONE
      :101-112 (NO 1/sqrt(d))                                    launch, per-sample key counts read on the device
    o-projection + residual :113-117, :169                       ovb_gemm_bf16 RESID_BF16
    fc1(x) * GELU_tanh(gate(x)) :133-138                         ovb_gemm_bf16 GELU_BF16 -> fp32 gate table ->
                                                                 ovb_gemm_bf16 GATE_RESID_F32 on a zero residual
    fc2 + residual :170                                          ovb_gemm_bf16 RESID_BF16

A block is 10 launches whatever the head count (round 1 looped over the heads: 4 launches x 64 heads x 24 blocks).
The relative-position bias (t5.py:229-263) depends on (key - query) only: per block it is an fp32 [heads, 2L-1]
vector table, gathered once per sequence length from the bucket embedding (integer arithmetic on positions, on the
host, like the reference does) and cached next to the packed weights.
"""
from __future__ import annotations

import logging
import math
from typing import List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..ops import EPI_BF16, EPI_GATE_RESID_F32, EPI_GELU_BF16, EPI_RESID_BF16

__all__ = ["T5Encoder", "T5EncoderModel", "umt5_xxl"]

bf16 = torch.bfloat16
f32 = torch.float32


class T5LayerNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))


class T5Attention(nn.Module):
    def __init__(self, dim, dim_attn, num_heads, dropout=0.1):
        assert dim_attn % num_heads == 0
        super().__init__()
        self.dim, self.dim_attn, self.num_heads, self.head_dim = dim, dim_attn, num_heads, dim_attn // num_heads
        self.q = nn.Linear(dim, dim_attn, bias=False)
        self.k = nn.Linear(dim, dim_attn, bias=False)
        self.v = nn.Linear(dim, dim_attn, bias=False)
        self.o = nn.Linear(dim_attn, dim, bias=False)


class T5FeedForward(nn.Module):
    def __init__(self, dim, dim_ffn, dropout=0.1):
        super().__init__()
        self.dim, self.dim_ffn = dim, dim_ffn
        self.gate = nn.Sequential(nn.Linear(dim, dim_ffn, bias=False), nn.GELU(approximate="tanh"))
        self.fc1 = nn.Linear(dim, dim_ffn, bias=False)
        self.fc2 = nn.Linear(dim_ffn, dim, bias=False)


class T5RelativeEmbedding(nn.Module):
    def __init__(self, num_buckets, num_heads, bidirectional, max_dist=128):
        super().__init__()
        self.num_buckets, self.num_heads, self.bidirectional, self.max_dist = num_buckets, num_heads, bidirectional, max_dist
        self.embedding = nn.Embedding(num_buckets, num_heads)

    def buckets(self, lq: int, lk: int) -> torch.Tensor:
        """[lq, lk] int64 bucket index of (key position - query position); reference :241-263."""
        if not self.bidirectional:
            raise NotImplementedError("only the encoder's bidirectional table is on the path")
        rel = torch.arange(lk).unsqueeze(0) - torch.arange(lq).unsqueeze(1)
        half = self.num_buckets // 2
        out = (rel > 0).long() * half
        rel = rel.abs()
        exact = half // 2
        far = exact + (torch.log(rel.float() / exact) / math.log(self.max_dist / exact) * (half - exact)).long()
        far = torch.minimum(far, torch.full_like(far, half - 1))
        return out + torch.where(rel < exact, rel, far)

    def relative_buckets(self, L: int) -> torch.Tensor:
        """[2L-1] int64: bucket of (key - query) = -(L-1) .. L-1, i.e. `buckets(L, L)[i, j]` at index (j - i) + L - 1."""
        b = self.buckets(L, L)
        return torch.cat([b[L - 1, :L - 1], b[0]])


class T5SelfAttention(nn.Module):
    def __init__(self, dim, dim_attn, dim_ffn, num_heads, num_buckets, shared_pos=True, dropout=0.1):
        super().__init__()
        self.shared_pos = shared_pos
        self.norm1 = T5LayerNorm(dim)
        self.attn = T5Attention(dim, dim_attn, num_heads, dropout)
        self.norm2 = T5LayerNorm(dim)
        self.ffn = T5FeedForward(dim, dim_ffn, dropout)
        self.pos_embedding = None if shared_pos else T5RelativeEmbedding(num_buckets, num_heads, bidirectional=True)


def _w(p: torch.Tensor) -> torch.Tensor:
    return p.detach().to(bf16).contiguous()


class T5Encoder(nn.Module):
    """reference t5.py:266-311; `forward(ids [B, L], mask [B, L] | None) -> bf16 [B, L, dim]`."""

    def __init__(self, vocab, dim, dim_attn, dim_ffn, num_heads, num_layers, num_buckets, shared_pos=True,
                 dropout=0.1):
        super().__init__()
        self.dim, self.dim_attn, self.dim_ffn = dim, dim_attn, dim_ffn
        self.num_heads, self.num_layers, self.num_buckets, self.shared_pos = num_heads, num_layers, num_buckets, shared_pos
        self.token_embedding = vocab if isinstance(vocab, nn.Embedding) else nn.Embedding(vocab, dim)
        self.pos_embedding = T5RelativeEmbedding(num_buckets, num_heads, bidirectional=True) if shared_pos else None
        self.blocks = nn.ModuleList([T5SelfAttention(dim, dim_attn, dim_ffn, num_heads, num_buckets, shared_pos, dropout)
                                     for _ in range(num_layers)])
        self.norm = T5LayerNorm(dim)
        self._packed = None
        self._bucket_cache = {}
        self._bias_cache = {}

    # packed bf16 weights (fused [Wq; Wk; Wv]): a derived cache, rebuilt when parameters move or change
    def _weights(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or self._packed[0] != key:
            layers = []
            for blk in self.blocks:
                a, f = blk.attn, blk.ffn
                layers.append(dict(n1=_w(blk.norm1.weight), qkv=_w(torch.cat([a.q.weight, a.k.weight, a.v.weight], 0)),
                                   o=_w(a.o.weight), n2=_w(blk.norm2.weight), gate=_w(f.gate[0].weight),
                                   fc1=_w(f.fc1.weight), fc2=_w(f.fc2.weight)))
            self._packed = (key, layers, _w(self.norm.weight))
            self._bias_cache = {}
        return self._packed[1], self._packed[2]

    def _rel_bias(self, L: int) -> List[torch.Tensor]:
        """Per block: fp32 [heads, 2L-1], the position bias of key j for query i at [h, (j - i) + L - 1]
        (reference t5.py:229-263).  Derived from the bucket embeddings: cached with the packed weights."""
        self._weights()
        tabs = self._bias_cache.get(L)
        if tabs is None:
            emb = self.pos_embedding or self.blocks[0].pos_embedding
            idx = self._bucket_cache.get(L)
            if idx is None:
                idx = self._bucket_cache[L] = emb.relative_buckets(L)

            def table(e):
                w = e.embedding.weight.detach()
                return w.float()[idx.to(w.device)].t().contiguous()

            if self.pos_embedding is not None:
                tabs = [table(self.pos_embedding)] * len(self.blocks)
            else:
                tabs = [table(blk.pos_embedding) for blk in self.blocks]
            self._bias_cache = {L: tabs}
        return tabs

    @torch.no_grad()
    def _encode(self, ids: torch.Tensor, k_lens: Optional[torch.Tensor]) -> torch.Tensor:
        """ids [B, L] int64, k_lens int32 [B] on the same device (valid tokens per sample) or None -> bf16 [B, L, dim]."""
        layers, norm_w = self._weights()
        B, L = ids.shape
        d, da, nh, eps = self.dim, self.dim_attn, self.num_heads, 1e-6
        hd = da // nh
        dev = ids.device
        if hd not in (64, 128):
            raise NotImplementedError(f"head_dim {hd}: ovb_attn_relbias_fwd has 64- and 128-wide heads (umT5-XXL: 64)")
        rel_bias = self._rel_bias(L)
        M = B * L
        x = self.token_embedding.weight.detach()[ids.reshape(-1)].to(bf16).contiguous()          # [B*L, d]
        h = torch.empty_like(x)
        qkv = torch.empty((M, 3 * da), device=dev, dtype=bf16)
        att = torch.empty((M, da), device=dev, dtype=bf16)
        zero = torch.zeros((M, self.dim_ffn), device=dev, dtype=f32)
        u = torch.empty_like(zero)
        q, k, v = (qkv[:, i * da:(i + 1) * da].unflatten(1, (nh, hd)).unflatten(0, (B, L)) for i in range(3))
        for W, bias in zip(layers, rel_bias):
            # ---- self attention (no 1/sqrt(d) scaling in T5)
            ops.rmsnorm_rope_(x, W["n1"], eps, out=h)
            ops.gemm(h, W["qkv"], None, epilogue=EPI_BF16, out=qkv)
            ops.attention_relbias(q, k, v, bias, k_lens, 1.0, out=att.view(B, L, nh, hd))
            ops.gemm(att, W["o"], None, epilogue=EPI_RESID_BF16, out=x, resid=x)
            # ---- gated-GELU feed forward
            ops.rmsnorm_rope_(x, W["n2"], eps, out=h)
            g = ops.cast_bf16_to_f32(ops.gemm(h, W["gate"], None, epilogue=EPI_GELU_BF16))
            ops.gemm(h, W["fc1"], None, epilogue=EPI_GATE_RESID_F32, out=u, resid=zero, gate=g)
            ops.gemm(ops.cast_f32_to_bf16(u), W["fc2"], None, epilogue=EPI_RESID_BF16, out=x, resid=x)
        ops.rmsnorm_rope_(x, norm_w, eps)
        return x.view(B, L, d)

    def forward(self, ids, mask=None):
        if ids.device.type != "cuda":
            raise RuntimeError("T5Encoder (B200 path) needs CUDA tensors; there is no CPU fallback")
        # the tokenizer pads at the end (tokenizers.py:60-70): the mask is a prefix, its row sums are the key counts;
        # they stay on the device (no host synchronisation)
        k_lens = None if mask is None else mask.gt(0).sum(dim=1).clamp_(min=1).to(torch.int32).contiguous()
        return self._encode(ids, k_lens)


def clean_whitespace(text: str) -> str:
    """`whitespace_clean(basic_clean(text))` of reference tokenizers.py:12-22: ftfy mojibake repair (when ftfy is
    installed; it is a no-op on clean text), HTML entities unescaped twice, runs of whitespace collapsed."""
    import html
    import re
    try:
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    text = html.unescape(html.unescape(text)).strip()
    return re.sub(r"\s+", " ", text).strip()


def umt5_xxl(encoder_only=True, return_tokenizer=False, dtype=bf16, device="cuda", **kwargs):
    """reference t5.py:454-467 (encoder only: the decoder is not on the path)."""
    if not encoder_only or return_tokenizer:
        raise NotImplementedError("only `umt5_xxl(encoder_only=True, return_tokenizer=False)` is on the path")
    cfg = dict(vocab=256384, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=24, num_buckets=32,
               shared_pos=False, dropout=0.1)
    cfg.update(kwargs)
    with torch.device(device):
        model = T5Encoder(**cfg)
    return model.to(dtype=dtype, device=device)


class T5EncoderModel:
    """reference t5.py:471-512: `T5EncoderModel(text_len, dtype, device, checkpoint_path, tokenizer_path)(texts, device)`
    -> list of [n_tokens_i, 4096]."""

    def __init__(self, text_len, dtype=bf16, device="cuda", checkpoint_path=None, tokenizer_path=None, shard_fn=None,
                 tokenizer=None):
        self.text_len, self.dtype, self.device = text_len, dtype, device
        self.checkpoint_path, self.tokenizer_path = checkpoint_path, tokenizer_path
        model = umt5_xxl(encoder_only=True, return_tokenizer=False, dtype=dtype, device=device).eval().requires_grad_(False)
        if checkpoint_path is not None:
            logging.info("loading %s", checkpoint_path)
            model.load_state_dict(torch.load(checkpoint_path, map_location="cpu", weights_only=True))
        self.model = model
        if tokenizer is None:
            try:        # the reference's HuggingfaceTokenizer(name, seq_len, clean='whitespace'), tokenizers.py:37-82
                from transformers import AutoTokenizer
                hf = AutoTokenizer.from_pretrained(tokenizer_path)

                def tokenizer(texts, return_mask=True, add_special_tokens=True):
                    texts = [clean_whitespace(t) for t in ([texts] if isinstance(texts, str) else texts)]
                    enc = hf(texts, return_tensors="pt", padding="max_length", truncation=True,
                             max_length=text_len, add_special_tokens=add_special_tokens)
                    return enc.input_ids, enc.attention_mask
            except Exception as ex:  # noqa: BLE001
                raise RuntimeError(f"cannot build the umT5 tokenizer from {tokenizer_path!r}: {ex}") from ex
        self.tokenizer = tokenizer

    def __call__(self, texts: List[str], device) -> List[torch.Tensor]:
        ids, mask = self.tokenizer(texts, return_mask=True, add_special_tokens=True)
        ids, mask = ids.to(device), mask.to(device)
        seq_lens = mask.gt(0).sum(dim=1).long()
        context = self.model(ids, mask)
        return [u[:v] for u, v in zip(context, seq_lens)]
