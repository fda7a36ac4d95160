"""ORACLE (test infrastructure only — never imported by the product path).

CPU fp32 restatement of the conditioning front-end of the reference,
`UnifiedWanWithMixedConditionModel.forward` steps 2-4 (omnivideo/modules/unified_model.py:192-276) and
`VisualContextAdapter.forward` (omnivideo/modules/visual_context_adapter.py:77-115).  Pinned to
outputs of the reference itself: tests/golden/unified_*.npz (oracle/gen_golden_unified.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """WanRMSNorm, reference model.py:76-84 (fp32 here, so `.type_as` is the identity)."""
    x = x.float()
    return x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps) * weight


def layer_norm(x: torch.Tensor, eps: float) -> torch.Tensor:
    """nn.LayerNorm(elementwise_affine=False): biased variance over the last dim."""
    x = x.float()
    mu = x.mean(dim=-1, keepdim=True)
    var = (x - mu).pow(2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps)


def vlm_features(sd: Dict[str, torch.Tensor], tokens: torch.Tensor, eps: float) -> torch.Tensor:
    """unified_model.py:195-206: vlm_proj(vlm_norm(tokens))."""
    h = rms_norm(tokens, sd["vlm_norm.weight"], eps)
    return h @ sd["vlm_proj.weight"].T + sd["vlm_proj.bias"]


def visual_context(sd: Dict[str, torch.Tensor], latent: torch.Tensor, eps: float) -> torch.Tensor:
    """visual_context_adapter.py:77-115 for one [C, F, H, W] latent -> [N, out_dim].
    The stride=kernel Conv3d is written as a patch gather + matmul (floors like the conv does)."""
    w = sd["visual_context_adapter.patch_embedding.weight"]
    d, c, pt, ph, pw = w.shape
    C, F, H, W = latent.shape
    Fp, Hp, Wp = F // pt, H // ph, W // pw
    x = latent.float()[:, :Fp * pt, :Hp * ph, :Wp * pw]
    x = x.reshape(C, Fp, pt, Hp, ph, Wp, pw).permute(1, 3, 5, 0, 2, 4, 6).reshape(Fp * Hp * Wp, C * pt * ph * pw)
    h = x @ w.reshape(d, -1).T + sd["visual_context_adapter.patch_embedding.bias"]
    h = layer_norm(h, eps)
    h = h @ sd["visual_context_adapter.projection.weight"].T + sd["visual_context_adapter.projection.bias"]
    return layer_norm(h, eps)


def mixed_context(sd: Dict[str, torch.Tensor], batch_size: int, context: Optional[List], ar_vision_input: Optional[List],
                  visual_emb: Optional[List], special_tokens: Optional[Dict[str, torch.Tensor]],
                  max_context_len: Optional[int], text_dim: int, eps: float = 1e-6) -> List[torch.Tensor]:
    """unified_model.py:192-276: order VLM | <prp_st> text <prp_ed> | <img_st> visual <img_ed>, then truncate."""
    out = []
    has_adapter = "visual_context_adapter.projection.weight" in sd
    for i in range(batch_size):
        vlm = None if ar_vision_input is None or ar_vision_input[i] is None else vlm_features(sd, ar_vision_input[i], eps)
        txt = None if context is None or context[i] is None else context[i].float()
        vis = None
        if visual_emb is not None and has_adapter and visual_emb[i] is not None:
            vis = visual_context(sd, visual_emb[i], eps)
        parts = []
        if special_tokens is not None:
            if vlm is not None:
                parts.append(vlm)
            if txt is not None:
                parts += [special_tokens["<prp_st>"].float(), txt, special_tokens["<prp_ed>"].float()]
            if vis is not None:
                parts += [special_tokens["<img_st>"].float(), vis, special_tokens["<img_ed>"].float()]
        else:
            parts = [p for p in (vlm, txt, vis) if p is not None]
        if parts:
            c = torch.cat([p.reshape(-1, p.shape[-1]) for p in parts], dim=0)
            if max_context_len is not None and c.shape[0] > max_context_len:
                c = c[:max_context_len]
        else:
            c = torch.zeros(1, text_dim)
        out.append(c)
    return out
