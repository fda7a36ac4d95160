"""ORACLE support — deterministic synthetic weights and inputs (test infrastructure only).

Everything is drawn from numpy's PCG64 (`np.random.default_rng(seed)`), which is bit-stable across
numpy versions and platforms, so the build container (where the goldens are generated from the
reference) and the GPU box (where the CUDA path is checked) see identical tensors without shipping
weights.  Scales mimic the reference's initialisers (model.py:551-573) except that the two tensors
the reference zero-initialises (`head.head.weight`, VAE `proj.weight`) get real values, otherwise
the output would be identically zero (SURVEY.md §8(c) "Random-init pitfalls").
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

# reduced DiT used by tests and goldens (head_dim must stay 128)
TINY_DIT = dict(model_type="t2v", patch_size=(1, 2, 2), text_len=32, in_dim=16, dim=256, ffn_dim=512,
                freq_dim=256, text_dim=64, out_dim=16, num_heads=2, num_layers=2,
                window_size=(-1, -1), qk_norm=True, cross_attn_norm=True, eps=1e-6)

# OmniVideo2 configurations (reference omnivideo/configs/wan_t2v_1_3B.py:19-40, wan_t2v_A14B.py:19-42)
DIT_1_3B = dict(model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=1536,
                ffn_dim=8960, freq_dim=256, text_dim=4096, out_dim=16, num_heads=12, num_layers=30,
                window_size=(-1, -1), qk_norm=True, cross_attn_norm=True, eps=1e-6)
DIT_A14B = dict(model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=5120,
                ffn_dim=13824, freq_dim=256, text_dim=4096, out_dim=16, num_heads=40, num_layers=40,
                window_size=(-1, -1), qk_norm=True, cross_attn_norm=True, eps=1e-6)


def _normal(rng, shape, std):
    return torch.from_numpy((rng.standard_normal(shape, dtype=np.float32) * np.float32(std)))


def _xavier(rng, out_f, in_f):
    a = math.sqrt(6.0 / (in_f + out_f))
    return torch.from_numpy(rng.uniform(-a, a, size=(out_f, in_f)).astype(np.float32))


def dit_state_dict(cfg: dict, seed: int = 1818, num_layers: int = None) -> Dict[str, torch.Tensor]:
    """fp32 state dict with the reference's parameter names/shapes."""
    rng = np.random.default_rng(seed)
    d, f, td, fd = cfg["dim"], cfg["ffn_dim"], cfg["text_dim"], cfg["freq_dim"]
    cin, cout = cfg["in_dim"], cfg["out_dim"]
    nl = cfg["num_layers"] if num_layers is None else num_layers
    sd = {}

    def lin(name, out_f, in_f, std=None, bias_std=0.02):
        sd[name + ".weight"] = _normal(rng, (out_f, in_f), std) if std else _xavier(rng, out_f, in_f)
        sd[name + ".bias"] = _normal(rng, (out_f,), bias_std)

    sd["patch_embedding.weight"] = _xavier(rng, d, cin * 4).view(d, cin, 1, 2, 2).contiguous()
    sd["patch_embedding.bias"] = _normal(rng, (d,), 0.02)
    lin("text_embedding.0", d, td, std=0.02)
    lin("text_embedding.2", d, d, std=0.02)
    lin("time_embedding.0", d, fd, std=0.02)
    lin("time_embedding.2", d, d, std=0.02)
    lin("time_projection.1", 6 * d, d)
    for i in range(nl):
        p = f"blocks.{i}."
        for att in ("self_attn", "cross_attn"):
            for w in ("q", "k", "v", "o"):
                lin(p + f"{att}.{w}", d, d)
            sd[p + f"{att}.norm_q.weight"] = 1.0 + _normal(rng, (d,), 0.05)
            sd[p + f"{att}.norm_k.weight"] = 1.0 + _normal(rng, (d,), 0.05)
        sd[p + "norm3.weight"] = 1.0 + _normal(rng, (d,), 0.05)
        sd[p + "norm3.bias"] = _normal(rng, (d,), 0.02)
        lin(p + "ffn.0", f, d)
        lin(p + "ffn.2", d, f)
        sd[p + "modulation"] = _normal(rng, (1, 6, d), 1.0 / math.sqrt(d))
    lin("head.head", cout * 4, d)  # NOT zero (reference zero-inits it, model.py:573)
    sd["head.modulation"] = _normal(rng, (1, 2, d), 1.0 / math.sqrt(d))
    return sd


def dit_inputs(cfg: dict, grid_lat: Tuple[int, int, int], ctx_len: int, seed: int = 7,
               batch: int = 1, timestep: int = 999):
    """x: list of [16, F, H, W] N(0,1); context: list of [ctx_len, text_dim] N(0,1); t: [B] int64."""
    rng = np.random.default_rng(seed)
    Fr, H, W = grid_lat
    xs = [_normal(rng, (cfg["in_dim"], Fr, H, W), 1.0) for _ in range(batch)]
    ctx = [_normal(rng, (ctx_len, cfg["text_dim"]), 1.0) for _ in range(batch)]
    t = torch.full((batch,), timestep, dtype=torch.int64)
    return xs, t, ctx


def round_bf16(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k: v.to(torch.bfloat16).to(torch.float32) for k, v in sd.items()}
