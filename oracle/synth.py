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
from typing import Dict, Tuple

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


# ------------------------------------------------------------------------------------------ VAE
def vae_state_dict(seed: int = 1818, dim: int = 96, z_dim: int = 16) -> Dict[str, torch.Tensor]:
    """fp32 state dict with the parameter names/shapes of the reference's `WanVAE_` for the one
    configuration the pipeline instantiates (vae2_1.py:596-603: dim 96, dim_mult [1,2,4,4], two
    res blocks, temperal_downsample [False, True, True]).  Weights ~ N(0, 1/fan_in) so activations
    stay O(1) through ~30 convs; the zero-initialised attention `proj.weight` (vae2_1.py:237) gets
    real values."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, co, ci, *k, gain=1.0):
        fan_in = ci * int(np.prod(k))
        sd[name + ".weight"] = _normal(rng, (co, ci) + tuple(k), gain / math.sqrt(fan_in))
        sd[name + ".bias"] = _normal(rng, (co,), 0.02)

    def gamma(name, c, images=False):
        sd[name + ".gamma"] = (1.0 + _normal(rng, (c,), 0.05)).reshape((c, 1, 1) if images else (c, 1, 1, 1))

    def res(p, ci, co):
        gamma(p + "residual.0", ci)
        conv(p + "residual.2", co, ci, 3, 3, 3)
        gamma(p + "residual.3", co)
        conv(p + "residual.6", co, co, 3, 3, 3, gain=0.5)
        if ci != co:
            conv(p + "shortcut", co, ci, 1, 1, 1)

    def attn(p, c):
        gamma(p + "norm", c, images=True)
        conv(p + "to_qkv", 3 * c, c, 1, 1)
        conv(p + "proj", c, c, 1, 1, gain=0.5)

    # ---- encoder (vae2_1.py:264-315)
    dims = [dim * u for u in (1, 1, 2, 4, 4)]
    conv("encoder.conv1", dims[0], 3, 3, 3, 3)
    j = 0
    t_down = [False, True, True]
    for i, (ci, co) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(2):
            res(f"encoder.downsamples.{j}.", ci, co)
            ci = co
            j += 1
        if i != 3:
            conv(f"encoder.downsamples.{j}.resample.1", co, co, 3, 3)
            if t_down[i]:
                conv(f"encoder.downsamples.{j}.time_conv", co, co, 3, 1, 1)
            j += 1
    res("encoder.middle.0.", dims[-1], dims[-1])
    attn("encoder.middle.1.", dims[-1])
    res("encoder.middle.2.", dims[-1], dims[-1])
    gamma("encoder.head.0", dims[-1])
    conv("encoder.head.2", z_dim * 2, dims[-1], 3, 3, 3)
    conv("conv1", z_dim * 2, z_dim * 2, 1, 1, 1)
    conv("conv2", z_dim, z_dim, 1, 1, 1)
    # ---- decoder (vae2_1.py:368-420)
    ddims = [dim * u for u in (4, 4, 4, 2, 1)]
    conv("decoder.conv1", ddims[0], z_dim, 3, 3, 3)
    res("decoder.middle.0.", ddims[0], ddims[0])
    attn("decoder.middle.1.", ddims[0])
    res("decoder.middle.2.", ddims[0], ddims[0])
    j = 0
    t_up = [True, True, False]
    for i, (ci, co) in enumerate(zip(ddims[:-1], ddims[1:])):
        if i in (1, 2, 3):
            ci = ci // 2
        for _ in range(3):
            res(f"decoder.upsamples.{j}.", ci, co)
            ci = co
            j += 1
        if i != 3:
            conv(f"decoder.upsamples.{j}.resample.1", co // 2, co, 3, 3)
            if t_up[i]:
                conv(f"decoder.upsamples.{j}.time_conv", co * 2, co, 3, 1, 1)
            j += 1
    gamma("decoder.head.0", ddims[-1])
    conv("decoder.head.2", 3, ddims[-1], 3, 3, 3)
    return sd


def vae_latent(grid: Tuple[int, int, int], seed: int = 7) -> torch.Tensor:
    """Normalised latent z [16, T, h, w] ~ N(0,1) (what the sampler hands to decode)."""
    rng = np.random.default_rng(seed)
    return _normal(rng, (16,) + tuple(grid), 1.0)


def vae_video(shape: Tuple[int, int, int], seed: int = 7) -> torch.Tensor:
    """Smooth synthetic video [3, T, H, W] in [-1, 1]: low-frequency sinusoids + 0.05 noise
    (SURVEY.md §8(d) config 5)."""
    rng = np.random.default_rng(seed)
    T, H, W = shape
    t = torch.arange(T, dtype=torch.float32).view(1, T, 1, 1)
    y = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
    x = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    ph = torch.from_numpy(rng.uniform(0, 6.28, size=(3, 1, 1, 1)).astype(np.float32))
    v = 0.5 * torch.sin(0.11 * x + 0.07 * y + 0.3 * t + ph) + 0.3 * torch.cos(0.05 * x - 0.13 * y + ph)
    v = v + 0.05 * _normal(rng, (3, T, H, W), 1.0)
    return v.clamp(-1, 1)


# ------------------------------------------------------------------------- conditioning front-end
def unified_state_dict(cfg: dict, vlm_in_dim: int, adapter_patch=(1, 4, 4), seed: int = 1818,
                       with_adapter: bool = True) -> Dict[str, torch.Tensor]:
    """fp32 tensors for UnifiedWanWithMixedConditionModel's OWN parameters (reference
    unified_model.py:78-92): vlm_norm, vlm_proj and the VisualContextAdapter."""
    rng = np.random.default_rng(seed + 101)
    td, d, cin = cfg["text_dim"], cfg["dim"], cfg["in_dim"]
    sd = {"vlm_norm.weight": 1.0 + _normal(rng, (vlm_in_dim,), 0.05),
          "vlm_proj.weight": _xavier(rng, td, vlm_in_dim), "vlm_proj.bias": _normal(rng, (td,), 0.02)}
    if with_adapter:
        k = cin * adapter_patch[0] * adapter_patch[1] * adapter_patch[2]
        sd["visual_context_adapter.patch_embedding.weight"] = _xavier(rng, d, k).view(d, cin, *adapter_patch).contiguous()
        sd["visual_context_adapter.patch_embedding.bias"] = _normal(rng, (d,), 0.02)
        sd["visual_context_adapter.projection.weight"] = _xavier(rng, td, d)
        sd["visual_context_adapter.projection.bias"] = _normal(rng, (td,), 0.02)
    return sd


def unified_inputs(cfg: dict, vlm_in_dim: int, n_vlm, n_text, visual_grid, seed: int = 7,
                   special_tokens: bool = True):
    """Per-sample conditions: n_vlm / n_text are lists (None = condition absent for that sample),
    visual_grid is (F, H, W) of the conditioning latent or None."""
    rng = np.random.default_rng(seed + 13)
    B = len(n_text)
    ar = [None if n is None else _normal(rng, (n, vlm_in_dim), 1.0) for n in n_vlm]
    ctx = [None if n is None else _normal(rng, (n, cfg["text_dim"]), 1.0) for n in n_text]
    vis = None if visual_grid is None else [_normal(rng, (cfg["in_dim"], *visual_grid), 1.0) for _ in range(B)]
    st = None
    if special_tokens:
        st = {k: _normal(rng, (1, cfg["text_dim"]), 1.0) for k in ("<img_st>", "<img_ed>", "<prp_st>", "<prp_ed>")}
    if all(a is None for a in ar):
        ar = None
    return ar, ctx, vis, st


# ------------------------------------------------------------------------------- umT5 encoder
TINY_T5 = dict(vocab=97, dim=128, dim_attn=128, dim_ffn=256, num_heads=2, num_layers=2, num_buckets=32,
               shared_pos=False)          # head_dim 64 like umT5-XXL (4096 / 64 heads)


def t5_state_dict(cfg: dict, seed: int = 1818) -> Dict[str, torch.Tensor]:
    """fp32 state dict of the reference's T5Encoder (t5.py:266-297); scales follow its init_weights (:26-44)
    except that norm weights are perturbed around 1 so that they matter."""
    rng = np.random.default_rng(seed + 7)
    d, da, df, nh, nb = cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"], cfg["num_heads"], cfg["num_buckets"]
    sd = {"token_embedding.weight": _normal(rng, (cfg["vocab"], d), 1.0)}
    if cfg.get("shared_pos"):
        sd["pos_embedding.embedding.weight"] = _normal(rng, (nb, nh), 0.5)
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + _normal(rng, (d,), 0.05)
        sd[p + "attn.q.weight"] = _normal(rng, (da, d), (d * (da // nh)) ** -0.5 * 4)
        sd[p + "attn.k.weight"] = _normal(rng, (da, d), d ** -0.5)
        sd[p + "attn.v.weight"] = _normal(rng, (da, d), d ** -0.5)
        sd[p + "attn.o.weight"] = _normal(rng, (d, da), da ** -0.5)
        sd[p + "norm2.weight"] = 1.0 + _normal(rng, (d,), 0.05)
        sd[p + "ffn.gate.0.weight"] = _normal(rng, (df, d), d ** -0.5)
        sd[p + "ffn.fc1.weight"] = _normal(rng, (df, d), d ** -0.5)
        sd[p + "ffn.fc2.weight"] = _normal(rng, (d, df), df ** -0.5)
        if not cfg.get("shared_pos"):
            sd[p + "pos_embedding.embedding.weight"] = _normal(rng, (nb, nh), 0.5)
    sd["norm.weight"] = 1.0 + _normal(rng, (d,), 0.05)
    return sd


def t5_inputs(cfg: dict, lengths, seq_len: int, seed: int = 7):
    """ids [B, seq_len] (0-padded, as the tokenizer pads to text_len) and mask [B, seq_len]."""
    rng = np.random.default_rng(seed + 3)
    B = len(lengths)
    ids = torch.zeros(B, seq_len, dtype=torch.int64)
    mask = torch.zeros(B, seq_len, dtype=torch.int64)
    for b, n in enumerate(lengths):
        ids[b, :n] = torch.from_numpy(rng.integers(1, cfg["vocab"], size=n))
        mask[b, :n] = 1
    return ids, mask
