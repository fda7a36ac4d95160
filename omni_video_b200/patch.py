"""Rebind a LIVE reference model onto the B200 kernels — the reference's own seam.

The reference has no plugin registry: its pipeline swaps operators by attribute rebinding
(`types.MethodType`, omnivideo/x2x_gen_unified.py:176-182) and by the module globals
`model.flash_attention` (omnivideo/modules/model.py:8) / `ulysses.flash_attention`
(omnivideo/distributed/ulysses.py:4).  `patch_reference_*` use exactly that seam, so a maintainer
keeps the reference's classes, checkpoints and pipeline and only the arithmetic moves:

    import omnivideo.modules.model as ref_model_py
    from omni_video_b200 import patch
    patch.patch_reference_module_globals(ref_model_py)            # flash_attention -> ovb_attn_fwd
    patch.patch_reference_model(pipe.low_noise_model.wan_model)   # WanModel.forward -> DiTEngine
    patch.patch_reference_vae(pipe.vae)                            # Wan2_1_VAE.encode/decode

The DiTEngine only needs the parameter tree of the reference `WanModel` (same names and `[out,in]`
layouts), which is why it can be driven from the reference's own module instance.
"""
from __future__ import annotations

import types

from .modules import attention as _attention
from .modules.model import DiTEngine

__all__ = ["patch_reference_model", "patch_reference_module_globals", "patch_reference_vae",
           "patch_reference_unified", "unpatch_reference_model"]


def patch_reference_model(wan_model, sp_group=None):
    """`wan_model`: an instance of the reference's `WanModel` (or ours).  After this call
    `wan_model(x, t, context, seq_len, y=None)` runs on libovb200 and returns what the reference
    returns: a list of fp32 [C_out, F, H, W].  Parameters stay owned by `wan_model`; packed bf16
    copies are rebuilt when they move or change (`DiTEngine.packed`)."""
    for name in ("patch_embedding", "text_embedding", "time_embedding", "time_projection", "blocks",
                 "head"):
        if not hasattr(wan_model, name):
            raise TypeError(f"patch_reference_model: object has no `{name}`; expected the "
                            "reference's WanModel (omnivideo/modules/model.py:293)")
    engine = DiTEngine(wan_model)
    wan_model._ovb_engine = engine
    if sp_group is not None:
        wan_model.sp_group = sp_group

    def forward(self, x, t, context, seq_len, y=None):
        return self._ovb_engine.forward(x, t, context, seq_len, y)

    wan_model.forward = types.MethodType(forward, wan_model)
    return wan_model


def unpatch_reference_model(wan_model):
    """Remove the instance-level `forward` again (the class's own forward reappears)."""
    wan_model.__dict__.pop("forward", None)
    wan_model.__dict__.pop("_ovb_engine", None)
    return wan_model


def patch_reference_module_globals(*modules):
    """Point the `flash_attention` / `attention` globals of the given reference modules
    (omnivideo.modules.model, omnivideo.distributed.ulysses, omnivideo.modules.attention) at the
    B200 entry points, for callers that keep the reference's block-level Python."""
    for mod in modules:
        if hasattr(mod, "flash_attention"):
            mod.flash_attention = _attention.flash_attention
        if hasattr(mod, "attention"):
            mod.attention = _attention.attention


def patch_reference_vae(vae):
    """`vae`: the reference's `Wan2_1_VAE` wrapper (omnivideo/modules/vae2_1.py:620-662).  Rebinds
    `.encode` / `.decode` to the B200 VAE engine, which reads the parameters of `vae.model`."""
    from .modules.vae2_1 import VAEEngine
    engine = VAEEngine(vae.model, vae.mean, vae.std)
    vae._ovb_engine = engine

    def encode(self, videos):
        return self._ovb_engine.encode(videos)

    def decode(self, zs):
        return self._ovb_engine.decode(zs)

    vae.encode = types.MethodType(encode, vae)
    vae.decode = types.MethodType(decode, vae)
    return vae


def patch_reference_unified(unified, sp_group=None):
    """`unified`: the reference's `UnifiedWanWithMixedConditionModel`
    (omnivideo/modules/unified_model.py:13).  Its `wan_model` is patched as above, and its own
    `forward` — the conditioning front-end (vlm_norm/vlm_proj, VisualContextAdapter, special-token
    concatenation, truncation) — is rebound to the B200 implementation, which reads the reference
    instance's parameters and caches the context per sample across the sampling loop."""
    from .modules import unified_model as U
    from .modules.visual_context_adapter import visual_context_forward
    for name in ("wan_model", "vlm_norm", "vlm_proj", "max_context_len"):
        if not hasattr(unified, name):
            raise TypeError(f"patch_reference_unified: object has no `{name}`; expected the reference's "
                            "UnifiedWanWithMixedConditionModel")
    patch_reference_model(unified.wan_model, sp_group)
    unified._ctx_cache = {}
    unified.context_cache_hits = 0
    unified.context_cache_misses = 0
    cls = U.UnifiedWanWithMixedConditionModel
    for name in ("_vlm", "_params_key", "_per_sample", "mixed_context", "clear_context_cache", "forward"):
        setattr(unified, name, types.MethodType(getattr(cls, name), unified))
    unified._condition_dropout = cls._condition_dropout        # a staticmethod: no binding
    adapter = getattr(unified, "visual_context_adapter", None)
    if adapter is not None:
        adapter.forward = types.MethodType(visual_context_forward, adapter)
    return unified
