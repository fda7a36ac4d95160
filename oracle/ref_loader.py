"""ORACLE support — imports the UNMODIFIED reference modules from /root/reference by file path.

Only usable in the build container (the GPU box has no /root/reference); used by
`oracle/gen_golden.py` to produce tests/golden/*.npz and by the optional cross-checks in
tests/test_oracle_golden.py (skipped when the reference tree is absent).  Recipe: SURVEY.md
Appendix B — empty `omnivideo` packages, a 10-line `diffusers` shim, and on CPU a masked-SDPA
restatement for `flash_attention` (reference attention.py:54 asserts CUDA).
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("OVB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "omnivideo", "modules", "model.py"))


def _shim_diffusers():
    if "diffusers" in sys.modules:
        return
    d = types.ModuleType("diffusers")
    cu = types.ModuleType("diffusers.configuration_utils")
    mu_pkg = types.ModuleType("diffusers.models")
    mu = types.ModuleType("diffusers.models.modeling_utils")

    class ConfigMixin:  # noqa: D401
        pass

    def register_to_config(fn):
        return fn

    class ModelMixin(torch.nn.Module):
        pass

    cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
    mu.ModelMixin = ModelMixin
    d.configuration_utils, d.models = cu, mu_pkg
    mu_pkg.modeling_utils = mu
    sys.modules.update({"diffusers": d, "diffusers.configuration_utils": cu,
                        "diffusers.models": mu_pkg, "diffusers.models.modeling_utils": mu})


def _load(modname: str, relpath: str):
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def cpu_flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None,
                        q_scale=None, causal=False, window_size=(-1, -1), deterministic=False,
                        dtype=torch.bfloat16, version=None):
    """Masked SDPA with the semantics of reference attention.py:24-133 in the input precision."""
    b, lq, lk = q.size(0), q.size(1), k.size(1)
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(q.size(-1))
    if q_scale is not None:
        q = q * q_scale
    qh, kh, vh = (t.transpose(1, 2).float() for t in (q, k, v))
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    if k_lens is not None:
        mask = torch.arange(lk)[None, :] >= torch.as_tensor(k_lens)[:, None]
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    out = torch.matmul(torch.softmax(s, dim=-1), vh).transpose(1, 2).contiguous()
    return out.type(q.dtype)


def load_reference(cpu_attention: bool = True):
    """Returns a namespace with the reference modules: .model, .attention, .vae, .ulysses, .fsp, .util"""
    assert available(), f"reference tree not found under {REF_ROOT}"
    _shim_diffusers()
    for pkg, sub in (("omnivideo", "omnivideo"), ("omnivideo.modules", "omnivideo/modules"),
                     ("omnivideo.distributed", "omnivideo/distributed")):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REF_ROOT, sub)]
            sys.modules[pkg] = m
    ns = types.SimpleNamespace()
    ns.attention = _load("omnivideo.modules.attention", "omnivideo/modules/attention.py")
    ns.model = _load("omnivideo.modules.model", "omnivideo/modules/model.py")
    ns.vae = _load("omnivideo.modules.vae2_1", "omnivideo/modules/vae2_1.py")
    ns.util = _load("omnivideo.distributed.util", "omnivideo/distributed/util.py")
    ns.ulysses = _load("omnivideo.distributed.ulysses", "omnivideo/distributed/ulysses.py")
    ns.fsp = _load("omnivideo.distributed.full_sequence_parallel",
                   "omnivideo/distributed/full_sequence_parallel.py")
    if cpu_attention:
        ns.model.flash_attention = cpu_flash_attention
        ns.ulysses.flash_attention = cpu_flash_attention
    return ns
