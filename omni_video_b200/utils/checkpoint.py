"""Checkpoint formats of the reference, read into the B200 module trees.

Parameter names and shapes of `omni_video_b200.modules.*` are the reference's, so loading is key-for-key;
what this module restates is the reference's FILE handling:

* trained Omni-Video checkpoints, `high_noise_model/model.pt` / `low_noise_model/model.pt`
  (tools/inference/generate_omni_v2v.py:495-557): `torch.load` on CPU, unwrap a `'module'` (DeepSpeed) or
  `'model'` wrapper, cast every tensor to the run's precision BEFORE `load_state_dict(strict=False)`,
  report missing / unexpected keys;
* the diffusers directory layout that `WanModel.from_pretrained(dir, subfolder=...)` reads
  (omnivideo/modules/unified_model.py:43; `ModelMixin`): `config.json` holding the constructor arguments
  that are not in `WanModel.ignore_for_config`, and `diffusion_pytorch_model.safetensors`, its sharded form
  with `diffusion_pytorch_model.safetensors.index.json`, or `diffusion_pytorch_model.bin`;
* `special_tokens.pkl` (omnivideo/x2x_gen_unified.py:215-223): a pickled dict of `[1, text_dim]` tensors,
  moved to the run's dtype and device;
* `Wan2.1_VAE.pth` is handled where the reference handles it (`modules/vae2_1.py::_video_vae`).

Packed bf16 weight copies (fused Wq|Wk|Wv, conv weights in tap-major layout ...) are derived caches of the
engines, rebuilt from the parameters when those change — nothing here needs to know about them.
"""
from __future__ import annotations

import json
import logging
import os
import pickle
from typing import Dict, Iterable, Optional, Tuple

import torch

WEIGHTS_NAME = "diffusion_pytorch_model.bin"
SAFETENSORS_NAME = "diffusion_pytorch_model.safetensors"
SAFETENSORS_INDEX = SAFETENSORS_NAME + ".index.json"
CONFIG_NAME = "config.json"

__all__ = ["read_state_dict", "load_trained_checkpoint", "wan_from_pretrained", "save_wan_pretrained",
           "load_special_tokens"]


def _unwrap(obj) -> Dict[str, torch.Tensor]:
    """generate_omni_v2v.py:501-505 (+ the DistributedDataParallel `module.` key prefix)."""
    if isinstance(obj, dict):
        for wrapper in ("module", "model", "state_dict"):
            inner = obj.get(wrapper)
            if isinstance(inner, dict) and inner and all(isinstance(k, str) for k in inner):
                obj = inner
                break
    if not isinstance(obj, dict):
        raise TypeError(f"checkpoint holds a {type(obj).__name__}, expected a state dict")
    if obj and all(k.startswith("module.") for k in obj):
        obj = {k[len("module."):]: v for k, v in obj.items()}
    return obj


def read_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """One checkpoint file -> {name: CPU tensor}.  `.safetensors` via safetensors, anything else via
    `torch.load(map_location='cpu', weights_only=True)` (memory-mapped when the file allows it)."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True, mmap=True)
    except (RuntimeError, ValueError):      # legacy (non-zip) serialisation cannot be mmapped
        obj = torch.load(path, map_location="cpu", weights_only=True)
    return _unwrap(obj)


def load_trained_checkpoint(model: torch.nn.Module, path: str, dtype: Optional[torch.dtype] = None,
                            strict: bool = False) -> Tuple[Iterable[str], Iterable[str]]:
    """The reference's loading of a trained expert (generate_omni_v2v.py:499-520) into `model` — a
    `UnifiedWanWithMixedConditionModel` (keys `wan_model.*`, `vlm_norm.*`, `vlm_proj.*`,
    `visual_context_adapter.*`) or a bare `WanModel`.  Returns (missing_keys, unexpected_keys)."""
    sd = read_state_dict(path)
    if dtype is not None:
        sd = {k: (v.to(dtype) if isinstance(v, torch.Tensor) and v.is_floating_point() else v)
              for k, v in sd.items()}
    res = model.load_state_dict(sd, strict=strict)
    logging.info("[%s] load_state_dict: missing_keys=%d, unexpected_keys=%d", os.path.basename(path),
                 len(res.missing_keys), len(res.unexpected_keys))
    return list(res.missing_keys), list(res.unexpected_keys)


# ------------------------------------------------------------------------------ diffusers layout
def _model_dir(ckpt_dir: str, subfolder: Optional[str]) -> str:
    d = os.path.join(ckpt_dir, subfolder) if subfolder else ckpt_dir
    if not os.path.isdir(d):
        raise FileNotFoundError(f"no checkpoint directory {d}")
    return d


def wan_from_pretrained(ckpt_dir: str, subfolder: Optional[str] = None,
                        torch_dtype: Optional[torch.dtype] = None, **overrides):
    """`WanModel.from_pretrained` (diffusers `ModelMixin`): constructor arguments from `config.json`
    (keys starting with `_` are diffusers bookkeeping), weights from the safetensors file, its shards,
    or the `.bin`; every checkpoint key must exist in the model and vice versa."""
    from ..modules.model import WanModel
    d = _model_dir(ckpt_dir, subfolder)
    with open(os.path.join(d, CONFIG_NAME)) as fh:
        cfg = {k: v for k, v in json.load(fh).items() if not k.startswith("_")}
    cfg.update(overrides)
    for k in ("patch_size", "window_size"):
        if k in cfg:
            cfg[k] = tuple(cfg[k])
    model = WanModel(**cfg, skip_init=True)

    index = os.path.join(d, SAFETENSORS_INDEX)
    if os.path.isfile(index):
        with open(index) as fh:
            shards = sorted(set(json.load(fh)["weight_map"].values()))
        files = [os.path.join(d, s) for s in shards]
    elif os.path.isfile(os.path.join(d, SAFETENSORS_NAME)):
        files = [os.path.join(d, SAFETENSORS_NAME)]
    elif os.path.isfile(os.path.join(d, WEIGHTS_NAME)):
        files = [os.path.join(d, WEIGHTS_NAME)]
    else:
        raise FileNotFoundError(f"{d}: none of {SAFETENSORS_INDEX}, {SAFETENSORS_NAME}, {WEIGHTS_NAME}")
    expected = set(model.state_dict().keys())
    seen = set()
    for f in files:                                   # shard by shard: peak host memory = one shard
        sd = read_state_dict(f)
        unexpected = set(sd) - expected
        if unexpected:
            raise KeyError(f"{f}: keys not in WanModel: {sorted(unexpected)[:5]} ...")
        model.load_state_dict(sd, strict=False)
        seen |= set(sd)
    if seen != expected:
        raise KeyError(f"{d}: checkpoint lacks {sorted(expected - seen)[:5]} ...")
    if torch_dtype is not None:
        model = model.to(torch_dtype)
    return model.eval()


def save_wan_pretrained(model, save_dir: str, max_shard_bytes: int = 5 << 30, safe_serialization: bool = True):
    """Writer of the same layout (`ModelMixin.save_pretrained`): used by
    `UnifiedWanWithMixedConditionModel.save_pretrained` and by the tests."""
    os.makedirs(save_dir, exist_ok=True)
    cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in model.config.items()
           if k not in getattr(model, "ignore_for_config", ())}
    cfg.update(_class_name="WanModel", _diffusers_version="0.35.1")
    with open(os.path.join(save_dir, CONFIG_NAME), "w") as fh:
        json.dump(cfg, fh, indent=2)
    sd = {k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()}
    if not safe_serialization:
        torch.save(sd, os.path.join(save_dir, WEIGHTS_NAME))
        return
    from safetensors.torch import save_file
    shards, cur, size = [], {}, 0
    for k, v in sd.items():
        nb = v.numel() * v.element_size()
        if cur and size + nb > max_shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = v
        size += nb
    shards.append(cur)
    if len(shards) == 1:
        save_file(shards[0], os.path.join(save_dir, SAFETENSORS_NAME))
        return
    weight_map = {}
    for i, shard in enumerate(shards):
        name = f"diffusion_pytorch_model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(shard, os.path.join(save_dir, name))
        weight_map.update({k: name for k in shard})
    total = sum(v.numel() * v.element_size() for v in sd.values())
    with open(os.path.join(save_dir, SAFETENSORS_INDEX), "w") as fh:
        json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, fh, indent=2)


def load_special_tokens(path: str, dtype: torch.dtype, device) -> Optional[Dict[str, torch.Tensor]]:
    """x2x_gen_unified.py:215-226: `special_tokens.pkl` next to the experts, or None when it is absent.
    The file is a plain pickle (that is how the reference writes and reads it): only load files you trust."""
    if not os.path.exists(path):
        logging.info("No special_tokens.pkl found at %s, special_tokens disabled", path)
        return None
    with open(path, "rb") as fh:
        toks = pickle.load(fh)
    assert isinstance(toks, dict), "special_tokens should be a dictionary"
    return {k: v.to(dtype).to(device) for k, v in toks.items()}
