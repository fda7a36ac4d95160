"""B200-native UnifiedWanWithMixedConditionModel (drop-in for reference omnivideo/modules/unified_model.py).

Builds the DiT's cross-attention context from the mixed conditions ("tight concatenation",
reference unified_model.py:138-278) and hands it to `WanModel.forward`:

    VLM tokens   [Nv, vlm_in_dim] : WanRMSNorm -> Linear(vlm_in_dim, text_dim)   ovb_rmsnorm_rope (rope=0) + ovb_gemm_bf16
    text prompt  [Lt, text_dim]   : as is, between <prp_st> ... <prp_ed>
    visual clip  [16, F, H, W]    : VisualContextAdapter, between <img_st> ... <img_ed>
    -> concat in the order VLM | text | visual, truncated to max_context_len

The context does not depend on the timestep, and the sampling loop calls the model twice per step
for 50 steps with the SAME condition tensors (x2x_gen_unified.py:452-481), so the result is cached
per sample, keyed on the identity/version of the condition tensors: the
front-end runs once per prompt instead of 100 times.  The training-time condition dropout
(`classifier_free_ratio`, unified_model.py:157-190) is kept, including its consumption of two
`random.random()` draws per sample per call, which keeps Python's RNG stream identical to the
reference's.
"""
from __future__ import annotations

import logging
import os
import random
from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn

from .. import ops
from .model import WanModel, WanRMSNorm
from .visual_context_adapter import VisualContextAdapter

bf16 = torch.bfloat16
f32 = torch.float32

__all__ = ["UnifiedWanWithMixedConditionModel", "build_mixed_context"]


def _item(item, idx, batch_size):
    """reference unified_model.py:222-240 (`get_item`): pick sample idx, force 2-D [L, D]."""
    if item is None:
        return None
    if isinstance(item, list) and idx is not None:
        item = item[idx]
    if isinstance(item, torch.Tensor):
        if idx is not None and item.dim() >= 3 and item.size(0) == batch_size:
            item = item[idx]
        if item.dim() == 3:
            item = item[0]
        elif item.dim() == 1:
            item = item.unsqueeze(0)
    return item


def _ident(t):
    if t is None:
        return None
    if isinstance(t, (list, tuple)):
        return tuple(_ident(u) for u in t)
    if isinstance(t, dict):
        return tuple((k, _ident(v)) for k, v in sorted(t.items()))
    if isinstance(t, torch.Tensor):
        return (t.data_ptr(), t._version, tuple(t.shape), t.dtype, str(t.device))
    return t


def _tensors(t, out):
    """Flat list of the tensors inside a condition argument, in `_ident`'s traversal order."""
    if isinstance(t, (list, tuple)):
        for u in t:
            _tensors(u, out)
    elif isinstance(t, dict):
        for _, v in sorted(t.items()):
            _tensors(v, out)
    elif isinstance(t, torch.Tensor):
        out.append(t)
    return out


# WanModel constructor defaults the reference applies when a key is missing from `wan_config`
# (unified_model.py:50-61); the remaining keys are forwarded only when present (:63-67).
_WAN_CONFIG_DEFAULTS = dict(patch_size=(1, 2, 2), dim=2048, ffn_dim=8192, freq_dim=256, num_heads=16,
                            num_layers=32, window_size=(-1, -1), qk_norm=True, cross_attn_norm=True, eps=1e-6)
_WAN_CONFIG_OPTIONAL = ("model_type", "text_len", "in_dim", "text_dim", "out_dim")


def _wan_from_config(wan_config: Dict, skip_init: bool) -> WanModel:
    kw = {k: wan_config.get(k, dflt) for k, dflt in _WAN_CONFIG_DEFAULTS.items()}
    kw.update({k: wan_config[k] for k in _WAN_CONFIG_OPTIONAL if k in wan_config})
    logging.info("WanModel from config, random weights (skip_init=%s)", skip_init)
    return WanModel(**kw, skip_init=skip_init)


class UnifiedWanWithMixedConditionModel(nn.Module):
    # the signature is the reference's (unified_model.py:22-38): pipelines construct it by keyword
    def __init__(self, wan_model_or_ckpt_dir: Union[WanModel, str, None] = None, wan_subfolder: Optional[str] = None,
                 vlm_in_dim: int = 4096, precision_dtype: torch.dtype = torch.float32, device_id: int = 0,
                 rank: int = 0, dit_fsdp: bool = False, use_usp: bool = False,
                 use_visual_context_adapter: bool = False, visual_context_adapter_patch_size: tuple = None,
                 max_context_len: int = None, eps: float = 1e-6, wan_config: Optional[Dict] = None,
                 skip_init: bool = False):
        super().__init__()
        src = wan_model_or_ckpt_dir
        if isinstance(src, WanModel):
            wan = src
        elif isinstance(src, str):
            wan = WanModel.from_pretrained(src, subfolder=wan_subfolder)
        elif wan_config is not None:
            wan = _wan_from_config(wan_config, skip_init)
        else:
            raise ValueError("Either wan_model_or_ckpt_dir or wan_config must be provided")
        self.wan_model = wan
        if max_context_len is not None:            # the DiT pads/attends over this many context rows
            wan.text_len = max_context_len
        self.max_context_len = wan.text_len

        # VLM tokens -> text space; parameters only, the arithmetic is in `_vlm`
        self.vlm_norm = WanRMSNorm(vlm_in_dim, eps=eps)
        self.vlm_proj = nn.Linear(vlm_in_dim, wan.text_dim)
        self.visual_context_adapter = None
        if use_visual_context_adapter:
            patch = visual_context_adapter_patch_size or wan.patch_size
            self.visual_context_adapter = VisualContextAdapter(patch_size=patch, in_channels=wan.in_dim,
                                                               hidden_dim=wan.dim, out_dim=wan.text_dim, eps=eps)
        self._ctx_cache: Dict = {}
        self.context_cache_hits = 0
        self.context_cache_misses = 0

    @classmethod
    def from_pretrained(cls, wan_ckpt_dir, **kw):
        return cls(wan_model_or_ckpt_dir=wan_ckpt_dir, **kw)

    def enable_gradient_checkpointing(self):
        self.wan_model.enable_gradient_checkpointing()

    def reset_wan_text_len(self, new_len: int):
        self.wan_model.text_len = new_len

    # ------------------------------------------------------------------ the front-end proper
    @torch.no_grad()
    def _vlm(self, tokens: torch.Tensor) -> torch.Tensor:
        """vlm_proj(vlm_norm(tokens)) (reference unified_model.py:195-206): bf16 [.., text_dim]."""
        lead, width = tokens.shape[:-1], tokens.shape[-1]
        h = tokens.reshape(-1, width).to(bf16).contiguous().clone()        # normalised in place below
        ops.rmsnorm_rope_(h, self.vlm_norm.weight.data.to(bf16).contiguous(), self.vlm_norm.eps)
        out = ops.gemm(h, self.vlm_proj.weight.data.to(bf16).contiguous(),
                       self.vlm_proj.bias.data.to(bf16).contiguous())
        return out.view(*lead, out.shape[-1])

    def _params_key(self):
        ps = [self.vlm_norm.weight, self.vlm_proj.weight, self.vlm_proj.bias]
        if self.visual_context_adapter is not None:
            ps += list(self.visual_context_adapter.parameters())
        return tuple((p.data_ptr(), p._version) for p in ps)

    def _per_sample(self, fn, batched, batch_size):
        """Apply `fn` to a per-sample list (None entries stay None) or to a batched tensor as a whole —
        the two input conventions of reference unified_model.py:195-217."""
        if batched is None:
            return None
        if isinstance(batched, list):
            return [None if batched[i] is None else fn(batched[i]) for i in range(batch_size)]
        return fn(batched)

    @torch.no_grad()
    def mixed_context(self, batch_size, context=None, ar_vision_input=None, visual_emb=None,
                      special_token_dict=None, device=None) -> List[torch.Tensor]:
        """Steps 2-4 of reference unified_model.py:192-276 -> list of bf16 [<= max_context_len, text_dim].
        (The reference's list entries are fp32 after `torch.cat` promotion; `text_embedding` rounds
        them to bf16 under autocast before its first GEMM — the values that reach the GEMM are the same.)"""
        key = (batch_size, _ident(context), _ident(ar_vision_input), _ident(visual_emb),
               _ident(special_token_dict), self._params_key(), self.max_context_len)
        # `_ident` keys on storage address + version, which identifies a tensor only while it is
        # alive: a freed condition tensor's block is handed to the next prompt's tensor of the same
        # size by the caching allocator.  Every entry therefore HOLDS its input tensors (their
        # addresses cannot be recycled while the entry lives) and a hit is confirmed object by object.
        inputs = _tensors((context, ar_vision_input, visual_emb, special_token_dict), [])
        hit = self._ctx_cache.get(key)
        if hit is not None and len(hit[1]) == len(inputs) and all(a is b for a, b in zip(hit[1], inputs)):
            self.context_cache_hits += 1
            return hit[0]
        self.context_cache_misses += 1

        vlm = self._per_sample(self._vlm, ar_vision_input, batch_size)
        visual = None
        if visual_emb is not None:
            adapter = self.visual_context_adapter
            if adapter is None:
                visual = [None] * batch_size       # the reference keeps the slots and fills none (:209-217)
            elif isinstance(visual_emb, list):
                visual = self._per_sample(lambda v: adapter(v).squeeze(0), visual_emb, batch_size)
            else:
                visual = adapter(visual_emb)

        tok = (lambda name: _item(special_token_dict.get(name), None, batch_size)) if special_token_dict else None
        mixed = []
        for idx in range(batch_size):
            parts = {"vlm": _item(vlm, idx, batch_size), "text": _item(context, idx, batch_size),
                     "visual": _item(visual, idx, batch_size)}
            seq = []
            if tok is None:                         # no markers: VLM | text | visual
                seq = [parts[k] for k in ("vlm", "text", "visual")]
            else:                                   # VLM | <prp_st> text <prp_ed> | <img_st> visual <img_ed>
                seq.append(parts["vlm"])
                if parts["text"] is not None:
                    seq += [tok("<prp_st>"), parts["text"], tok("<prp_ed>")]
                if parts["visual"] is not None:
                    seq += [tok("<img_st>"), parts["visual"], tok("<img_ed>")]
            seq = [c for c in seq if c is not None]
            if not seq:                             # nothing to attend to: one zero row (:273-274)
                mixed.append(torch.zeros(1, self.wan_model.text_dim, device=device, dtype=bf16))
                continue
            dev = device if device is not None else seq[0].device
            rows = torch.cat([c.to(device=dev, dtype=bf16) for c in seq], dim=0)
            mixed.append(rows[:self.max_context_len] if self.max_context_len is not None else rows)
        if len(self._ctx_cache) >= 8:          # two entries per prompt (cond / uncond) are live at a time
            self._ctx_cache.pop(next(iter(self._ctx_cache)))
        self._ctx_cache.pop(key, None)         # a same-address entry of a dead tensor, if any
        self._ctx_cache[key] = (mixed, inputs)
        return mixed

    def clear_context_cache(self):
        """Drop the cached mixed contexts (and the condition tensors they keep alive)."""
        self._ctx_cache.clear()

    @staticmethod
    def _condition_dropout(batch_size, context, ar_vision_input, classifier_free_ratio, unconditioned_context):
        """Training-time CFG (reference unified_model.py:157-190).  Two `random.random()` draws per sample
        are consumed on EVERY call, used or not, so Python's RNG stream matches the reference's; the second
        set belongs to a branch whose threshold upstream is 0.00 (never taken)."""
        drop = [random.random() for _ in range(batch_size)]
        _unused = [random.random() for _ in range(batch_size)]
        if not classifier_free_ratio or classifier_free_ratio <= 0:
            return
        for idx, r in enumerate(drop):
            if r >= classifier_free_ratio:
                continue
            if unconditioned_context is None:
                raise ValueError("unconditioned_context must be provided when classifier_free_ratio > 0")
            if not isinstance(unconditioned_context, dict):
                continue
            if ar_vision_input is not None and ar_vision_input[idx] is not None:
                v = ar_vision_input[idx]            # null VLM condition: two near-zero tokens
                ar_vision_input[idx] = torch.zeros(2, v.size(1), device=v.device, dtype=v.dtype) + 1e-6
            if context is not None and context[idx] is not None and "uncond_context" in unconditioned_context:
                context[idx] = unconditioned_context["uncond_context"]

    def forward(self, x: List[torch.Tensor], t: torch.Tensor, context: List[torch.Tensor] = None,
                aligned_emb: Optional[torch.Tensor] = None, ar_vision_input: Optional[List[torch.Tensor]] = None,
                visual_emb: Optional[torch.Tensor] = None, seq_len: Optional[int] = None,
                special_token_dict: Optional[Dict[str, torch.Tensor]] = None,
                classifier_free_ratio: Optional[float] = None,
                unconditioned_context: Optional[Dict[str, torch.Tensor]] = None, condition_mode: str = "auto"):
        """Same keywords as reference unified_model.py:138-152; returns what `WanModel.forward` returns."""
        if condition_mode not in ("auto", "full"):
            # the reference's forward has no other branch: it falls off the end and returns None, which
            # its caller then indexes.  Fail here with the reason instead.
            raise ValueError(f"condition_mode={condition_mode!r}: only 'full' / 'auto' are implemented "
                             "(reference unified_model.py:152-158 has no other branch either)")
        batch_size = x.size(0) if isinstance(x, torch.Tensor) else len(x)
        self._condition_dropout(batch_size, context, ar_vision_input, classifier_free_ratio, unconditioned_context)
        mixed = self.mixed_context(batch_size, context, ar_vision_input, visual_emb, special_token_dict, x[0].device)
        return self.wan_model(x, t=t, context=mixed, seq_len=seq_len)

    # ------------------------------------------------------------------ persistence (reference :280-285)
    def save_pretrained(self, save_dir: str):
        os.makedirs(save_dir, exist_ok=True)
        self.wan_model.save_pretrained(os.path.join(save_dir, "wan_model"))
        parts = {"vlm_norm": self.vlm_norm, "vlm_proj": self.vlm_proj}
        if self.visual_context_adapter is not None:
            parts["visual_context_adapter"] = self.visual_context_adapter
        for name, mod in parts.items():
            torch.save(mod.state_dict(), os.path.join(save_dir, f"{name}_pytorch_model.bin"))

    def enable_eval(self):
        self.eval()

    def enable_train(self):
        raise NotImplementedError("the B200 path is inference-only (SURVEY.md §8(b) Threading)")


def build_mixed_context(model: UnifiedWanWithMixedConditionModel, **kw):
    """Functional form used by patch.patch_reference_unified()."""
    return model.mixed_context(**kw)
