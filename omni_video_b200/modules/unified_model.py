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


class UnifiedWanWithMixedConditionModel(nn.Module):
    def __init__(
        self,
        wan_model_or_ckpt_dir: Union[WanModel, str, None] = None,
        wan_subfolder: Optional[str] = None,
        vlm_in_dim: int = 4096,
        precision_dtype: torch.dtype = torch.float32,
        device_id: int = 0,
        rank: int = 0,
        dit_fsdp: bool = False,
        use_usp: bool = False,
        use_visual_context_adapter: bool = False,
        visual_context_adapter_patch_size: tuple = None,
        max_context_len: int = None,
        eps: float = 1e-6,
        wan_config: Optional[Dict] = None,
        skip_init: bool = False,
    ):
        super().__init__()
        if isinstance(wan_model_or_ckpt_dir, str):
            self.wan_model = WanModel.from_pretrained(wan_model_or_ckpt_dir, subfolder=wan_subfolder)
        elif isinstance(wan_model_or_ckpt_dir, WanModel):
            self.wan_model = wan_model_or_ckpt_dir
        elif wan_config is not None:
            logging.info("Initializing WanModel with random weights from config (skip_init=%s)", skip_init)
            # reference unified_model.py:48-69
            model_params = {
                'patch_size': wan_config.get('patch_size', (1, 2, 2)),
                'dim': wan_config.get('dim', 2048),
                'ffn_dim': wan_config.get('ffn_dim', 8192),
                'freq_dim': wan_config.get('freq_dim', 256),
                'num_heads': wan_config.get('num_heads', 16),
                'num_layers': wan_config.get('num_layers', 32),
                'window_size': wan_config.get('window_size', (-1, -1)),
                'qk_norm': wan_config.get('qk_norm', True),
                'cross_attn_norm': wan_config.get('cross_attn_norm', True),
                'eps': wan_config.get('eps', 1e-6),
            }
            for k in ('model_type', 'text_len', 'in_dim', 'text_dim', 'out_dim'):
                if k in wan_config:
                    model_params[k] = wan_config[k]
            self.wan_model = WanModel(**model_params, skip_init=skip_init)
        else:
            raise ValueError("Either wan_model_or_ckpt_dir or wan_config must be provided")

        if max_context_len is not None:
            self.wan_model.text_len = max_context_len
        self.max_context_len = self.wan_model.text_len

        self.vlm_norm = WanRMSNorm(vlm_in_dim, eps=eps)
        self.vlm_proj = nn.Linear(vlm_in_dim, self.wan_model.text_dim)

        if use_visual_context_adapter:
            self.visual_context_adapter = VisualContextAdapter(
                patch_size=(self.wan_model.patch_size if visual_context_adapter_patch_size is None
                            else visual_context_adapter_patch_size),
                in_channels=self.wan_model.in_dim,
                hidden_dim=self.wan_model.dim,
                out_dim=self.wan_model.text_dim,
                eps=eps,
            )
        else:
            self.visual_context_adapter = None
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
        shp = tokens.shape
        h = tokens.reshape(-1, shp[-1]).to(bf16).contiguous().clone()
        ops.rmsnorm_rope_(h, self.vlm_norm.weight.data.to(bf16).contiguous(), self.vlm_norm.eps)
        out = ops.gemm(h, self.vlm_proj.weight.data.to(bf16).contiguous(),
                       self.vlm_proj.bias.data.to(bf16).contiguous())
        return out.view(*shp[:-1], out.shape[-1])

    def _params_key(self):
        ps = [self.vlm_norm.weight, self.vlm_proj.weight, self.vlm_proj.bias]
        if self.visual_context_adapter is not None:
            ps += list(self.visual_context_adapter.parameters())
        return tuple((p.data_ptr(), p._version) for p in ps)

    @torch.no_grad()
    def mixed_context(self, batch_size, context=None, ar_vision_input=None, visual_emb=None,
                      special_token_dict=None, device=None) -> List[torch.Tensor]:
        """Steps 2-4 of reference unified_model.py:192-276 -> list of bf16 [<= max_context_len, text_dim].
        (The reference's list entries are fp32 after `torch.cat` promotion; `text_embedding` rounds
        them to bf16 under autocast before its first GEMM — the values that reach the GEMM are the same.)"""
        key = (batch_size, _ident(context), _ident(ar_vision_input), _ident(visual_emb),
               _ident(special_token_dict), self._params_key(), self.max_context_len)
        hit = self._ctx_cache.get(key)
        if hit is not None:
            self.context_cache_hits += 1
            return hit
        self.context_cache_misses += 1

        vlm_output = None if ar_vision_input is None else [None] * batch_size
        if ar_vision_input is not None:
            if isinstance(ar_vision_input, list):
                for idx in range(batch_size):
                    if ar_vision_input[idx] is not None:
                        vlm_output[idx] = self._vlm(ar_vision_input[idx])
            else:
                vlm_output = self._vlm(ar_vision_input)

        processed_visual = None if visual_emb is None else [None] * batch_size
        if visual_emb is not None and self.visual_context_adapter is not None:
            if isinstance(visual_emb, list):
                for idx in range(batch_size):
                    if visual_emb[idx] is not None:
                        processed_visual[idx] = self.visual_context_adapter(visual_emb[idx]).squeeze(0)
            else:
                processed_visual = self.visual_context_adapter(visual_emb)

        mixed = []
        for idx in range(batch_size):
            vlm_item = _item(vlm_output, idx, batch_size)
            context_item = _item(context, idx, batch_size)
            visual_item = _item(processed_visual, idx, batch_size)
            components = []
            if special_token_dict is not None:
                img_st = _item(special_token_dict.get('<img_st>'), None, batch_size)
                img_ed = _item(special_token_dict.get('<img_ed>'), None, batch_size)
                prp_st = _item(special_token_dict.get('<prp_st>'), None, batch_size)
                prp_ed = _item(special_token_dict.get('<prp_ed>'), None, batch_size)
                if vlm_item is not None:
                    components.append(vlm_item)
                if context_item is not None:
                    components.extend([prp_st, context_item, prp_ed])
                if visual_item is not None:
                    components.extend([img_st, visual_item, img_ed])
            else:
                components = [c for c in (vlm_item, context_item, visual_item) if c is not None]
            components = [c for c in components if c is not None]
            if components:
                dev = device if device is not None else components[0].device
                new_context = torch.cat([c.to(device=dev, dtype=bf16) for c in components], dim=0)
                if self.max_context_len is not None and new_context.shape[0] > self.max_context_len:
                    new_context = new_context[:self.max_context_len]
                mixed.append(new_context)
            else:
                mixed.append(torch.zeros(1, self.wan_model.text_dim, device=device, dtype=bf16))
        if len(self._ctx_cache) >= 8:          # two entries per prompt (cond / uncond) are live at a time
            self._ctx_cache.pop(next(iter(self._ctx_cache)))
        self._ctx_cache[key] = mixed
        return mixed

    def forward(
        self,
        x: List[torch.Tensor],
        t: torch.Tensor,
        context: List[torch.Tensor] = None,
        aligned_emb: Optional[torch.Tensor] = None,
        ar_vision_input: Optional[List[torch.Tensor]] = None,
        visual_emb: Optional[torch.Tensor] = None,
        seq_len: Optional[int] = None,
        special_token_dict: Optional[Dict[str, torch.Tensor]] = None,
        classifier_free_ratio: Optional[float] = None,
        unconditioned_context: Optional[Dict[str, torch.Tensor]] = None,
        condition_mode: str = "auto",
    ):
        if condition_mode == "auto":
            condition_mode = "full"
        batch_size = x.size(0) if isinstance(x, torch.Tensor) else len(x)
        if condition_mode != "full":
            # the reference's forward has no other branch: it falls off the end and returns None, which
            # its caller then indexes.  Fail here with the reason instead.
            raise ValueError(f"condition_mode={condition_mode!r}: only 'full' / 'auto' are implemented "
                             "(reference unified_model.py:152-158 has no other branch either)")

        # 1. condition dropout (training-time CFG).  Both draws happen for every sample on every
        #    call, whether used or not (reference unified_model.py:160-162).
        cfg_rands = [random.random() for _ in range(batch_size)]
        _ar_only_rands = [random.random() for _ in range(batch_size)]  # threshold is 0.00 upstream: never taken
        for idx in range(batch_size):
            r = cfg_rands[idx]
            if classifier_free_ratio is not None and classifier_free_ratio > 0 and r < classifier_free_ratio:
                if unconditioned_context is None:
                    raise ValueError("unconditioned_context must be provided when classifier_free_ratio > 0")
                if isinstance(unconditioned_context, dict):
                    if ar_vision_input is not None and ar_vision_input[idx] is not None:
                        v = ar_vision_input[idx]
                        ar_vision_input[idx] = torch.zeros(2, v.size(1), device=v.device, dtype=v.dtype) + 1e-6
                    if context is not None and context[idx] is not None and 'uncond_context' in unconditioned_context:
                        context[idx] = unconditioned_context['uncond_context']

        dev = x[0].device
        mixed = self.mixed_context(batch_size, context, ar_vision_input, visual_emb, special_token_dict, dev)
        return self.wan_model(x, t=t, context=mixed, seq_len=seq_len)

    # ------------------------------------------------------------------ persistence (reference :280-285)
    def save_pretrained(self, save_dir: str):
        os.makedirs(save_dir, exist_ok=True)
        self.wan_model.save_pretrained(os.path.join(save_dir, "wan_model"))
        if self.visual_context_adapter is not None:
            torch.save(self.visual_context_adapter.state_dict(),
                       os.path.join(save_dir, "visual_context_adapter_pytorch_model.bin"))
        torch.save(self.vlm_norm.state_dict(), os.path.join(save_dir, "vlm_norm_pytorch_model.bin"))
        torch.save(self.vlm_proj.state_dict(), os.path.join(save_dir, "vlm_proj_pytorch_model.bin"))

    def enable_eval(self):
        self.eval()

    def enable_train(self):
        raise NotImplementedError("the B200 path is inference-only (SURVEY.md §8(b) Threading)")


def build_mixed_context(model: UnifiedWanWithMixedConditionModel, **kw):
    """Functional form used by patch.patch_reference_unified()."""
    return model.mixed_context(**kw)
