"""ORACLE support — goldens of the conditioning front-end from the UNMODIFIED reference
(omnivideo/modules/unified_model.py + visual_context_adapter.py, imported by file path).
Run in the build container:   python -m oracle.gen_golden_unified

The reference's `UnifiedWanWithMixedConditionModel.forward` is run on CPU in fp32 with its
`wan_model.forward` replaced by a recorder, so the fixture holds exactly the `mixed_context` the
reference hands to `WanModel.forward` (unified_model.py:278).
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader, synth  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # all three conditions + special tokens, nothing truncated
    dict(name="unified_full", vlm_in_dim=128, n_vlm=[9], n_text=[12], visual=(3, 8, 12), special=True,
         max_len=None, adapter=True, patch=(1, 4, 4), text_len=64),
    # truncation inside the visual part, no special tokens
    dict(name="unified_truncate", vlm_in_dim=128, n_vlm=[5], n_text=[7], visual=(2, 8, 8), special=False,
         max_len=16, adapter=True, patch=(1, 4, 4)),
    # batch of 2 with ragged text, one sample without VLM tokens, no adapter, special tokens
    dict(name="unified_batch2_noadapter", vlm_in_dim=64, n_vlm=[6, None], n_text=[10, 3], visual=None,
         special=True, max_len=None, adapter=False, patch=None),
    # adapter with the DiT's own patch size and a latent that the strided conv floors (H=9 -> 4 rows)
    dict(name="unified_patch122_floor", vlm_in_dim=128, n_vlm=[None], n_text=[4], visual=(2, 9, 6), special=True,
         max_len=None, adapter=True, patch=(1, 2, 2)),
]


def run_case(ref_unified, case):
    cfg = dict(synth.TINY_DIT, num_layers=1, text_len=case.get("text_len", 32))
    wan_config = dict(cfg)
    random.seed(0)
    torch.manual_seed(0)
    model = ref_unified.UnifiedWanWithMixedConditionModel(
        wan_config=wan_config, vlm_in_dim=case["vlm_in_dim"], use_visual_context_adapter=case["adapter"],
        visual_context_adapter_patch_size=case["patch"], max_context_len=case["max_len"], skip_init=True)
    sd = synth.unified_state_dict(cfg, case["vlm_in_dim"], case["patch"] or (1, 4, 4), seed=1818,
                                  with_adapter=case["adapter"])
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("wan_model.") for k in missing), (missing, unexpected)
    model.eval().requires_grad_(False)
    captured = {}

    def recorder(x, t=None, context=None, seq_len=None):
        captured["context"] = [c.clone() for c in context]
        return [torch.zeros(1)]

    model.wan_model.forward = recorder
    ar, ctx, vis, st = synth.unified_inputs(cfg, case["vlm_in_dim"], case["n_vlm"], case["n_text"], case["visual"],
                                            seed=7, special_tokens=case["special"])
    B = len(case["n_text"])
    x = [torch.zeros(16, 1, 2, 2) for _ in range(B)]
    with torch.no_grad():
        model(x, t=torch.zeros(B, dtype=torch.int64), context=ctx, ar_vision_input=ar, visual_emb=vis, seq_len=1,
              special_token_dict=st)
    return cfg, [c.float().numpy() for c in captured["context"]], model.max_context_len


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref_loader.load_reference(cpu_attention=True)
    ref_loader._load("omnivideo.modules.visual_context_adapter", "omnivideo/modules/visual_context_adapter.py")
    ref_unified = ref_loader._load("omnivideo.modules.unified_model", "omnivideo/modules/unified_model.py")
    for case in CASES:
        cfg, outs, max_len = run_case(ref_unified, case)
        meta = dict(kind="unified_context", cfg=cfg, vlm_in_dim=case["vlm_in_dim"], n_vlm=case["n_vlm"],
                    n_text=case["n_text"], visual=case["visual"], special=case["special"], max_len=max_len,
                    adapter=case["adapter"], patch=case["patch"], weight_seed=1818, input_seed=7,
                    reference="omnivideo/modules/unified_model.py UnifiedWanWithMixedConditionModel.forward, fp32 CPU, "
                              "mixed_context captured at the wan_model call (:278)")
        path = os.path.join(GOLDEN_DIR, case["name"] + ".npz")
        np.savez_compressed(path, meta=json.dumps(meta), **{f"out{i}": o for i, o in enumerate(outs)})
        print(f"wrote {path}: shapes {[o.shape for o in outs]}")


if __name__ == "__main__":
    main()
