"""ORACLE support — goldens of the umT5 encoder from the UNMODIFIED reference (omnivideo/modules/t5.py,
imported by file path).  Run in the build container:   python -m oracle.gen_golden_t5

Import needs two shims that do not touch the arithmetic: `omnivideo.modules.tokenizers` (ftfy / HF
tokenizers are absent here; only the encoder is exercised) and `torch.cuda.current_device`, which the
reference evaluates as a default argument at class-definition time (t5.py:477)."""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader, synth  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    dict(name="t5_tiny_ragged", cfg={}, lengths=[9, 20], seq_len=24),          # padded keys must be masked out
    dict(name="t5_tiny_long", cfg={}, lengths=[200], seq_len=200),             # distances beyond max_dist = 128
    dict(name="t5_tiny_shared_pos", cfg=dict(shared_pos=True, num_layers=1), lengths=[7], seq_len=12),
]


def load_t5():
    ref_loader.load_reference(cpu_attention=True)
    tok = types.ModuleType("omnivideo.modules.tokenizers")
    tok.HuggingfaceTokenizer = object
    sys.modules["omnivideo.modules.tokenizers"] = tok
    real = torch.cuda.current_device
    torch.cuda.current_device = lambda: 0
    try:
        return ref_loader._load("omnivideo.modules.t5", "omnivideo/modules/t5.py")
    finally:
        torch.cuda.current_device = real


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    t5 = load_t5()
    for case in CASES:
        cfg = dict(synth.TINY_T5)
        cfg.update(case["cfg"])
        model = t5.T5Encoder(**cfg, dropout=0.1).eval().requires_grad_(False)
        sd = synth.t5_state_dict(cfg, seed=1818)
        model.load_state_dict(sd, strict=True)
        ids, mask = synth.t5_inputs(cfg, case["lengths"], case["seq_len"], seed=7)
        with torch.no_grad():
            ctx = model(ids, mask)
        outs = [u[:int(v)].numpy() for u, v in zip(ctx, mask.gt(0).sum(dim=1))]    # T5EncoderModel.__call__ :510-512
        meta = dict(kind="t5_encode", cfg=cfg, lengths=case["lengths"], seq_len=case["seq_len"], weight_seed=1818,
                    input_seed=7, reference="omnivideo/modules/t5.py T5Encoder.forward, fp32 CPU, eval mode")
        path = os.path.join(GOLDEN_DIR, case["name"] + ".npz")
        np.savez_compressed(path, meta=json.dumps(meta), **{f"out{i}": o for i, o in enumerate(outs)})
        print(f"wrote {path}: shapes {[o.shape for o in outs]} absmax {max(float(np.abs(o).max()) for o in outs):.3f}")


if __name__ == "__main__":
    main()
