"""Shared helpers: load a golden fixture and rebuild its weights/inputs from the recorded seeds."""
import glob
import json
import os

import numpy as np
import torch

from oracle import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(kind):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        with np.load(p, allow_pickle=False) as z:
            if json.loads(str(z["meta"]))["kind"] == kind:
                out.append(os.path.basename(p)[:-4])
    return out


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    outs = [torch.from_numpy(z[f"out{i}"]) for i in range(len([k for k in z.files if k.startswith("out")]))]
    return meta, outs


def dit_case(meta):
    cfg = dict(meta["cfg"])
    cfg["patch_size"] = tuple(cfg["patch_size"])
    cfg["window_size"] = tuple(cfg["window_size"])
    sd = synth.dit_state_dict(cfg, seed=meta["weight_seed"])
    xs, t, ctx = synth.dit_inputs(cfg, tuple(meta["grid"]), meta["ctx_len"], seed=meta["input_seed"],
                                  batch=meta["batch"], timestep=meta["t"])
    return cfg, sd, xs, t, ctx, meta["seq_len"]
