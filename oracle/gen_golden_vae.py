"""ORACLE support — VAE goldens from the UNMODIFIED reference (its chunked, cached implementation,
vae2_1.py:515-567) on oracle/synth.py's weights.  Run through `python -m oracle.gen_golden`."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle import synth
from oracle.vae import LATENT_MEAN, LATENT_STD, VAE_CFG

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    dict(name="vae_dec_t3", kind="vae_decode", grid=(3, 4, 6)),      # -> [3, 9, 32, 48]
    dict(name="vae_dec_t1", kind="vae_decode", grid=(1, 6, 4)),      # single frame: no time_conv
    dict(name="vae_dec_t5", kind="vae_decode", grid=(5, 2, 4)),
    dict(name="vae_enc_t9", kind="vae_encode", shape=(9, 32, 48)),   # -> [16, 3, 4, 6]
    dict(name="vae_enc_t1", kind="vae_encode", shape=(1, 48, 32)),
    dict(name="vae_enc_t13", kind="vae_encode", shape=(13, 16, 32)),
]


# BASELINE.json configs[0]'s video: latent [16, 5, 32, 32] -> [3, 17, 256, 256].  The fixture keeps every
# 4th pixel in H and W (all frames, all channels): 0.8 MB instead of 13; PSNR over the kept positions.
BIG_CASE = dict(name="vae_dec_config1", kind="vae_decode_strided", grid=(5, 32, 32), stride=4)


def main_big(ref):
    sd = synth.vae_state_dict(seed=1818)
    model = ref.vae.WanVAE_(**VAE_CFG)
    model.load_state_dict(sd, strict=True)
    model.eval().requires_grad_(False)
    scale = [torch.tensor(LATENT_MEAN), 1.0 / torch.tensor(LATENT_STD)]
    case = BIG_CASE
    with torch.no_grad():
        z = synth.vae_latent(case["grid"], seed=7)
        out = model.decode(z.unsqueeze(0), scale).float().clamp_(-1, 1).squeeze(0)
    st = case["stride"]
    kept = out[:, :, ::st, ::st].contiguous()
    meta = dict(case, weight_seed=1818, input_seed=7, full_shape=list(out.shape),
                reference="omnivideo/modules/vae2_1.py WanVAE_.decode (chunked, feat_cache), fp32 CPU; "
                          "out0 = decoded[:, :, ::4, ::4]")
    path = os.path.join(GOLDEN_DIR, case["name"] + ".npz")
    np.savez_compressed(path, meta=json.dumps(meta), out0=kept.numpy())
    print(f"wrote {path}: full {tuple(out.shape)} kept {tuple(kept.shape)} absmax {out.abs().max():.4f} std {out.std():.4f}")


def main(ref):
    sd = synth.vae_state_dict(seed=1818)
    model = ref.vae.WanVAE_(**VAE_CFG)
    model.load_state_dict(sd, strict=True)
    model.eval().requires_grad_(False)
    scale = [torch.tensor(LATENT_MEAN), 1.0 / torch.tensor(LATENT_STD)]
    for case in CASES:
        with torch.no_grad():
            if case["kind"] == "vae_decode":
                z = synth.vae_latent(case["grid"], seed=7)
                out = model.decode(z.unsqueeze(0), scale).float().clamp_(-1, 1).squeeze(0)
            else:
                v = synth.vae_video(case["shape"], seed=7)
                out = model.encode(v.unsqueeze(0), scale).float().squeeze(0)
        meta = dict(case, weight_seed=1818, input_seed=7,
                    reference="omnivideo/modules/vae2_1.py WanVAE_.decode/encode (chunked, feat_cache), fp32 CPU")
        path = os.path.join(GOLDEN_DIR, case["name"] + ".npz")
        np.savez_compressed(path, meta=json.dumps(meta), out0=out.numpy())
        print(f"wrote {path}: shape {tuple(out.shape)} absmax {out.abs().max():.4f} std {out.std():.4f}")
