"""ORACLE support — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported by oracle/ref_loader.py) on the synthetic weights/inputs of
oracle/synth.py.  Run in the build container:   python -m oracle.gen_golden
(`python -m oracle.gen_golden --full` writes only the three big fixtures: full-depth 1.3B at BASELINE
config 1, the A14B width at 2 layers, and the VAE decode of config 1's latent.)

Each fixture stores only the recipe (config, seeds, shapes) and the reference's OUTPUT, so that
tests regenerate identical weights/inputs from the seeds and compare (a) the oracle restatement in
fp32 and (b) the CUDA path in bf16 against what the reference itself computed.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader, synth  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

DIT_CASES = [
    # name, cfg overrides, latent grid (F, H, W), ctx_len, seq_len (None = exact), batch, timestep
    dict(name="dit_tiny_a", cfg={}, grid=(3, 8, 12), ctx_len=20, seq_len=None, batch=1, t=999),
    dict(name="dit_tiny_pad", cfg={}, grid=(2, 6, 10), ctx_len=32, seq_len=40, batch=1, t=500),
    dict(name="dit_tiny_batch2", cfg={}, grid=(2, 8, 8), ctx_len=11, seq_len=None, batch=2, t=17),
    dict(name="dit_tiny_4head", cfg=dict(dim=512, num_heads=4, ffn_dim=768, num_layers=1),
         grid=(5, 8, 8), ctx_len=7, seq_len=None, batch=1, t=875),
]


def run_dit_case(ref, case):
    cfg = dict(synth.TINY_DIT)
    cfg.update(case["cfg"])
    sd = synth.dit_state_dict(cfg, seed=1818)
    model = ref.model.WanModel(**cfg)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    model.eval().requires_grad_(False)
    xs, t, ctx = synth.dit_inputs(cfg, case["grid"], case["ctx_len"], seed=7, batch=case["batch"],
                                  timestep=case["t"])
    Fr, H, W = case["grid"]
    L = Fr * (H // 2) * (W // 2)
    seq_len = case["seq_len"] or L
    with torch.no_grad():
        out = model(xs, t=t, context=ctx, seq_len=seq_len)
    return cfg, seq_len, [o.numpy() for o in out]


# image-to-video flavour (model.py:417-424, :446-447): `y` is concatenated to `x` along the channels before
# the patch embedding, `model_type='i2v'` makes it mandatory
I2V_CASE = dict(name="dit_tiny_i2v", cfg=dict(model_type="i2v", in_dim=20), grid=(2, 6, 8), ctx_len=9, y_channels=4,
                t=640)


def run_i2v_case(ref):
    case = I2V_CASE
    cfg = dict(synth.TINY_DIT)
    cfg.update(case["cfg"])
    model = ref.model.WanModel(**cfg)
    model.load_state_dict(synth.dit_state_dict(cfg, seed=1818), strict=True)
    model.eval().requires_grad_(False)
    xs, t, ctx = synth.dit_inputs(cfg, case["grid"], case["ctx_len"], seed=7, batch=1, timestep=case["t"])
    nx = cfg["in_dim"] - case["y_channels"]
    Fr, H, W = case["grid"]
    seq_len = Fr * (H // 2) * (W // 2)
    with torch.no_grad():
        out = model([u[:nx] for u in xs], t=t, context=ctx, seq_len=seq_len, y=[u[nx:] for u in xs])
    meta = dict(kind="dit_forward_i2v", cfg=cfg, grid=case["grid"], ctx_len=case["ctx_len"], seq_len=seq_len, batch=1,
                t=case["t"], y_channels=case["y_channels"], weight_seed=1818, input_seed=7,
                reference="omnivideo/modules/model.py WanModel.forward with y (model_type i2v), fp32 CPU")
    path = os.path.join(GOLDEN_DIR, case["name"] + ".npz")
    np.savez_compressed(path, meta=json.dumps(meta), out0=out[0].numpy())
    print(f"wrote {path}: out0 shape {tuple(out[0].shape)} absmax {out[0].abs().max():.4f}")


def run_ragged_case(ref):
    """Two samples with DIFFERENT latent grids in one call: every sample is zero-padded to seq_len
    (model.py:455-464), attends to its own tokens only and is rotated with its own (f, h, w)."""
    cfg = dict(synth.TINY_DIT)
    model = ref.model.WanModel(**cfg)
    model.load_state_dict(synth.dit_state_dict(cfg, seed=1818), strict=True)
    model.eval().requires_grad_(False)
    grids = [(2, 6, 8), (3, 4, 4)]                       # 24 and 12 tokens
    xa, t, ctx_a = synth.dit_inputs(cfg, grids[0], 9, seed=7, batch=1, timestep=321)
    xb, _, ctx_b = synth.dit_inputs(cfg, grids[1], 17, seed=8, batch=1, timestep=321)
    seq_len = 28
    with torch.no_grad():
        out = model(xa + xb, t=torch.cat([t, t]), context=ctx_a + ctx_b, seq_len=seq_len)
    meta = dict(kind="dit_forward_ragged", cfg=cfg, grids=grids, ctx_lens=[9, 17], seq_len=seq_len, t=321,
                weight_seed=1818, input_seeds=[7, 8],
                reference="omnivideo/modules/model.py WanModel.forward on a batch of two different grids, fp32 CPU")
    path = os.path.join(GOLDEN_DIR, "dit_tiny_ragged.npz")
    np.savez_compressed(path, meta=json.dumps(meta), out0=out[0].numpy(), out1=out[1].numpy())
    print(f"wrote {path}: shapes {[tuple(o.shape) for o in out]}")


# BASELINE.json configs[0] exactly: OmniVideo2-1.3B (30 layers, dim 1536, 12 heads), one forward on a
# 17x256x256 video = latent [16, 5, 32, 32] = 1280 tokens, T5 context [77, 4096], fp32 on CPU
# (reference model.py:417-524 end to end at FULL DEPTH: this is where bf16 error accumulates).
FULL_1_3B = dict(name="dit_1_3B_config1", grid=(5, 32, 32), ctx_len=77, t=999)


def run_full_1_3b(ref):
    case = FULL_1_3B
    cfg = dict(synth.DIT_1_3B)
    model = ref.model.WanModel(**cfg)
    model.load_state_dict(synth.dit_state_dict(cfg, seed=1818), strict=True)
    model.eval().requires_grad_(False)
    xs, t, ctx = synth.dit_inputs(cfg, case["grid"], case["ctx_len"], seed=7, batch=1, timestep=case["t"])
    Fr, H, W = case["grid"]
    seq_len = Fr * (H // 2) * (W // 2)
    with torch.no_grad():
        out = model(xs, t=t, context=ctx, seq_len=seq_len)
    meta = dict(kind="dit_forward_full", cfg=cfg, grid=case["grid"], ctx_len=case["ctx_len"], seq_len=seq_len,
                batch=1, t=case["t"], weight_seed=1818, input_seed=7,
                reference="omnivideo/modules/model.py WanModel.forward, OmniVideo2-1.3B config "
                          "(configs/wan_t2v_1_3B.py), BASELINE.json configs[0], fp32 CPU")
    path = os.path.join(GOLDEN_DIR, case["name"] + ".npz")
    np.savez_compressed(path, meta=json.dumps(meta), out0=out[0].numpy())
    print(f"wrote {path}: out0 shape {tuple(out[0].shape)} absmax {out[0].abs().max():.4f} "
          f"rms {out[0].pow(2).mean().sqrt():.4f}")


# OmniVideo2-A14B WIDTH (dim 5120, 40 heads, ffn 13824: configs/wan_t2v_A14B.py) at 2 layers and
# 2160 tokens (latent [16, 5, 36, 48]): the GEMM shapes (5120 x 15360 / 5120 / 13824) and the 40-head
# attention of BASELINE.json configs[2-3], at a depth the CPU reference finishes in a minute.
WIDE_A14B = dict(name="dit_A14B_width_2layers", grid=(5, 36, 48), ctx_len=40, t=400, num_layers=2, text_len=64)


def run_wide_a14b(ref):
    case = WIDE_A14B
    cfg = dict(synth.DIT_A14B, num_layers=case["num_layers"], text_len=case["text_len"])
    model = ref.model.WanModel(**cfg)
    model.load_state_dict(synth.dit_state_dict(cfg, seed=1818), strict=True)
    model.eval().requires_grad_(False)
    xs, t, ctx = synth.dit_inputs(cfg, case["grid"], case["ctx_len"], seed=7, batch=1, timestep=case["t"])
    Fr, H, W = case["grid"]
    seq_len = Fr * (H // 2) * (W // 2)
    with torch.no_grad():
        out = model(xs, t=t, context=ctx, seq_len=seq_len)
    meta = dict(kind="dit_forward_full", cfg=cfg, grid=case["grid"], ctx_len=case["ctx_len"], seq_len=seq_len,
                batch=1, t=case["t"], weight_seed=1818, input_seed=7,
                reference="omnivideo/modules/model.py WanModel.forward at the OmniVideo2-A14B width "
                          "(configs/wan_t2v_A14B.py: dim 5120, 40 heads, ffn 13824), 2 layers, fp32 CPU")
    path = os.path.join(GOLDEN_DIR, case["name"] + ".npz")
    np.savez_compressed(path, meta=json.dumps(meta), out0=out[0].numpy())
    print(f"wrote {path}: out0 shape {tuple(out[0].shape)} absmax {out[0].abs().max():.4f} "
          f"rms {out[0].pow(2).mean().sqrt():.4f}")


def run_pertoken_case(ref):
    """t of shape [B, seq_len] with a different timestep per token (model.py:466-476): every token gets
    its own time embedding and modulation row."""
    cfg = dict(synth.TINY_DIT)
    model = ref.model.WanModel(**cfg)
    model.load_state_dict(synth.dit_state_dict(cfg, seed=1818), strict=True)
    model.eval().requires_grad_(False)
    grid, seq_len = (2, 6, 8), 28                                   # 24 tokens + 4 padded rows
    xs, _, ctx = synth.dit_inputs(cfg, grid, 9, seed=7, batch=2, timestep=0)
    t = torch.stack([(torch.arange(seq_len) * 37 + 11 + 400 * b) % 1000 for b in range(2)]).to(torch.int64)
    with torch.no_grad():
        out = model(xs, t=t, context=ctx, seq_len=seq_len)
    meta = dict(kind="dit_forward_pertoken", cfg=cfg, grid=grid, ctx_len=9, seq_len=seq_len, batch=2,
                t_rule="(arange(seq_len) * 37 + 11 + 400 * b) % 1000", weight_seed=1818, input_seed=7,
                reference="omnivideo/modules/model.py WanModel.forward with t [B, seq_len], fp32 CPU")
    path = os.path.join(GOLDEN_DIR, "dit_tiny_pertoken.npz")
    np.savez_compressed(path, meta=json.dumps(meta), out0=out[0].numpy(), out1=out[1].numpy())
    print(f"wrote {path}: shapes {[tuple(o.shape) for o in out]}")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_loader.load_reference(cpu_attention=True)
    if "--full" in sys.argv:                 # ~6 GB of fp32 weights, ~1 minute: on request only
        run_full_1_3b(ref)
        run_wide_a14b(ref)
        from oracle import gen_golden_vae
        gen_golden_vae.main_big(ref)
        return
    run_i2v_case(ref)
    run_ragged_case(ref)
    run_pertoken_case(ref)
    for case in DIT_CASES:
        cfg, seq_len, outs = run_dit_case(ref, case)
        meta = dict(kind="dit_forward", cfg=cfg, grid=case["grid"], ctx_len=case["ctx_len"],
                    seq_len=seq_len, batch=case["batch"], t=case["t"], weight_seed=1818, input_seed=7,
                    reference="omnivideo/modules/model.py WanModel.forward, fp32 CPU, "
                              "flash_attention -> masked SDPA (oracle/ref_loader.py)")
        path = os.path.join(GOLDEN_DIR, case["name"] + ".npz")
        np.savez_compressed(path, meta=json.dumps(meta), **{f"out{i}": o for i, o in enumerate(outs)})
        print(f"wrote {path}: out0 shape {outs[0].shape} absmax {np.abs(outs[0]).max():.4f}")
    if "--no-vae" not in sys.argv:
        try:
            from oracle import gen_golden_vae
            gen_golden_vae.main(ref)
        except ImportError:
            pass


if __name__ == "__main__":
    main()
