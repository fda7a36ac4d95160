"""Time the two warp-per-row norm kernels alone at the bench shapes (L2 flushed between launches) and check them
against fp32 torch: `python scripts/norm_time.py`.  Prints GB/s of algorithmic bytes (LayerNorm 6 B/elem,
RMSNorm+RoPE 4 B/elem) next to the measured copy bandwidth of MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omni_video_b200 import ops  # noqa: E402

bf16 = torch.bfloat16
_flush = None


def time_fn(fn, iters=20, warmup=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    peak = 6531.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    torch.manual_seed(0)
    for L, d, grid in ((65520, 1536, (21, 30, 52)), (32760, 1536, (21, 30, 52)), (75600, 5120, (21, 45, 80))):
        x = torch.randn(L, d, device="cuda")
        tbl = torch.randn(1, 6, d, device="cuda") * 0.1
        out = torch.empty(L, d, device="cuda", dtype=bf16)
        ms = time_fn(lambda: ops.layernorm_modulate(x, tbl[:, 1], tbl[:, 0], 1e-6, out=out))
        ref = torch.nn.functional.layer_norm(x[:4096], (d,), eps=1e-6) * (1 + tbl[0, 1]) + tbl[0, 0]
        err = (out[:4096].float() - ref).abs().max().item()
        gbs = L * d * 6 / ms / 1e6
        print(f"layernorm_modulate [{L},{d}]: {ms*1e3:.1f} us {gbs:.0f} GB/s = {gbs/peak:.2f} of copy bandwidth; max|err| {err:.2e}", flush=True)
        rows_per_sample = grid[0] * grid[1] * grid[2]
        if L % rows_per_sample:
            rows_per_sample = L
        qkv = torch.randn(L, 3 * d, device="cuda").to(bf16)
        w = (1 + 0.1 * torch.randn(d, device="cuda")).to(bf16)
        cs = ops.rope_cos_sin_table().cuda()
        ms = time_fn(lambda: ops.rmsnorm_rope_(qkv[:, :d], w, 1e-6, cs_table=cs, grid=grid, rows_per_sample=rows_per_sample))
        gbs = L * d * 4 / ms / 1e6
        print(f"rmsnorm_rope [{L},{d}] (in a [L,3d] buffer): {ms*1e3:.1f} us {gbs:.0f} GB/s = {gbs/peak:.2f}", flush=True)


if __name__ == "__main__":
    main()
