"""VAE decode/encode timing on the B200 with a per-launch breakdown (CUDA events).

    python scripts/gpu_vae_bench.py [--video 480p|720p|256] [--frames 21] [--encode]
"""
import collections
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omni_video_b200 import ops  # noqa: E402
from omni_video_b200.modules.vae2_1 import Wan2_1_VAE  # noqa: E402

GRIDS = {"480p": (60, 104), "720p": (90, 160), "256": (32, 32)}


def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


def vae_decode_flops(T, h, w):
    """conv FLOPs of decode, 0.688 GFLOP/latent pixel (frame 0) + 2.162 per later frame
    (FlopCounterMode on the reference) + attention 4*(hw)^2*384 per frame."""
    return (0.688e9 + 2.162e9 * (T - 1)) * h * w + 4.0 * (h * w) ** 2 * 384 * T


def print_timeline():
    tl, ops.timeline = ops.timeline, None
    agg = collections.OrderedDict()
    for kind, work, shape, s, e in tl:
        a = agg.setdefault((kind,) + tuple(shape), [0, 0.0, 0.0])
        a[0] += 1; a[1] += s.elapsed_time(e); a[2] += work
    tot = sum(a[1] for a in agg.values())
    print(f"timed launches total {tot:.1f} ms")
    for k, (n, ms, work) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        rate = work / ms / 1e9
        unit = "TF/s" if k[0] in ("conv", "gemm", "attn") else "GB/s*1e-3"
        print(f"  {ms:8.2f} ms {100 * ms / tot:5.1f}% n={n:3d} {rate:8.1f} {unit}  {k}")


def main():
    video = arg("--video", "480p")
    T = int(arg("--frames", "21"))
    h, w = GRIDS[video]
    torch.manual_seed(1818)
    vae = Wan2_1_VAE(vae_pth=None, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1818)
    for name, p in vae.model.named_parameters():
        if p.dim() > 1 and "gamma" not in name:
            fan_in = p[0].numel()
            p.data.copy_(torch.randn(p.shape, generator=g, device="cuda") / fan_in ** 0.5)
    z = torch.randn((16, T, h, w), generator=g, device="cuda")
    res = {}
    for it in range(3):
        torch.cuda.synchronize()
        ops.timeline = [] if it == 2 else None
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = vae.decode([z])[0]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"decode {video} T={T}: {ms:.1f} ms (wall {1e3 * (time.perf_counter() - t0):.1f}) "
              f"{vae_decode_flops(T, h, w) / ms / 1e9:.0f} TF/s algorithmic, out {tuple(out.shape)} "
              f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
        res["decode_ms"] = ms
    print_timeline()
    del out
    if "--encode" in sys.argv:
        Tv = 1 + 4 * (T - 1)
        v = torch.rand((3, Tv, 8 * h, 8 * w), generator=g, device="cuda") * 2 - 1
        for it in range(3):
            torch.cuda.synchronize()
            ops.timeline = [] if it == 2 else None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mu = vae.encode([v])[0]
            e1.record()
            torch.cuda.synchronize()
            print(f"encode {video} T={Tv}: {e0.elapsed_time(e1):.1f} ms out {tuple(mu.shape)}", flush=True)
            res["encode_ms"] = e0.elapsed_time(e1)
        print_timeline()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"vae_bench_{video}.json"), "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    main()
