"""VAE H-strip parity on real GPUs (torchrun, one rank per GPU): decode and encode split into strips
with halo exchange must reproduce the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29541 scripts/vae_strip_check.py [--big | --config5 [--video 480p|720p]]

--config5: BASELINE.json configs[4] — encode + decode of an 81-frame clip (default 720x1280) as H-strips
over all ranks: device-timed (CUDA events, max over ranks), against the same work whole on one GPU, with
the strips-vs-whole difference.  Prints one JSON line on rank 0.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omni_video_b200.modules.vae2_1 import Wan2_1_VAE  # noqa: E402


def _timed(fn, world, dev, reps=3):
    ts = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return min(ts[1:]), out


def config5(vae, g, rank, world, dev):
    video = sys.argv[sys.argv.index("--video") + 1] if "--video" in sys.argv else "720p"
    T, h, w = {"480p": (21, 60, 104), "720p": (21, 90, 160)}[video]
    z = torch.randn((16, T, h, w), generator=g, device=dev)
    v = torch.rand((3, 1 + 4 * (T - 1), 8 * h, 8 * w), generator=g, device=dev) * 2 - 1
    res = dict(config=f"Wan2.1 VAE encode+decode of 81x{8 * h}x{8 * w} (latent 16x{T}x{h}x{w}), bf16 kernels, "
                      f"{world} H-strips", n_gpus=world)
    one_dec, ref_dec = _timed(lambda: vae.decode([z])[0], 1, dev)
    one_enc, ref_enc = _timed(lambda: vae.encode([v])[0], 1, dev)
    res.update(decode_ms_1gpu=round(one_dec, 1), encode_ms_1gpu=round(one_enc, 1))
    if world > 1:
        vae.set_strip_group(dist.new_group(list(range(world))))
        dec, out = _timed(lambda: vae.decode([z])[0], world, dev)
        enc, mu = _timed(lambda: vae.encode([v])[0], world, dev)
        d = torch.stack([(out - ref_dec).abs().max(), (mu - ref_enc).abs().max()])
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        res.update(decode_ms=round(dec, 1), encode_ms=round(enc, 1),
                   decode_efficiency=round(one_dec / dec / world, 3), encode_efficiency=round(one_enc / enc / world, 3),
                   strips_vs_whole_max_abs=dict(decode=float(f"{d[0].item():.3e}"), encode=float(f"{d[1].item():.3e}")))
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1818)
    vae = Wan2_1_VAE(vae_pth=None, device=dev)
    g = torch.Generator(device="cuda").manual_seed(1818)
    for name, p in vae.model.named_parameters():  # same weights on every rank
        if p.dim() > 1 and "gamma" not in name:
            p.data.copy_(torch.randn(p.shape, generator=g, device=dev) / p[0].numel() ** 0.5)
    if "--config5" in sys.argv:
        return config5(vae, g, rank, world, dev)
    big = "--big" in sys.argv
    grid = (21, 60, 104) if big else (3, 6 * world + 1, 8)      # uneven strips in the small case
    z = torch.randn((16,) + grid, generator=g, device=dev)
    ref = vae.decode([z])[0]
    grp = dist.new_group(list(range(world)))
    vae.set_strip_group(grp)
    out = vae.decode([z])[0]
    d_dec = (out - ref).abs().max().item()
    ok = out.shape == ref.shape and d_dec <= 2e-2
    if big:
        for _ in range(2):
            torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
            vae.decode([z]); torch.cuda.synchronize(); dist.barrier()
            t_strip = time.perf_counter() - t0
        vae.set_strip_group(None)
        torch.cuda.synchronize(); t0 = time.perf_counter(); vae.decode([z]); torch.cuda.synchronize()
        t_one = time.perf_counter() - t0
        if rank == 0:
            print(f"decode 81x480x832: 1 GPU {1e3 * t_one:.1f} ms, {world} GPUs (strips) {1e3 * t_strip:.1f} ms", flush=True)
    else:
        vae.set_strip_group(None)
        v = torch.rand((3, 9, 16 * world, 32), generator=g, device=dev) * 2 - 1
        mu_ref = vae.encode([v])[0]
        vae.set_strip_group(grp)
        mu = vae.encode([v])[0]
        d_enc = (mu - mu_ref).abs().max().item()
        ok = ok and mu.shape == mu_ref.shape and d_enc <= 2e-2
        print(f"[rank {rank}/{world}] strips vs single GPU: decode max|diff| {d_dec:.3e}, encode {d_enc:.3e}", flush=True)
    if big:
        print(f"[rank {rank}/{world}] big decode max|diff| {d_dec:.3e}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
