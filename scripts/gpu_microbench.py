"""Kernel-level timings on the B200 (CUDA events on torch's current stream, which is the stream the
C ABI launches on): our tcgen05 GEMM / attention against the library kernels the reference hits
(cuBLASLt through torch.matmul, flash_attn 2.8.3) on the DiT shapes of BASELINE.json.

    python scripts/gpu_microbench.py [--quick] [--json gpurun_out/microbench.json]
"""
from __future__ import annotations

import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omni_video_b200 import ops  # noqa: E402

bf16 = torch.bfloat16
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def time_fn(fn, iters=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_gemm(M, N, K, res):
    a = torch.randn(M, K, device="cuda").to(bf16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bf16)
    b = torch.randn(N, device="cuda").to(bf16)
    out = torch.empty(M, N, device="cuda", dtype=bf16)
    fl = 2.0 * M * N * K
    med, best = time_fn(lambda: ops.gemm(a, w, b, out=out))
    med_c, best_c = time_fn(lambda: torch.nn.functional.linear(a, w, b))
    r = dict(kind="gemm", M=M, N=N, K=K, ovb_ms=med, ovb_tflops=fl / med / 1e9, cublas_ms=med_c,
             cublas_tflops=fl / med_c / 1e9)
    print(f"gemm {M}x{N}x{K}: ovb {med:.3f} ms {r['ovb_tflops']:.0f} TF/s | cuBLAS {med_c:.3f} ms "
          f"{r['cublas_tflops']:.0f} TF/s", flush=True)
    res.append(r)


def bench_gemm_resid(M, N, K, res):
    """O-projection / FFN2 form: fp32 residual stream updated in place with a per-sample gate."""
    a = torch.randn(M, K, device="cuda").to(bf16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bf16)
    b = torch.randn(N, device="cuda").to(bf16)
    x = torch.randn(M, N, device="cuda")
    gate = torch.randn(1, N, device="cuda")
    fl = 2.0 * M * N * K
    med, _ = time_fn(lambda: ops.gemm(a, w, b, epilogue=ops.EPI_GATE_RESID_F32, out=x, resid=x, gate=gate,
                                      rows_per_sample=M))
    r = dict(kind="gemm_gate_resid_f32", M=M, N=N, K=K, ovb_ms=med, ovb_tflops=fl / med / 1e9)
    print(f"gemm+gate+resid(fp32) {M}x{N}x{K}: ovb {med:.3f} ms {r['ovb_tflops']:.0f} TF/s "
          f"(epilogue moves {M * N * 8 / 1e6:.0f} MB)", flush=True)
    res.append(r)


def bench_attn(L, H, Lk, res):
    q = torch.randn(1, L, H, 128, device="cuda").to(bf16)
    k = torch.randn(1, Lk, H, 128, device="cuda").to(bf16)
    v = torch.randn(1, Lk, H, 128, device="cuda").to(bf16)
    out = torch.empty_like(q)
    fl = 4.0 * L * Lk * H * 128
    med, best = time_fn(lambda: ops.attention(q, k, v, out=out), iters=5, warmup=2)
    r = dict(kind="attn", L=L, H=H, Lk=Lk, ovb_ms=med, ovb_tflops=fl / med / 1e9)
    line = f"attn L={L} Lk={Lk} H={H}: ovb {med:.3f} ms {r['ovb_tflops']:.0f} TF/s"
    items = ((L + 255) // 256) * H
    if items % 74 and Lk >= 4096:
        med_w, _ = time_fn(lambda: ops.attention(q, k, v, out=out, split_keys=False), iters=5, warmup=2)
        r.update(whole_items_ms=med_w, whole_items_tflops=fl / med_w / 1e9)
        line += f" (one CTA per item: {med_w:.3f} ms {fl / med_w / 1e9:.0f} TF/s; {items} items on 74 CTA pairs)"
    try:
        import flash_attn
        fa = lambda: flash_attn.flash_attn_func(q, k, v)  # noqa: E731
        med_f, _ = time_fn(fa, iters=5, warmup=2)
        r.update(fa2_ms=med_f, fa2_tflops=fl / med_f / 1e9)
        line += f" | flash_attn2 {med_f:.3f} ms {r['fa2_tflops']:.0f} TF/s"
        err = (fa().float() - out.float()).abs().max().item()
        line += f" | max|diff| vs FA2 {err:.2e}"
        r["max_abs_diff_vs_fa2"] = err
    except Exception as ex:  # pragma: no cover
        line += f" | flash_attn unavailable: {type(ex).__name__}"
    print(line, flush=True)
    res.append(r)


def bench_elementwise(L, d, res):
    x = torch.randn(L, d, device="cuda")
    tbl = torch.randn(1, 6, d, device="cuda")
    out = torch.empty(L, d, device="cuda", dtype=bf16)
    med, _ = time_fn(lambda: ops.layernorm_modulate(x, tbl[:, 1], tbl[:, 0], 1e-6, out=out))
    gbs = L * d * 6 / med / 1e6
    print(f"layernorm_modulate [{L},{d}]: {med:.3f} ms {gbs:.0f} GB/s (6 B/elem)", flush=True)
    res.append(dict(kind="ln_modulate", L=L, d=d, ms=med, gbs=gbs))
    qkv = torch.randn(L, 3 * d, device="cuda").to(bf16)
    w = torch.ones(d, device="cuda", dtype=bf16)
    cs = ops.rope_cos_sin_table().cuda()
    grid = (21, 30, 52) if L >= 32760 else (1, 1, 1)
    med, _ = time_fn(lambda: ops.rmsnorm_rope_(qkv[:, :d], w, 1e-6, cs_table=cs, grid=grid, rows_per_sample=L))
    gbs = L * d * 4 / med / 1e6
    print(f"rmsnorm_rope [{L},{d}]: {med:.3f} ms {gbs:.0f} GB/s (4 B/elem)", flush=True)
    res.append(dict(kind="rmsnorm_rope", L=L, d=d, ms=med, gbs=gbs))


def main():
    quick = "--quick" in sys.argv
    res = []
    L = 32760
    shapes = [(L, 4608, 1536), (L, 1536, 1536), (L, 8960, 1536), (L, 1536, 8960)]
    if not quick:
        shapes += [(8192, 8192, 8192), (L, 15360, 5120), (L, 5120, 5120), (L, 13824, 5120), (L, 5120, 13824)]
    attn_only = "--attn-only" in sys.argv
    if not attn_only:
        for s in shapes:
            bench_gemm(*s, res)
        bench_gemm_resid(L, 1536, 1536, res)
        bench_gemm_resid(L, 1536, 8960, res)
    bench_attn(4096, 12, 4096, res)
    bench_attn(L, 12, L, res)
    bench_attn(L, 12, 512, res)
    bench_attn(L, 3, L, res)    # per-rank shape of cfg2 x ulysses4 on 8 GPUs: 384 items = 2.6 waves
    bench_attn(L, 6, L, res)    # cfg2 x ulysses2 on 4 GPUs: 768 items = 5.2 waves
    if not quick:
        bench_attn(L, 40, L, res)
    if not attn_only:
        bench_elementwise(L, 1536, res)
    else:
        bench_attn(L, 12, 6144, res)
        bench_attn(75600, 5, 75600, res)   # A14B 720p per-rank shape on 8 GPUs
    if "--json" in sys.argv:
        path = sys.argv[sys.argv.index("--json") + 1]
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
