"""Times ovb_gemm_bf16 epilogue variants at the 1.3B FFN shapes (is the epilogue hidden behind the main loop?)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omni_video_b200 import ops  # noqa: E402

bf16 = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, n=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[n // 2]


for M, N, K in [(65520, 8960, 1536), (65520, 4608, 1536), (65520, 1536, 8960), (65520, 1536, 1536)]:
    a = torch.randn(M, K, device="cuda").to(bf16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bf16)
    b = torch.randn(N, device="cuda").to(bf16)
    o = torch.empty(M, N, device="cuda", dtype=bf16)
    x = torch.randn(M, N, device="cuda")
    g = torch.randn(2, N, device="cuda")
    fl = 2.0 * M * N * K
    for name, fn in [("bf16", lambda: ops.gemm(a, w, b, out=o)),
                     ("gelu", lambda: ops.gemm(a, w, b, epilogue=ops.EPI_GELU_BF16, out=o)),
                     ("gate+resid f32", lambda: ops.gemm(a, w, b, epilogue=ops.EPI_GATE_RESID_F32, out=x, resid=x, gate=g,
                                                         rows_per_sample=M // 2))]:
        if os.environ.get("GEMM_ONLY") and os.environ["GEMM_ONLY"] not in name:
            continue
        ms = t(fn)
        print(f"gemm [{M},{N},{K}] {name}: {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s", flush=True)


def back_to_back(fn, n=20):
    """Per-launch time of n launches issued back to back (no L2 flush: in the step the operands of these short
    GEMMs come out of the previous kernel and mostly sit in L2)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# Per-rank shapes of the sequence-parallel layouts (1.3B, 32760 tokens): M = 8190 rows at 8 GPUs (ulysses4 x cfg2),
# 16380 at 4.  `GEMM_RANK_SHAPES=1 python scripts/gemm_time.py`
if os.environ.get("GEMM_RANK_SHAPES"):
    only = os.environ.get("GEMM_EPI", "").split(",") if os.environ.get("GEMM_EPI") else None
    for M in (8190, 16380, 32760) + ((65520,) if only else ()):
        tot_ms, tot_fl = 0.0, 0.0
        for name, N, K, epi in [("qkv", 4608, 1536, "bf16"), ("o", 1536, 1536, "gate"), ("cross q", 1536, 1536, "bf16"),
                                ("cross o", 1536, 1536, "resid"), ("ffn1", 8960, 1536, "gelu"), ("ffn2", 1536, 8960, "gate")]:
            if only and epi not in only:
                continue
            a = torch.randn(M, K, device="cuda").to(bf16)
            w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bf16)
            b = torch.randn(N, device="cuda").to(bf16)
            o = torch.empty(M, N, device="cuda", dtype=bf16)
            x = torch.randn(M, N, device="cuda")
            g = torch.randn(1, N, device="cuda")
            fn = {"bf16": lambda: ops.gemm(a, w, b, out=o),
                  "gelu": lambda: ops.gemm(a, w, b, epilogue=ops.EPI_GELU_BF16, out=o),
                  "resid": lambda: ops.gemm(a, w, b, epilogue=ops.EPI_RESID_F32, out=x, resid=x),
                  "gate": lambda: ops.gemm(a, w, b, epilogue=ops.EPI_GATE_RESID_F32, out=x, resid=x, gate=g,
                                           rows_per_sample=M)}[epi]
            ms = back_to_back(fn)
            wt = w.t()
            ms_lib = ms if only else back_to_back(lambda: torch.matmul(a, wt, out=o))
            fl = 2.0 * M * N * K
            tot_ms += ms
            tot_fl += fl
            print(f"rank gemm M={M} {name:8s} [{N},{K}] {epi:5s}: {ms * 1e3:7.1f} us {fl / ms / 1e9:5.0f} TF/s | "
                  f"cuBLAS plain {ms_lib * 1e3:7.1f} us {fl / ms_lib / 1e9:5.0f} TF/s", flush=True)
        print(f"rank gemm M={M} layer total {tot_ms * 1e3:.1f} us = {tot_fl / tot_ms / 1e9:.0f} TF/s", flush=True)
