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
