"""Pipeline trace of one attention CTA (debug build scripts/microbench/build/libovb_trace.so, compiled with
-DOVB_ATTN_TRACE by scripts/build_trace_lib.sh): clock64() stamps of the MMA thread and of one softmax lane
for KV tiles 64..71, printed relative to the first stamp.
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omni_video_b200 import _lib  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "scripts", "microbench", "build", os.environ.get("OVB_TRACE_LIB", "libovb_trace.so")))
lib.ovb_attn_fwd.restype = C.c_int
lib.ovb_attn_fwd.argtypes = [C.POINTER(_lib.AttnArgs), C.c_void_p]
lib.ovb_attn_set_trace.argtypes = [C.c_void_p]
bf16 = torch.bfloat16
L, H = int(os.environ.get("OVB_TRACE_L", "32760")), int(os.environ.get("OVB_TRACE_H", "12"))
LQ = int(os.environ.get("OVB_TRACE_LQ", str(L)))
q = torch.randn(1, LQ, H, 128, device="cuda").to(bf16)
k, v = (torch.randn(1, L, H, 128, device="cuda").to(bf16) for _ in range(2))
out = torch.empty_like(q)
trace = torch.zeros(8 * 32, dtype=torch.int64, device="cuda")
lib.ovb_attn_set_trace(C.c_void_p(trace.data_ptr()))
a = _lib.AttnArgs()
for name, t in (("q", q), ("k", k), ("v", v), ("out", out)):
    setattr(a, name, t.data_ptr())
a.q_row_stride = a.k_row_stride = a.v_row_stride = a.o_row_stride = H * 128
a.k_batch_stride = a.v_batch_stride = L * H * 128
a.q_batch_stride = a.o_batch_stride = LQ * H * 128
a.B, a.H, a.Lq, a.Lk = 1, H, LQ, L
for _ in range(2):
    rc = lib.ovb_attn_fwd(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for i in range(5):
    evs[i].record()
    lib.ovb_attn_fwd(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream))
evs[5].record()
torch.cuda.synchronize()
ms = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(5))[2]
print(f"{os.environ.get('OVB_TRACE_LIB', 'libovb_trace.so')}: L={L} H={H}: {ms:.3f} ms, {4.0 * LQ * L * H * 128 / ms / 1e9:.0f} TF/s (trace build)")
t = trace.cpu().view(8, 32)
t0 = int(t[0, 0])
names = {0: "MMA  QK(j) issued", 1: "MMA  P(j) NOT ready after QK issue", 2: "MMA  P(j) ready, issuing PV", 3: "MMA  PV(j) issued", 4: "MMA  spin for P(j) done", 5: "MMA  fenced, P(j) usable"}
for q in range(4):
    names[16 + q] = f"SM{q}  S(j) seen"
    names[20 + q] = f"SM{q}  max known"
    names[24 + q] = f"SM{q}  decided"
    names[28 + q] = f"SM{q}  P published"
ev = []
for j in range(8):
    for slot, nm in names.items():
        if int(t[j, slot]) != 0:
            ev.append((int(t[j, slot]) - t0, j + 64, nm))
ev.sort()
prev = 0
for ts, j, nm in ev:
    print(f"{ts:8d} (+{ts - prev:5d})  j={j}  {nm}")
    prev = ts
per_step = (int(t[7, 2]) - int(t[0, 2])) / 7.0
print(f"cycles per KV tile (PV issue -> PV issue): {per_step:.0f}")
for j in range(1, 7):
    for q in range(4):
        print(f"j={j+64} quad {q}: S seen->max known {int(t[j,20+q])-int(t[j,16+q])}, ->decided {int(t[j,24+q])-int(t[j,20+q])}, "
              f"->published {int(t[j,28+q])-int(t[j,24+q])}; published -> MMA issues PV {int(t[j,2])-int(t[j,28+q])}")
