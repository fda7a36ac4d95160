"""Pipeline trace of one attention CTA (debug build scripts/microbench/build/libovb_trace.so, compiled with
-DOVB_ATTN_TRACE): clock64() stamps of the MMA thread and of one softmax lane per Q tile for KV
tiles 64..71, printed relative to the first stamp.
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
t = trace.cpu().view(8, 32)
t0 = int(t[0, 0])
names = {0: "MMA  issue QK0", 1: "MMA  QK0 issued", 2: "MMA  PV1(j-1) issued", 3: "MMA  QK1 issued",
         4: "MMA  P0 part seen", 5: "MMA  PV0 issued", 6: "MMA  P1(j-1) part seen",
         8: "SM0  S seen", 9: "SM0  chunk0 loaded", 10: "SM0  3/4 done", 11: "SM0  exps issued",
         12: "SM0  P published", 16: "SM1  S seen", 17: "SM1  chunk0 loaded", 18: "SM1  3/4 done",
         19: "SM1  exps issued", 20: "SM1  P published"}
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
per_step = (int(t[7, 0]) - int(t[0, 0])) / 7.0
print(f"cycles per KV step: {per_step:.0f}")
for j in range(1, 7):
    sm0 = int(t[j, 12]) - int(t[j, 8]); sm1 = int(t[j, 20]) - int(t[j, 16])
    print(f"j={j+64}: softmax0 S-seen->P-published {sm0}, (load0 {int(t[j,9])-int(t[j,8])}, 3/4 {int(t[j,10])-int(t[j,9])}, "
          f"rest {int(t[j,11])-int(t[j,10])}, publish {int(t[j,12])-int(t[j,11])}); softmax1 {sm1}")
