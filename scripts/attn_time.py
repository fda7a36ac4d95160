"""Times ovb_attn_fwd of one or more builds of attn.cu (scripts/microbench/build/<lib>) at the bench shapes."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omni_video_b200 import _lib  # noqa: E402

bf16 = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(libname, B, L, H):
    lib = C.CDLL(os.path.join(ROOT, "scripts", "microbench", "build", libname))
    lib.ovb_attn_fwd.restype = C.c_int
    lib.ovb_attn_fwd.argtypes = [C.POINTER(_lib.AttnArgs), C.c_void_p]
    lib.ovb_attn_workspace_size.restype = C.c_int64
    q, k, v = (torch.randn(B, L, H, 128, device="cuda").to(bf16) for _ in range(3))
    out = torch.empty_like(q)
    a = _lib.AttnArgs()
    for name, t in (("q", q), ("k", k), ("v", v), ("out", out)):
        setattr(a, name, t.data_ptr())
    a.q_row_stride = a.k_row_stride = a.v_row_stride = a.o_row_stride = H * 128
    a.q_batch_stride = a.k_batch_stride = a.v_batch_stride = a.o_batch_stride = L * H * 128
    a.B, a.H, a.Lq, a.Lk = B, H, L, L
    need = int(lib.ovb_attn_workspace_size(C.byref(a)))
    ws = torch.empty(max(need, 16), dtype=torch.uint8, device="cuda")
    a.workspace, a.workspace_bytes = ws.data_ptr(), need
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        assert lib.ovb_attn_fwd(C.byref(a), st) == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.ovb_attn_fwd(C.byref(a), st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    ref = torch.nn.functional.scaled_dot_product_attention(q[:, :512].transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    err = (ref.transpose(1, 2).float() - out[:, :512].float()).abs().max().item()
    print(f"{libname} V1={os.environ.get('OVB_ATTN_V1', '0')}: B={B} L={L} H={H}: {ms:.3f} ms {4.0 * B * L * L * H * 128 / ms / 1e9:.0f} TF/s  max|diff| vs SDPA(512 rows) {err:.2e}", flush=True)


if __name__ == "__main__":
    for lib in sys.argv[1:]:
        run(lib, 1, 32760, 12)
        run(lib, 2, 32760, 12)
