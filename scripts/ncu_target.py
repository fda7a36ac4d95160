"""Small single-kernel targets for `ncu --set full` captures.
    python scripts/ncu_target.py attn|attn_b2|gemm|gemm_resid|conv96|conv96_dual|conv192|head|norm|ln|rope|t5attn
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omni_video_b200 import ops, vae_ops  # noqa: E402

bf16 = torch.bfloat16
what = sys.argv[1] if len(sys.argv) > 1 else "attn"
torch.manual_seed(0)
if what in ("attn", "attn_b2"):
    L, H = 32760, 12
    B = 2 if what == "attn_b2" else 1      # attn_b2: the launch bench.py times at N = 1 (cond + uncond batched)
    q, k, v = (torch.randn(B, L, H, 128, device="cuda").to(bf16) for _ in range(3))
    out = torch.empty_like(q)
    fn = lambda: ops.attention(q, k, v, out=out)  # noqa: E731
elif what in ("gemm", "gemm_resid"):
    M, N, K = 32760, 1536, 1536
    a = torch.randn(M, K, device="cuda").to(bf16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bf16)
    b = torch.randn(N, device="cuda").to(bf16)
    if what == "gemm":
        o = torch.empty(M, N, device="cuda", dtype=bf16)
        fn = lambda: ops.gemm(a, w, b, out=o)  # noqa: E731
    else:
        x = torch.randn(M, N, device="cuda")
        g = torch.randn(1, N, device="cuda")
        fn = lambda: ops.gemm(a, w, b, epilogue=ops.EPI_GATE_RESID_F32, out=x, resid=x, gate=g, rows_per_sample=M)  # noqa: E731
elif what in ("conv96", "conv192"):
    c = 96 if what == "conv96" else 192
    T, H, W = (9, 480, 832) if c == 96 else (9, 240, 416)
    x = torch.randn(T, H, W, c, device="cuda").to(bf16)
    wt = torch.randn(c, c, 3, 3, 3, device="cuda") / math.sqrt(27 * c)
    pc = vae_ops.pack_conv3d(wt, torch.zeros(c, device="cuda"))
    fn = lambda: vae_ops.conv3d(x, pc, (T, H, W))  # noqa: E731
elif what == "conv96_dual":
    # second conv of a 96-channel residual block: + residual, raw sum AND its RMS_norm.SiLU written (two tensors)
    T, H, W, c = 9, 480, 832, 96
    x = torch.randn(T, H, W, c, device="cuda").to(bf16)
    res = torch.randn(T, H, W, c, device="cuda").to(bf16)
    pc = vae_ops.pack_conv3d(torch.randn(c, c, 3, 3, 3, device="cuda") / math.sqrt(27 * c), torch.zeros(c, device="cuda"))
    gamma = torch.ones(c, device="cuda")
    o, on = torch.empty_like(x), torch.empty_like(x)
    fn = lambda: vae_ops.conv3d(x, pc, (T, H, W), out=o, resid=res, norm_gamma=gamma, norm_out=on)  # noqa: E731
elif what == "head":
    # the decoder head as per-tap planes (1x1x1 conv, n_group = 4) + tap_sum: two launches
    T, H, W = 9, 480, 832
    x = torch.randn(T, H, W, 96, device="cuda").to(bf16)
    pc, taps, bias = vae_ops.pack_tap_planes(torch.randn(3, 96, 3, 3, 3, device="cuda") / math.sqrt(27 * 96),
                                             torch.zeros(3, device="cuda"))
    y = torch.empty((27, T, H, W, 4), device="cuda", dtype=bf16)
    o = torch.empty((3, T, H, W), device="cuda")

    def fn():
        vae_ops.conv3d(x, pc, (T, H, W), out=y, out_strides=(H * W * 4, W * 4, 4), n_group=4, group_stride=T * H * W * 4)
        vae_ops.tap_sum(y, taps, bias, 3, (T, H, W), clamp=True, out=o)
elif what == "ln":
    L, d = 65520, 1536
    x = torch.randn(L, d, device="cuda")
    tbl = torch.randn(2, 6, d, device="cuda")
    o = torch.empty(L, d, device="cuda", dtype=bf16)
    fn = lambda: ops.layernorm_modulate(x, tbl[:, 1], tbl[:, 0], 1e-6, rows_per_sample=32760, out=o)  # noqa: E731
elif what == "t5attn":
    # umT5-XXL self-attention of the pipeline's prompt batch: 2 x 512 tokens, 64 heads of 64, one launch
    B, L, H, D = 2, 512, 64, 64
    qkv = torch.randn(B, L, 3, H, D, device="cuda").to(bf16)
    rb = torch.randn(H, 2 * L - 1, device="cuda")
    kl = torch.tensor([77, 512], dtype=torch.int32, device="cuda")
    o = torch.empty(B, L, H, D, device="cuda", dtype=bf16)
    fn = lambda: ops.attention_relbias(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], rb, kl, 1.0, out=o)  # noqa: E731
elif what == "rope":
    L, d = 65520, 1536
    qkv = torch.randn(L, 3 * d, device="cuda").to(bf16)
    w = torch.ones(d, device="cuda", dtype=bf16)
    cs = ops.rope_cos_sin_table().cuda()
    fn = lambda: ops.rmsnorm_rope_(qkv[:, :d], w, 1e-6, cs_table=cs, grid=(21, 30, 52), rows_per_sample=32760)  # noqa: E731
else:
    x = torch.randn(9 * 480 * 832, 96, device="cuda").to(bf16)
    g = torch.ones(96, device="cuda")
    fn = lambda: vae_ops.rmsnorm_silu(x, g)  # noqa: E731
for _ in range(3):
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", what)
