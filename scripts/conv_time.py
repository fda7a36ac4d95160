"""Times ovb_conv3d at the VAE's full-resolution shapes (experiments: OVB_CONV_NO_BRICK, OVB_CONV_BOX_L2)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omni_video_b200 import vae_ops  # noqa: E402

bf16 = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(cin, cout, T, H, W, tag):
    x = torch.randn(T, H, W, cin, device="cuda").to(bf16)
    wt = torch.randn(cout, cin, 3, 3, 3, device="cuda") / math.sqrt(27 * cin)
    pc = vae_ops.pack_conv3d(wt, torch.zeros(cout, device="cuda"))
    gamma = torch.ones(cout, device="cuda")
    res = torch.randn(T, H, W, cout, device="cuda").to(bf16)
    yn = torch.empty(T, H, W, cout, device="cuda", dtype=bf16)
    out = torch.empty(T, H, W, cout, device="cuda", dtype=bf16)
    mode = os.environ.get("CONV_MODE", "plain")
    fn = {"plain": lambda: vae_ops.conv3d(x, pc, (T, H, W), out=out),
          "resid": lambda: vae_ops.conv3d(x, pc, (T, H, W), out=out, resid=res),
          "norm": lambda: vae_ops.conv3d(x, pc, (T, H, W), out=out, norm_gamma=gamma),
          "resid_dual": lambda: vae_ops.conv3d(x, pc, (T, H, W), out=out, resid=res, norm_gamma=gamma, norm_out=yn)}[mode]
    tag = tag + " " + mode
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    fl = 2.0 * T * H * W * cout * 27 * cin
    kc = 64 if cin % 64 == 0 else (32 if cin % 32 == 0 else 16)
    n_mma = 27 * cin // 16 * (T * H * W / 256.0) / 74.0          # pair MMAs per CTA pair
    print(f"{tag}: conv {cin}->{cout} [{T},{H},{W}]: {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s  "
          f"(KC {kc}: ~{ms * 1e-3 * 1.9e9 / n_mma:.0f} cycles per M=256,N={cout},K=16 MMA at 1.9 GHz; floor {cout // 2})", flush=True)


if __name__ == "__main__":
    tag = f"brick={'0' if os.environ.get('OVB_CONV_NO_BRICK') else '1'} box_l2={os.environ.get('OVB_CONV_BOX_L2', 'auto')}"
    if "--sweep" in sys.argv:
        # cycles per MMA vs channel-chunk width (swizzle mode) and N: Cin 32/96 -> 64-byte rows, 64/128 -> 128-byte
        for cin, cout in [(32, 96), (64, 96), (96, 96), (128, 96), (64, 64), (128, 128), (96, 192), (192, 96)]:
            run(cin, cout, 9, 480, 832, tag)
        sys.exit(0)
    run(96, 96, 21, 480, 832, tag)
    if "--all" in sys.argv:
        run(192, 192, 21, 240, 416, tag)
