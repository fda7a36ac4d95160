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
    print(f"{tag}: conv {cin}->{cout} [{T},{H},{W}]: {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s", flush=True)


if __name__ == "__main__":
    tag = f"brick={'0' if os.environ.get('OVB_CONV_NO_BRICK') else '1'} box_l2={os.environ.get('OVB_CONV_BOX_L2', 'auto')}"
    run(96, 96, 21, 480, 832, tag)
    if "--all" in sys.argv:
        run(192, 192, 21, 240, 416, tag)
