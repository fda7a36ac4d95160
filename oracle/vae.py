"""ORACLE — test infrastructure only.  Not imported by the product (omni_video_b200/*).

CPU restatement of the reference's Wan2.1 causal-3D VAE (SAIS-FUXI/Omni-Video,
omnivideo/modules/vae2_1.py) as plain functions over a state dict with the reference's parameter
names.  Unlike the reference, which walks the video in temporal chunks (decode: one latent frame
per call, encode: 1,4,4,... frames; vae2_1.py:515-567) and carries a two-frame `feat_cache`
through every CausalConv3d (:201-219), this restatement processes ALL frames in one pass:

  * a cached CausalConv3d over chunks == one causal conv with zero left padding of 2 (:16-35);
  * `upsample3d` (:100-140): frame 0 bypasses `time_conv` (the 'Rep' marker); frames 1.. go through
    `time_conv` as a causal conv that does NOT see frame 0 (zero left context), and its 2C output
    channels become two consecutive frames;
  * `downsample3d` (:142-158): frame 0 is kept; output frame j>=1 is the stride-2 (3,1,1) conv over
    input frames (2j-2, 2j-1, 2j).

Pinning: the reference has no tests/goldens for this path; the restatement is
pinned against OUTPUTS OF THE REFERENCE'S OWN chunked implementation, run in the build container by
`oracle/gen_golden_vae.py` and committed under tests/golden/vae_*.npz
(`tests/test_oracle_vae_golden.py`).

Two arithmetic modes, as in oracle/dit.py: fp32 throughout (the reference runs the VAE in fp32,
SURVEY.md §5), or `emulate_bf16=True`, which rounds activations to bf16 where the sm_100a path
stores them (conv / norm outputs) and uses bf16-rounded weights.
"""
from __future__ import annotations

import math
import re
from typing import Dict

import torch
import torch.nn.functional as F

f32 = torch.float32

# vae2_1.py:596-603 (the only configuration the pipeline instantiates)
VAE_CFG = dict(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
               temperal_downsample=[False, True, True], dropout=0.0)
# vae2_1.py:628-635
LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
               0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
              3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


def _r(x, em):
    return x.to(torch.bfloat16).to(x.dtype) if em else x


# ---------------------------------------------------------------- primitives (x: [C, T, H, W])
def causal_conv3d(x, w, b, em=False, stride_t=1, causal_pad=None):
    """vae2_1.py:16-35.  w [Co, Ci, kt, kh, kw]; spatial zero pad (k-1)/2, temporal zero pad on the
    LEFT only: 2*pt where pt is the module's temporal padding ((kt-1)/2 for the 3x3x3 and (3,1,1)
    convs of ResidualBlock/upsample3d; 0 for downsample3d's strided time_conv)."""
    kt, kh, kw = w.shape[2:]
    if causal_pad is None:
        causal_pad = kt - 1
    xp = F.pad(x.unsqueeze(0), (kw // 2, kw // 2, kh // 2, kh // 2, causal_pad, 0))
    y = F.conv3d(xp, w, b, stride=(stride_t, 1, 1))
    return _r(y.squeeze(0), em)


def rms_norm(x, gamma, em=False, silu=False):
    """vae2_1.py:38-53: F.normalize over channels * sqrt(C) * gamma; optional SiLU after it
    (the nn.SiLU that always follows in ResidualBlock/head, :194-196)."""
    C = x.shape[0]
    y = F.normalize(x, dim=0) * (C ** 0.5) * gamma.reshape(C, 1, 1, 1)
    if silu:
        y = F.silu(y)
    return _r(y, em)


def conv2d_frames(x, w, b, em=False, stride=1, pad=(1, 1, 1, 1)):
    """nn.Conv2d applied per frame ('b c t h w -> (b t) c h w', vae2_1.py:137-140)."""
    xt = x.permute(1, 0, 2, 3)                                  # [T, C, H, W]
    y = F.conv2d(F.pad(xt, pad), w, b, stride=stride)
    return _r(y.permute(1, 0, 2, 3).contiguous(), em)


def upsample_nearest2x(x):
    """vae2_1.py:56-62,77: nearest-exact, scale 2 == 2x2 pixel replication."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def residual_block(sd, p, x, em):
    """vae2_1.py:185-219."""
    h = causal_conv3d(x, sd[p + "shortcut.weight"], sd[p + "shortcut.bias"], em) \
        if (p + "shortcut.weight") in sd else x
    y = rms_norm(x, sd[p + "residual.0.gamma"], em, silu=True)
    y = causal_conv3d(y, sd[p + "residual.2.weight"], sd[p + "residual.2.bias"], em)
    y = rms_norm(y, sd[p + "residual.3.gamma"], em, silu=True)
    y = causal_conv3d(y, sd[p + "residual.6.weight"], sd[p + "residual.6.bias"], em)
    return _r(y + h, em)


def attention_block(sd, p, x, em):
    """vae2_1.py:222-261: per frame, RMS_norm -> 1x1 qkv -> single-head attention over h*w tokens
    with head_dim = C -> 1x1 proj -> + identity."""
    C, T, H, W = x.shape
    xn = rms_norm(x, sd[p + "norm.gamma"], em)
    tok = xn.permute(1, 2, 3, 0).reshape(T, H * W, C)            # [T, hw, C]
    wq = sd[p + "to_qkv.weight"].reshape(3 * C, C)
    qkv = _r(tok @ wq.T + sd[p + "to_qkv.bias"], em)
    q, k, v = qkv.chunk(3, dim=-1)
    s = (q @ k.transpose(1, 2)) / math.sqrt(C)
    pr = torch.softmax(s, dim=-1)
    o = _r(_r(pr, em) @ v, em)
    o = _r(o @ sd[p + "proj.weight"].reshape(C, C).T + sd[p + "proj.bias"], em)
    return _r(o.reshape(T, H, W, C).permute(3, 0, 1, 2) + x, em)


def resample(sd, p, x, mode, em):
    """vae2_1.py:100-159 in single-pass form (see module docstring)."""
    C = x.shape[0]
    if mode == "upsample3d":
        if x.shape[1] > 1:
            y = causal_conv3d(x[:, 1:], sd[p + "time_conv.weight"], sd[p + "time_conv.bias"], em)  # [2C, T-1, H, W]
            T1 = y.shape[1]
            y = torch.stack((y[:C], y[C:]), dim=2).reshape(C, 2 * T1, *x.shape[2:])    # :133-136
            x = torch.cat([x[:, :1], y], dim=1)
    if mode in ("upsample2d", "upsample3d"):
        return conv2d_frames(upsample_nearest2x(x), sd[p + "resample.1.weight"], sd[p + "resample.1.bias"], em)
    # downsample: ZeroPad2d((0,1,0,1)) + conv 3x3 stride 2 (:86-95)
    x = conv2d_frames(x, sd[p + "resample.1.weight"], sd[p + "resample.1.bias"], em, stride=2,
                      pad=(0, 1, 0, 1))
    if mode == "downsample3d" and x.shape[1] > 1:
        y = causal_conv3d(x, sd[p + "time_conv.weight"], sd[p + "time_conv.bias"], em, stride_t=2,
                          causal_pad=0)                                                 # frames (2j-2,2j-1,2j)
        x = torch.cat([x[:, :1], y], dim=1)
    return x


# ---------------------------------------------------------------- module walkers
def _indices(sd, prefix):
    pat = re.compile(re.escape(prefix) + r"(\d+)\.")
    return sorted({int(m.group(1)) for k in sd for m in [pat.match(k)] if m})


def _stack(sd, prefix, x, em, up: bool):
    for i in _indices(sd, prefix):
        p = f"{prefix}{i}."
        if (p + "residual.2.weight") in sd:
            x = residual_block(sd, p, x, em)
        elif (p + "to_qkv.weight") in sd:
            x = attention_block(sd, p, x, em)
        elif (p + "resample.1.weight") in sd:
            three_d = (p + "time_conv.weight") in sd
            mode = ("upsample" if up else "downsample") + ("3d" if three_d else "2d")
            x = resample(sd, p, x, mode, em)
        else:  # pragma: no cover
            raise KeyError(f"unknown module at {p}")
    return x


def decoder3d(sd, x, em=False, p="decoder."):
    """vae2_1.py:422-471."""
    x = causal_conv3d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], em)
    x = _stack(sd, p + "middle.", x, em, up=True)
    x = _stack(sd, p + "upsamples.", x, em, up=True)
    x = rms_norm(x, sd[p + "head.0.gamma"], em, silu=True)
    w, b = sd[p + "head.2.weight"], sd[p + "head.2.bias"]
    xp = F.pad(x.unsqueeze(0), (1, 1, 1, 1, 2, 0))
    return F.conv3d(xp, w, b).squeeze(0)                        # final output stays fp32


def encoder3d(sd, x, em=False, p="encoder."):
    """vae2_1.py:317-365."""
    x = causal_conv3d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], em)
    x = _stack(sd, p + "downsamples.", x, em, up=False)
    x = _stack(sd, p + "middle.", x, em, up=False)
    x = rms_norm(x, sd[p + "head.0.gamma"], em, silu=True)
    w, b = sd[p + "head.2.weight"], sd[p + "head.2.bias"]
    xp = F.pad(x.unsqueeze(0), (1, 1, 1, 1, 2, 0))
    return F.conv3d(xp, w, b).squeeze(0)


def vae_decode(sd: Dict[str, torch.Tensor], z: torch.Tensor, emulate_bf16=False) -> torch.Tensor:
    """Wan2_1_VAE.decode for one sample (vae2_1.py:656-662 -> :543-567): z [16, T, h, w] ->
    fp32 [3, 1+4(T-1), 8h, 8w] clamped to [-1, 1]."""
    em = emulate_bf16
    sd = {k: (_r(v.float(), em)) for k, v in sd.items()}
    mean = torch.tensor(LATENT_MEAN).view(-1, 1, 1, 1)
    std = torch.tensor(LATENT_STD).view(-1, 1, 1, 1)
    z = z.float() / (1.0 / std) + mean                          # :546-550
    x = F.conv3d(z.unsqueeze(0), sd["conv2.weight"], sd["conv2.bias"]).squeeze(0)   # 1x1x1 (:552)
    x = _r(x, em)
    return decoder3d(sd, x, em).float().clamp_(-1, 1)


def vae_encode(sd: Dict[str, torch.Tensor], video: torch.Tensor, emulate_bf16=False) -> torch.Tensor:
    """Wan2_1_VAE.encode for one sample (vae2_1.py:646-654 -> :515-541): video [3, T, H, W] with
    T = 1 + 4k -> normalised mu, fp32 [16, 1+k, H/8, W/8]."""
    em = emulate_bf16
    sd = {k: (_r(v.float(), em)) for k, v in sd.items()}
    out = encoder3d(sd, _r(video.float(), em), em)              # [32, T', h, w]
    out = F.conv3d(out.unsqueeze(0), sd["conv1.weight"], sd["conv1.bias"]).squeeze(0)  # :534
    mu = out[:16]
    mean = torch.tensor(LATENT_MEAN).view(-1, 1, 1, 1)
    std = torch.tensor(LATENT_STD).view(-1, 1, 1, 1)
    return ((mu - mean) * (1.0 / std)).float()                  # :535-539


def psnr(a: torch.Tensor, b: torch.Tensor, peak: float = 2.0) -> float:
    """PSNR in dB for videos in [-1, 1] (peak-to-peak 2)."""
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
