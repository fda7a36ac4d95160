"""Parity of the sm_100a VAE path (ovb_conv3d & friends, omni_video_b200/modules/vae2_1.py):
  * kernel by kernel against fp32 PyTorch/oracle restatements on seeded inputs;
  * whole decode/encode against the committed goldens = outputs of the reference's own chunked
    implementation (north-star tolerance: decoded video PSNR >= 40 dB), and against the
    bf16-emulating oracle (tighter).
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import synth
from oracle import vae as V
from tests.golden_util import GOLDEN_DIR, golden_names

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def vops():
    from omni_video_b200 import vae_ops
    return vae_ops


def _rng(seed):
    return np.random.default_rng(seed)


def _randn(rng, *shape, scale=1.0):
    return torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(scale))


def _cl(x):  # [C,T,H,W] -> channels-last bf16 cuda
    return x.permute(1, 2, 3, 0).contiguous().to(bf16).cuda()


def _planar(y):  # channels-last cuda -> [C,T,H,W] fp32 cpu
    return y.float().cpu().permute(3, 0, 1, 2).contiguous()


def _close(got, ref, tol=2e-2):
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= tol * max(scale, 1e-3), f"max-abs err {err} vs scale {scale}"
    rel = ((got - ref).norm() / ref.norm().clamp_min(1e-20)).item()
    assert rel <= 6e-3, f"rel-L2 {rel}"


@pytest.mark.parametrize("cin,cout,T,H,W", [(16, 16, 2, 5, 9), (96, 96, 3, 9, 20), (192, 384, 2, 8, 16),
                                            (384, 192, 2, 6, 10), (384, 768, 1, 4, 8), (96, 16, 4, 17, 33),
                                            (32, 32, 1, 3, 3)])
def test_causal_conv3d_3x3x3(vops, cin, cout, T, H, W):
    rng = _rng(cin + cout)
    x = _randn(rng, cin, T, H, W).to(bf16).float()
    w = _randn(rng, cout, cin, 3, 3, 3, scale=1.0 / (27 * cin) ** 0.5).to(bf16).float()
    b = _randn(rng, cout, scale=0.1)
    ref = V.causal_conv3d(x, w, b)
    pc = vops.pack_conv3d(w.cuda(), b.cuda())
    y = vops.conv3d(_cl(x), pc, (T, H, W))
    _close(_planar(y), ref.to(bf16).float())


@pytest.mark.parametrize("c,T,H,W,with_resid,dual", [(96, 2, 9, 20, False, False), (96, 3, 16, 24, True, True),
                                                      (192, 2, 8, 16, True, False), (32, 1, 5, 9, False, True)])
def test_conv_with_fused_rmsnorm_silu(vops, c, T, H, W, with_resid, dual):
    """The conv epilogue applies the RMS_norm·SiLU that consumes its output (reference vae2_1.py:38-53,
    :194-196): same result as conv followed by ovb_rmsnorm_silu_cl, for the single-output form (only the
    normalised tensor is kept) and the dual form (raw sum + normalised tensor)."""
    rng = _rng(7 * c + T)
    x = _randn(rng, c, T, H, W).to(bf16).float()
    w = _randn(rng, c, c, 3, 3, 3, scale=1.0 / (27 * c) ** 0.5).to(bf16).float()
    b = _randn(rng, c, scale=0.1)
    gamma = (1.0 + _randn(rng, c, scale=0.1)).cuda()
    res = _cl(_randn(rng, c, T, H, W)) if with_resid else None
    pc = vops.pack_conv3d(w.cuda(), b.cuda())
    assert vops.can_fuse_norm(pc)
    raw = vops.conv3d(_cl(x), pc, (T, H, W), resid=res)
    ref_n = vops.rmsnorm_silu(raw, gamma)
    if dual:
        yn = torch.empty_like(raw)
        y = vops.conv3d(_cl(x), pc, (T, H, W), resid=res, norm_gamma=gamma, norm_out=yn)
        assert torch.equal(y, raw)
    else:
        yn = vops.conv3d(_cl(x), pc, (T, H, W), resid=res, norm_gamma=gamma)
    # the two differ only in the summation order of the squares (and then by at most one bf16 ulp)
    d = (yn.float() - ref_n.float()).abs()
    assert d.max().item() <= 2 ** -7 * max(1.0, ref_n.float().abs().max().item())
    assert (d > 0).float().mean().item() < 0.02
    # and against the fp32 oracle composition
    ref = V.causal_conv3d(x, w, b).to(bf16).float()
    if with_resid:
        ref = (ref + _planar(res)).to(bf16).float()
    refn = V.rms_norm(ref, gamma.cpu().view(c, 1, 1, 1), silu=True)
    _close(_planar(yn), refn.to(bf16).float())


def test_conv_brick_mode_matches_tap_mode(vops, monkeypatch):
    """3x3x3 stride-1 convs fetch one (bh+2)-row brick per (dt, dw) instead of one box per tap; results
    are the same products accumulated in another tap order: equal to the per-tap path (OVB_CONV_NO_BRICK=1)
    up to fp32 accumulation order, i.e. a rare last-bit difference after the bf16 rounding."""
    rng = _rng(11)
    for cin, cout, T, H, W in [(96, 96, 3, 40, 72), (96, 16, 2, 33, 50), (16, 384, 2, 12, 20)]:
        x = _cl(_randn(rng, cin, T, H, W))
        w = _randn(rng, cout, cin, 3, 3, 3, scale=1.0 / (27 * cin) ** 0.5)
        pc = vops.pack_conv3d(w.cuda(), _randn(rng, cout, scale=0.1).cuda())
        monkeypatch.delenv("OVB_CONV_NO_BRICK", raising=False)
        a = vops.conv3d(x, pc, (T, H, W))
        monkeypatch.setenv("OVB_CONV_NO_BRICK", "1")
        b = vops.conv3d(x, pc, (T, H, W))
        monkeypatch.delenv("OVB_CONV_NO_BRICK", raising=False)
        # same products, summed in another tap order: fp32 accumulation-order differences only
        d = (a.float() - b.float()).abs()
        assert d.max().item() <= 2 ** -7 * max(1.0, b.float().abs().max().item()), (cin, cout)
        assert (d > 0).float().mean().item() < 0.05, (cin, cout)


def test_conv_1x1x1_time_conv_and_residual(vops):
    rng = _rng(3)
    cin, cout, T, H, W = 192, 384, 3, 7, 12
    x = _randn(rng, cin, T, H, W).to(bf16).float()
    w1 = _randn(rng, cout, cin, 1, 1, 1, scale=cin ** -0.5).to(bf16).float()
    b1 = _randn(rng, cout, scale=0.1)
    r = _randn(rng, cout, T, H, W).to(bf16).float()
    ref = (V.causal_conv3d(x, w1, b1).to(bf16).float() + r).to(bf16).float()
    y = vops.conv3d(_cl(x), vops.pack_conv3d(w1.cuda(), b1.cuda()), (T, H, W), resid=_cl(r))
    _close(_planar(y), ref)
    wt = _randn(rng, cin, cin, 3, 1, 1, scale=(3 * cin) ** -0.5).to(bf16).float()
    ref = V.causal_conv3d(x, wt, None)
    y = vops.conv3d(_cl(x), vops.pack_conv3d(wt.cuda(), None), (T, H, W))
    _close(_planar(y), ref.to(bf16).float())


def test_upsample_phases_match_upsample_then_conv(vops):
    rng = _rng(4)
    c, T, H, W = 192, 2, 5, 11
    x = _randn(rng, c, T, H, W).to(bf16).float()
    w = _randn(rng, c // 2, c, 3, 3, scale=(9 * c) ** -0.5).to(bf16).float()
    b = _randn(rng, c // 2, scale=0.1)
    ref = V.conv2d_frames(V.upsample_nearest2x(x), w, b)
    phases = vops.pack_upsample_conv2d(w.cuda(), b.cuda())
    co = c // 2
    out = torch.zeros((T, 2 * H, 2 * W, co), device="cuda", dtype=bf16)
    for (a, bb), pc in phases.items():
        vops.conv3d(_cl(x), pc, (T, H, W), out=out, out_strides=(4 * H * W * co, 4 * W * co, 2 * co),
                    out_base=(a * 2 * W + bb) * co)
    _close(_planar(out), ref.to(bf16).float())


def test_time_conv_interleave_and_strided_convs(vops):
    rng = _rng(5)
    sd = {}
    c, T, H, W = 96, 4, 6, 8
    x = _randn(rng, c, T, H, W).to(bf16).float()
    sd["p.time_conv.weight"] = _randn(rng, 2 * c, c, 3, 1, 1, scale=(3 * c) ** -0.5).to(bf16).float()
    sd["p.time_conv.bias"] = _randn(rng, 2 * c, scale=0.1)
    # upsample3d temporal part (oracle: frames 1.. -> time_conv -> interleave)
    y = V.causal_conv3d(x[:, 1:], sd["p.time_conv.weight"], sd["p.time_conv.bias"])
    ref = torch.cat([x[:, :1], torch.stack((y[:c], y[c:]), dim=2).reshape(c, 2 * (T - 1), H, W)], dim=1)
    pc = vops.pack_conv3d(sd["p.time_conv.weight"].cuda(), sd["p.time_conv.bias"].cuda())
    xc = _cl(x)
    out = torch.empty((1 + 2 * (T - 1), H, W, c), device="cuda", dtype=bf16)
    out[0].copy_(xc[0])
    fr = H * W * c
    vops.conv3d(xc[1:], pc, (T - 1, H, W), out=out, out_strides=(2 * fr, W * c, c), out_base=fr,
                n_group=c, group_stride=fr)
    _close(_planar(out), ref.to(bf16).float())
    # downsample: ZeroPad2d((0,1,0,1)) + conv 3x3 stride 2, odd and even sizes
    for (Hh, Ww) in ((6, 8), (7, 9)):
        xx = _randn(rng, c, 2, Hh, Ww).to(bf16).float()
        w2 = _randn(rng, c, c, 3, 3, scale=(9 * c) ** -0.5).to(bf16).float()
        b2 = _randn(rng, c, scale=0.1)
        ref = V.conv2d_frames(xx, w2, b2, stride=2, pad=(0, 1, 0, 1))
        pc2 = vops.pack_conv3d(w2.cuda(), b2.cuda(), stride=(1, 2, 2), spatial_pad=(0, 0))
        y2 = vops.conv3d(_cl(xx), pc2, (2,) + tuple(ref.shape[2:]))
        _close(_planar(y2), ref.to(bf16).float())
    # downsample3d temporal part: out[j] = conv(x[2j-2], x[2j-1], x[2j]), stride 2
    x5 = _randn(rng, c, 5, H, W).to(bf16).float()
    wt = _randn(rng, c, c, 3, 1, 1, scale=(3 * c) ** -0.5).to(bf16).float()
    bt = _randn(rng, c, scale=0.1)
    ref = V.causal_conv3d(x5, wt, bt, stride_t=2, causal_pad=0)                    # 2 frames
    pt = vops.pack_conv3d(wt.cuda(), bt.cuda(), causal_pad=0, stride=(2, 1, 1))
    y5 = vops.conv3d(_cl(x5), pt, (2, H, W))
    _close(_planar(y5), ref.to(bf16).float())


def test_head_conv_planar_clamped(vops):
    rng = _rng(6)
    c, T, H, W = 96, 2, 9, 14
    x = _randn(rng, c, T, H, W).to(bf16).float()
    w = _randn(rng, 3, c, 3, 3, 3, scale=3.0 * (27 * c) ** -0.5).to(bf16).float()
    b = _randn(rng, 3, scale=0.1)
    ref = V.causal_conv3d(x, w, b).clamp(-1, 1)
    out = torch.empty((3, T, H, W), device="cuda", dtype=torch.float32)
    vops.conv3d(_cl(x), vops.pack_conv3d(w.cuda(), b.cuda()), (T, H, W), out=out, out_strides=(H * W, W, 1),
                out_mode=vops.CONV_OUT_F32_PLANAR_CLAMP, plane_stride=T * H * W)
    assert (ref.abs() == 1).float().mean() > 0.01  # the clamp is exercised
    assert (out.cpu() - ref).abs().max().item() <= 2e-3


@pytest.mark.parametrize("T,H,W,pad", [(2, 9, 14, 0), (5, 17, 40, 0), (3, 10, 24, 1)])
def test_head_conv_as_tap_planes_and_tap_sum(vops, T, H, W, pad):
    """The 96 -> 3 head conv (vae2_1.py:164-165) as one 1x1x1 conv writing per-tap partial products as planes
    (n_group = 4) + ovb_tap_sum: against the fp32 conv, against the implicit-GEMM form of the same conv (the partial
    products are rounded to bf16: a few 1e-3 of the value range), and with one halo row above and below the
    strip (pad = 1: the conv runs over the halo rows as well, tap_sum shifts its row offsets)."""
    rng = _rng(60 + T)
    c = 96
    x = _randn(rng, c, T, H, W).to(bf16).float()
    w = _randn(rng, 3, c, 3, 3, 3, scale=3.0 * (27 * c) ** -0.5).to(bf16).float()
    b = _randn(rng, 3, scale=0.1)
    ref = V.causal_conv3d(x, w, b)
    xin = x
    if pad:   # strip layout: one halo row above and below; a halo row of zeros == the image border
        xin = F.pad(x, (0, 0, 1, 1))
    Hp = H + 2 * pad
    pc, taps, bias = vops.pack_tap_planes(w.cuda(), b.cuda())
    assert pc.w.shape == (128, c) and pc.cout == 108 and len(taps) == 27
    y = torch.full((27, T, Hp, W, 4), float("nan"), device="cuda", dtype=bf16)
    vops.conv3d(_cl(xin), pc, (T, Hp, W), out=y, out_strides=(Hp * W * 4, W * 4, 4), n_group=4,
                group_stride=T * Hp * W * 4)
    assert bool(torch.isfinite(y).all())                      # every plane element written
    # plane k is x . W[:, :, tap k]
    k = 14
    wk = w.reshape(3, c, 27)[:, :, k]
    yk = torch.einsum("cthw,oc->thwo", xin, wk)
    assert (y[k, ..., :3].float().cpu() - yk).abs().max().item() <= 2e-2 * yk.abs().max().item()
    assert float(y[..., 3].abs().max()) == 0.0
    for clamp in (False, True):
        out = vops.tap_sum(y, taps, bias, 3, (T, H, W), dh_shift=pad, clamp=clamp)
        want = ref.clamp(-1, 1) if clamp else ref
        assert (out.cpu() - want).abs().max().item() <= 6e-3 * max(1.0, ref.abs().max().item())
    direct = torch.empty((3, T, H, W), device="cuda", dtype=torch.float32)
    vops.conv3d(_cl(xin), vops.pack_conv3d(w.cuda(), b.cuda()), (T, H, W), out=direct, out_strides=(H * W, W, 1),
                dh_shift=pad, out_mode=vops.CONV_OUT_F32_PLANAR, plane_stride=T * H * W)
    out = vops.tap_sum(y, taps, bias, 3, (T, H, W), dh_shift=pad)
    rel = ((out - direct).norm() / direct.norm()).item()
    print(f"head conv as tap planes + tap_sum vs implicit GEMM: rel-L2 {rel:.2e}")
    assert rel <= 3e-3



@pytest.mark.parametrize("cin,cout,k,stride,pad,cache", [
    (96, 96, 3, (1, 1, 1), 1, 2), (3, 96, 3, (1, 1, 1), 1, 0), (192, 192, (3, 1, 1), (2, 1, 1), (0, 0, 0), 0),
    (384, 32, 1, (1, 1, 1), 0, 0)])
def test_leaf_causal_conv3d_forward(vops, cin, cout, k, stride, pad, cache):
    """`CausalConv3d.forward(x, cache_x)` (reference vae2_1.py:26-35) at leaf granularity on the kernels: planar
    [B, C, T, H, W] in and out, cached frames in place of the causal zero padding; against torch's fp32 conv on the
    same bf16-rounded operands."""
    import torch.nn.functional as F
    from omni_video_b200.modules.vae2_1 import CausalConv3d, RMS_norm
    torch.manual_seed(5)
    torch.backends.cudnn.allow_tf32 = False
    m = CausalConv3d(cin, cout, k, stride=stride, padding=pad).cuda()
    with torch.no_grad():
        m.weight.copy_(m.weight.bfloat16().float())
    x = torch.randn(2, cin, 5, 24, 40, device="cuda").bfloat16().float()
    cache_x = torch.randn(2, cin, cache, 24, 40, device="cuda").bfloat16().float() if cache else None
    pt, ph, pw = m.padding
    xp = x if cache_x is None else torch.cat([cache_x, x], dim=2)
    want = F.conv3d(F.pad(xp, [pw, pw, ph, ph, 2 * pt - cache, 0]), m.weight, m.bias, stride=m.stride)
    got = m(x, cache_x)
    assert got.shape == want.shape and got.dtype == torch.float32
    assert (got - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
    if cin == 96:                                            # and the norm leaf on the result
        n = RMS_norm(cout, images=False).cuda()
        y = n(got)
        ref = F.normalize(got.bfloat16().float(), dim=1) * n.scale * n.gamma
        assert y.shape == got.shape and (y - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("c", [16, 96, 192, 384, 768])
def test_rmsnorm_silu_cl(vops, c):
    rng = _rng(c)
    n = 1000 + c
    x = _randn(rng, n, c).to(bf16)
    g = 1.0 + _randn(rng, c, scale=0.1)
    for silu in (True, False):
        ref = V.rms_norm(x.float().T.reshape(c, n, 1, 1), g, silu=silu).reshape(c, n).T
        y = vops.rmsnorm_silu(x.cuda(), g.cuda(), silu=silu)
        d = (y.float().cpu() - ref.to(bf16).float()).abs()
        assert d.max().item() <= ref.abs().max().item() * 2 ** -7
        assert (d > 0).float().mean().item() < 0.05


def test_layout_converters_and_softmax(vops):
    rng = _rng(8)
    z = _randn(rng, 16, 2, 3, 5)
    mean, std = torch.tensor(V.LATENT_MEAN), torch.tensor(V.LATENT_STD)
    cl = vops.planar_to_cl(z.cuda(), std.cuda(), mean.cuda(), 16)
    ref = (z * std.view(-1, 1, 1, 1) + mean.view(-1, 1, 1, 1)).permute(1, 2, 3, 0)
    assert torch.equal(cl.float().cpu(), ref.to(bf16).float())
    v = _randn(rng, 3, 2, 4, 6)
    cl3 = vops.planar_to_cl(v.cuda(), torch.ones(3).cuda(), torch.zeros(3).cuda(), 16)
    assert torch.equal(cl3[..., :3].float().cpu(), v.permute(1, 2, 3, 0).to(bf16).float())
    assert float(cl3[..., 3:].abs().max()) == 0.0
    back = vops.cl_to_planar(cl, (1 / std).cuda(), (-mean / std).cuda(), 16)
    assert (back.cpu() - z).abs().max().item() < 0.05
    S = _randn(rng, 37, 1000, scale=4.0)
    P = vops.softmax_rows(S.cuda(), 0.25)
    ref = torch.softmax(S * 0.25, dim=-1)
    assert (P.float().cpu() - ref).abs().max().item() <= 2 ** -8 * ref.max().item() + 1e-6


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return json.loads(str(z["meta"])), torch.from_numpy(z["out0"])


def _vae(sd):
    from omni_video_b200.modules.vae2_1 import Wan2_1_VAE
    vae = Wan2_1_VAE(vae_pth=None, device="cuda")
    vae.model.load_state_dict(sd, strict=True)     # reference parameter names, strict
    return vae


@pytest.mark.parametrize("name", golden_names("vae_decode"))
def test_decode_vs_reference_golden(name):
    meta, gold = _load(name)
    sd = synth.vae_state_dict(seed=meta["weight_seed"])
    z = synth.vae_latent(tuple(meta["grid"]), seed=meta["input_seed"])
    out = _vae(sd).decode([z.cuda()])[0].cpu()
    assert out.shape == gold.shape and out.dtype == torch.float32
    assert float(out.abs().max()) <= 1.0
    p_ref = V.psnr(out, gold)
    emu = V.vae_decode(sd, z, emulate_bf16=True)
    p_emu = V.psnr(out, emu)
    print(f"{name}: PSNR vs reference fp32 {p_ref:.1f} dB, vs bf16 oracle {p_emu:.1f} dB")
    assert p_ref >= 40.0, "north-star tolerance: decoded-video PSNR >= 40 dB"
    assert p_emu >= 45.0


def test_decode_at_baseline_config1_vs_reference_golden():
    """latent [16, 5, 32, 32] -> video [3, 17, 256, 256] (BASELINE.json configs[0]) against every 4th
    pixel of the reference's own decode, and the whole frame against the fp32 oracle."""
    meta, gold = _load("vae_dec_config1")
    sd = synth.vae_state_dict(seed=meta["weight_seed"])
    z = synth.vae_latent(tuple(meta["grid"]), seed=meta["input_seed"])
    out = _vae(sd).decode([z.cuda()])[0].cpu()
    assert list(out.shape) == meta["full_shape"] and float(out.abs().max()) <= 1.0
    st = meta["stride"]
    p_ref = V.psnr(out[:, :, ::st, ::st], gold)
    full = V.vae_decode(sd, z)
    p_full = V.psnr(out, full)
    print(f"vae_dec_config1: PSNR vs reference fp32 (strided) {p_ref:.1f} dB, vs fp32 oracle (all pixels) {p_full:.1f} dB")
    assert p_ref >= 40.0 and p_full >= 40.0, "north-star tolerance: decoded-video PSNR >= 40 dB"


@pytest.mark.parametrize("name", golden_names("vae_encode"))
def test_encode_vs_reference_golden(name):
    meta, gold = _load(name)
    sd = synth.vae_state_dict(seed=meta["weight_seed"])
    v = synth.vae_video(tuple(meta["shape"]), seed=meta["input_seed"])
    out = _vae(sd).encode([v.cuda()])[0].cpu()
    assert out.shape == gold.shape and out.dtype == torch.float32
    rel = ((out - gold).norm() / gold.norm()).item()
    print(f"{name}: encode rel-L2 vs reference fp32 {rel:.3e}")
    assert rel <= 2e-2


def test_vae_encode_decode_roundtrip_shapes_and_determinism():
    sd = synth.vae_state_dict(seed=1818)
    vae = _vae(sd)
    v = synth.vae_video((5, 32, 48), seed=3).cuda()
    z1 = vae.encode([v])[0]
    z2 = vae.encode([v])[0]
    assert z1.shape == (16, 2, 4, 6) and torch.equal(z1, z2)
    x = vae.decode([z1])[0]
    assert x.shape == (3, 5, 32, 48) and bool(torch.isfinite(x).all())
    # reference-style low-level entry (WanVAE_.decode returns the unclamped video, batch dim kept)
    raw = vae.model.decode(z1.unsqueeze(0), vae.scale)
    assert raw.shape == (1, 3, 5, 32, 48)
    assert torch.equal(raw[0].clamp(-1, 1), x)


def test_vae_strips_two_gpus_match_single_gpu():
    """Runs scripts/vae_strip_check.py under torchrun when the box has >= 2 GPUs (host logic on CPU:
    tests/test_distributed_cpu.py::test_vae_strip_halo_and_gather_two_ranks_gloo)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547",
                        os.path.join(root, "scripts", "vae_strip_check.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
