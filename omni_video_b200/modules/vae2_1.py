"""B200-native Wan2.1 causal-3D VAE (drop-in for reference omnivideo/modules/vae2_1.py).

`Wan2_1_VAE(z_dim, vae_pth, dtype, device)` with `.encode(videos)` / `.decode(zs)` and the
attributes the pipeline touches (`.model`, `.mean`, `.std`, `.scale`, `.model.z_dim`;
x2x_gen_unified.py:328,506-510) keeps the reference's surface; the module tree below carries the
reference's parameter names so `Wan2.1_VAE.pth` loads unchanged.  No leaf `forward` runs on the hot path
(`CausalConv3d` and `RMS_norm` have one, on the same kernels, for callers that walk the tree themselves):
`VAEEngine` walks the parameters and drives libovb200:

    reference op (vae2_1.py)                               kernel
    -----------------------------------------------------  ---------------------------------------
    CausalConv3d 3x3x3 / 1x1x1 / (3,1,1) (:16-35)           ovb_conv3d  (implicit GEMM, tcgen05 + TMA)
    RMS_norm + SiLU (:38-53, :194-196)                      ovb_rmsnorm_silu_cl
    Upsample(nearest 2x) + Conv2d 3x3 (:75-84)              ovb_conv3d x4 (one 2x2 conv per output phase)
    upsample3d time_conv + frame interleave (:100-136)      ovb_conv3d (channel-group -> frame epilogue)
    ZeroPad2d(0,1,0,1) + Conv2d s2 (:86-95)                 ovb_conv3d (element-strided TMA boxes)
    downsample3d time_conv stride 2 (:142-158)              ovb_conv3d (st = 2)
    AttentionBlock, one head, d = C (:222-261)              ovb_gemm_bf16 + ovb_softmax_rows
    latent (de)normalisation, clamp (:535-550, :660)        ovb_planar_to_cl / ovb_cl_to_planar / conv epilogue

All frames of a clip are processed in ONE pass over channels-last bf16 tensors; the reference's
temporal chunk loop with `feat_cache` (:515-567) is exactly causal zero padding plus the first-frame
rules of upsample3d / downsample3d (pinned by tests/test_oracle_vae_golden.py).
The reference computes the VAE in fp32; bf16 tensor-core math here is held to PSNR >= 40 dB.
"""
from __future__ import annotations

import logging
import math
import os
import re
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops, vae_ops
from ..ops import EPI_BF16, EPI_F32, EPI_RESID_BF16
from ..vae_ops import CONV_OUT_F32_PLANAR_CLAMP

__all__ = ["Wan2_1_VAE", "WanVAE_", "VAEEngine"]

bf16 = torch.bfloat16
f32 = torch.float32

# reference vae2_1.py:628-635
_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
         0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
        3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]


# ------------------------------------------------------------------------------------------------
# Parameter containers with the reference's names (vae2_1.py:16-471)
# ------------------------------------------------------------------------------------------------
def _leaf_device_check(x: torch.Tensor, what: str):
    if x.device.type != "cuda":
        raise RuntimeError(f"{what} (B200 path) needs CUDA tensors; there is no CPU fallback")


def _unit_affine(c: int, dev):
    return torch.ones(c, device=dev, dtype=f32), torch.zeros(c, device=dev, dtype=f32)


class CausalConv3d(nn.Conv3d):
    """reference :16-35.  Inside `Wan2_1_VAE.encode/decode` the arithmetic is VAEEngine's (fused norms, strips,
    resident channels-last activations); `forward` below is the same kernel at LEAF granularity, for callers that
    drive the reference's module tree themselves: planar [B, C, T, H, W] in and out like the reference, one
    ovb_planar_to_cl + ovb_conv3d (fp32 planar output mode) per sample."""

    def _packed(self, causal_pad: int):
        key = (self.weight.data_ptr(), self.weight._version, None if self.bias is None else self.bias._version,
               causal_pad)
        if getattr(self, "_pack_key", None) != key:
            _, ph, pw = self.padding
            self._pack = vae_ops.pack_conv3d(self.weight, self.bias, causal_pad=causal_pad, stride=self.stride,
                                             spatial_pad=(ph, pw))
            self._pack_key = key
        return self._pack

    @torch.no_grad()
    def forward(self, x, cache_x=None):
        _leaf_device_check(x, "CausalConv3d")
        kt, kh, kw = self.kernel_size
        st, sh, sw = self.stride
        pt, ph, pw = self.padding
        cp = 2 * pt                                      # all of the time padding goes in front (:23-24)
        if cache_x is not None and cp > 0:               # :29-32: cached frames replace part of the zero padding
            x = torch.cat([cache_x.to(x.device), x], dim=2)
            cp -= cache_x.shape[2]
        if cp < 0:
            raise ValueError(f"cache_x holds {cache_x.shape[2]} frames, more than the causal padding {2 * pt}")
        B, Ci, Ti, H, W = x.shape
        To, Ho, Wo = (Ti + cp - kt) // st + 1, (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
        pc = self._packed(cp)
        one, zero = _unit_affine(Ci, x.device)
        out = torch.empty((B, self.out_channels, To, Ho, Wo), device=x.device, dtype=f32)
        for b in range(B):
            xc = vae_ops.planar_to_cl(x[b].to(f32).contiguous(), one, zero, pc.cin)
            vae_ops.conv3d(xc, pc, (To, Ho, Wo), out=out[b], out_strides=(Ho * Wo, Wo, 1),
                           out_mode=vae_ops.CONV_OUT_F32_PLANAR, plane_stride=To * Ho * Wo)
        return out.to(x.dtype)


class RMS_norm(nn.Module):
    """reference :38-53; `forward` = ovb_rmsnorm_silu_cl (without the SiLU) at leaf granularity, see CausalConv3d."""

    def __init__(self, dim, channel_first=True, images=True, bias=False):
        super().__init__()
        shape = (dim, 1, 1) if images else (dim, 1, 1, 1)
        self.gamma = nn.Parameter(torch.ones(shape if channel_first else (dim,)))
        self.scale = dim ** 0.5
        self.channel_first = channel_first

    @torch.no_grad()
    def forward(self, x):
        _leaf_device_check(x, "RMS_norm")
        gamma = self.gamma.detach().reshape(-1).to(f32).contiguous()
        c = gamma.numel()
        if c % 16 != 0:
            raise NotImplementedError(f"RMS_norm over {c} channels: the kernel takes multiples of 16")
        if not self.channel_first:                       # [..., C]
            return vae_ops.rmsnorm_silu(x.to(bf16).contiguous(), gamma, silu=False).to(x.dtype)
        one, zero = _unit_affine(c, x.device)
        lead = x.shape[0]
        x4 = x.reshape((lead, c, -1) + tuple(x.shape[-2:]))                # [B, C, T | 1, H, W]
        outs = []
        for b in range(lead):
            xc = vae_ops.planar_to_cl(x4[b].to(f32).contiguous(), one, zero, c)
            outs.append(vae_ops.cl_to_planar(vae_ops.rmsnorm_silu(xc, gamma, silu=False), one, zero, c))
        return torch.stack(outs).reshape(x.shape).to(x.dtype)


class Resample(nn.Module):
    """reference :65-182"""

    def __init__(self, dim, mode):
        assert mode in ("upsample2d", "upsample3d", "downsample2d", "downsample3d")
        super().__init__()
        self.dim, self.mode = dim, mode
        if mode.startswith("upsample"):
            self.resample = nn.Sequential(nn.Identity(), nn.Conv2d(dim, dim // 2, 3, padding=1))
            if mode == "upsample3d":
                self.time_conv = CausalConv3d(dim, dim * 2, (3, 1, 1), padding=(1, 0, 0))
        else:
            self.resample = nn.Sequential(nn.Identity(), nn.Conv2d(dim, dim, 3, stride=(2, 2)))
            if mode == "downsample3d":
                self.time_conv = CausalConv3d(dim, dim, (3, 1, 1), stride=(2, 1, 1), padding=(0, 0, 0))


class ResidualBlock(nn.Module):
    """reference :185-219"""

    def __init__(self, in_dim, out_dim, dropout=0.0):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.residual = nn.Sequential(
            RMS_norm(in_dim, images=False), nn.SiLU(), CausalConv3d(in_dim, out_dim, 3, padding=1),
            RMS_norm(out_dim, images=False), nn.SiLU(), nn.Dropout(dropout),
            CausalConv3d(out_dim, out_dim, 3, padding=1))
        self.shortcut = CausalConv3d(in_dim, out_dim, 1) if in_dim != out_dim else nn.Identity()


class AttentionBlock(nn.Module):
    """reference :222-261"""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.norm = RMS_norm(dim)
        self.to_qkv = nn.Conv2d(dim, dim * 3, 1)
        self.proj = nn.Conv2d(dim, dim, 1)
        nn.init.zeros_(self.proj.weight)


class Encoder3d(nn.Module):
    """reference :264-365"""

    def __init__(self, dim=128, z_dim=4, dim_mult=(1, 2, 4, 4), num_res_blocks=2, attn_scales=(),
                 temperal_downsample=(True, True, False), dropout=0.0):
        super().__init__()
        dims = [dim * u for u in [1] + list(dim_mult)]
        scale = 1.0
        self.conv1 = CausalConv3d(3, dims[0], 3, padding=1)
        downs = []
        for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(num_res_blocks):
                downs.append(ResidualBlock(in_dim, out_dim, dropout))
                if scale in attn_scales:
                    downs.append(AttentionBlock(out_dim))
                in_dim = out_dim
            if i != len(dim_mult) - 1:
                downs.append(Resample(out_dim, "downsample3d" if temperal_downsample[i] else "downsample2d"))
                scale /= 2.0
        self.downsamples = nn.Sequential(*downs)
        self.middle = nn.Sequential(ResidualBlock(out_dim, out_dim, dropout), AttentionBlock(out_dim),
                                    ResidualBlock(out_dim, out_dim, dropout))
        self.head = nn.Sequential(RMS_norm(out_dim, images=False), nn.SiLU(),
                                  CausalConv3d(out_dim, z_dim, 3, padding=1))


class Decoder3d(nn.Module):
    """reference :368-471"""

    def __init__(self, dim=128, z_dim=4, dim_mult=(1, 2, 4, 4), num_res_blocks=2, attn_scales=(),
                 temperal_upsample=(False, True, True), dropout=0.0):
        super().__init__()
        dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
        scale = 1.0 / 2 ** (len(dim_mult) - 2)
        self.conv1 = CausalConv3d(z_dim, dims[0], 3, padding=1)
        self.middle = nn.Sequential(ResidualBlock(dims[0], dims[0], dropout), AttentionBlock(dims[0]),
                                    ResidualBlock(dims[0], dims[0], dropout))
        ups = []
        for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
            if i in (1, 2, 3):
                in_dim = in_dim // 2
            for _ in range(num_res_blocks + 1):
                ups.append(ResidualBlock(in_dim, out_dim, dropout))
                if scale in attn_scales:
                    ups.append(AttentionBlock(out_dim))
                in_dim = out_dim
            if i != len(dim_mult) - 1:
                ups.append(Resample(out_dim, "upsample3d" if temperal_upsample[i] else "upsample2d"))
                scale *= 2.0
        self.upsamples = nn.Sequential(*ups)
        self.head = nn.Sequential(RMS_norm(out_dim, images=False), nn.SiLU(),
                                  CausalConv3d(out_dim, 3, 3, padding=1))


class WanVAE_(nn.Module):
    """reference :482-589 (constructor and encode/decode signatures)."""

    def __init__(self, dim=128, z_dim=4, dim_mult=(1, 2, 4, 4), num_res_blocks=2, attn_scales=(),
                 temperal_downsample=(True, True, False), dropout=0.0):
        super().__init__()
        self.dim, self.z_dim = dim, z_dim
        self.dim_mult, self.num_res_blocks = list(dim_mult), num_res_blocks
        self.attn_scales = list(attn_scales)
        self.temperal_downsample = list(temperal_downsample)
        self.temperal_upsample = list(temperal_downsample)[::-1]
        self.encoder = Encoder3d(dim, z_dim * 2, dim_mult, num_res_blocks, attn_scales,
                                 self.temperal_downsample, dropout)
        self.conv1 = CausalConv3d(z_dim * 2, z_dim * 2, 1)
        self.conv2 = CausalConv3d(z_dim, z_dim, 1)
        self.decoder = Decoder3d(dim, z_dim, dim_mult, num_res_blocks, attn_scales,
                                 self.temperal_upsample, dropout)
        self._engine = None

    def engine(self, scale) -> "VAEEngine":
        mean, inv_std = scale
        if self._engine is None:
            self._engine = VAEEngine(self, torch.as_tensor(mean), 1.0 / torch.as_tensor(inv_std))
        return self._engine

    def encode(self, x, scale):
        """x: [1, 3, T, H, W] -> normalised mu [1, z, T', h, w] (reference :515-541)."""
        assert x.shape[0] == 1
        return self.engine(scale).encode([x[0]])[0].unsqueeze(0)

    def decode(self, z, scale):
        """z: [1, z, T, h, w] -> [1, 3, 1+4(T-1), 8h, 8w], NOT clamped (reference :543-567)."""
        assert z.shape[0] == 1
        return self.engine(scale).decode([z[0]], clamp=False)[0].unsqueeze(0)

    def clear_cache(self):  # the single-pass engine keeps no temporal cache (reference :581-589)
        pass


def _video_vae(pretrained_path=None, z_dim=None, device="cpu", **kwargs):
    """reference :591-615"""
    cfg = dict(dim=96, z_dim=z_dim, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[],
               temperal_downsample=[False, True, True], dropout=0.0)
    cfg.update(**kwargs)
    if pretrained_path is None:
        return WanVAE_(**cfg)
    with torch.device("meta"):
        model = WanVAE_(**cfg)
    logging.info(f"loading {pretrained_path}")
    model.load_state_dict(torch.load(pretrained_path, map_location=device), assign=True)
    return model


class Wan2_1_VAE:
    """reference :618-662.  `vae_pth=None` builds a randomly initialised model (tests / benchmarks
    without checkpoints)."""

    def __init__(self, z_dim=16, vae_pth="cache/vae_step_411000.pth", dtype=torch.float, device="cuda"):
        self.dtype, self.device = dtype, device
        self.mean = torch.tensor(_MEAN, dtype=dtype, device=device)
        self.std = torch.tensor(_STD, dtype=dtype, device=device)
        self.scale = [self.mean, 1.0 / self.std]
        self.model = _video_vae(pretrained_path=vae_pth, z_dim=z_dim).eval().requires_grad_(False).to(device)

    def _engine(self) -> "VAEEngine":
        return self.model.engine(self.scale)

    def set_strip_group(self, group):
        """Multi-GPU: every rank of `group` calls encode/decode with the SAME full input and gets
        the full output; the work is split into H-strips with halo exchange (None: single GPU)."""
        self._engine().set_strip_group(group)

    def encode(self, videos):
        """videos: list of [3, T, H, W] -> list of fp32 [16, T', H/8, W/8]."""
        return self._engine().encode(videos)

    def decode(self, zs):
        """zs: list of [16, T, h, w] -> list of fp32 [3, 1+4(T-1), 8h, 8w] clamped to [-1, 1]."""
        return self._engine().decode(zs, clamp=True)


# ------------------------------------------------------------------------------------------------
# Spatial strips across GPUs: rank r of p owns latent rows [h0, h1) of every
# activation (x 2^k at finer levels).  Per-voxel work is strip-local; a convolution with a spatial
# extent needs one row of halo from each neighbour at ITS input resolution.  In strip mode every
# activation therefore lives in a tensor with one halo slot above and below its rows
# ([T, 1 + Hl + 1, W, C]; `VAEEngine._pad`): the kernels write the interior through their output
# strides, the two boundary rows are exchanged in place right before a spatial conv (`halo_fill`),
# and the conv reads the padded tensor as it is — the activation itself is never copied.  The
# single-head attention of the middle block gathers the (latent-resolution) rows of all strips.  The reference decodes on one rank only
# (x2x_gen_unified.py:512-524); this is the multi-GPU layout of BASELINE.json configs[4].
# ------------------------------------------------------------------------------------------------
class Strip:
    def __init__(self, group):
        import torch.distributed as dist
        self.group = group
        self.p = dist.get_world_size(group)
        self.r = dist.get_rank(group)
        self.ranks = dist.get_process_group_ranks(group)

    def rows(self, H: int, r: Optional[int] = None):
        """[h0, h1) of rank r when H latent rows are cut into p nearly equal strips."""
        r = self.r if r is None else r
        return (r * H) // self.p, ((r + 1) * H) // self.p


def halo_fill(xp: torch.Tensor, strip: Strip, top: bool, bottom: bool) -> None:
    """xp: this rank's rows WITH one halo slot above and below, [T, 1 + Hl + 1, W, C].  Fills slot 0 with
    the upper neighbour's last row (if `top`) and slot Hl+1 with the lower neighbour's first row (if
    `bottom`); zeros at the image border, which is exactly the convolution's zero padding there.
    In place: the activation is allocated padded by the kernel that produces it (the conv writes the
    interior rows through its output strides), so no copy of the activation is made for the halo —
    only the two boundary rows travel."""
    import torch.distributed as dist
    r, p = strip.r, strip.p
    ops_, recv_top, recv_bot = [], None, None
    # my first row is the upper neighbour's bottom halo; my last row the lower neighbour's top halo
    if r > 0:
        if bottom:
            ops_.append(dist.P2POp(dist.isend, xp[:, 1].contiguous(), strip.ranks[r - 1], group=strip.group))
        if top:
            recv_top = torch.empty_like(xp[:, 0])
            ops_.append(dist.P2POp(dist.irecv, recv_top, strip.ranks[r - 1], group=strip.group))
    if r < p - 1:
        if top:
            ops_.append(dist.P2POp(dist.isend, xp[:, -2].contiguous(), strip.ranks[r + 1], group=strip.group))
        if bottom:
            recv_bot = torch.empty_like(xp[:, 0])
            ops_.append(dist.P2POp(dist.irecv, recv_bot, strip.ranks[r + 1], group=strip.group))
    if ops_:
        for req in dist.batch_isend_irecv(ops_):
            req.wait()
    if top:
        if recv_top is not None:
            xp[:, 0].copy_(recv_top)
        else:
            xp[:, 0].zero_()
    if bottom:
        if recv_bot is not None:
            xp[:, -1].copy_(recv_bot)
        else:
            xp[:, -1].zero_()


def gather_rows(x: torch.Tensor, strip: Strip, dim: int, sizes: List[int]) -> torch.Tensor:
    """All-gather strips of unequal height along `dim` (padded to the tallest strip for NCCL)."""
    import torch.distributed as dist
    hmax = max(sizes)
    xm = x.movedim(dim, 0).contiguous()
    if xm.shape[0] < hmax:
        pad = torch.zeros((hmax - xm.shape[0],) + tuple(xm.shape[1:]), device=x.device, dtype=x.dtype)
        xm = torch.cat([xm, pad], dim=0)
    full = torch.empty((strip.p * hmax,) + tuple(xm.shape[1:]), device=x.device, dtype=x.dtype)
    dist.all_gather_into_tensor(full, xm, group=strip.group)   # concatenated form (gloo and NCCL)
    out = torch.cat([full[i * hmax:i * hmax + n] for i, n in enumerate(sizes)], dim=0)
    return out.movedim(0, dim).contiguous()


# ------------------------------------------------------------------------------------------------
# Engine
# ------------------------------------------------------------------------------------------------
class VAEEngine:
    """Runs encode/decode of a `WanVAE_`-shaped module (ours or the reference's) on libovb200."""

    def __init__(self, model: nn.Module, mean: torch.Tensor, std: torch.Tensor):
        self.m = model
        self.mean_host, self.std_host = mean.detach().float().cpu(), std.detach().float().cpu()
        self._packed: Dict[str, object] = {}
        self._key = None
        self.strip: Optional[Strip] = None   # set_strip_group()
        self._lat_H = None                   # latent rows of the clip being processed (strip mode)

    def set_strip_group(self, group):
        """Run encode/decode as H-strips over the ranks of `group` (None: single GPU)."""
        self.strip = None if group is None else Strip(group)
        if self.strip is not None and self.strip.p == 1:
            self.strip = None

    @staticmethod
    def _check_device(dev):
        """The product has no CPU path: parameters must be on the GPU the kernels run on."""
        if dev.type != "cuda":
            raise ops._lib.OvbError("VAE parameters must live on a CUDA device: the B200 path has "
                                    "no CPU fallback")

    # -------------------------------------------------------------------------------- weights
    def _params(self) -> Dict[str, torch.Tensor]:
        key = tuple((p.device, p._version, p.data_ptr()) for p in self.m.parameters())
        if key != self._key:
            self._packed, self._key = {}, key
            self._sd = {k: v for k, v in self.m.state_dict().items()}
            dev = next(self.m.parameters()).device
            self._check_device(dev)
            self.dev = dev
            z = self.mean_host.numel()
            self.mean, self.std = self.mean_host.to(dev), self.std_host.to(dev)
            self._ones = {n: torch.ones(n, device=dev, dtype=f32) for n in (3, z, 2 * z)}
            self._zeros = {n: torch.zeros(n, device=dev, dtype=f32) for n in (3, z, 2 * z)}
        return self._sd

    def _conv(self, name: str, **kw) -> vae_ops.PackedConv:
        if name not in self._packed:
            sd = self._params()
            self._packed[name] = vae_ops.pack_conv3d(sd[name + ".weight"], sd[name + ".bias"], **kw)
        return self._packed[name]

    def _head_planes(self, name: str):
        key = (name, "planes")
        if key not in self._packed:
            sd = self._params()
            self._packed[key] = vae_ops.pack_tap_planes(sd[name + ".weight"], sd[name + ".bias"])
        return self._packed[key]

    def _gamma(self, name: str) -> torch.Tensor:
        if name not in self._packed:
            self._packed[name] = self._params()[name].detach().float().reshape(-1).contiguous()
        return self._packed[name]

    def _mat(self, name: str, rows: Optional[slice] = None) -> torch.Tensor:
        key = (name, None if rows is None else (rows.start, rows.stop))
        if key not in self._packed:
            w = self._params()[name].detach()
            w = w.reshape(w.shape[0], -1)
            if rows is not None:
                w = w[rows]
            self._packed[key] = w.to(bf16).contiguous()
        return self._packed[key]

    # -------------------------------------------------------------------------------- blocks
    # Activations are [T, pad + H + pad, W, C] with pad = 1 in strip mode (one halo slot above and below
    # this rank's H rows), pad = 0 on a single GPU; everything below is written once for both.
    @property
    def _pad(self) -> int:
        return 0 if self.strip is None else 1

    def _new_act(self, T, H, W, C):
        return torch.empty((T, H + 2 * self._pad, W, C), device=self.dev, dtype=bf16)

    def _interior(self, x):
        return x if self._pad == 0 else x[:, 1:-1]

    def _to_act(self, x):
        """A dense [T, H, W, C] tensor -> the engine's activation layout."""
        if self._pad == 0:
            return x
        y = self._new_act(x.shape[0], x.shape[1], x.shape[2], x.shape[3])
        y[:, 1:-1].copy_(x)
        return y

    def _apply_conv(self, x, pc, out_hw=None, resid=None, want_norm_out=False, **kw):
        """ovb_conv3d on the activation layout: fills the halo rows `pc` reads (strip mode), lets the conv
        read the padded tensor as it is (every tap's row offset shifted by the pad) and write the interior
        rows of a new padded tensor through its output strides."""
        pad = self._pad
        T, Hp, W, _ = x.shape
        lo, hi = pc.spatial_rows
        if pad and (lo != 0 or hi != 0):
            halo_fill(x, self.strip, lo < 0, hi > 0)
        H, Wo = out_hw if out_hw is not None else (Hp - 2 * pad, W)
        co = pc.cout
        out = self._new_act(T, H, Wo, co)
        if want_norm_out:
            kw["norm_out"] = self._new_act(T, H, Wo, co)
        vae_ops.conv3d(x, pc, (T, H, Wo), out=out, out_strides=((H + 2 * pad) * Wo * co, Wo * co, co),
                       out_base=pad * Wo * co, resid=resid, dh_shift=pad, **kw)
        return (out, kw["norm_out"]) if want_norm_out else out

    def _causal_conv(self, x, name, resid=None, **kw):
        return self._apply_conv(x, self._conv(name), resid=resid, **kw)

    def _residual_block(self, x, p, x_norm=None, next_gamma=None, raw_needed=True):
        """reference vae2_1.py:185-219: x + conv(act(norm(conv(act(norm(x)))))) (+ 1x1x1 shortcut).
        Both RMS_norm·SiLU passes disappear into conv epilogues where one N tile holds all channels
        (OVB_FUSE_NORM=0 keeps them as separate launches):
          * residual.3 (norm of the first conv's output, used by nothing else) is applied by that
            conv's epilogue;
          * the norm that CONSUMES this block's output (the next block's residual.0, or the decoder/
            encoder head norm: `next_gamma`) is written by the second conv's epilogue as a second
            tensor, next to the raw sum the next block needs as its residual input.
        x_norm: act(norm(x)) already produced by the previous block.  Returns (y, act(norm_next(y)) or
        None); y is None when the caller declared it unused (raw_needed=False)."""
        sd = self._params()
        fuse = os.environ.get("OVB_FUSE_NORM", "1") != "0"
        h = self._causal_conv(x, p + "shortcut") if (p + "shortcut.weight") in sd else x
        a = x_norm if x_norm is not None else vae_ops.rmsnorm_silu(x, self._gamma(p + "residual.0.gamma"))
        g3 = self._gamma(p + "residual.3.gamma")
        if fuse and vae_ops.can_fuse_norm(self._conv(p + "residual.2")):
            b = self._causal_conv(a, p + "residual.2", norm_gamma=g3)
        else:
            b = self._causal_conv(a, p + "residual.2")
            vae_ops.rmsnorm_silu(b, g3, out=b)                                  # in place: elementwise
        del a
        pc2 = self._conv(p + "residual.6")
        # the two-tensor form keeps a tile's values in registers between the statistics and the
        # normalised store, which only fits up to 96 channels; wider layers keep the separate pass
        dual_ok = pc2.cout <= 96
        if (next_gamma is not None and fuse and vae_ops.can_fuse_norm(pc2) and next_gamma.numel() == pc2.cout
                and (dual_ok or not raw_needed)):
            if raw_needed:
                return self._causal_conv(b, p + "residual.6", resid=h, norm_gamma=next_gamma, want_norm_out=True)
            return None, self._causal_conv(b, p + "residual.6", resid=h, norm_gamma=next_gamma)
        y = self._causal_conv(b, p + "residual.6", resid=h)
        return y, None

    def _attention_block(self, x, p):
        if self._pad:
            # latent resolution, 3 % of the FLOPs: on a dense copy of this rank's rows
            return self._to_act(self._attention_dense(self._interior(x).contiguous(), p))
        return self._attention_dense(x, p)

    def _attention_dense(self, x, p):
        T, H, W, Cc = x.shape
        hw = H * W
        if hw % 8 != 0 and self.strip is None:
            raise NotImplementedError(f"VAE attention needs h*w % 8 == 0 (got {H}x{W})")
        x2 = x.view(T * hw, Cc)
        xn = vae_ops.rmsnorm_silu(x2, self._gamma(p + "norm.gamma"), silu=False)
        sd = self._params()
        bias = sd[p + "to_qkv.bias"].detach().to(bf16)
        wqk = self._mat(p + "to_qkv.weight", slice(0, 2 * Cc))
        wv = self._mat(p + "to_qkv.weight", slice(2 * Cc, 3 * Cc))
        # queries: this rank's rows; keys/values: every row of the frame (strip mode gathers the
        # normalised latent-resolution rows of all ranks; K and V^T are then computed replicated)
        if self.strip is not None:
            sizes = [self.strip.rows(self._lat_H, r)[1] - self.strip.rows(self._lat_H, r)[0]
                     for r in range(self.strip.p)]
            xk = gather_rows(xn.view(T, H, W, Cc), self.strip, 1, sizes)        # [T, H_all, W, C]
            hk = xk.shape[1] * W
            if hk % 8 != 0:
                raise NotImplementedError(f"VAE attention needs h*w % 8 == 0 (got {xk.shape[1]}x{W})")
            xk = xk.view(T * hk, Cc)
        else:
            xk, hk = xn, hw
        q = ops.gemm(xn, wqk[:Cc], bias[:Cc].contiguous(), epilogue=EPI_BF16)           # [T*hw, C]
        k = ops.gemm(xk, wqk[Cc:], bias[Cc:2 * Cc].contiguous(), epilogue=EPI_BF16)     # [T*hk, C]
        o = torch.empty((T * hw, Cc), device=x.device, dtype=bf16)
        S = torch.empty((hw, hk), device=x.device, dtype=f32)
        P = torch.empty((hw, hk), device=x.device, dtype=bf16)
        vt = torch.empty((Cc, hk), device=x.device, dtype=bf16)
        bv = bias[2 * Cc:].contiguous()
        scale = 1.0 / math.sqrt(Cc)
        for f in range(T):
            rows = slice(f * hw, (f + 1) * hw)
            krows = slice(f * hk, (f + 1) * hk)
            ops.gemm(wv, xk[krows], None, epilogue=EPI_BF16, out=vt)             # V^T = Wv Xn^T   [C, hk]
            ops.gemm(q[rows], k[krows], None, epilogue=EPI_F32, out=S)            # Q K^T          [hw, hk]
            vae_ops.softmax_rows(S, scale, out=P)
            ops.gemm(P, vt, bv, epilogue=EPI_BF16, out=o[rows])                   # P V + b_v (rows of P sum to 1)
        wp = self._mat(p + "proj.weight")
        bp = sd[p + "proj.bias"].detach().to(bf16).contiguous()
        y = ops.gemm(o, wp, bp, epilogue=EPI_RESID_BF16, resid=x2)
        return y.view(T, H, W, Cc)

    def _upsample(self, x, p, three_d: bool):
        pad = self._pad
        T, Hp, W, Cc = x.shape
        H = Hp - 2 * pad
        sd = self._params()
        if three_d and T > 1:
            # time_conv over frames 1.. (zero left context), channel halves -> consecutive frames.  No
            # spatial extent: the halo slots are processed like rows (and refilled before the next use)
            pc = self._conv(p + "time_conv")
            y = torch.empty((1 + 2 * (T - 1), Hp, W, Cc), device=x.device, dtype=bf16)
            y[0].copy_(x[0])
            fr = Hp * W * Cc
            vae_ops.conv3d(x[1:], pc, (T - 1, Hp, W), out=y, out_strides=(2 * fr, W * Cc, Cc), out_base=fr,
                           n_group=Cc, group_stride=fr)
            x, T = y, y.shape[0]
        key = p + "resample.1#phases"
        if key not in self._packed:
            self._packed[key] = vae_ops.pack_upsample_conv2d(sd[p + "resample.1.weight"],
                                                             sd[p + "resample.1.bias"])
        phases = self._packed[key]
        Co = Cc // 2
        out = self._new_act(T, 2 * H, 2 * W, Co)
        if pad:  # one halo fill (both neighbours) serves the four phases
            halo_fill(x, self.strip, True, True)
        row = 2 * W * Co                        # one output row
        for (a, b), pc in phases.items():
            vae_ops.conv3d(x, pc, (T, H, W), out=out, out_strides=((2 * H + 2 * pad) * row, 2 * row, 2 * Co),
                           out_base=(pad + a) * row + b * Co, dh_shift=pad)
        return out

    def _downsample(self, x, p, three_d: bool):
        pad = self._pad
        T, Hp, W, Cc = x.shape
        H = Hp - 2 * pad
        Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1                     # ZeroPad2d(0,1,0,1) + k3 s2
        pc = self._conv(p + "resample.1", stride=(1, 2, 2), spatial_pad=(0, 0))
        if pad:
            assert H % 2 == 0, "strip heights must be even at every stride-2 level"
            Ho = H // 2                          # rows 2h, 2h+1, 2h+2: the last one may be the lower halo
        x = self._apply_conv(x, pc, out_hw=(Ho, Wo))
        if three_d and T > 1:
            # out[0] = x[0]; out[j] = time_conv(x[2j-2], x[2j-1], x[2j])  (stride 2, no padding); no spatial
            # extent: halo slots are processed like rows
            To = 1 + (T - 1) // 2
            pt = self._conv(p + "time_conv", causal_pad=0, stride=(2, 1, 1))
            Hq = Ho + 2 * pad
            y = torch.empty((To, Hq, Wo, Cc), device=x.device, dtype=bf16)
            y[0].copy_(x[0])
            fr = Hq * Wo * Cc
            vae_ops.conv3d(x, pt, (To - 1, Hq, Wo), out=y, out_strides=(fr, Wo * Cc, Cc), out_base=fr)
            x = y
        return x

    def _stack(self, x, prefix, up: bool, tail_gamma=None):
        """Runs the numbered modules under `prefix`.  tail_gamma: gamma of the RMS_norm·SiLU that follows
        the stack and is the ONLY consumer of its output (the head norm): when the last module is a
        residual block whose conv can fuse it, the stack returns that normalised tensor and the caller
        skips the norm (`self._tail_fused` tells)."""
        sd = self._params()
        pat = re.compile(re.escape(prefix) + r"(\d+)\.")
        idx = sorted({int(mm.group(1)) for k in sd for mm in [pat.match(k)] if mm})
        x_norm = None                      # act(norm(x)) for the next residual block, if already made
        self._tail_fused = False
        for n, i in enumerate(idx):
            p = f"{prefix}{i}."
            if (p + "residual.2.weight") in sd:
                nxt = f"{prefix}{idx[n + 1]}." if n + 1 < len(idx) else None
                if nxt is not None and (nxt + "residual.0.gamma") in sd:
                    x, x_norm = self._residual_block(x, p, x_norm, self._gamma(nxt + "residual.0.gamma"))
                elif nxt is None and tail_gamma is not None:
                    y, yn = self._residual_block(x, p, x_norm, tail_gamma, raw_needed=False)
                    x, x_norm = (yn, None) if yn is not None else (y, None)
                    self._tail_fused = yn is not None
                else:
                    x, x_norm = self._residual_block(x, p, x_norm)
                continue
            x_norm = None
            if (p + "to_qkv.weight") in sd:
                x = self._attention_block(x, p)
            elif (p + "resample.1.weight") in sd:
                three_d = (p + "time_conv.weight") in sd
                x = self._upsample(x, p, three_d) if up else self._downsample(x, p, three_d)
            else:  # pragma: no cover
                raise KeyError(f"unknown VAE module at {p}")
        return x

    # -------------------------------------------------------------------------------- entry points
    @torch.no_grad()
    def decode(self, zs: List[torch.Tensor], clamp: bool = True) -> List[torch.Tensor]:
        self._params()
        outs = []
        for z in zs:
            z = z.to(self.dev, dtype=f32).contiguous()
            zc = z.shape[0]
            x = vae_ops.planar_to_cl(z, self.std, self.mean, 16)                 # z*std + mean (:546-550)
            assert zc == 16, "Wan2.1 VAE has z_dim 16"
            if self.strip is not None:
                self._lat_H = x.shape[1]
                h0, h1 = self.strip.rows(self._lat_H)
                assert h1 > h0, "more strips than latent rows"
                x = x[:, h0:h1]
            x = self._to_act(x)
            x = self._causal_conv(x, "conv2")                                    # 1x1x1 (:552)
            x = self._causal_conv(x, "decoder.conv1")
            x = self._stack(x, "decoder.middle.", up=True)
            head_gamma = self._gamma("decoder.head.0.gamma")
            x = self._stack(x, "decoder.upsamples.", up=True, tail_gamma=head_gamma)
            if not self._tail_fused:
                vae_ops.rmsnorm_silu(x, head_gamma, out=x)
            pad = self._pad
            T, Hp, W, _ = x.shape
            H = Hp - 2 * pad
            out = torch.empty((3, T, H, W), device=self.dev, dtype=f32)
            if pad:
                halo_fill(x, self.strip, True, True)
            if os.environ.get("OVB_HEAD_TAPSUM", "1") != "0":
                # 96 -> 3 channels: as an implicit GEMM every MMA would fetch a full A tile for 3 useful columns.
                # Instead ONE 1x1x1 conv (K = 96, N = 27 taps x 4) writes the per-tap partial products as planes
                # and an HBM-bound pass shifts and adds them (ovb200.h: ovb_tap_sum)
                pc_planes, taps, bias = self._head_planes("decoder.head.2")
                y = torch.empty((len(taps), T, Hp, W, 4), device=self.dev, dtype=bf16)
                vae_ops.conv3d(x, pc_planes, (T, Hp, W), out=y, out_strides=(Hp * W * 4, W * 4, 4), n_group=4,
                               group_stride=T * Hp * W * 4)
                vae_ops.tap_sum(y, taps, bias, 3, (T, H, W), dh_shift=pad, clamp=clamp, out=out)
                del y
            else:
                vae_ops.conv3d(x, self._conv("decoder.head.2"), (T, H, W), out=out, out_strides=(H * W, W, 1),
                               dh_shift=pad,
                               out_mode=CONV_OUT_F32_PLANAR_CLAMP if clamp else vae_ops.CONV_OUT_F32_PLANAR,
                               plane_stride=T * H * W)
            if self.strip is not None:  # every rank ends up with the whole clip
                scale_h = H // (self.strip.rows(self._lat_H)[1] - self.strip.rows(self._lat_H)[0])
                sizes = [(self.strip.rows(self._lat_H, r)[1] - self.strip.rows(self._lat_H, r)[0]) * scale_h
                         for r in range(self.strip.p)]
                out = gather_rows(out, self.strip, 2, sizes)
            outs.append(out)
        return outs

    @torch.no_grad()
    def encode(self, videos: List[torch.Tensor]) -> List[torch.Tensor]:
        self._params()
        outs = []
        for v in videos:
            v = v.to(self.dev, dtype=f32).contiguous()
            assert v.shape[0] == 3 and (v.shape[1] - 1) % 4 == 0, "video must be [3, 1+4k, H, W]"
            if self.strip is not None:
                assert v.shape[2] % 8 == 0
                self._lat_H = v.shape[2] // 8
                h0, h1 = self.strip.rows(self._lat_H)
                assert h1 > h0, "more strips than latent rows"
                v = v[:, :, 8 * h0:8 * h1].contiguous()
            x = self._to_act(vae_ops.planar_to_cl(v, self._ones[3], self._zeros[3], 16))
            x = self._causal_conv(x, "encoder.conv1")
            x = self._stack(x, "encoder.downsamples.", up=False)
            head_gamma = self._gamma("encoder.head.0.gamma")
            x = self._stack(x, "encoder.middle.", up=False, tail_gamma=head_gamma)
            if not self._tail_fused:
                vae_ops.rmsnorm_silu(x, head_gamma, out=x)
            x = self._causal_conv(x, "encoder.head.2")                            # [T', h, w, 32]
            x = self._causal_conv(x, "conv1")                                     # 1x1x1 (:534)
            z = self.mean.numel()
            inv = 1.0 / self.std
            x = self._interior(x)
            mu = vae_ops.cl_to_planar(x if x.is_contiguous() else x.contiguous(), inv, -self.mean * inv, z)  # (:535-539)
            if self.strip is not None:
                sizes = [self.strip.rows(self._lat_H, r)[1] - self.strip.rows(self._lat_H, r)[0]
                         for r in range(self.strip.p)]
                mu = gather_rows(mu, self.strip, 2, sizes)
            outs.append(mu)
        return outs
