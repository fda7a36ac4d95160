"""Output video path (drop-in for `cache_video` of reference omnivideo/utils/utils.py:160-198).

The reference clamps, grids, normalises and quantises the decoded video with a chain of ATen ops and only then
copies it to the host; here that chain is ONE launch (`ovb_video_to_uint8`) that reads the VAE's planar fp32 output
once and writes interleaved uint8 frames, so the device->host copy moves 3 bytes per pixel instead of 12 and no
fp32 intermediate is materialised.  Encoding the frames (libx264 through imageio) stays host-side library work,
exactly as in the reference.
"""
from __future__ import annotations

import binascii
import ctypes as C
import math
import os
import os.path as osp
from typing import Optional

import torch

from .. import _lib

__all__ = ["video_frames_uint8", "cache_video"]


def video_frames_uint8(tensor: torch.Tensor, nrow: int = 8, normalize: bool = True, value_range=(-1, 1),
                       padding: int = 2) -> torch.Tensor:
    """tensor: fp32 [B, C, T, H, W] (or [C, T, H, W]) on the device -> uint8 [T, Hg, Wg, C] on the device, laid out
    as `torchvision.utils.make_grid(nrow, normalize=True, value_range)` lays out the B videos frame by frame
    (2-pixel zero border and gutters; a single video is its own grid)."""
    if not normalize:
        raise NotImplementedError("cache_video is only called with normalize=True on the path")
    if tensor.dim() == 4:
        tensor = tensor.unsqueeze(0)
    if not tensor.is_cuda:
        raise RuntimeError("video_frames_uint8 needs a CUDA tensor; there is no CPU fallback")
    B, Cc, T, H, W = tensor.shape
    lo, hi = float(min(value_range)), float(max(value_range))
    if B == 1:
        ymaps = xmaps = 1
        Hg, Wg, pad = H, W, 0
        frames = torch.empty((T, Hg, Wg, Cc), device=tensor.device, dtype=torch.uint8)
    else:
        xmaps = min(nrow, B)
        ymaps = int(math.ceil(B / xmaps))
        pad = padding
        Hg, Wg = (H + pad) * ymaps + pad, (W + pad) * xmaps + pad
        frames = torch.zeros((T, Hg, Wg, Cc), device=tensor.device, dtype=torch.uint8)      # pad_value 0
    lib = _lib.load()
    stream = C.c_void_p(torch.cuda.current_stream(tensor.device).cuda_stream)
    for k in range(B):
        y, x = divmod(k, xmaps)
        origin = ((y * (H + pad) + pad) * Wg + x * (W + pad) + pad) * Cc
        src = tensor[k].to(torch.float32).contiguous()
        _lib.check(lib.ovb_video_to_uint8(C.c_void_p(src.data_ptr()), C.c_void_p(frames.data_ptr() + origin), Cc, T, H, W,
                                          Hg * Wg * Cc, Wg * Cc, lo, hi, stream), "ovb_video_to_uint8")
    return frames


def cache_video(tensor, save_file: Optional[str] = None, fps: int = 30, suffix: str = ".mp4", nrow: int = 8,
                normalize: bool = True, value_range=(-1, 1), retry: int = 5):
    """Same signature and return value as the reference's `cache_video`: writes `tensor` ([B, C, T, H, W]) as a video
    and returns the file name, or None when writing fails `retry` times."""
    if save_file is None:
        name = binascii.b2a_hex(os.urandom(8)).decode("utf-8")
        save_file = osp.join("/tmp", name + (suffix if suffix.startswith(".") else "." + suffix))
    frames = video_frames_uint8(tensor, nrow=nrow, normalize=normalize, value_range=value_range).cpu().numpy()
    error = None
    for _ in range(retry):
        try:
            import imageio
            writer = imageio.get_writer(save_file, fps=fps, codec="libx264", quality=8)
            for frame in frames:
                writer.append_data(frame)
            writer.close()
            return save_file
        except Exception as ex:  # noqa: BLE001 - the reference retries on anything and reports the last error
            error = ex
    print(f"cache_video failed, error: {error}", flush=True)
    return None
