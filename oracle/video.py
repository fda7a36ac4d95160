"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the tensor part of the reference's `cache_video` (omnivideo/utils/utils.py:177-185): clamp to
`value_range`, per frame `torchvision.utils.make_grid(nrow, normalize=True, value_range)` over the batch of videos
(padding 2, pad value 0; a single video is returned as is), `* 255`, cast to uint8 (truncation).  Pinned to frames
captured from the reference's own `cache_video` (tests/golden/video_*.npz, oracle/gen_golden_video.py)."""
from __future__ import annotations

import math

import numpy as np


def frames_uint8(video: np.ndarray, nrow: int = 8, value_range=(-1, 1), padding: int = 2) -> np.ndarray:
    """video: fp32 [B, C, T, H, W] -> uint8 [T, Hg, Wg, C]."""
    lo, hi = float(min(value_range)), float(max(value_range))
    B, C, T, H, W = video.shape
    v = np.clip(video.astype(np.float32), np.float32(lo), np.float32(hi))
    v = (v - np.float32(lo)) / np.float32(max(hi - lo, 1e-5))          # norm_ip of make_grid
    if B == 1:
        grid = v[0]                                                      # make_grid returns the single image
    else:
        xmaps = min(nrow, B)
        ymaps = int(math.ceil(B / xmaps))
        h, w = H + padding, W + padding
        grid = np.zeros((C, T, h * ymaps + padding, w * xmaps + padding), dtype=np.float32)
        for k in range(B):
            y, x = divmod(k, xmaps)
            grid[:, :, y * h + padding:y * h + padding + H, x * w + padding:x * w + padding + W] = v[k]
    return (grid * np.float32(255)).astype(np.uint8).transpose(1, 2, 3, 0)      # [T, Hg, Wg, C]
