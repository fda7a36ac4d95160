"""ORACLE support — frames of the UNMODIFIED reference `cache_video` (omnivideo/utils/utils.py:160-198).
Run in the build container:   python -m oracle.gen_golden_video

`imageio` (absent here) is replaced by a recorder that keeps what `writer.append_data` receives: the uint8
frames the reference hands to libx264.  Inputs are regenerated from seeds by the tests."""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader, synth  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [dict(name="video_single", batch=1, shape=(5, 24, 40), nrow=1, scale=1.6),      # the pipeline's call: video[None], nrow=1
         dict(name="video_grid3", batch=3, shape=(2, 10, 14), nrow=2, scale=1.3)]      # a grid with an empty cell


def main():
    frames = []
    fake = types.ModuleType("imageio")

    class Writer:
        def append_data(self, f):
            frames.append(np.array(f))

        def close(self):
            pass

    fake.get_writer = lambda *a, **k: Writer()
    sys.modules["imageio"] = fake
    utils = ref_loader._load("omnivideo_utils_utils", "omnivideo/utils/utils.py")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for case in CASES:
        del frames[:]
        vids = torch.stack([synth.vae_video(case["shape"], seed=31 + b) * case["scale"] for b in range(case["batch"])])
        out = utils.cache_video(vids, save_file="/tmp/unused.mp4", nrow=case["nrow"], normalize=True, value_range=(-1, 1))
        assert out is not None and len(frames) == case["shape"][0]
        meta = dict(kind="video_frames", batch=case["batch"], shape=case["shape"], nrow=case["nrow"], scale=case["scale"],
                    seed0=31, reference="omnivideo/utils/utils.py cache_video, frames captured at writer.append_data")
        path = os.path.join(GOLDEN_DIR, case["name"] + ".npz")
        np.savez_compressed(path, meta=json.dumps(meta), out0=np.stack(frames))
        print(f"wrote {path}: frames {np.stack(frames).shape} uint8, range {np.stack(frames).min()}..{np.stack(frames).max()}")


if __name__ == "__main__":
    main()
