"""The output-video restatement (oracle/video.py) is held to frames captured from the reference's own
`cache_video` (tests/golden/video_*.npz, oracle/gen_golden_video.py).  CPU only, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import synth
from oracle import video as V
from tests.golden_util import golden_names, load_golden


def video_case(meta):
    return torch.stack([synth.vae_video(tuple(meta["shape"]), seed=meta["seed0"] + b) * meta["scale"]
                        for b in range(meta["batch"])])


@pytest.mark.parametrize("name", golden_names("video_frames"))
def test_oracle_frames_match_reference_cache_video(name):
    meta, gold = load_golden(name)
    vids = video_case(meta)
    got = V.frames_uint8(vids.numpy(), nrow=meta["nrow"], value_range=(-1, 1))
    assert got.dtype == np.uint8 and got.shape == tuple(gold[0].shape)
    assert np.array_equal(got, gold[0].numpy())
    assert (vids.abs() > 1).any(), "the fixture must exercise the clamp"
