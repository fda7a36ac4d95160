"""Output video path on the GPU (omni_video_b200/utils/video.py, `ovb_video_to_uint8`) against frames captured
from the reference's own `cache_video` — bit-exact (passed on B200 at the end of round 1)."""
import pytest
import torch

from tests.golden_util import golden_names, load_golden
from tests.test_oracle_video_golden import video_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


@pytest.mark.parametrize("name", golden_names("video_frames"))
def test_frames_match_reference_cache_video(name):
    from omni_video_b200.utils.video import video_frames_uint8
    meta, gold = load_golden(name)
    vids = video_case(meta).cuda()
    frames = video_frames_uint8(vids, nrow=meta["nrow"], normalize=True, value_range=(-1, 1))
    assert frames.dtype == torch.uint8 and frames.shape == tuple(gold[0].shape)
    assert torch.equal(frames.cpu(), gold[0])


def test_cache_video_reports_failure_like_the_reference(tmp_path, capsys):
    """imageio is not installed in this image: the writer fails, the function prints and returns None
    (reference utils.py:191-198) after having produced the frames on the device."""
    from omni_video_b200.utils.video import cache_video
    try:
        import imageio  # noqa: F401
        pytest.skip("imageio present")
    except ImportError:
        pass
    out = cache_video(torch.zeros(1, 3, 2, 8, 8, device="cuda"), save_file=str(tmp_path / "x.mp4"), nrow=1, retry=2)
    assert out is None and "cache_video failed" in capsys.readouterr().out
