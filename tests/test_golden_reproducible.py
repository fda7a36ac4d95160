"""Build container only (skipped where /root/reference is absent): the committed fixtures are what the
committed generator scripts produce from the reference TODAY, bit for bit — the fast generators here, the DiT /
VAE ones (a minute of CPU) were checked the same way by hand at the end of round 1."""
import glob
import importlib
import json
import os

import numpy as np
import pytest

from oracle import ref_loader
from tests.golden_util import GOLDEN_DIR

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.mark.timeout(600)
@pytest.mark.parametrize("module,prefix", [("oracle.gen_golden_unified", "unified_"), ("oracle.gen_golden_t5", "t5_"),
                                           ("oracle.gen_golden_video", "video_"), ("oracle.gen_golden_unipc", "unipc_"),
                                           ("oracle.gen_golden_dpm", "dpm_")])
def test_generator_reproduces_committed_fixtures(tmp_path, monkeypatch, module, prefix):
    import sys
    gen = importlib.import_module(module)
    monkeypatch.setattr(gen, "GOLDEN_DIR", str(tmp_path))
    before = set(sys.modules)
    try:
        gen.main()
    finally:                      # the generators register stand-ins (imageio, tokenizers): do not leak them
        for name in ("imageio", "omnivideo.modules.tokenizers"):
            if name in sys.modules and name not in before:
                del sys.modules[name]
    made = sorted(glob.glob(os.path.join(str(tmp_path), prefix + "*.npz")))
    committed = sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))
    assert [os.path.basename(p) for p in made] == [os.path.basename(p) for p in committed] and made
    for new, old in zip(made, committed):
        a, b = np.load(new, allow_pickle=False), np.load(old, allow_pickle=False)
        assert set(a.files) == set(b.files)
        for k in a.files:
            if k == "meta":
                assert json.loads(str(a[k])) == json.loads(str(b[k]))
            else:
                assert np.array_equal(a[k], b[k]), (os.path.basename(new), k)
