"""omni_video_b200 — B200-native (sm_100a) denoising hot path of Omni-Video 2.

Public surface mirrors the reference's operator names:
  modules.model.WanModel              (reference omnivideo/modules/model.py)
  modules.attention.flash_attention   (reference omnivideo/modules/attention.py)
  modules.vae2_1.Wan2_1_VAE           (reference omnivideo/modules/vae2_1.py)
  distributed.*                       (reference omnivideo/distributed/{util,ulysses,full_sequence_parallel}.py)
  patch.patch_reference_model         rebinds a live reference model onto these kernels

All arithmetic runs in `libovb200.so` (include/ovb200.h).  There is no CPU or PyTorch fallback.
"""
from . import _lib, ops  # noqa: F401

__version__ = "0.1.0"
