"""omni_video_b200 — B200-native (sm_100a) denoising hot path of Omni-Video 2.

Public surface mirrors the reference's operator names:
  modules.model.WanModel                                   (reference omnivideo/modules/model.py)
  modules.unified_model.UnifiedWanWithMixedConditionModel  (reference omnivideo/modules/unified_model.py)
  modules.visual_context_adapter.VisualContextAdapter      (reference omnivideo/modules/visual_context_adapter.py)
  modules.attention.flash_attention                        (reference omnivideo/modules/attention.py)
  modules.vae2_1.Wan2_1_VAE                                (reference omnivideo/modules/vae2_1.py)
  utils.fm_solvers_unipc.FlowUniPCMultistepScheduler       (reference omnivideo/utils/fm_solvers_unipc.py)
  utils.checkpoint                                         (the reference's checkpoint files, read into these trees)
  sampling.DenoiseLoop                                     (the loop of reference omnivideo/x2x_gen_unified.py:474-492)
  distributed.{util,ulysses,full_sequence_parallel}        (reference omnivideo/distributed/*.py); distributed.peer is
                                                           the NVLink peer-memory transport underneath
  patch.patch_reference_{model,unified,vae,module_globals} rebind a live reference pipeline onto these kernels

All arithmetic runs in `libovb200.so` (include/ovb200.h).  There is no CPU or PyTorch fallback.
"""
from . import _lib, ops  # noqa: F401

__version__ = "0.1.0"
