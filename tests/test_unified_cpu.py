"""Host logic of the conditioning front-end (omni_video_b200/modules/unified_model.py) on CPU: order of
concatenation, special tokens, truncation, absent conditions, condition dropout, RNG consumption and the
per-sample cache — with the two CUDA pieces (`_vlm`, the adapter's forward) replaced by the oracle's fp32
arithmetic.  Compared with what the reference's own forward produced (tests/golden/unified_*.npz)."""
import random
import types

import pytest
import torch

from oracle import dit as O
from oracle import unified as U
from tests.golden_util import golden_names, load_golden
from tests.test_oracle_unified_golden import unified_case


def build_cpu(meta, cfg, sd):
    from omni_video_b200.modules.unified_model import UnifiedWanWithMixedConditionModel
    m = UnifiedWanWithMixedConditionModel(
        wan_config=dict(cfg), vlm_in_dim=meta["vlm_in_dim"], use_visual_context_adapter=meta["adapter"],
        visual_context_adapter_patch_size=tuple(meta["patch"]) if meta["patch"] else None,
        max_context_len=meta["max_len"], skip_init=True)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("wan_model.") for k in missing)
    sd32 = {k: v.float() for k, v in sd.items()}
    m._vlm = types.MethodType(lambda self, tokens: U.vlm_features(sd32, tokens.float(), 1e-6), m)
    if m.visual_context_adapter is not None:
        m.visual_context_adapter.forward = lambda x: U.visual_context(sd32, x.float(), 1e-6).unsqueeze(0)
    return m


@pytest.mark.parametrize("name", golden_names("unified_context"))
def test_mixed_context_host_logic_matches_reference_golden(name):
    meta, gold = load_golden(name)
    cfg, sd, ar, ctx, vis, st = unified_case(meta)
    m = build_cpu(meta, cfg, sd)
    outs = m.mixed_context(len(meta["n_text"]), ctx, ar, vis, st, device=torch.device("cpu"))
    assert len(outs) == len(gold)
    for o, g in zip(outs, gold):
        assert o.shape == g.shape and o.dtype == torch.bfloat16
        assert O.rel_l2(o.float(), g) <= 5e-3          # only the final bf16 cast separates the two
    again = m.mixed_context(len(meta["n_text"]), ctx, ar, vis, st, device=torch.device("cpu"))
    assert again[0] is outs[0] and (m.context_cache_hits, m.context_cache_misses) == (1, 1)


def test_context_cache_is_not_fooled_by_recycled_addresses():
    """Two successive prompts of equal length: the second must get ITS context even when its tensors
    land on the addresses the first prompt's (freed) tensors had.  The cache entry keeps its inputs alive
    (so a live entry's key cannot collide) and confirms every hit by object identity."""
    meta, _ = load_golden("unified_batch2_noadapter")
    cfg, sd, ar, ctx, vis, st = unified_case(meta)
    m = build_cpu(meta, cfg, sd)
    n = len(meta["n_text"])
    dev = torch.device("cpu")
    ctx_a = [c.clone() for c in ctx]
    out_a = [o.clone() for o in m.mixed_context(n, ctx_a, ar, vis, st, device=dev)]
    held = next(iter(m._ctx_cache.values()))[1]
    assert any(t is ctx_a[0] for t in held)             # the entry pins its inputs
    # forge the collision the allocator would produce: an entry under prompt B's key that was computed
    # from other (dead) tensors
    ctx_b = [c.clone() + 1.0 for c in ctx]
    from omni_video_b200.modules import unified_model as UM
    key_b = (n, UM._ident(ctx_b), UM._ident(ar), UM._ident(vis), UM._ident(st), m._params_key(), m.max_context_len)
    m._ctx_cache[key_b] = (out_a, [torch.empty(0)] * len(held))
    out_b = m.mixed_context(n, ctx_b, ar, vis, st, device=dev)
    assert out_b is not out_a
    assert not all(torch.equal(a, b) for a, b in zip(out_a, out_b))
    ref_b = [o.clone() for o in out_b]
    m.clear_context_cache()
    again = m.mixed_context(n, ctx_b, ar, vis, st, device=dev)
    assert all(torch.equal(a, b) for a, b in zip(ref_b, again))


def test_forward_dropout_rng_and_errors_on_cpu():
    meta, _ = load_golden("unified_batch2_noadapter")
    cfg, sd, ar, ctx, vis, st = unified_case(meta)
    m = build_cpu(meta, cfg, sd)
    seen = {}
    m.wan_model.forward = lambda x, t=None, context=None, seq_len=None: seen.setdefault("ctx", context)
    xs = [torch.zeros(16, 1, 2, 2)] * 2
    unc = {"uncond_context": torch.zeros(3, cfg["text_dim"])}
    random.seed(5)
    out = m(xs, t=torch.zeros(2), context=list(ctx), ar_vision_input=list(ar), seq_len=1, special_token_dict=st,
            classifier_free_ratio=1.0, unconditioned_context=unc)
    after = random.random()
    random.seed(5)
    [random.random() for _ in range(4)]                     # 2 draws per sample per call
    assert after == random.random()
    assert out is seen["ctx"] and len(out) == 2
    # sample 0 had VLM tokens: 2 null tokens | <prp_st> 3 uncond rows <prp_ed>; sample 1 had none
    assert out[0].shape[0] == 2 + 1 + 3 + 1 and out[1].shape[0] == 1 + 3 + 1
    assert torch.equal(out[0][3:6].float(), unc["uncond_context"])
    with pytest.raises(ValueError):
        m(xs, t=torch.zeros(2), context=list(ctx), seq_len=1, classifier_free_ratio=0.5 + 0.5)
    with pytest.raises(ValueError):
        m(xs, t=torch.zeros(2), context=list(ctx), seq_len=1, condition_mode="text_only")
    # no conditions at all: the reference's safety fallback, one zero token
    seen.clear()
    out = m([xs[0]], t=torch.zeros(1), seq_len=1)
    assert out[0].shape == (1, cfg["text_dim"]) and float(out[0].abs().sum()) == 0.0


def test_patch_reference_unified_binds_every_method_the_forward_needs():
    """A reference-shaped instance (same attribute names, forwards that must not run) after
    patch.patch_reference_unified: the rebound front-end produces the reference's context."""
    import torch.nn as nn
    from omni_video_b200 import patch
    from omni_video_b200.modules.model import WanModel, WanRMSNorm
    meta, gold = load_golden("unified_full")
    cfg, sd, ar, ctx, vis, st = unified_case(meta)

    class Adapter(nn.Module):
        def __init__(self):
            super().__init__()
            self.patch_size, self.hidden_dim, self.out_dim, self.eps = tuple(meta["patch"]), cfg["dim"], cfg["text_dim"], 1e-6
            self.patch_embedding = nn.Conv3d(16, cfg["dim"], kernel_size=self.patch_size, stride=self.patch_size)
            self.projection = nn.Linear(cfg["dim"], cfg["text_dim"])

        def forward(self, x):
            raise AssertionError("reference arithmetic must not run after patching")

    class RefLike(nn.Module):
        def __init__(self):
            super().__init__()
            self.wan_model = WanModel(**{**cfg, "patch_size": tuple(cfg["patch_size"]), "window_size": (-1, -1)},
                                      skip_init=True)
            self.max_context_len = self.wan_model.text_len
            self.vlm_norm = WanRMSNorm(meta["vlm_in_dim"], eps=1e-6)
            self.vlm_proj = nn.Linear(meta["vlm_in_dim"], cfg["text_dim"])
            self.visual_context_adapter = Adapter()

        def forward(self, *a, **k):
            raise AssertionError("reference forward must not run after patching")

    r = RefLike()
    r.load_state_dict(sd, strict=False)
    patch.patch_reference_unified(r)
    sd32 = {k: v.float() for k, v in sd.items()}
    r._vlm = types.MethodType(lambda self, tokens: U.vlm_features(sd32, tokens.float(), 1e-6), r)
    r.visual_context_adapter.forward = lambda x: U.visual_context(sd32, x.float(), 1e-6).unsqueeze(0)
    seen = {}
    r.wan_model.forward = lambda x, t=None, context=None, seq_len=None: seen.setdefault("ctx", context)
    out = r([torch.zeros(16, 1, 2, 2)], t=torch.zeros(1), context=ctx, ar_vision_input=ar, visual_emb=vis, seq_len=1,
            special_token_dict=st)
    assert out is seen["ctx"] and O.rel_l2(out[0].float(), gold[0]) <= 5e-3
    with pytest.raises(TypeError):
        patch.patch_reference_unified(nn.Linear(2, 2))


def test_patch_seam_on_the_reference_own_unified_class():
    """Build container only: the REFERENCE's `UnifiedWanWithMixedConditionModel` instance after
    `patch.patch_reference_unified`, with the two CUDA pieces replaced by the oracle's arithmetic, builds the
    context its own forward built (the golden) — names and conventions line up with the real class."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    from omni_video_b200 import patch
    ref_loader.load_reference(cpu_attention=True)
    ref_loader._load("omnivideo.modules.visual_context_adapter", "omnivideo/modules/visual_context_adapter.py")
    ref_unified = ref_loader._load("omnivideo.modules.unified_model", "omnivideo/modules/unified_model.py")
    meta, gold = load_golden("unified_full")
    cfg, sd, ar, ctx, vis, st = unified_case(meta)
    r = ref_unified.UnifiedWanWithMixedConditionModel(
        wan_config=dict(cfg), vlm_in_dim=meta["vlm_in_dim"], use_visual_context_adapter=True,
        visual_context_adapter_patch_size=tuple(meta["patch"]), max_context_len=meta["max_len"], skip_init=True)
    r.load_state_dict(sd, strict=False)
    assert type(r).__module__ == "omnivideo.modules.unified_model"
    patch.patch_reference_unified(r)
    sd32 = {k: v.float() for k, v in sd.items()}
    r._vlm = types.MethodType(lambda self, tokens: U.vlm_features(sd32, tokens.float(), 1e-6), r)
    r.visual_context_adapter.forward = lambda x: U.visual_context(sd32, x.float(), 1e-6).unsqueeze(0)
    seen = {}
    r.wan_model.forward = lambda x, t=None, context=None, seq_len=None: seen.setdefault("ctx", context)
    out = r([torch.zeros(16, 1, 2, 2)], t=torch.zeros(1), context=ctx, ar_vision_input=ar, visual_emb=vis, seq_len=1,
            special_token_dict=st, classifier_free_ratio=0.0, unconditioned_context=None, condition_mode="full")
    assert out is seen["ctx"] and O.rel_l2(out[0].float(), gold[0]) <= 5e-3
