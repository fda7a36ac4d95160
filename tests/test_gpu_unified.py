"""Conditioning front-end on the GPU: UnifiedWanWithMixedConditionModel
builds the same cross-attention context as the reference (goldens captured from the reference's own
forward), caches it across the sampling loop, and hands it to WanModel exactly like the reference."""
import random

import pytest
import torch
import torch.nn as nn

from oracle import dit as O
from oracle import synth
from oracle import unified as U
from tests.golden_util import golden_names, load_golden
from tests.test_oracle_unified_golden import unified_case

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def build(meta, cfg, sd):
    from omni_video_b200.modules.unified_model import UnifiedWanWithMixedConditionModel
    with torch.device("cuda"):
        m = UnifiedWanWithMixedConditionModel(
            wan_config=dict(cfg), vlm_in_dim=meta["vlm_in_dim"], use_visual_context_adapter=meta["adapter"],
            visual_context_adapter_patch_size=tuple(meta["patch"]) if meta["patch"] else None,
            max_context_len=meta["max_len"])
    missing, unexpected = m.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=False)
    assert not unexpected and all(k.startswith("wan_model.") for k in missing)
    return m.to(bf16).eval().requires_grad_(False)


def to_cuda(ar, ctx, vis, st):
    f = lambda t: None if t is None else t.cuda().to(bf16)  # noqa: E731
    return (None if ar is None else [f(a) for a in ar], [f(c) for c in ctx],
            None if vis is None else [v.cuda() for v in vis],
            None if st is None else {k: f(v) for k, v in st.items()})


@pytest.mark.parametrize("name", golden_names("unified_context"))
def test_mixed_context_matches_reference_golden(name):
    meta, gold = load_golden(name)
    cfg, sd, ar, ctx, vis, st = unified_case(meta)
    m = build(meta, cfg, sd)
    ar, ctx, vis, st = to_cuda(ar, ctx, vis, st)
    outs = m.mixed_context(len(meta["n_text"]), ctx, ar, vis, st, device=torch.device("cuda"))
    assert len(outs) == len(gold)
    for o, g in zip(outs, gold):
        assert o.shape == g.shape and o.dtype == bf16
        assert O.rel_l2(o.float().cpu(), g) <= 1e-2          # north-star bf16 tolerance
        assert (o.float().cpu() - g).abs().max().item() <= 0.06 * max(1.0, g.abs().max().item())


def test_context_is_cached_across_steps_and_invalidated_by_writes():
    meta, _ = load_golden("unified_full")
    cfg, sd, ar, ctx, vis, st = unified_case(meta)
    m = build(meta, cfg, sd)
    ar, ctx, vis, st = to_cuda(ar, ctx, vis, st)
    a = m.mixed_context(1, ctx, ar, vis, st, device=torch.device("cuda"))
    b = m.mixed_context(1, ctx, ar, vis, st, device=torch.device("cuda"))
    assert (m.context_cache_misses, m.context_cache_hits) == (1, 1) and a[0] is b[0]
    ctx[0].mul_(2.0)                                          # in-place write bumps the version
    c = m.mixed_context(1, ctx, ar, vis, st, device=torch.device("cuda"))
    assert m.context_cache_misses == 2 and not torch.equal(a[0], c[0])
    with torch.no_grad():
        m.vlm_proj.bias.add_(1.0)                             # so does touching a parameter (load_state_dict, .to())
    m.mixed_context(1, ctx, ar, vis, st, device=torch.device("cuda"))
    assert m.context_cache_misses == 3


def test_forward_equals_wan_model_on_mixed_context_and_consumes_rng_like_reference():
    meta, _ = load_golden("unified_full")
    cfg, sd, ar, ctx, vis, st = unified_case(meta)
    m = build(meta, cfg, sd)
    m.wan_model.load_state_dict({k: v.cuda() for k, v in synth.dit_state_dict(dict(cfg), seed=3).items()})
    m = m.to(bf16)
    ar, ctx, vis, st = to_cuda(ar, ctx, vis, st)
    xs, t, _ = synth.dit_inputs(cfg, (2, 8, 8), 4, seed=5)
    xs = [x.cuda() for x in xs]
    random.seed(123)
    out = m(xs, t=t.cuda(), context=ctx, ar_vision_input=ar, visual_emb=vis, seq_len=32, special_token_dict=st)
    after = random.random()
    random.seed(123)
    random.random(), random.random()                          # 2 draws per sample per call (unified_model.py:160-162)
    assert after == random.random()
    mixed = m.mixed_context(1, ctx, ar, vis, st, device=torch.device("cuda"))
    ref = m.wan_model(xs, t=t.cuda(), context=mixed, seq_len=32)
    assert torch.equal(out[0], ref[0]) and out[0].shape == (16, 2, 8, 8)
    # and against the fp32 oracle of the whole chain: front-end restatement -> DiT restatement
    sd_all = {k[len("wan_model."):]: v.float().cpu() for k, v in m.state_dict().items() if k.startswith("wan_model.")}
    sd_u = {k: v.float().cpu() for k, v in m.state_dict().items() if not k.startswith("wan_model.")}
    f = lambda z: None if z is None else [None if u is None else u.float().cpu() for u in z]  # noqa: E731
    mixed_o = U.mixed_context(sd_u, 1, f(ctx), f(ar), f(vis), {k: v.float().cpu() for k, v in st.items()},
                              m.max_context_len, cfg["text_dim"])
    want = O.dit_forward(sd_all, dict(cfg, text_len=m.wan_model.text_len), [x.cpu() for x in xs], t, mixed_o, 32,
                         emulate_bf16=True)
    assert O.rel_l2(out[0].cpu(), want[0]) <= 1e-2


def test_condition_dropout_replaces_context_and_vlm_tokens():
    meta, _ = load_golden("unified_batch2_noadapter")
    cfg, sd, ar, ctx, vis, st = unified_case(meta)
    m = build(meta, cfg, sd)
    ar, ctx, vis, st = to_cuda(ar, ctx, vis, st)
    captured = {}
    m.wan_model.forward = lambda x, t=None, context=None, seq_len=None: captured.setdefault("c", context)
    unc = {"uncond_context": torch.zeros(3, cfg["text_dim"], device="cuda", dtype=bf16)}
    xs = [torch.zeros(16, 1, 2, 2, device="cuda")] * 2
    m(xs, t=torch.zeros(2, device="cuda"), context=ctx, ar_vision_input=ar, seq_len=1, special_token_dict=st,
      classifier_free_ratio=1.0, unconditioned_context=unc)
    c0 = captured["c"][0]
    # sample 0: 2 null VLM tokens + <prp_st> + 3 uncond rows + <prp_ed>
    assert c0.shape[0] == 2 + 1 + 3 + 1 and torch.equal(c0[3:6], unc["uncond_context"])
    with pytest.raises(ValueError):
        m(xs, t=torch.zeros(2, device="cuda"), context=ctx, seq_len=1, classifier_free_ratio=1.0)


class _RefLikeAdapter(nn.Module):
    """Parameter tree of the reference's VisualContextAdapter with a forward that must NOT run."""

    def __init__(self, patch, cin, hidden, out):
        super().__init__()
        self.patch_size, self.hidden_dim, self.out_dim, self.eps = patch, hidden, out, 1e-6
        self.patch_embedding = nn.Conv3d(cin, hidden, kernel_size=patch, stride=patch)
        self.projection = nn.Linear(hidden, out)

    def forward(self, x):
        raise AssertionError("reference arithmetic must not run after patching")


def test_patch_reference_unified_rebinds_a_reference_shaped_instance():
    from omni_video_b200 import patch
    from omni_video_b200.modules.model import WanModel, WanRMSNorm
    meta, gold = load_golden("unified_full")
    cfg, sd, ar, ctx, vis, st = unified_case(meta)

    class RefLike(nn.Module):
        def __init__(self):
            super().__init__()
            self.wan_model = WanModel(**cfg)
            self.max_context_len = self.wan_model.text_len
            self.vlm_norm = WanRMSNorm(meta["vlm_in_dim"], eps=1e-6)
            self.vlm_proj = nn.Linear(meta["vlm_in_dim"], cfg["text_dim"])
            self.visual_context_adapter = _RefLikeAdapter(tuple(meta["patch"]), 16, cfg["dim"], cfg["text_dim"])

        def forward(self, *a, **k):
            raise AssertionError("reference forward must not run after patching")

    with torch.device("cuda"):
        r = RefLike()
    r.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=False)
    r = r.to(bf16).eval().requires_grad_(False)
    patch.patch_reference_unified(r)
    ar, ctx, vis, st = to_cuda(ar, ctx, vis, st)
    out = r.mixed_context(1, ctx, ar, vis, st, device=torch.device("cuda"))
    assert O.rel_l2(out[0].float().cpu(), gold[0]) <= 1e-2
    xs = [torch.randn(16, 2, 8, 8, device="cuda")]
    y = r(xs, t=torch.tensor([500], device="cuda"), context=ctx, ar_vision_input=ar, visual_emb=vis, seq_len=32,
          special_token_dict=st)
    assert y[0].shape == (16, 2, 8, 8) and bool(torch.isfinite(y[0]).all())
