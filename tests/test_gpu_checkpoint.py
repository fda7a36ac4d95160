"""Checkpoint formats of the reference all the way to the kernels: the weights the reference's
own golden forward was computed with are written in the formats the reference reads — the diffusers directory
behind `WanModel.from_pretrained` (sharded safetensors, x2x_gen_unified.py:120-126) and a trained-expert
`model.pt` with a wrapped, DDP-prefixed state dict (generate_omni_v2v.py:499-520) —, loaded through
`omni_video_b200.utils.checkpoint`, moved to the GPU in bf16 like the pipeline does, and the forward is compared
with the reference's output (tests/golden/dit_tiny_a.npz).  (Real checkpoints are not reachable from this sandbox;
tests/test_checkpoint_cpu.py covers the format corner cases.)"""
import pytest
import torch

from oracle import dit as O
from tests.golden_util import dit_case, load_golden

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]
bf16 = torch.bfloat16


def _forward(m, xs, t, ctx, seq_len):
    m = m.to(bf16).cuda().eval().requires_grad_(False)
    return m([x.cuda() for x in xs], t=t.cuda(), context=[c.cuda().to(bf16) for c in ctx], seq_len=seq_len)[0].cpu()


def test_diffusers_directory_to_gpu_forward_vs_reference_golden(tmp_path):
    from omni_video_b200.modules.model import WanModel
    from omni_video_b200.utils import checkpoint as ck
    meta, gold = load_golden("dit_tiny_a")
    cfg, sd, xs, t, ctx, seq_len = dit_case(meta)
    src = WanModel(**cfg, skip_init=True)
    src.load_state_dict(sd, strict=True)
    d = tmp_path / "ckpt" / "low_noise_model"
    ck.save_wan_pretrained(src, str(d), max_shard_bytes=300_000)          # several shards + index
    # (text_dim is not part of config.json — reference model.py:298-300 `ignore_for_config` — and the toy's differs
    # from the default 4096)
    m = WanModel.from_pretrained(str(tmp_path / "ckpt"), subfolder="low_noise_model", torch_dtype=bf16,
                                 text_dim=cfg["text_dim"])
    assert next(m.parameters()).dtype == bf16
    out = _forward(m, xs, t, ctx, seq_len)
    assert out.shape == gold[0].shape and O.rel_l2(out, gold[0]) <= 1e-2


def test_trained_expert_file_to_gpu_forward_vs_reference_golden(tmp_path):
    from omni_video_b200.modules.model import WanModel
    from omni_video_b200.utils import checkpoint as ck
    meta, gold = load_golden("dit_tiny_a")
    cfg, sd, xs, t, ctx, seq_len = dit_case(meta)
    path = tmp_path / "model.pt"
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}, "step": 411000}, path)
    m = WanModel(**cfg, skip_init=True)
    missing, unexpected = ck.load_trained_checkpoint(m, str(path), dtype=bf16)
    assert not missing and not unexpected
    out = _forward(m, xs, t, ctx, seq_len)
    assert O.rel_l2(out, gold[0]) <= 1e-2
