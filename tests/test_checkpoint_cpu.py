"""Checkpoint formats of the reference read into our module trees: trained-expert
`model.pt` files (wrapped state dicts, cast before load, strict=False), the diffusers directory layout
behind `WanModel.from_pretrained`, `special_tokens.pkl`, `Wan2.1_VAE.pth`.  CPU only: nothing here launches
a kernel — parameters are plain torch tensors whatever the engine does with them later."""
import json
import os
import pickle

import pytest
import torch

from oracle import synth

TINY = dict(synth.TINY_DIT, num_layers=1)


def _wan():
    from omni_video_b200.modules.model import WanModel
    m = WanModel(**TINY)
    m.load_state_dict(synth.dit_state_dict(TINY, seed=3))
    return m


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("shard_bytes,safe", [(5 << 30, True), (200_000, True), (5 << 30, False)])
def test_wan_from_pretrained_roundtrip(tmp_path, shard_bytes, safe):
    from omni_video_b200.modules.model import WanModel
    from omni_video_b200.utils import checkpoint as ck
    m = _wan()
    d = tmp_path / "ckpt" / "high_noise_model"
    m.save_pretrained(str(d), max_shard_size=shard_bytes, safe_serialization=safe)
    files = sorted(os.listdir(d))
    assert ck.CONFIG_NAME in files
    if safe and shard_bytes < (1 << 30):
        assert ck.SAFETENSORS_INDEX in files and sum(f.endswith(".safetensors") for f in files) > 1
    cfg = json.load(open(d / ck.CONFIG_NAME))
    # diffusers does not store the arguments listed in ignore_for_config (reference model.py:299-301)
    assert not (set(WanModel.ignore_for_config) & set(cfg)) and cfg["dim"] == TINY["dim"]
    back = WanModel.from_pretrained(str(tmp_path / "ckpt"), subfolder="high_noise_model", text_dim=TINY["text_dim"])
    _same(m.state_dict(), back.state_dict())
    assert back.num_layers == 1 and back.dim == TINY["dim"] and not back.training
    half = WanModel.from_pretrained(str(d), torch_dtype=torch.bfloat16, text_dim=TINY["text_dim"])
    assert all(v.dtype == torch.bfloat16 for v in half.state_dict().values())


def test_from_pretrained_rejects_incomplete_or_foreign_checkpoints(tmp_path):
    from safetensors.torch import save_file
    from omni_video_b200.modules.model import WanModel
    from omni_video_b200.utils import checkpoint as ck
    m = _wan()
    d = tmp_path / "m"
    m.save_pretrained(str(d))
    sd = {k: v.contiguous() for k, v in m.state_dict().items()}
    del sd["head.head.bias"]
    save_file(sd, str(d / ck.SAFETENSORS_NAME))
    with pytest.raises(KeyError):
        WanModel.from_pretrained(str(d), text_dim=TINY["text_dim"])
    sd["head.head.bias"] = m.state_dict()["head.head.bias"].contiguous()
    sd["not.a.parameter"] = torch.zeros(1)
    save_file(sd, str(d / ck.SAFETENSORS_NAME))
    with pytest.raises(KeyError):
        WanModel.from_pretrained(str(d), text_dim=TINY["text_dim"])
    with pytest.raises(FileNotFoundError):
        WanModel.from_pretrained(str(tmp_path), subfolder="missing")


@pytest.mark.parametrize("wrapper", [None, "module", "model"])
def test_trained_expert_checkpoint_into_unified_model(tmp_path, wrapper):
    """generate_omni_v2v.py:499-520: unwrap, cast to the run's precision, load with strict=False."""
    from omni_video_b200.modules.unified_model import UnifiedWanWithMixedConditionModel
    from omni_video_b200.utils.checkpoint import load_trained_checkpoint
    sd = {"wan_model." + k: v for k, v in synth.dit_state_dict(TINY, seed=5).items()}
    sd.update(synth.unified_state_dict(TINY, 128, (1, 4, 4), seed=5))
    sd["optimizer_leftover"] = torch.zeros(3)
    path = tmp_path / "model.pt"
    torch.save({wrapper: sd, "global_step": 7} if wrapper else sd, str(path))
    model = UnifiedWanWithMixedConditionModel(wan_config=dict(TINY), vlm_in_dim=128, use_visual_context_adapter=True,
                                              visual_context_adapter_patch_size=(1, 4, 4), skip_init=True)
    missing, unexpected = load_trained_checkpoint(model, str(path), dtype=torch.bfloat16)
    assert missing == [] and unexpected == ["optimizer_leftover"]
    got = model.state_dict()
    for k, v in sd.items():
        if k != "optimizer_leftover":
            assert torch.equal(got[k].float(), v.to(torch.bfloat16).float()), k
    # a checkpoint trained without the adapter leaves the adapter's parameters as they were, and says so
    no_adapter = {k: v for k, v in sd.items() if not k.startswith("visual_context_adapter.")}
    torch.save(no_adapter, str(path))
    missing, _ = load_trained_checkpoint(model, str(path))
    assert missing and all(k.startswith("visual_context_adapter.") for k in missing)
    with pytest.raises(RuntimeError):
        load_trained_checkpoint(model, str(path), strict=True)


def test_ddp_prefix_is_stripped(tmp_path):
    from omni_video_b200.utils.checkpoint import read_state_dict
    sd = synth.dit_state_dict(TINY, seed=1)
    torch.save({"module." + k: v for k, v in sd.items()}, str(tmp_path / "ddp.pt"))
    _same(read_state_dict(str(tmp_path / "ddp.pt")), sd)


def test_unified_model_from_checkpoint_dir_and_save(tmp_path):
    from omni_video_b200.modules.unified_model import UnifiedWanWithMixedConditionModel
    m = _wan()
    m.save_pretrained(str(tmp_path / "ckpt" / "low_noise_model"))
    # text_dim is not in config.json (ignore_for_config): the diffusers default 4096 applies, as in the reference
    cfg = json.load(open(tmp_path / "ckpt" / "low_noise_model" / "config.json"))
    assert "text_dim" not in cfg
    u = UnifiedWanWithMixedConditionModel(wan_model_or_ckpt_dir=m, vlm_in_dim=64, max_context_len=48)
    assert u.wan_model is m and u.max_context_len == 48 and m.text_len == 48
    u.save_pretrained(str(tmp_path / "out"))
    assert sorted(os.listdir(tmp_path / "out")) == ["vlm_norm_pytorch_model.bin", "vlm_proj_pytorch_model.bin",
                                                     "wan_model"]
    with pytest.raises(ValueError):
        UnifiedWanWithMixedConditionModel()


def test_special_tokens_pkl(tmp_path):
    from omni_video_b200.utils.checkpoint import load_special_tokens
    toks = {k: torch.randn(1, 64) for k in ("<img_st>", "<img_ed>", "<prp_st>", "<prp_ed>")}
    with open(tmp_path / "special_tokens.pkl", "wb") as fh:
        pickle.dump(toks, fh)
    got = load_special_tokens(str(tmp_path / "special_tokens.pkl"), torch.bfloat16, "cpu")
    assert set(got) == set(toks) and all(v.dtype == torch.bfloat16 and v.shape == (1, 64) for v in got.values())
    assert load_special_tokens(str(tmp_path / "absent.pkl"), torch.bfloat16, "cpu") is None


def test_vae_checkpoint_file(tmp_path):
    """Wan2.1_VAE.pth -> Wan2_1_VAE(vae_pth=...) (reference vae2_1.py:591-615: meta-device build + assign)."""
    from omni_video_b200.modules.vae2_1 import Wan2_1_VAE
    sd = synth.vae_state_dict(seed=9)
    torch.save(sd, str(tmp_path / "Wan2.1_VAE.pth"))
    vae = Wan2_1_VAE(vae_pth=str(tmp_path / "Wan2.1_VAE.pth"), device="cpu")
    _same(vae.model.state_dict(), sd)
    assert vae.model.z_dim == 16 and not any(p.requires_grad for p in vae.model.parameters())
    assert vae.mean.shape == (16,) and torch.allclose(vae.scale[1], 1.0 / vae.std)
