"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/ovb200.h declares; the ctypes binding table covers exactly the same set.  No compute call
is made here."""
import ctypes
import os
import re

from omni_video_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ovb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ovb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_a_nontrivial_abi():
    syms = _declared_symbols()
    for must in ("ovb_gemm_bf16", "ovb_attn_fwd", "ovb_layernorm", "ovb_rmsnorm_rope", "ovb_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/ovb200.h but not exported: {missing}"


def test_binding_table_matches_header():
    _lib.load()
    try:  # optional parts of the binding register their symbols on import
        import omni_video_b200.vae_ops  # noqa: F401
    except ImportError:
        pass
    assert sorted(_lib._SIGNATURES) == _declared_symbols()


def test_no_fallback_without_cuda():
    import pytest
    import torch
    from omni_video_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(_lib.OvbError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16), None)
    lib = _lib.load()
    assert lib.ovb_check_device() != 0  # no sm_100 device here -> loud failure, not a silent path
    assert lib.ovb_last_error()


def _plan(B, H, Lq, Lk, ws_bytes, k_lens=None, no_split=0):
    import ctypes as C
    from omni_video_b200 import _lib
    lib = _lib.load()
    a = _lib.AttnArgs()
    a.B, a.H, a.Lq, a.Lk = B, H, Lq, Lk
    if k_lens is not None:
        arr = (C.c_int32 * B)(*k_lens)
        a.k_lens = C.cast(arr, C.POINTER(C.c_int32))
    a.no_split = no_split
    if ws_bytes:
        a.workspace, a.workspace_bytes = 0x1000, ws_bytes      # never dereferenced by the planner
    need = int(lib.ovb_attn_workspace_size(C.byref(a)))
    nw, nt, parts = C.c_int32(), C.c_int32(), C.c_int32()
    assert lib.ovb_attn_plan(C.byref(a), C.byref(nw), C.byref(nt), C.byref(parts)) == 0
    return nw.value, nt.value, parts.value, need


def test_attention_work_decomposition_host_logic():
    """ovb_attn_plan is host-only (no launch): the last partial wave of (q block, head, sample) items is
    key-split only when a workspace is there, the slices stay long, and the saving is worth it."""
    big = 1 << 30
    # per-rank shape of cfg2 x ulysses4 on 8 GPUs: 128 q blocks x 3 heads = 384 items = 2 waves of 148 + 88
    nw, nt, parts, need = _plan(1, 3, 32760, 32760, big)
    assert (nw, nt) == (296, 88) and parts >= 3 and nw + nt == 384
    assert need >= nt * parts * 256 * (128 * 4 + 8)
    # rounds with the split are fewer than the 3 of the unsplit launch
    assert 2 + -(-nt * parts // 148) / parts < 3
    # no workspace / too small / disabled -> one CTA per item
    assert _plan(1, 3, 32760, 32760, 0)[1:3] == (0, 1)
    assert _plan(1, 3, 32760, 32760, 1024)[1:3] == (0, 1)
    assert _plan(1, 3, 32760, 32760, big, no_split=1)[1:3] == (0, 1)
    # single GPU, CFG-batched: 3072 items = 20.76 waves, the tail is not worth a split
    assert _plan(2, 12, 32760, 32760, big)[1:3] == (0, 1)
    # a full last wave needs nothing, and reports a zero workspace
    assert _plan(1, 1, 148 * 256, 8192, big) == (148, 0, 1, 0)
    # cross-attention: 512 keys are too few to cut
    assert _plan(1, 3, 32760, 512, big)[1:3] == (0, 1)
    # the shortest sample of a ragged batch decides how far the keys can be cut
    long_ = _plan(2, 1, 1280, 8192, big, k_lens=[8192, 8192])
    short = _plan(2, 1, 1280, 8192, big, k_lens=[8192, 1100])
    assert long_[2] > short[2] >= 1


def test_argument_errors_are_reported_before_any_launch():
    """Bad shapes / alignments / counts come back as negative codes with a message (never an exception,
    never a launch): the checks run before the first CUDA call, so this holds on a box without a GPU."""
    import ctypes as C
    from omni_video_b200 import _lib
    lib = _lib.load()

    def err():
        return lib.ovb_last_error().decode()

    g = _lib.GemmArgs()
    g.A = g.W = g.out = 0x1000
    g.M, g.N, g.K, g.lda, g.ldw, g.ldo = 128, 100, 64, 64, 64, 100      # N not a multiple of 8
    assert lib.ovb_gemm_bf16(C.byref(g), None) < 0 and "multiples of 8" in err()

    a = _lib.AttnArgs()
    a.q = a.k = a.v = a.out = 0x1000
    a.B, a.H, a.Lq, a.Lk = 9, 1, 128, 128                              # B > 8
    assert lib.ovb_attn_fwd(C.byref(a), None) < 0 and "B <= 8" in err()
    a.B = 1
    a.q_row_stride = a.k_row_stride = a.v_row_stride = a.o_row_stride = 130   # not a multiple of 8
    assert lib.ovb_attn_fwd(C.byref(a), None) < 0 and "strides" in err()

    t = _lib.AttnRelBiasArgs()
    t.q = t.k = t.v = t.out = t.rel_bias = 0x1000
    t.B, t.H, t.L, t.head_dim = 1, 2, 64, 32                            # no 32-wide heads
    assert lib.ovb_attn_relbias_fwd(C.byref(t), None) < 0 and "head_dim" in err()
    t.head_dim = 64
    t.q_row_stride = t.k_row_stride = t.v_row_stride = t.o_row_stride = 120      # rows narrower than H * head_dim
    assert lib.ovb_attn_relbias_fwd(C.byref(t), None) < 0 and "row pitch" in err()
    t.q_row_stride = t.k_row_stride = t.v_row_stride = t.o_row_stride = 128
    t.bias_head_stride = 100                                            # table shorter than 2L - 1
    assert lib.ovb_attn_relbias_fwd(C.byref(t), None) < 0 and "2L-1" in err()

    r = _lib.RmsRopeArgs()
    r.x = r.weight = 0x1000
    r.rows, r.d, r.ldx = 4, 100, 104                                   # d not a multiple of 8
    assert lib.ovb_rmsnorm_rope(C.byref(r), None) < 0 and "multiple of 8" in err()
    r.d, r.ldx, r.copy_only = 128, 128, 1                              # copy_only without destinations
    assert lib.ovb_rmsnorm_rope(C.byref(r), None) < 0 and "copy_only" in err()

    ln = _lib.LnArgs()
    ln.x = ln.out = ln.a = ln.b = 0x1000
    ln.rows, ln.d, ln.ldx, ln.ldo = 4, 8200, 8200, 8200                # d > 8192
    assert lib.ovb_layernorm(C.byref(ln), None) < 0

    assert lib.ovb_patchify3d(0x1000, 0, 0x1000, 256, 16, 1, 2, 2, 1, 4, 4, None) < 0 and "smaller than patch" in err()
    pads = (C.c_void_p * 9)(*([0x1000] * 9))
    assert lib.ovb_peer_signal_wait(pads, 0, 9, 1, None, None, None) < 0 and "rank/world" in err()
    assert lib.ovb_attn_plan(None, None, None, None) < 0
