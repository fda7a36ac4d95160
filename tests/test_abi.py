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
