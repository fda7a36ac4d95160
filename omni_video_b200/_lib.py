"""ctypes binding of ``libovb200.so`` (C ABI declared in ``include/ovb200.h``).

There is no fallback: if the shared library is missing or the device is not sm_100, every op
raises.  The oracle under ``oracle/`` is test infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OVB_LIB_PATH: load another build of the same ABI (kernel experiments: scripts/build_variant_lib.sh)
LIB_PATH = os.environ.get("OVB_LIB_PATH") or os.path.join(_HERE, "libovb200.so")

OVB_OK = 0
ABI_VERSION = 4   # include/ovb200.h OVB_ABI_VERSION

# epilogues / modes (mirror include/ovb200.h)
EPI_BF16, EPI_GELU_BF16, EPI_RESID_F32, EPI_GATE_RESID_F32, EPI_F32, EPI_RESID_BF16 = range(6)
CONV_OUT_BF16, CONV_OUT_F32_PLANAR, CONV_OUT_F32_PLANAR_CLAMP = range(3)
LN_MODULATE, LN_AFFINE = 0, 1


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("W", C.c_void_p), ("ldw", C.c_int64),
        ("bias", C.c_void_p),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("epilogue", C.c_int32),
        ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("gate", C.c_void_p), ("gate_row_stride", C.c_int64), ("gate_sample_stride", C.c_int64),
        ("rows_per_sample", C.c_int32),
        ("n_split", C.c_int32), ("out_split_stride", C.c_int64),
        ("k_split", C.c_int32), ("a_split_stride", C.c_int64),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_row_stride", C.c_int64), ("q_batch_stride", C.c_int64),
        ("k", C.c_void_p), ("k_row_stride", C.c_int64), ("k_batch_stride", C.c_int64),
        ("v", C.c_void_p), ("v_row_stride", C.c_int64), ("v_batch_stride", C.c_int64),
        ("out", C.c_void_p), ("o_row_stride", C.c_int64), ("o_batch_stride", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32),
        ("k_lens", C.POINTER(C.c_int32)),
        ("softmax_scale", C.c_float),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("no_split", C.c_int32),
        ("out_ptrs", C.POINTER(C.c_void_p)), ("n_out_ptrs", C.c_int32), ("rows_per_out", C.c_int32),
    ]


class AttnRelBiasArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_row_stride", C.c_int64), ("q_batch_stride", C.c_int64),
        ("k", C.c_void_p), ("k_row_stride", C.c_int64), ("k_batch_stride", C.c_int64),
        ("v", C.c_void_p), ("v_row_stride", C.c_int64), ("v_batch_stride", C.c_int64),
        ("out", C.c_void_p), ("o_row_stride", C.c_int64), ("o_batch_stride", C.c_int64),
        ("rel_bias", C.c_void_p), ("bias_head_stride", C.c_int64),
        ("k_lens", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("L", C.c_int32), ("head_dim", C.c_int32),
        ("softmax_scale", C.c_float),
    ]


class LnArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64), ("x_is_bf16", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_is_f32", C.c_int32),
        ("rows", C.c_int32), ("d", C.c_int32),
        ("eps", C.c_float),
        ("mode", C.c_int32),
        ("a", C.c_void_p), ("b", C.c_void_p),
        ("mod_row_stride", C.c_int64), ("mod_sample_stride", C.c_int64),
        ("rows_per_sample", C.c_int32),
        ("round_ln_bf16", C.c_int32),
    ]


class RmsRopeArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("rows", C.c_int32), ("d", C.c_int32),
        ("weight", C.c_void_p),
        ("eps", C.c_float),
        ("rope", C.c_int32),
        ("cs_table", C.c_void_p), ("table_len", C.c_int32),
        ("F", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("rows_per_sample", C.c_int32),
        ("token_offset", C.c_int32),
        ("col_split", C.c_int32), ("col_split_stride", C.c_int64),
        ("dst_ptrs", C.POINTER(C.c_void_p)), ("dst_ld", C.c_int64), ("copy_only", C.c_int32),
    ]


class ConvArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("Ti", C.c_int32), ("Hi", C.c_int32), ("Wi", C.c_int32), ("Cin", C.c_int32),
        ("w", C.c_void_p), ("w_rows", C.c_int32),
        ("bias", C.c_void_p),
        ("Cout", C.c_int32), ("n_taps", C.c_int32),
        ("tap_dt", C.POINTER(C.c_int8)), ("tap_dh", C.POINTER(C.c_int8)), ("tap_dw", C.POINTER(C.c_int8)),
        ("st", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("To", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("out", C.c_void_p),
        ("os_t", C.c_int64), ("os_h", C.c_int64), ("os_w", C.c_int64), ("out_base", C.c_int64),
        ("n_group", C.c_int32), ("group_stride", C.c_int64),
        ("resid", C.c_void_p),
        ("out_mode", C.c_int32), ("plane_stride", C.c_int64),
        ("norm_gamma", C.c_void_p), ("norm_out", C.c_void_p), ("norm_silu", C.c_int32),
    ]


class TapSumArgs(C.Structure):
    _fields_ = [
        ("y", C.c_void_p), ("n_taps", C.c_int32),
        ("tap_dt", C.POINTER(C.c_int8)), ("tap_dh", C.POINTER(C.c_int8)), ("tap_dw", C.POINTER(C.c_int8)),
        ("Ty", C.c_int32), ("Hy", C.c_int32), ("Wy", C.c_int32),
        ("bias", C.c_void_p), ("out", C.c_void_p),
        ("C", C.c_int32), ("To", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("clamp", C.c_int32),
    ]


class LinCombArgs(C.Structure):
    _fields_ = [("x", C.c_void_p * 5), ("c", C.c_float * 5), ("is_bf16", C.c_int32 * 5),
                ("n_terms", C.c_int32), ("round_mask", C.c_int32), ("out", C.c_void_p),
                ("out_is_bf16", C.c_int32), ("n", C.c_int64)]


# symbol -> (restype, argtypes); every symbol declared in include/ovb200.h must be listed here
# (tests/test_abi.py cross-checks this table against the header and the built library).
_SIGNATURES = {
    "ovb_abi_version": (C.c_int, []),
    "ovb_last_error": (C.c_char_p, []),
    "ovb_check_device": (C.c_int, []),
    "ovb_launch_count": (C.c_int64, []),
    "ovb_add_launch_count": (None, [C.c_int64]),
    "ovb_video_to_uint8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                     C.c_int64, C.c_float, C.c_float, C.c_void_p]),
    "ovb_gemm_bf16": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "ovb_attn_fwd": (C.c_int, [C.POINTER(AttnArgs), C.c_void_p]),
    "ovb_attn_workspace_size": (C.c_int64, [C.POINTER(AttnArgs)]),
    "ovb_attn_plan": (C.c_int, [C.POINTER(AttnArgs), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                C.POINTER(C.c_int32)]),
    "ovb_attn_relbias_fwd": (C.c_int, [C.POINTER(AttnRelBiasArgs), C.c_void_p]),
    "ovb_peer_signal_wait": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_uint32, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "ovb_enable_peer_access": (C.c_int, [C.c_int32]),
    "ovb_layernorm": (C.c_int, [C.POINTER(LnArgs), C.c_void_p]),
    "ovb_rmsnorm_rope": (C.c_int, [C.POINTER(RmsRopeArgs), C.c_void_p]),
    "ovb_patchify": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                               C.c_int32, C.c_int32, C.c_void_p]),
    "ovb_patchify3d": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ovb_unpatchify": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int32, C.c_void_p]),
    "ovb_linear_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_void_p]),
    "ovb_sinusoidal_embedding": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_void_p]),
    "ovb_modulation_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ovb_lincomb": (C.c_int, [C.POINTER(LinCombArgs), C.c_void_p]),
    "ovb_conv3d": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "ovb_tap_sum": (C.c_int, [C.POINTER(TapSumArgs), C.c_void_p]),
    "ovb_rmsnorm_silu_cl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                      C.c_int32, C.c_void_p]),
    "ovb_planar_to_cl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                   C.c_int32, C.c_int64, C.c_void_p]),
    "ovb_cl_to_planar": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_int64, C.c_void_p]),
    "ovb_softmax_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                   C.c_float, C.c_void_p]),
    "ovb_cast_f32_to_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ovb_cast_bf16_to_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
}

_lib = None


class OvbError(RuntimeError):
    pass


def load():
    """Load libovb200.so (once).  Raises OvbError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OvbError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C omni_video_b200/csrc`. There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.ovb_abi_version() != ABI_VERSION:
        raise OvbError(f"ABI version mismatch: library {lib.ovb_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def register(name, restype, argtypes):
    """Used by optional parts of the binding (VAE kernels) to add their symbols."""
    _SIGNATURES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


def check(rc: int, what: str):
    if rc != OVB_OK:
        msg = load().ovb_last_error().decode("utf-8", "replace")
        raise OvbError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().ovb_launch_count())
