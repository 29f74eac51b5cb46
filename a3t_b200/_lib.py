"""ctypes binding of the C-ABI library `liba3t_b200.so` (include/a3t_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, an
exception is raised.  `build()` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liba3t_b200.so")
CSRC = os.path.join(_HERE, "csrc")

A3T_F32, A3T_BF16 = 0, 1
ACT_NONE, ACT_SWISH, ACT_TANH = 0, 1, 2
GEMM_PLAIN, GEMM_CONV, GEMM_WGRAD = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_TC, IMPL_TC_PAIR = 0, 1, 2, 3


class A3TError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("mode", C.c_int32),
        ("taps", C.c_int32), ("pad", C.c_int32), ("seq", C.c_int32), ("cin", C.c_int32),
        ("batch1", C.c_int32), ("batch2", C.c_int32),
        ("dtype_a", C.c_int32), ("dtype_b", C.c_int32), ("dtype_c", C.c_int32), ("dtype_mask", C.c_int32),
        ("relu", C.c_int32),
        ("impl", C.c_int32),
        ("alpha", C.c_float), ("out_scale", C.c_float), ("mask_scale", C.c_float),
        ("drop_p", C.c_float),
        ("drop_site", C.c_uint32),
        ("c_zeroed", C.c_int32),
        ("sa_m", C.c_int64), ("sa_k", C.c_int64), ("sa_b1", C.c_int64), ("sa_b2", C.c_int64),
        ("sb_n", C.c_int64), ("sb_k", C.c_int64), ("sb_b1", C.c_int64), ("sb_b2", C.c_int64), ("sb_tap", C.c_int64),
        ("sc_m", C.c_int64), ("sc_n", C.c_int64), ("sc_b1", C.c_int64), ("sc_b2", C.c_int64), ("sc_tap", C.c_int64),
        ("sr_m", C.c_int64), ("sr_n", C.c_int64), ("sr_b1", C.c_int64), ("sr_b2", C.c_int64),
    ]


class PackItem(C.Structure):
    _fields_ = [("w", C.c_void_p * 4), ("fwd", C.c_void_p), ("dgrad", C.c_void_p),
                ("N", C.c_int32), ("C", C.c_int32), ("taps", C.c_int32), ("seg_rows", C.c_int32),
                ("tile_start", C.c_int32), ("tiles_c", C.c_int32), ("_pad", C.c_int32 * 2)]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_F = C.c_float
_U = C.c_uint32

# name -> argtypes (restype is int unless listed in _RESTYPE)
_SIGS = {
    "a3t_version": [],
    "a3t_gemm": [C.POINTER(GemmDesc), _P, _P, _P, _P, _P, _P, _P, _P],
    "a3t_gemm_tc_supported": [C.POINTER(GemmDesc), _P, _P, _P],
    "a3t_gemm_fallback_count": [_I],
    "a3t_pack_conv_weight": [_P, _I, _I, _I, _P, _P, _P],
    "a3t_pack_conv_weights": [_P, _I, _I, _I, _P],
    "a3t_qkv4_bias": [_P, _P, _P, _P, _P, _P, _I, _P],
    "a3t_layernorm_fwd": [_P, _P, _P, _P, _I, _P, _P, _L, _I, _F, _I, _F, _F, _P, _U, _P],
    "a3t_layernorm_bwd_blocks": [_L],
    "a3t_layernorm_bwd": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _F, _F, _P, _U, _P, _I, _F, _F, _U, _P, _P],
    "a3t_colsum_blocks": [_L],
    "a3t_colsum": [_P, _I, _P, _P, _L, _I, _L, _P],
    "a3t_scale_dropout": [_P, _P, _I, _L, _F, _F, _P, _U, _P],
    "a3t_mask_input_fwd": [_P, _P, _P, _P, _I, _L, _I, _P],
    "a3t_mask_input_bwd": [_P, _P, _P, _P, _L, _I, _P],
    "a3t_embed_assemble_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _U, _U, _I, _I, _P, _P],
    "a3t_embed_assemble_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _I, _F, _P, _U, _U, _I, _I, _P],
    "a3t_relpos_softmax_fwd": [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _P, _U, _P],
    "a3t_relpos_softmax_bwd": [_P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _I, _F, _F, _P, _U, _P],
    "a3t_attn_fused_supported": [_I, _I, _I, _I],
    "a3t_attn_set_trace": [_P],
    "a3t_relpos_attn_fwd": [_P, _P, _L, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _U, _P],
    "a3t_relpos_attn_bwd": [_P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _U, _P],
    "a3t_glu_dwconv_fwd": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _P],
    "a3t_dwconv_bwd_blocks": [_I, _I],
    "a3t_glu_dwconv_bwd": [_P, _P, _I, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P],
    "a3t_bn_stats": [_P, _P, _P, _P, _P, _P, _P, _L, _I, _F, _F, _I, _P],
    "a3t_bn_act_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _P, _U, _P],
    "a3t_bn_act_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _F, _P, _U, _P],
    "a3t_masked_l1_fwd": [_P, _P, _P, _P, _P, _P, _L, _I, _P],
    "a3t_masked_l1_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P],
    "a3t_grad_sqnorm": [_P, _L, _P, _P, _P],
    "a3t_adam_step": [_P, _P, _P, _P, _L, _P, _P, _F, _F, _F, _F, _F, _F, _F, _F, _P, _P],
    "a3t_seed_advance": [_P, _P],
    "a3t_stft_logmel": [_P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _I, _I, _P],
    "a3t_align_to_frames": [_P, _P, _L, _F, _F, _P],
    "a3t_expand_phone_mask": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "a3t_segment_pos": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "a3t_pwg_upsample": [_P, _P, _P, _I, _L, _I, _P],
    "a3t_pwg_conv1d": [_P, _P, _P, _P, _I, _I, _I, _L, _I, _I, _I, _I, _F, _P],
    "a3t_pwg_last": [_P, _P, _P, _P, _P, _P, _I, _L, _F, _P],
    "a3t_pwg_split_planes": [_P, _P, _P, _I, _I, _L, _P],
    "a3t_pwg_resblock_tc": [_P] * 13 + [_I, _L, _I, _I, _P],
    "a3t_pwg_resblock": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _I, _I, _I, _I, _P],
}
# entry points that return a count / flag rather than a status code
_PLAIN_INT = {"a3t_version", "a3t_layernorm_bwd_blocks", "a3t_colsum_blocks", "a3t_dwconv_bwd_blocks",
              "a3t_gemm_tc_supported", "a3t_gemm_fallback_count", "a3t_attn_fused_supported"}

EXPORTED_SYMBOLS = sorted(list(_SIGS) + ["a3t_last_error"])

_lock = threading.Lock()
_lib = None
launch_count = 0  # number of C-ABI compute calls made (bench.py reports kernel launches from it)


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise A3TError("building liba3t_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def load():
    """Load the library (once).  Raises if it has not been built: there is no CPU fallback."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise A3TError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the a3t_b200 product path has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.a3t_last_error.restype = C.c_char_p
        lib.a3t_last_error.argtypes = []
        for name, args in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = lib
        return lib


def call(name: str, *args) -> int:
    """Invoke a status-returning entry point; raise A3TError with the library's message on failure."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    if name in _PLAIN_INT:
        return rc
    launch_count += 1
    if rc != 0:
        msg = lib.a3t_last_error()
        raise A3TError(f"{name} failed (rc={rc}): {msg.decode() if msg else '?'}")
    return rc
