"""The C-ABI library builds, loads without a GPU, and exports every symbol include/a3t_b200.h declares."""
import ctypes
import os
import re

from a3t_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "a3t_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(a3t_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    _lib.build()
    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/a3t_b200.h but not exported"
    # the python binding declares the same set
    assert set(syms) == set(_lib.EXPORTED_SYMBOLS), set(syms) ^ set(_lib.EXPORTED_SYMBOLS)


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.a3t_version() == 100
    assert isinstance(lib.a3t_last_error(), bytes)


def test_gemm_desc_struct_layout_matches_header():
    # 16 int32 + 4 float + uint32 + int32 pad = 88 bytes, then 18 int64
    assert ctypes.sizeof(_lib.GemmDesc) == 88 + 18 * 8
    assert _lib.GemmDesc.sa_m.offset == 88


def test_argument_validation_without_gpu():
    """Argument errors are reported before any launch, so they can be exercised on the CPU box."""
    lib = _lib.load()
    rc = lib.a3t_layernorm_fwd(None, None, None, None, 0, None, None, 4, 384, 1e-5, 0, 1.0, 0.0, None, 0, None)
    assert rc == -1 and b"null" in lib.a3t_last_error()
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.mode, d.batch1, d.batch2 = 8, 8, 7, _lib.GEMM_CONV, 1, 1
    d.taps, d.cin, d.seq = 3, 2, 4
    rc = lib.a3t_gemm(ctypes.byref(d), 1, 1, 1, None, None, None, None, None)
    assert rc == -1 and b"taps*cin" in lib.a3t_last_error()
