"""ctypes binding of libsclgpu.so (the C ABI declared in include/sclgpu.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a only) at
``secure-computation-library_b200/csrc/libsclgpu.so``.  There is no fallback of
any kind: if the shared object is missing, ``load()`` raises, and every compute
entry point of the library itself fails with SCLGPU_ECUDA when no B200 is usable.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "csrc", "libsclgpu.so")
HEADER_PATH = os.path.join(REPO, "include", "sclgpu.h")

OK, EINVAL, ELOGIC, EDETECT, ECUDA, ENOMEM, ECORRECT = 0, -1, -2, -3, -4, -5, -6
SECRET_MAJOR, PARTY_MAJOR = 0, 1

_vp = C.c_void_p
_u64 = C.c_uint64
_u32 = C.c_uint32
_int = C.c_int


class SclGpuMissing(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Every function name include/sclgpu.h declares."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sclgpu_[a-z0-9_]+)\s*\(", src)))


def _sig(lib, name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)


_LIB = None


def load() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise SclGpuMissing(
            f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    _sig(lib, "sclgpu_init", _int, _int, C.POINTER(_vp))
    _sig(lib, "sclgpu_destroy", None, _vp)
    _sig(lib, "sclgpu_set_stream", _int, _vp, _vp)
    _sig(lib, "sclgpu_sync", _int, _vp)
    _sig(lib, "sclgpu_last_error", C.c_char_p, _vp)
    _sig(lib, "sclgpu_strerror", C.c_char_p, _int)
    _sig(lib, "sclgpu_launch_count", _u64, _vp)
    _sig(lib, "sclgpu_device_info", _int, _vp, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int),
         C.POINTER(C.c_size_t), C.POINTER(C.c_size_t))
    _sig(lib, "sclgpu_malloc", _int, _vp, C.c_size_t, C.POINTER(_vp))
    _sig(lib, "sclgpu_free", _int, _vp, _vp)
    _sig(lib, "sclgpu_host_alloc", _int, _vp, C.c_size_t, C.POINTER(_vp))
    _sig(lib, "sclgpu_host_free", _int, _vp, _vp)
    _sig(lib, "sclgpu_memcpy_h2d", _int, _vp, _vp, _vp, C.c_size_t)
    _sig(lib, "sclgpu_memcpy_d2h", _int, _vp, _vp, _vp, C.c_size_t)
    for suf in ("", "_dev", "_bitsliced_dev"):
        _sig(lib, "sclgpu_prg_expand" + suf, _int, _vp, _vp, _u64, _u64, _vp)
    for f in ("fp61", "fp127"):
        for suf in ("", "_dev"):
            _sig(lib, f"sclgpu_{f}_from_bytes{suf}", _int, _vp, _vp, _u64, _vp)
            _sig(lib, f"sclgpu_{f}_random{suf}", _int, _vp, _vp, _u64, _u64, _vp)
            _sig(lib, f"sclgpu_{f}_ff_random{suf}", _int, _vp, _vp, _u64, _u64, _vp)
            for op in ("add", "sub", "mul", "scale"):
                _sig(lib, f"sclgpu_{f}_vec_{op}{suf}", _int, _vp, _vp, _vp, _u64, _vp)
            _sig(lib, f"sclgpu_{f}_vec_muladd{suf}", _int, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp)
            _sig(lib, f"sclgpu_{f}_vec_equal{suf}", _int, _vp, _vp, _vp, _u64, C.POINTER(_int))
            _sig(lib, f"sclgpu_{f}_dot{suf}", _int, _vp, _vp, _vp, _u64, _vp)
            _sig(lib, f"sclgpu_{f}_sum{suf}", _int, _vp, _vp, _u64, _vp)
            _sig(lib, f"sclgpu_{f}_matvec{suf}", _int, _vp, _vp, _u32, _u32, _vp, _vp)
            _sig(lib, f"sclgpu_{f}_matmul{suf}", _int, _vp, _vp, _u32, _u32, _vp, _u32, _vp)
        _sig(lib, f"sclgpu_{f}_shamir_share", _int, _vp, _vp, _u64, _u32, _u32, _vp, _u64, _vp)
        _sig(lib, f"sclgpu_{f}_shamir_share_dev", _int, _vp, _vp, _u64, _u32, _u32, _vp, _u64, _vp, _int)
        _sig(lib, f"sclgpu_{f}_shamir_share_coeffs_dev", _int, _vp, _vp, _u64, _u32, _u32, _vp, _int)
        _sig(lib, f"sclgpu_{f}_recover_c", _int, _vp, _vp, _u64, _u32, _vp, _vp, _vp, _vp, C.POINTER(_u64))
        _sig(lib, f"sclgpu_{f}_recover_c_dev", _int, _vp, _vp, _u64, _u32, _int, _vp, _vp, _vp, _vp, C.POINTER(_u64))
        _sig(lib, f"sclgpu_{f}_shamir_share_packets", _int, _vp, _vp, _u64, _u32, _u32, _vp, _u64, _vp)
        _sig(lib, f"sclgpu_{f}_recover_p_packets", _int, _vp, _vp, _u64, _u32, _vp, _vp, _vp)
        _sig(lib, f"sclgpu_{f}_additive_share", _int, _vp, _vp, _u64, _u32, _vp, _u64, _vp)
        _sig(lib, f"sclgpu_{f}_additive_share_dev", _int, _vp, _vp, _u64, _u32, _vp, _u64, _vp, _int)
        _sig(lib, f"sclgpu_{f}_additive_recover", _int, _vp, _vp, _u64, _u32, _vp)
        _sig(lib, f"sclgpu_{f}_additive_recover_dev", _int, _vp, _vp, _u64, _u32, _int, _vp)
        _sig(lib, f"sclgpu_{f}_lagrange_basis", _int, _vp, _vp, _u32, _vp, _vp)
        _sig(lib, f"sclgpu_{f}_shamir_share_array", _int, _vp, _vp, _u64, _u32, _u32, _u32, _vp, _u64, _vp)
        _sig(lib, f"sclgpu_{f}_shamir_share_array_dev", _int, _vp, _vp, _u64, _u32, _u32, _u32, _vp, _u64, _vp, _int)
        _sig(lib, f"sclgpu_{f}_recover_p_array", _int, _vp, _vp, _u64, _u32, _u32, _vp)
        _sig(lib, f"sclgpu_{f}_recover_p_array_dev", _int, _vp, _vp, _u64, _u32, _u32, _int, _vp)
        _sig(lib, f"sclgpu_{f}_hyper_invertible", _int, _vp, _u32, _u32, _vp)
        _sig(lib, f"sclgpu_{f}_recover_p", _int, _vp, _vp, _u64, _u32, _vp, _vp, _vp)
        _sig(lib, f"sclgpu_{f}_recover_p_dev", _int, _vp, _vp, _u64, _u32, _int, _vp, _vp, _vp)
        _sig(lib, f"sclgpu_{f}_recover_d", _int, _vp, _vp, _u64, _u32, _u32, _vp, _u32, _u32, _vp, _vp,
             _vp, C.POINTER(_u64))
        _sig(lib, f"sclgpu_{f}_recover_d_dev", _int, _vp, _vp, _u64, _u32, _int, _u32, _vp, _u32, _u32,
             _vp, _vp, _vp, C.POINTER(_u64))
        _sig(lib, f"sclgpu_{f}_vandermonde", _int, _vp, _u32, _u32, _vp)
        _sig(lib, f"sclgpu_{f}_vandermonde_xs", _int, _vp, _u32, _u32, _vp, _u32, _vp)
        _sig(lib, f"sclgpu_{f}_poly_evaluate", _int, _vp, _vp, _u64, _u32, _vp, _u32, _vp)
        _sig(lib, f"sclgpu_{f}_poly_evaluate_dev", _int, _vp, _vp, _u64, _u32, _vp, _u32, _vp, _int)
        _sig(lib, f"sclgpu_{f}_transpose", _int, _vp, _vp, _u64, _u64, _vp)
        _sig(lib, f"sclgpu_{f}_transpose_dev", _int, _vp, _vp, _u64, _u64, _vp)
    _sig(lib, "sclgpu_fp61_shamir_share_recover_dev", _int, _vp, _vp, _u64, _u32, _u32, _vp, _u64, _vp, _vp, _vp, _vp, _vp)
    _sig(lib, "sclgpu_fp61_shamir_share_recover_gather_dev", _int, _vp, _vp, _u64, _u32, _u32, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _u32, _u64)
    _sig(lib, "sclgpu_fp61_recover_p_gather_dev", _int, _vp, _vp, _u64, _u32, _vp, _vp, _vp, _u32, _u64)
    _sig(lib, "sclgpu_memcpy_d2d", _int, _vp, _vp, _vp, C.c_size_t)
    _sig(lib, "sclgpu_ipc_export", _int, _vp, _vp, _vp)
    _sig(lib, "sclgpu_ipc_open", _int, _vp, _vp, C.POINTER(_vp))
    _sig(lib, "sclgpu_ipc_close", _int, _vp, _vp)
    _sig(lib, "sclgpu_enable_peer", _int, _vp, _int)
    _sig(lib, "sclgpu_multi_init", _int, _vp, _int, C.POINTER(_vp))
    _sig(lib, "sclgpu_multi_destroy", None, _vp)
    _sig(lib, "sclgpu_multi_device_count", _int, _vp)
    _sig(lib, "sclgpu_multi_slice", _int, _u64, _int, _int, _u64, C.POINTER(_u64), C.POINTER(_u64))
    _sig(lib, "sclgpu_multi_context", _vp, _vp, _int)
    _sig(lib, "sclgpu_multi_last_error", C.c_char_p, _vp)
    _sig(lib, "sclgpu_multi_fp61_random", _int, _vp, _vp, _u64, _u64, _vp)
    for f in ("fp61", "fp127"):
        _sig(lib, f"sclgpu_multi_{f}_shamir_share", _int, _vp, _vp, _u64, _u32, _u32, _vp, _u64, _vp)
        _sig(lib, f"sclgpu_multi_{f}_recover_p", _int, _vp, _vp, _u64, _u32, _vp, _vp, _vp)
        _sig(lib, f"sclgpu_multi_{f}_recover_d", _int, _vp, _vp, _u64, _u32, _u32, _vp, _u32, _u32, _vp, _vp, _vp, C.POINTER(_u64))
        _sig(lib, f"sclgpu_{f}_shamir_share_async", _int, _vp, _vp, _u64, _u32, _u32, _vp, _u64, _vp)
        _sig(lib, f"sclgpu_{f}_recover_p_async", _int, _vp, _vp, _u64, _u32, _vp, _vp, _vp)
    _sig(lib, "sclgpu_wait", _int, _vp)
    _sig(lib, "sclgpu_device_index", _int, _vp, C.POINTER(_int))
    _sig(lib, "sclgpu_set_error", None, _vp, C.c_char_p)
    _sig(lib, "sclgpu_packet_bytes", _u64, _u32, _u64)
    _sig(lib, "sclgpu_share_array_blocks", _u64, _u32, _u32, _u32)
    _sig(lib, "sclgpu_pipe_microbench", _int, _vp, _int, _u32, C.POINTER(C.c_double))
    _LIB = lib
    return lib
