"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes loaders for the two CPU checkers:

* ``PortOracle``  -> oracle/libscloracle.so  (oracle/scl_oracle.c, the plain-C
  restatement; built on demand with gcc, runs anywhere)
* ``RefOracle``   -> oracle/_ref/libsclref.so (the reference's own, unmodified
  sources + oracle/ref_driver.cc; can only be BUILT where /root/reference exists,
  but the prebuilt .so travels to the GPU box)

Both expose the same methods so a test can be parametrised over them.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product never does.

Element conventions (SCL's FF::write bytes): Fp<61> -> numpy uint64; Fp<127> ->
numpy uint64 with a trailing axis of 2 (low word, high word) = 16 LE bytes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SCL_REFERENCE_ROOT", "/root/reference")
PORT_SO = os.path.join(HERE, "libscloracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsclref.so")
REF_SO_V3 = os.path.join(HERE, "_ref", "libsclref_v3.so")  # same sources, -march=x86-64-v3
REF_FLAGS = {REF_SO: "-O3 -maes -msse4.1", REF_SO_V3: "-O3 -march=x86-64-v3 -maes -mpclmul"}


def _cpu_has_v3() -> bool:
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split()
    except (OSError, StopIteration):
        return False
    return all(f in flags for f in ("avx2", "bmi2", "fma", "aes"))

P61 = (1 << 61) - 1
P127 = (1 << 127) - 1
PRIME = {61: P61, 127: P127}
SUFFIX = {61: "fp61", 127: "fp127"}
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p


def build_port(force: bool = False) -> str:
    src = [os.path.join(HERE, f) for f in ("scl_oracle.c", "scl_oracle.h", "scl_oracle_field.inc")]
    if force or not os.path.exists(PORT_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(PORT_SO) for s in src
    ):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return PORT_SO


def build_ref(force: bool = False) -> str | None:
    """Build oracle/_ref from the reference sources if they are present here."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "scl")):
        return REF_SO if os.path.exists(REF_SO) else None
    drv = os.path.join(HERE, "ref_driver.cc")
    if (force or not os.path.exists(REF_SO) or not os.path.exists(REF_SO_V3)
            or os.path.getmtime(drv) > min(os.path.getmtime(REF_SO), os.path.getmtime(REF_SO_V3))):
        subprocess.check_call(["make", "-C", HERE, "ref", f"REF={REFERENCE_ROOT}"], stdout=subprocess.DEVNULL)
    return REF_SO


def ref_available() -> bool:
    return os.path.exists(REF_SO)


# ------------------------------------------------------------------ helpers


def seed16(seed) -> bytes:
    """PRG::create(seed): zero-pad / truncate to 16 bytes (prg.cc:88-101)."""
    if isinstance(seed, str):
        seed = seed.encode()
    seed = bytes(seed)[:16]
    return seed + b"\0" * (16 - len(seed))


def elem_shape(field: int):
    return () if field == 61 else (2,)


def empty(field: int, *shape) -> np.ndarray:
    return np.zeros(tuple(shape) + elem_shape(field), dtype=np.uint64)


def from_ints(vals, field: int) -> np.ndarray:
    vals = np.asarray(vals, dtype=object)
    flat = [int(v) for v in vals.reshape(-1)]
    if field == 61:
        out = np.array(flat, dtype=np.uint64)
        return out.reshape(vals.shape)
    out = np.array([[v & 0xFFFFFFFFFFFFFFFF, v >> 64] for v in flat], dtype=np.uint64)
    return out.reshape(vals.shape + (2,))


def to_ints(arr: np.ndarray, field: int):
    arr = np.asarray(arr, dtype=np.uint64)
    if field == 61:
        return np.array([int(v) for v in arr.reshape(-1)], dtype=object).reshape(arr.shape)
    flat = arr.reshape(-1, 2)
    vals = [int(lo) | (int(hi) << 64) for lo, hi in flat]
    return np.array(vals, dtype=object).reshape(arr.shape[:-1])


def _c(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _nelem(a: np.ndarray, field: int) -> int:
    return a.size if field == 61 else a.size // 2


# ------------------------------------------------------------------ oracles


class _Base:
    kind = "?"

    def to_ints(self, arr, field):
        return to_ints(arr, field)

    def from_ints(self, vals, field):
        return from_ints(vals, field)


class PortOracle(_Base):
    """oracle/scl_oracle.c through ctypes."""

    kind = "port"

    def __init__(self):
        self.lib = C.CDLL(build_port())
        L = self.lib
        L.sclo_prg_next.argtypes = [_vp, C.c_uint64, C.c_uint64, _vp]
        L.sclo_prg_next.restype = None
        for f in ("fp61", "fp127"):
            g = lambda n: getattr(L, f"sclo_{f}_{n}")
            g("scalar_op").argtypes = [C.c_int, _vp, _vp, _vp]
            g("scalar_op").restype = C.c_int
            g("from_bytes").argtypes = [_vp, C.c_uint64, _vp]
            g("from_bytes").restype = None
            for n in ("vector_random", "ff_random"):
                g(n).argtypes = [_vp, C.c_uint64, C.c_uint64, _vp]
                g(n).restype = None
            g("shamir_share").argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp, C.c_uint64, _vp]
            g("shamir_share").restype = None
            g("shamir_share_array").argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _vp, C.c_uint64, _vp]
            g("shamir_share_array").restype = None
            g("recover_p_array").argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp]
            g("recover_p_array").restype = C.c_int
            g("hyper_invertible").argtypes = [C.c_uint64, C.c_uint64, _vp]
            g("hyper_invertible").restype = C.c_int
            g("recover_c").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp]
            g("recover_c").restype = C.c_int64
            g("additive_share").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, C.c_uint64, _vp]
            g("additive_share").restype = None
            g("additive_recover").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp]
            g("additive_recover").restype = None
            g("lagrange").argtypes = [_vp, C.c_uint64, _vp, _vp]
            g("lagrange").restype = C.c_int
            g("recover_p").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp]
            g("recover_p").restype = C.c_int
            g("recover_d").argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp]
            g("recover_d").restype = C.c_int64
            g("vec_op").argtypes = [C.c_int, _vp, _vp, C.c_uint64, _vp]
            g("vec_op").restype = C.c_int
            g("beaver").argtypes = [_vp] * 5 + [C.c_uint64, _vp]
            g("beaver").restype = None
            g("matvec").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, _vp]
            g("matvec").restype = None
            g("matmul").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, C.c_uint64, _vp]
            g("matmul").restype = None
            g("vandermonde").argtypes = [C.c_uint64, C.c_uint64, _vp]
            g("vandermonde").restype = None
            g("bench_share_recover").argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int]
            g("bench_share_recover").restype = C.c_double

    def _f(self, field, name):
        return getattr(self.lib, f"sclo_{SUFFIX[field]}_{name}")

    def prg_next(self, seed, first_block: int, n_bytes: int) -> np.ndarray:
        out = np.zeros(n_bytes, dtype=np.uint8)
        self.lib.sclo_prg_next(seed16(seed), first_block, n_bytes, _p(out))
        return out

    def scalar_op(self, field, op, a: int, b: int | None = None):
        A = from_ints([a], field)
        B = from_ints([b if b is not None else 0], field)
        out = empty(field, 1)
        rc = self._f(field, "scalar_op")(op, _p(A), _p(B), _p(out))
        return rc, int(to_ints(out, field)[0])

    def from_bytes(self, field, raw: bytes):
        bs = 8 if field == 61 else 16
        n = len(raw) // bs
        src = np.frombuffer(raw, dtype=np.uint8).copy()
        out = empty(field, n)
        self._f(field, "from_bytes")(_p(src), n, _p(out))
        return out

    def vector_random(self, field, seed, first_block, n):
        out = empty(field, n)
        self._f(field, "vector_random")(seed16(seed), first_block, n, _p(out))
        return out

    def ff_random(self, field, seed, first_block, n):
        out = empty(field, n)
        self._f(field, "ff_random")(seed16(seed), first_block, n, _p(out))
        return out

    def shamir_share(self, field, secrets, t, n, seed, first_block=0):
        secrets = _c(secrets)
        N = _nelem(secrets, field)
        out = empty(field, N, n)
        self._f(field, "shamir_share")(_p(secrets), N, t, n, seed16(seed), first_block, _p(out))
        return out

    def shamir_share_array(self, field, secrets, t, n, seed, first_block=0):
        """secrets [N, W] -> shares [N, n, W] (shamirSecretShare on math::Array<FF, W>)."""
        secrets = _c(secrets)
        N, W = secrets.shape[0], secrets.shape[1]
        out = empty(field, N, n, W)
        self._f(field, "shamir_share_array")(_p(secrets), N, W, t, n, seed16(seed), first_block, _p(out))
        return out

    def recover_p_array(self, field, shares):
        shares = _c(shares)
        N, n, W = shares.shape[0], shares.shape[1], shares.shape[2]
        out = empty(field, N, W)
        if self._f(field, "recover_p_array")(_p(shares), N, W, n, _p(out)):
            raise ValueError("0 not invertible modulo prime")
        return out

    def hyper_invertible(self, field, n, m):
        out = empty(field, n, m)
        if self._f(field, "hyper_invertible")(n, m, _p(out)):
            raise ValueError("0 not invertible modulo prime")
        return out

    def additive_share(self, field, secrets, n, seed, first_block=0):
        secrets = _c(secrets)
        N = _nelem(secrets, field)
        out = empty(field, N, n)
        self._f(field, "additive_share")(_p(secrets), N, n, seed16(seed), first_block, _p(out))
        return out

    def additive_recover(self, field, shares):
        shares = _c(shares)
        N, n = shares.shape[0], shares.shape[1]
        out = empty(field, N)
        self._f(field, "additive_recover")(_p(shares), N, n, _p(out))
        return out

    def recover_c(self, field, shares, alphas=None):
        """shamirRecoverC per sharing -> (f [N][3t+1], err [N][t+1], status [N] uint8, n_failed)"""
        shares = _c(shares)
        N, n = shares.shape[0], shares.shape[1]
        t = (n - 1) // 3
        f = empty(field, N, 3 * t + 1)
        e = empty(field, N, t + 1)
        st = np.zeros(N, dtype=np.uint8)
        A = None if alphas is None else _c(alphas)
        nf = self._f(field, "recover_c")(_p(shares), N, n, _p(A), _p(f), _p(e), _p(st))
        return f, e, st, int(nf)

    def lagrange(self, field, nodes, x: int):
        nodes = _c(nodes)
        n = _nelem(nodes, field)
        X = from_ints([x], field)
        out = empty(field, n)
        rc = self._f(field, "lagrange")(_p(nodes), n, _p(X), _p(out))
        if rc:
            raise ValueError("0 not invertible modulo prime")
        return out

    def recover_p(self, field, shares, alphas=None, x: int | None = None):
        shares = _c(shares)
        N, n = shares.shape[0], shares.shape[1]
        out = empty(field, N)
        A = None if alphas is None else _c(alphas)
        X = from_ints([x or 0], field)
        rc = self._f(field, "recover_p")(_p(shares), N, n, _p(A), _p(X), _p(out))
        if rc:
            raise ValueError("0 not invertible modulo prime")
        return out

    def recover_d(self, field, shares, t, alphas=None, d=None, x: int | None = None):
        """-> (secrets, err[N] uint8, rc) ; rc == -1 <=> 'not enough shares...'"""
        shares = _c(shares)
        N, n_given = shares.shape[0], shares.shape[1]
        out = empty(field, N)
        err = np.zeros(N, dtype=np.uint8)
        A = None if alphas is None else _c(alphas)
        n_alphas = 0 if alphas is None else _nelem(A, field)
        X = from_ints([x or 0], field)
        rc = self._f(field, "recover_d")(
            _p(shares), N, n_given, t, _p(A), n_alphas, d if d is not None else t, _p(X), _p(out), _p(err)
        )
        return out, err, int(rc)

    def vec_op(self, field, op, a, b=None):
        a = _c(a)
        n = _nelem(a, field)
        b = _c(b) if b is not None else empty(field, 1)
        out = empty(field, 1 if op in (4, 5) else n)
        rc = self._f(field, "vec_op")(op, _p(a), _p(b), n, _p(out))
        assert rc == 0
        return out

    def beaver(self, field, e, b, d, a, c):
        e, b, d, a, c = map(_c, (e, b, d, a, c))
        n = _nelem(e, field)
        z = empty(field, n)
        self._f(field, "beaver")(_p(e), _p(b), _p(d), _p(a), _p(c), n, _p(z))
        return z

    def matvec(self, field, A, x):
        A, x = _c(A), _c(x)
        rows, cols = A.shape[0], A.shape[1]
        y = empty(field, rows)
        self._f(field, "matvec")(_p(A), rows, cols, _p(x), _p(y))
        return y

    def matmul(self, field, A, Bm):
        A, Bm = _c(A), _c(Bm)
        rows, inner, cols = A.shape[0], A.shape[1], Bm.shape[1]
        out = empty(field, rows, cols)
        self._f(field, "matmul")(_p(A), rows, inner, _p(Bm), cols, _p(out))
        return out

    def vandermonde(self, field, n, m):
        out = empty(field, n, m)
        self._f(field, "vandermonde")(n, m, _p(out))
        return out

    def bench_share_recover(self, field, N, t, n, detect, threads):
        return float(self._f(field, "bench_share_recover")(N, t, n, int(detect), threads))


class RefOracle(_Base):
    """The unmodified reference (oracle/_ref/libsclref.so) through ctypes."""

    kind = "reference"

    def __init__(self, prefer_v3: bool = True):
        so = build_ref()
        if so is None or not os.path.exists(so):
            raise FileNotFoundError("oracle/_ref/libsclref.so not built and /root/reference absent")
        if prefer_v3 and os.path.exists(REF_SO_V3) and _cpu_has_v3():
            so = REF_SO_V3
        self.build_flags = REF_FLAGS.get(so, "?")
        self.lib = C.CDLL(so)
        L = self.lib
        L.sclref_prg_next.argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp]
        L.sclref_prg_next.restype = None
        for f in ("fp61", "fp127"):
            g = lambda n: getattr(L, f"sclref_{f}_{n}")
            for n in ("vector_random", "ff_random"):
                g(n).argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp]
                g(n).restype = None
            g("shamir_share").argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp, C.c_uint64, C.c_uint64, _vp]
            g("shamir_share").restype = None
            g("shamir_share_array").argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _vp, C.c_uint64, C.c_uint64, _vp]
            g("shamir_share_array").restype = C.c_int
            g("recover_p_array").argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp]
            g("recover_p_array").restype = C.c_int
            g("hyper_invertible").argtypes = [C.c_uint64, C.c_uint64, _vp]
            g("hyper_invertible").restype = None
            g("recover_c").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp]
            g("recover_c").restype = C.c_int64
            g("additive_share").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, C.c_uint64, C.c_uint64, _vp]
            g("additive_share").restype = None
            g("additive_recover").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp]
            g("additive_recover").restype = None
            g("recover_p").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp]
            g("recover_p").restype = None
            g("recover_d").argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, _vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp]
            g("recover_d").restype = C.c_int64
            g("lagrange").argtypes = [_vp, C.c_uint64, _vp, _vp]
            g("lagrange").restype = None
            g("vec_op").argtypes = [C.c_int, _vp, _vp, C.c_uint64, _vp]
            g("vec_op").restype = C.c_int
            g("beaver").argtypes = [_vp] * 5 + [C.c_uint64, _vp]
            g("beaver").restype = None
            g("matvec").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, _vp]
            g("matvec").restype = None
            g("matmul").argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, C.c_uint64, _vp]
            g("matmul").restype = None
            g("vandermonde").argtypes = [C.c_uint64, C.c_uint64, _vp]
            g("vandermonde").restype = None
            g("scalar_op").argtypes = [C.c_int, _vp, _vp, _vp]
            g("scalar_op").restype = C.c_int
            g("bench_share_recover").argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64, _vp]
            g("bench_share_recover").restype = C.c_double

    def _f(self, field, name):
        return getattr(self.lib, f"sclref_{SUFFIX[field]}_{name}")

    @staticmethod
    def _seed(seed):
        if isinstance(seed, str):
            seed = seed.encode()
        return bytes(seed), len(seed)

    def prg_next(self, seed, first_block, n_bytes):
        s, sl = self._seed(seed)
        out = np.zeros(n_bytes, dtype=np.uint8)
        self.lib.sclref_prg_next(s, sl, first_block, n_bytes, _p(out))
        return out

    def scalar_op(self, field, op, a: int, b: int | None = None):
        A = from_ints([a], field)
        B = from_ints([b if b is not None else 0], field)
        out = empty(field, 1)
        rc = self._f(field, "scalar_op")(op, _p(A), _p(B), _p(out))
        return rc, int(to_ints(out, field)[0])

    def from_bytes(self, field, raw: bytes):
        # FF::read through the vec path: add zero (reads both operands with T::read)
        bs = 8 if field == 61 else 16
        n = len(raw) // bs
        src = np.frombuffer(raw, dtype=np.uint8).copy()
        zero = np.zeros(n * bs, dtype=np.uint8)
        out = empty(field, n)
        self._f(field, "vec_op")(0, _p(src), _p(zero), n, _p(out))
        return out

    def vector_random(self, field, seed, first_block, n):
        s, sl = self._seed(seed)
        out = empty(field, n)
        self._f(field, "vector_random")(s, sl, first_block, n, _p(out))
        return out

    def ff_random(self, field, seed, first_block, n):
        s, sl = self._seed(seed)
        out = empty(field, n)
        self._f(field, "ff_random")(s, sl, first_block, n, _p(out))
        return out

    def shamir_share(self, field, secrets, t, n, seed, first_block=0):
        s, sl = self._seed(seed)
        secrets = _c(secrets)
        N = _nelem(secrets, field)
        out = empty(field, N, n)
        self._f(field, "shamir_share")(_p(secrets), N, t, n, s, sl, first_block, _p(out))
        return out

    def shamir_share_array(self, field, secrets, t, n, seed, first_block=0):
        """secrets [N, W] -> shares [N, n, W]; W in 1..5 (the widths instantiated in ref_driver.cc)."""
        s, sl = self._seed(seed)
        secrets = _c(secrets)
        N, W = secrets.shape[0], secrets.shape[1]
        out = empty(field, N, n, W)
        if self._f(field, "shamir_share_array")(_p(secrets), N, W, t, n, s, sl, first_block, _p(out)):
            raise ValueError("array width not instantiated in the reference driver")
        return out

    def recover_p_array(self, field, shares):
        shares = _c(shares)
        N, n, W = shares.shape[0], shares.shape[1], shares.shape[2]
        out = empty(field, N, W)
        if self._f(field, "recover_p_array")(_p(shares), N, W, n, _p(out)):
            raise ValueError("array width not instantiated in the reference driver")
        return out

    def hyper_invertible(self, field, n, m):
        out = empty(field, n, m)
        self._f(field, "hyper_invertible")(n, m, _p(out))
        return out

    def additive_share(self, field, secrets, n, seed, first_block=0):
        s, sl = self._seed(seed)
        secrets = _c(secrets)
        N = _nelem(secrets, field)
        out = empty(field, N, n)
        self._f(field, "additive_share")(_p(secrets), N, n, s, sl, first_block, _p(out))
        return out

    def additive_recover(self, field, shares):
        shares = _c(shares)
        N, n = shares.shape[0], shares.shape[1]
        out = empty(field, N)
        self._f(field, "additive_recover")(_p(shares), N, n, _p(out))
        return out

    def recover_c(self, field, shares, alphas=None):
        """shamirRecoverC per sharing -> (f [N][3t+1], err [N][t+1], status [N] uint8, n_failed)"""
        shares = _c(shares)
        N, n = shares.shape[0], shares.shape[1]
        t = (n - 1) // 3
        f = empty(field, N, 3 * t + 1)
        e = empty(field, N, t + 1)
        st = np.zeros(N, dtype=np.uint8)
        A = None if alphas is None else _c(alphas)
        nf = self._f(field, "recover_c")(_p(shares), N, n, _p(A), _p(f), _p(e), _p(st))
        return f, e, st, int(nf)

    def lagrange(self, field, nodes, x: int):
        nodes = _c(nodes)
        n = _nelem(nodes, field)
        X = from_ints([x], field)
        out = empty(field, n)
        self._f(field, "lagrange")(_p(nodes), n, _p(X), _p(out))
        return out

    def recover_p(self, field, shares, alphas=None, x: int | None = None):
        shares = _c(shares)
        N, n = shares.shape[0], shares.shape[1]
        out = empty(field, N)
        A = None if alphas is None else _c(alphas)
        X = from_ints([x or 0], field)
        self._f(field, "recover_p")(_p(shares), N, n, _p(A), _p(X), _p(out))
        return out

    def recover_d(self, field, shares, t, alphas=None, d=None, x: int | None = None):
        shares = _c(shares)
        N, n_given = shares.shape[0], shares.shape[1]
        out = empty(field, N)
        err = np.zeros(N, dtype=np.uint8)
        A = None if alphas is None else _c(alphas)
        n_alphas = 0 if alphas is None else _nelem(A, field)
        X = from_ints([x or 0], field)
        rc = self._f(field, "recover_d")(
            _p(shares), N, n_given, t, 0 if alphas is None else 1, _p(A), n_alphas,
            d if d is not None else t, _p(X), _p(out), _p(err),
        )
        return out, err, int(rc)

    def vec_op(self, field, op, a, b=None):
        a = _c(a)
        n = _nelem(a, field)
        b = _c(b) if b is not None else empty(field, max(n, 1))
        out = empty(field, 1 if op in (4, 5) else n)
        rc = self._f(field, "vec_op")(op, _p(a), _p(b), n, _p(out))
        assert rc == 0
        return out

    def beaver(self, field, e, b, d, a, c):
        e, b, d, a, c = map(_c, (e, b, d, a, c))
        n = _nelem(e, field)
        z = empty(field, n)
        self._f(field, "beaver")(_p(e), _p(b), _p(d), _p(a), _p(c), n, _p(z))
        return z

    def matvec(self, field, A, x):
        A, x = _c(A), _c(x)
        rows, cols = A.shape[0], A.shape[1]
        y = empty(field, rows)
        self._f(field, "matvec")(_p(A), rows, cols, _p(x), _p(y))
        return y

    def matmul(self, field, A, Bm):
        A, Bm = _c(A), _c(Bm)
        rows, inner, cols = A.shape[0], A.shape[1], Bm.shape[1]
        out = empty(field, rows, cols)
        self._f(field, "matmul")(_p(A), rows, inner, _p(Bm), cols, _p(out))
        return out

    def vandermonde(self, field, n, m):
        out = empty(field, n, m)
        self._f(field, "vandermonde")(n, m, _p(out))
        return out

    def vandermonde_xs(self, field, n, m, xs):
        """Matrix::vandermonde(n, m, xs); raises ValueError where the reference throws invalid_argument."""
        xs = _c(xs)
        out = empty(field, n, m)
        f = self._f(field, "vandermonde_xs")
        f.argtypes = [C.c_uint64, C.c_uint64, _vp, C.c_uint64, _vp]
        f.restype = C.c_int
        if f(n, m, _p(xs), xs.size if field == 61 else xs.size // 2, _p(out)):
            raise ValueError("|xs| != number of rows")
        return out

    def poly_evaluate(self, field, coeffs, xs):
        """Polynomial::create(coeffs[j]).evaluate(xs[i]) -> [N, n]."""
        coeffs, xs = _c(coeffs), _c(xs)
        N, m = coeffs.shape[0], coeffs.shape[1]
        n = xs.size if field == 61 else xs.size // 2
        out = empty(field, N, n)
        f = self._f(field, "poly_evaluate")
        f.argtypes = [_vp, C.c_uint64, C.c_uint64, _vp, C.c_uint64, _vp]
        f.restype = None
        f(_p(coeffs), N, m, _p(xs), n, _p(out))
        return out

    def transpose(self, field, A):
        A = _c(A)
        rows, cols = A.shape[0], A.shape[1]
        out = empty(field, cols, rows)
        f = self._f(field, "transpose")
        f.argtypes = [_vp, C.c_uint64, C.c_uint64, _vp]
        f.restype = None
        f(_p(A), rows, cols, _p(out))
        return out

    def bench_share_recover(self, field, N, t, n, detect, threads):
        bs = 8 if field == 61 else 16
        blocks = ((t + 1) * bs + 15) // 16
        chk = empty(field, 1)
        return float(self._f(field, "bench_share_recover")(N, t, n, int(detect), threads, blocks, _p(chk)))


def best_oracle():
    """The reference if its prebuilt .so is present, else the port."""
    try:
        return RefOracle()
    except (FileNotFoundError, OSError):
        return PortOracle()
