"""Host-side handle on libsclgpu.so: one `Context` per GPU per process.

Two families of methods, both thin wrappers over the C ABI (include/sclgpu.h):

* host methods take / return numpy arrays in SCL's FF::write layout (Fp<61> ->
  uint64[...]; Fp<127> -> uint64[..., 2] = low word, high word) and go through the
  library's own pinned / chunked H2D-compute-D2H pipelines;
* ``*_dev`` methods take torch CUDA tensors (int64 storage, same bit layout) and
  enqueue on torch's current stream -- torch is only the allocator / stream owner.

Errors mirror the reference: SCLGPU_EINVAL -> InvalidArgument (std::invalid_argument),
SCLGPU_ELOGIC / EDETECT -> LogicError (std::logic_error) carrying the reference's
what() string.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import binding as B

P61 = (1 << 61) - 1
P127 = (1 << 127) - 1
PRIME = {61: P61, 127: P127}
_SUF = {61: "fp61", 127: "fp127"}


class InvalidArgument(ValueError):
    """std::invalid_argument"""


class LogicError(RuntimeError):
    """std::logic_error"""


class CudaError(RuntimeError):
    pass


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
        return np.array(flat, dtype=np.uint64).reshape(vals.shape)
    out = np.array([[v & 0xFFFFFFFFFFFFFFFF, v >> 64] for v in flat], dtype=np.uint64)
    return out.reshape(vals.shape + (2,))


def to_ints(arr, field: int):
    arr = np.asarray(arr, dtype=np.uint64)
    if field == 61:
        return np.array([int(v) for v in arr.reshape(-1)], dtype=object).reshape(arr.shape)
    flat = arr.reshape(-1, 2)
    return np.array([int(lo) | (int(hi) << 64) for lo, hi in flat], dtype=object).reshape(arr.shape[:-1])


def blocks_per_share_call(field: int, t: int) -> int:
    """Keystream blocks one shamirSecretShare call consumes: ceil((t+1)*byteSize/16)."""
    bs = 8 if field == 61 else 16
    return ((t + 1) * bs + 15) // 16


def blocks_per_array_share_call(field: int, width: int, t: int) -> int:
    """Blocks one shamirSecretShare(math::Array<FF, width>, t, ...) consumes."""
    return ((t + 1) * width * (8 if field == 61 else 16) + 15) // 16


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _nelem(a: np.ndarray, field: int) -> int:
    return a.size if field == 61 else a.size // 2


def _dp(t):
    """device pointer of a torch tensor (or None)"""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise InvalidArgument("expected a contiguous CUDA tensor")
    return C.c_void_p(t.data_ptr())


class Context:
    def __init__(self, device: int = 0):
        self.lib = B.load()
        self._ctx = C.c_void_p()
        rc = self.lib.sclgpu_init(int(device), C.byref(self._ctx))
        if rc != B.OK:
            self._ctx = None
            raise CudaError(
                f"sclgpu_init(device={device}) failed: {self.lib.sclgpu_strerror(rc).decode()} "
                "(libsclgpu needs a CUDA device; there is no CPU fallback)"
            )
        self.device = int(device)

    # ------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.sclgpu_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, allow=()):
        if rc == B.OK or rc in allow:
            return rc
        msg = self.lib.sclgpu_last_error(self._ctx).decode(errors="replace")
        if rc == B.EINVAL:
            raise InvalidArgument(msg)
        if rc in (B.ELOGIC, B.EDETECT, B.ECORRECT):
            raise LogicError(msg)
        raise CudaError(f"{self.lib.sclgpu_strerror(rc).decode()}: {msg}")

    def _f(self, field: int, name: str):
        return getattr(self.lib, f"sclgpu_{_SUF[field]}_{name}")

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self.lib.sclgpu_set_stream(self._ctx, C.c_void_p(cuda_stream_ptr or 0)))

    def use_torch_stream(self):
        import torch

        self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def sync(self):
        self._check(self.lib.sclgpu_sync(self._ctx))

    @property
    def launch_count(self) -> int:
        return int(self.lib.sclgpu_launch_count(self._ctx))

    def device_info(self) -> dict:
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        fr, tot = C.c_size_t(), C.c_size_t()
        self._check(self.lib.sclgpu_device_info(self._ctx, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(fr), C.byref(tot)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "free": fr.value, "total": tot.value}

    def host_alloc(self, nbytes: int) -> np.ndarray:
        """pinned host buffer as a uint8 numpy array (release with host_free)"""
        p = C.c_void_p()
        self._check(self.lib.sclgpu_host_alloc(self._ctx, nbytes, C.byref(p)))
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint8)

    def host_free(self, arr: np.ndarray):
        """`arr` is the array host_alloc returned (or a view starting at its first byte)"""
        self._check(self.lib.sclgpu_host_free(self._ctx, C.c_void_p(arr.ctypes.data)))

    def pipe_microbench(self, kind: int, iters: int = 4096) -> float:
        v = C.c_double()
        self._check(self.lib.sclgpu_pipe_microbench(self._ctx, kind, iters, C.byref(v)))
        return v.value

    # ------------------------------------------------------------ PRG
    def prg_expand(self, seed, first_block: int, n_bytes: int) -> np.ndarray:
        out = np.zeros(n_bytes, dtype=np.uint8)
        self._check(self.lib.sclgpu_prg_expand(self._ctx, seed16(seed), first_block, n_bytes, _p(out)))
        return out

    def prg_expand_dev(self, seed, first_block: int, n_bytes: int, out):
        self._check(self.lib.sclgpu_prg_expand_dev(self._ctx, seed16(seed), first_block, n_bytes, _dp(out)))

    def prg_expand_bitsliced_dev(self, seed, first_block: int, n_bytes: int, out):
        """The same keystream from the bitsliced AES kernel (whole, aligned blocks): the comparison arm."""
        self._check(self.lib.sclgpu_prg_expand_bitsliced_dev(self._ctx, seed16(seed), first_block, n_bytes, _dp(out)))

    def from_bytes(self, field: int, raw) -> np.ndarray:
        bs = 8 if field == 61 else 16
        src = np.frombuffer(bytes(raw), dtype=np.uint8).copy() if not isinstance(raw, np.ndarray) else np.ascontiguousarray(raw, dtype=np.uint8)
        n = src.size // bs
        out = empty(field, n)
        self._check(self._f(field, "from_bytes")(self._ctx, _p(src), n, _p(out)))
        return out

    def random(self, field: int, seed, first_block: int, n: int) -> np.ndarray:
        out = empty(field, n)
        self._check(self._f(field, "random")(self._ctx, seed16(seed), first_block, n, _p(out)))
        return out

    # oracle-compatible aliases
    vector_random = random

    def ff_random(self, field: int, seed, first_block: int, n: int) -> np.ndarray:
        out = empty(field, n)
        self._check(self._f(field, "ff_random")(self._ctx, seed16(seed), first_block, n, _p(out)))
        return out

    def random_dev(self, field: int, seed, first_block: int, n: int, out):
        self._check(self._f(field, "random_dev")(self._ctx, seed16(seed), first_block, n, _dp(out)))

    def ff_random_dev(self, field: int, seed, first_block: int, n: int, out):
        self._check(self._f(field, "ff_random_dev")(self._ctx, seed16(seed), first_block, n, _dp(out)))

    # ------------------------------------------------------------ Shamir
    def shamir_share(self, field: int, secrets, t: int, n: int, seed, first_block: int = 0) -> np.ndarray:
        secrets = _c(secrets)
        N = _nelem(secrets, field)
        out = empty(field, N, n)
        self._check(self._f(field, "shamir_share")(self._ctx, _p(secrets), N, t, n, seed16(seed), first_block, _p(out)))
        return out

    def shamir_share_dev(self, field: int, secrets, N: int, t: int, n: int, seed, first_block: int, shares,
                         layout: int = B.PARTY_MAJOR):
        self._check(self._f(field, "shamir_share_dev")(self._ctx, _dp(secrets), N, t, n, seed16(seed), first_block, _dp(shares), layout))

    def shamir_share_recover_dev(self, secrets, N: int, t: int, n: int, seed, first_block: int, shares, out,
                                 rec_shares=None, alphas=None, x: int | None = None):
        """Fp61, party-major planes: shamirSecretShare of `secrets` into `shares` AND shamirRecoverP of
        `rec_shares` (default: the sharings just produced) into `out`, in one launch."""
        A = None if alphas is None else _c(alphas)
        X = from_ints([x or 0], 61)
        rs = shares if rec_shares is None else rec_shares
        self._check(self.lib.sclgpu_fp61_shamir_share_recover_dev(self._ctx, _dp(secrets), N, t, n, seed16(seed), first_block,
                                                                  _dp(shares), _dp(rs), _p(A), _p(X), _dp(out)))

    def shamir_share_recover_gather_dev(self, secrets, N: int, t: int, n: int, seed, first_block: int, shares, dst_ptrs,
                                        offset: int, rec_shares=None, alphas=None, x: int | None = None):
        """shamir_share_recover_dev with the all-gather fused in: reconstructed secret j goes to dst_ptrs[r] + 8*(offset + j)
        for every r (device addresses valid on this device: own or peer memory)."""
        A = None if alphas is None else _c(alphas)
        X = from_ints([x or 0], 61)
        rs = shares if rec_shares is None else rec_shares
        arr = (C.c_void_p * len(dst_ptrs))(*[int(p) for p in dst_ptrs])
        self._check(self.lib.sclgpu_fp61_shamir_share_recover_gather_dev(
            self._ctx, _dp(secrets), N, t, n, seed16(seed), first_block, _dp(shares), _dp(rs), _p(A), _p(X), arr, len(dst_ptrs), offset))

    def shamir_share_coeffs_dev(self, field: int, coeffs, N: int, t: int, n: int, shares, layout: int = B.PARTY_MAJOR):
        self._check(self._f(field, "shamir_share_coeffs_dev")(self._ctx, _dp(coeffs), N, t, n, _dp(shares), layout))

    def recover_c(self, field: int, shares, alphas=None):
        """ss::shamirRecoverC (shamir.h:203-258) per sharing ->
        (f [N][3t+1], err [N][t+1], status [N] uint8, n_failed); the secrets are f[:, 0]."""
        shares = _c(shares)
        N, n = shares.shape[0], shares.shape[1]
        if n < 1:
            raise InvalidArgument("shamirRecoverC needs at least one share")
        t = (n - 1) // 3
        f = empty(field, N, 3 * t + 1)
        e = empty(field, N, t + 1)
        st = np.zeros(N, dtype=np.uint8)
        nf = C.c_uint64(0)
        A = None if alphas is None else _c(alphas)
        rc = self._f(field, "recover_c")(self._ctx, _p(shares), N, n, _p(A), _p(f), _p(e), _p(st), C.byref(nf))
        self._check(rc, allow=(B.ECORRECT,))
        return f, e, st, int(nf.value)

    def recover_c_dev(self, field: int, shares, N: int, n: int, f, err, status, layout: int = B.PARTY_MAJOR,
                      alphas=None) -> int:
        nf = C.c_uint64(0)
        A = None if alphas is None else _c(alphas)
        rc = self._f(field, "recover_c_dev")(self._ctx, _dp(shares), N, n, layout, _p(A), _dp(f), _dp(err), _dp(status),
                                              C.byref(nf))
        self._check(rc, allow=(B.ECORRECT,))
        return int(nf.value)

    # ------------------------------------------------------------ per-party packets
    def shamir_share_packets(self, field: int, secrets, t: int, n: int, seed, first_block: int = 0) -> list:
        """n uint8 arrays: packet i = Serializer<Vector<FF>>::write of party i's shares (u32 count + elements)."""
        secrets = _c(secrets)
        N = _nelem(secrets, field)
        nbytes = int(self.lib.sclgpu_packet_bytes(8 if field == 61 else 16, N))
        packets = [np.zeros(nbytes, dtype=np.uint8) for _ in range(n)]
        ptrs = (C.c_void_p * max(n, 1))(*[p.ctypes.data for p in packets])
        self._check(self._f(field, "shamir_share_packets")(self._ctx, _p(secrets), N, t, n, seed16(seed), first_block, ptrs))
        return packets

    def recover_p_packets(self, field: int, packets, N: int, alphas=None, x: int | None = None) -> np.ndarray:
        packets = [np.ascontiguousarray(p, dtype=np.uint8) for p in packets]
        n = len(packets)
        ptrs = (C.c_void_p * max(n, 1))(*[p.ctypes.data for p in packets])
        out = empty(field, N)
        A = None if alphas is None else _c(alphas)
        X = from_ints([x or 0], field)
        self._check(self._f(field, "recover_p_packets")(self._ctx, ptrs, N, n, _p(A), _p(X), _p(out)))
        return out

    # ------------------------------------------------------------ array-valued secrets
    def shamir_share_array(self, field: int, secrets, t: int, n: int, seed, first_block: int = 0) -> np.ndarray:
        """shamirSecretShare on math::Array<FF, W> (shamir.h:52-68; pedersen.h:137-138 for W = 2):
        secrets [N, W] -> shares [N, n, W]; consumes N*ceil((t+1)*W*bs/16) blocks."""
        secrets = _c(secrets)
        if secrets.ndim != (2 if field == 61 else 3):
            raise InvalidArgument("secrets must be [N, W]")
        N, W = secrets.shape[0], secrets.shape[1]
        out = empty(field, N, n, W)
        self._check(self._f(field, "shamir_share_array")(self._ctx, _p(secrets), N, W, t, n, seed16(seed), first_block, _p(out)))
        return out

    def shamir_share_array_dev(self, field: int, secrets, N: int, W: int, t: int, n: int, seed, first_block: int,
                               shares, layout: int = B.PARTY_MAJOR):
        self._check(self._f(field, "shamir_share_array_dev")(self._ctx, _dp(secrets), N, W, t, n, seed16(seed), first_block, _dp(shares), layout))

    def recover_p_array(self, field: int, shares) -> np.ndarray:
        """shamirRecoverP on Vector<Array<FF, W>> (shamir.h:100-104): shares [N, n, W] -> [N, W]."""
        shares = _c(shares)
        N, n, W = shares.shape[0], shares.shape[1], shares.shape[2]
        out = empty(field, N, W)
        self._check(self._f(field, "recover_p_array")(self._ctx, _p(shares), N, W, n, _p(out)))
        return out

    def recover_p_array_dev(self, field: int, shares, N: int, W: int, n: int, out, layout: int = B.PARTY_MAJOR):
        self._check(self._f(field, "recover_p_array_dev")(self._ctx, _dp(shares), N, W, n, layout, _dp(out)))

    def hyper_invertible(self, field: int, n: int, m: int) -> np.ndarray:
        """Matrix::hyperInvertible(n, m) (matrix.h:462-475), row-major [n, m]."""
        out = empty(field, max(n, 0), max(m, 0))
        self._check(self._f(field, "hyper_invertible")(self._ctx, n, m, _p(out)))
        return out

    # ------------------------------------------------------------ additive sharing
    def additive_share(self, field: int, secrets, n: int, seed, first_block: int = 0) -> np.ndarray:
        """ss::additiveShare (additive.h:42-53) of every secret; consumes N*(n-1) blocks."""
        if n < 1:
            raise InvalidArgument("additiveShare needs n >= 1")
        secrets = _c(secrets)
        N = _nelem(secrets, field)
        out = empty(field, N, n)
        self._check(self._f(field, "additive_share")(self._ctx, _p(secrets), N, n, seed16(seed), first_block, _p(out)))
        return out

    def additive_share_dev(self, field: int, secrets, N: int, n: int, seed, first_block: int, shares,
                           layout: int = B.PARTY_MAJOR):
        self._check(self._f(field, "additive_share_dev")(self._ctx, _dp(secrets), N, n, seed16(seed), first_block, _dp(shares), layout))

    def additive_recover(self, field: int, shares) -> np.ndarray:
        shares = _c(shares)
        N, n = shares.shape[0], shares.shape[1]
        out = empty(field, N)
        self._check(self._f(field, "additive_recover")(self._ctx, _p(shares), N, n, _p(out)))
        return out

    def additive_recover_dev(self, field: int, shares, N: int, n: int, out, layout: int = B.PARTY_MAJOR):
        self._check(self._f(field, "additive_recover_dev")(self._ctx, _dp(shares), N, n, layout, _dp(out)))

    def lagrange(self, field: int, nodes, x: int, n: int | None = None) -> np.ndarray:
        nodes_a = None if nodes is None else _c(nodes)
        n = _nelem(nodes_a, field) if nodes_a is not None else int(n)
        X = from_ints([x], field)
        out = empty(field, n)
        self._check(self._f(field, "lagrange_basis")(self._ctx, _p(nodes_a), n, _p(X), _p(out)))
        return out

    def recover_p(self, field: int, shares, alphas=None, x: int | None = None) -> np.ndarray:
        shares = _c(shares)
        N, n = shares.shape[0], shares.shape[1]
        out = empty(field, N)
        A = None if alphas is None else _c(alphas)
        X = from_ints([x or 0], field)
        self._check(self._f(field, "recover_p")(self._ctx, _p(shares), N, n, _p(A), _p(X), _p(out)))
        return out

    def recover_p_dev(self, field: int, shares, N: int, n: int, out, layout: int = B.PARTY_MAJOR, alphas=None,
                      x: int | None = None):
        A = None if alphas is None else _c(alphas)
        X = from_ints([x or 0], field)
        self._check(self._f(field, "recover_p_dev")(self._ctx, _dp(shares), N, n, layout, _p(A), _p(X), _dp(out)))

    def recover_p_gather_dev(self, shares, N: int, n: int, dst_ptrs, offset: int, alphas=None, x: int | None = None):
        """Fp61 shamirRecoverP of this rank's N sharings (party-major planes), result j stored to dst_ptrs[r] + 8*(offset + j)
        for every r: the all-gather over peer memory happens inside the reconstruction kernel.  dst_ptrs: device
        addresses (ints) valid on this device -- own memory or peer memory (malloc / ipc_export / ipc_open)."""
        A = None if alphas is None else _c(alphas)
        X = from_ints([x or 0], 61)
        arr = (C.c_void_p * len(dst_ptrs))(*[int(p) for p in dst_ptrs])
        self._check(self.lib.sclgpu_fp61_recover_p_gather_dev(self._ctx, _dp(shares), N, n, _p(A), _p(X), arr, len(dst_ptrs), offset))

    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self.lib.sclgpu_malloc(self._ctx, nbytes, C.byref(p)))
        return int(p.value)

    def free(self, ptr: int):
        self._check(self.lib.sclgpu_free(self._ctx, C.c_void_p(ptr)))

    def memcpy_d2d(self, dst: int, src: int, nbytes: int):
        self._check(self.lib.sclgpu_memcpy_d2d(self._ctx, C.c_void_p(dst), C.c_void_p(src), nbytes))

    def ipc_export(self, ptr: int) -> bytes:
        h = (C.c_uint8 * 64)()
        self._check(self.lib.sclgpu_ipc_export(self._ctx, C.c_void_p(ptr), h))
        return bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        self._check(self.lib.sclgpu_ipc_open(self._ctx, h, C.byref(p)))
        return int(p.value)

    def ipc_close(self, ptr: int):
        self._check(self.lib.sclgpu_ipc_close(self._ctx, C.c_void_p(ptr)))

    def enable_peer(self, peer_device: int):
        self._check(self.lib.sclgpu_enable_peer(self._ctx, int(peer_device)))

    # ------------------------------------------------------------ asynchronous host calls
    def shamir_share_async(self, field: int, secrets: np.ndarray, t: int, n: int, seed, first_block: int, out: np.ndarray):
        """sclgpu_*_shamir_share_async: returns at once; `secrets` / `out` (caller-owned, e.g. pinned) must stay alive
        until wait() / sync()."""
        N = _nelem(secrets, field)
        self._check(self._f(field, "shamir_share_async")(self._ctx, _p(secrets), N, t, n, seed16(seed), first_block, _p(out)))

    def recover_p_async(self, field: int, shares: np.ndarray, N: int, n: int, out: np.ndarray, alphas=None, x: int | None = None):
        A = None if alphas is None else _c(alphas)
        X = None if alphas is None else from_ints([x or 0], field)
        self._check(self._f(field, "recover_p_async")(self._ctx, _p(shares), N, n, _p(A), _p(X), _p(out)))

    def wait(self):
        self._check(self.lib.sclgpu_wait(self._ctx))

    def recover_d(self, field: int, shares, t: int, alphas=None, d: int | None = None, x: int | None = None):
        """-> (secrets, err uint8[N], rc) with rc = -1 for "not enough shares provided to
        detect errors", else the number of secrets flagged (oracle-compatible)."""
        shares = _c(shares)
        N, n_given = shares.shape[0], shares.shape[1]
        out = empty(field, N)
        err = np.zeros(N, dtype=np.uint8)
        A = None if alphas is None else _c(alphas)
        n_alphas = 0 if A is None else _nelem(A, field)
        X = from_ints([x or 0], field)
        nd = C.c_uint64(0)
        rc = self._f(field, "recover_d")(self._ctx, _p(shares), N, n_given, t, _p(A), n_alphas,
                                          d if d is not None else t, _p(X), _p(out), _p(err), C.byref(nd))
        if rc == B.ELOGIC:
            return out, err, -1
        self._check(rc, allow=(B.EDETECT,))
        return out, err, int(nd.value)

    def recover_d_dev(self, field: int, shares, N: int, n_given: int, t: int, out, err, layout: int = B.PARTY_MAJOR,
                      alphas=None, d: int | None = None, x: int | None = None) -> int:
        A = None if alphas is None else _c(alphas)
        n_alphas = 0 if A is None else _nelem(A, field)
        X = from_ints([x or 0], field)
        nd = C.c_uint64(0)
        rc = self._f(field, "recover_d_dev")(self._ctx, _dp(shares), N, n_given, layout, t, _p(A), n_alphas,
                                              d if d is not None else t, _p(X), _dp(out), _dp(err), C.byref(nd))
        self._check(rc, allow=(B.EDETECT,))
        return int(nd.value)

    # ------------------------------------------------------------ Vector / Matrix
    _OPS = {0: "vec_add", 1: "vec_sub", 2: "vec_mul", 3: "vec_scale", 4: "dot", 5: "sum"}

    def vec_op(self, field: int, op: int, a, b=None) -> np.ndarray:
        """op: 0 add 1 subtract 2 multiplyEntryWise 3 scalarMultiply(b[0]) 4 dot 5 sum"""
        a = _c(a)
        n = _nelem(a, field)
        out = empty(field, 1 if op in (4, 5) else n)
        f = self._f(field, self._OPS[op])
        if op == 5:
            self._check(f(self._ctx, _p(a), n, _p(out)))
        else:
            b = _c(b)
            self._check(f(self._ctx, _p(a), _p(b), n, _p(out)))
        return out

    def vec_equal(self, field: int, a, b) -> bool:
        """Vector::equals (vector.h:358-375)."""
        a, b = _c(a), _c(b)
        if _nelem(a, field) != _nelem(b, field):
            return False
        eq = C.c_int(0)
        self._check(self._f(field, "vec_equal")(self._ctx, _p(a), _p(b), _nelem(a, field), C.byref(eq)))
        return bool(eq.value)

    def vec_equal_dev(self, field: int, a, b, n: int) -> bool:
        eq = C.c_int(0)
        self._check(self._f(field, "vec_equal_dev")(self._ctx, _dp(a), _dp(b), n, C.byref(eq)))
        return bool(eq.value)

    def vec_op_dev(self, field: int, op: int, a, b, n: int, out):
        f = self._f(field, self._OPS[op] + "_dev")
        if op == 5:
            self._check(f(self._ctx, _dp(a), n, _dp(out)))
        elif op == 3:
            self._check(f(self._ctx, _dp(a), _p(_c(b)), n, _dp(out)))
        else:
            self._check(f(self._ctx, _dp(a), _dp(b), n, _dp(out)))

    def beaver(self, field: int, e, b, d, a, c) -> np.ndarray:
        e, b, d, a, c = map(_c, (e, b, d, a, c))
        n = _nelem(e, field)
        z = empty(field, n)
        self._check(self._f(field, "vec_muladd")(self._ctx, _p(e), _p(b), _p(d), _p(a), _p(c), n, _p(z)))
        return z

    def beaver_dev(self, field: int, e, b, d, a, c, n: int, z):
        self._check(self._f(field, "vec_muladd_dev")(self._ctx, _dp(e), _dp(b), _dp(d), _dp(a), _dp(c), n, _dp(z)))

    def matvec(self, field: int, A, x) -> np.ndarray:
        A, x = _c(A), _c(x)
        rows, cols = A.shape[0], A.shape[1]
        y = empty(field, rows)
        self._check(self._f(field, "matvec")(self._ctx, _p(A), rows, cols, _p(x), _p(y)))
        return y

    def matmul(self, field: int, A, Bm) -> np.ndarray:
        """Matrix::multiply(Matrix) (matrix.h:476-495)."""
        A, Bm = _c(A), _c(Bm)
        if A.shape[1] != Bm.shape[0]:
            raise InvalidArgument("matmul: this->cols() != that->rows()")
        rows, inner, cols = A.shape[0], A.shape[1], Bm.shape[1]
        out = empty(field, rows, cols)
        self._check(self._f(field, "matmul")(self._ctx, _p(A), rows, inner, _p(Bm), cols, _p(out)))
        return out

    def matmul_dev(self, field: int, A, rows: int, inner: int, Bm, cols: int, out):
        self._check(self._f(field, "matmul_dev")(self._ctx, _dp(A), rows, inner, _dp(Bm), cols, _dp(out)))

    def matvec_dev(self, field: int, A, rows: int, cols: int, x, y):
        self._check(self._f(field, "matvec_dev")(self._ctx, _dp(A), rows, cols, _dp(x), _dp(y)))

    def vandermonde(self, field: int, n: int, m: int) -> np.ndarray:
        out = empty(field, n, m)
        self._check(self._f(field, "vandermonde")(self._ctx, n, m, _p(out)))
        return out

    def vandermonde_xs(self, field: int, n: int, m: int, xs) -> np.ndarray:
        """Matrix::vandermonde(n, m, xs) (matrix.h:445-460)."""
        xs = _c(xs)
        out = empty(field, n, m)
        self._check(self._f(field, "vandermonde_xs")(self._ctx, n, m, _p(xs), _nelem(xs, field), _p(out)))
        return out

    def poly_evaluate(self, field: int, coeffs, xs) -> np.ndarray:
        """Polynomial::evaluate (poly.h:56-64): coeffs [N, t+1] (constant term first), xs [n] -> [N, n]."""
        coeffs, xs = _c(coeffs), _c(xs)
        N, m = coeffs.shape[0], coeffs.shape[1]
        n = _nelem(xs, field)
        out = empty(field, N, n)
        self._check(self._f(field, "poly_evaluate")(self._ctx, _p(coeffs), N, m - 1, _p(xs), n, _p(out)))
        return out

    def poly_evaluate_dev(self, field: int, coeffs, N: int, t: int, xs, out, layout: int = B.PARTY_MAJOR):
        xs = _c(xs)
        self._check(self._f(field, "poly_evaluate_dev")(self._ctx, _dp(coeffs), N, t, _p(xs), _nelem(xs, field), _dp(out), layout))

    def transpose(self, field: int, A) -> np.ndarray:
        """Matrix::transpose (matrix.h:344-355)."""
        A = _c(A)
        rows, cols = A.shape[0], A.shape[1]
        out = empty(field, cols, rows)
        self._check(self._f(field, "transpose")(self._ctx, _p(A), rows, cols, _p(out)))
        return out

    def transpose_dev(self, field: int, src, rows: int, cols: int, dst):
        self._check(self._f(field, "transpose_dev")(self._ctx, _dp(src), rows, cols, _dp(dst)))


class MultiContext:
    """sclgpu_mctx: several GPUs of one box behind one handle, for a single-process caller.  Host arrays in SCL's
    layout; [0, N) is cut into contiguous slices, one per device, PRG counters offset to match."""

    def __init__(self, devices):
        self.lib = B.load()
        devs = list(devices)
        arr = (C.c_int * len(devs))(*devs)
        self._m = C.c_void_p()
        rc = self.lib.sclgpu_multi_init(arr, len(devs), C.byref(self._m))
        if rc != B.OK:
            self._m = None
            raise CudaError(f"sclgpu_multi_init({devs}) failed: {self.lib.sclgpu_strerror(rc).decode()} (no CPU fallback)")
        self.devices = devs

    def close(self):
        if getattr(self, "_m", None):
            self.lib.sclgpu_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, allow=()):
        if rc == B.OK or rc in allow:
            return rc
        msg = self.lib.sclgpu_multi_last_error(self._m).decode(errors="replace")
        if rc == B.EINVAL:
            raise InvalidArgument(msg)
        if rc in (B.ELOGIC, B.EDETECT, B.ECORRECT):
            raise LogicError(msg)
        raise CudaError(f"{self.lib.sclgpu_strerror(rc).decode()}: {msg}")

    def _f(self, field: int, name: str):
        return getattr(self.lib, f"sclgpu_multi_{_SUF[field]}_{name}")

    def random(self, seed, first_block: int, n: int) -> np.ndarray:
        out = empty(61, n)
        self._check(self.lib.sclgpu_multi_fp61_random(self._m, seed16(seed), first_block, n, _p(out)))
        return out

    def shamir_share(self, field: int, secrets, t: int, n: int, seed, first_block: int = 0, out=None) -> np.ndarray:
        secrets = _c(secrets)
        N = _nelem(secrets, field)
        if out is None:
            out = empty(field, N, n)
        self._check(self._f(field, "shamir_share")(self._m, _p(secrets), N, t, n, seed16(seed), first_block, _p(out)))
        return out

    def recover_p(self, field: int, shares, alphas=None, x: int | None = None, out=None) -> np.ndarray:
        shares = _c(shares)
        N, n = shares.shape[0], shares.shape[1]
        if out is None:
            out = empty(field, N)
        A = None if alphas is None else _c(alphas)
        X = from_ints([x or 0], field)
        self._check(self._f(field, "recover_p")(self._m, _p(shares), N, n, _p(A), _p(X), _p(out)))
        return out

    def recover_d(self, field: int, shares, t: int, alphas=None, d: int | None = None, x: int | None = None):
        shares = _c(shares)
        N, n_given = shares.shape[0], shares.shape[1]
        out = empty(field, N)
        err = np.zeros(N, dtype=np.uint8)
        A = None if alphas is None else _c(alphas)
        n_alphas = 0 if A is None else _nelem(A, field)
        X = from_ints([x or 0], field)
        nd = C.c_uint64(0)
        rc = self._f(field, "recover_d")(self._m, _p(shares), N, n_given, t, _p(A), n_alphas, d if d is not None else t,
                                          _p(X), _p(out), _p(err), C.byref(nd))
        if rc == B.ELOGIC:
            return out, err, -1
        self._check(rc, allow=(B.EDETECT,))
        return out, err, int(nd.value)
