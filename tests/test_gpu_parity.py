"""GPU parity tests: libsclgpu.so (through its C ABI) against the oracle on the
same seeded inputs, bit-exact (integer work: no tolerance anywhere), plus the
committed golden vectors recorded from the unmodified reference, plus
size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = {61: (1 << 61) - 1, 127: (1 << 127) - 1}


def ints(o, arr, field):
    return [int(v) for v in o.to_ints(arr, field).reshape(-1)]


def unhex(o, hexes, field, shape=None):
    a = o.from_ints([int(h, 16) for h in hexes], field)
    if shape is not None:
        a = a.reshape(tuple(shape) + (() if field == 61 else (2,)))
    return a


def test_device_is_b200(ctx):
    info = ctx.device_info()
    assert info["cc"][0] == 10, info
    assert info["sm_count"] >= 100


# ------------------------------------------------------------------ PRG
def test_prg_golden(ctx, golden):
    for c in golden["prg"]:
        got = ctx.prg_expand(c["seed"], c["first_block"], c["n_bytes"])
        assert bytes(got).hex() == c["hex"], c


@pytest.mark.parametrize("nbytes", [0, 1, 15, 16, 17, 4096, 1 << 20, (1 << 22) + 5])
def test_prg_vs_oracle(ctx, orc, nbytes):
    first = 12345 if orc.kind == "reference" else (1 << 40) + 77
    got = ctx.prg_expand("prg bench", first, nbytes)
    assert np.array_equal(got, orc.prg_next("prg bench", first, nbytes))


def test_prg_bitsliced_kernel_vs_oracle(ctx, pkg, port):
    """sclgpu_prg_expand_bitsliced_dev (k_prg_bitsliced): ranges that start and end inside a 32-counter group, cross the
    2^32 counter boundary, a single block; whole aligned blocks only."""
    import torch

    ctx.use_torch_stream()
    for seed, first, nblk in [("prg bench", 0, 4096), ("", 5, 1000), ("shamir bench", (1 << 32) - 40, 200), ("k", 31, 3), ("k", 64, 1),
                              ("k", (1 << 63) - 70, 64)]:
        d = torch.zeros(16 * nblk + 16, dtype=torch.uint8, device="cuda")
        ctx.prg_expand_bitsliced_dev(seed, first, 16 * nblk, d)
        got = d.cpu().numpy()
        assert np.array_equal(got[:16 * nblk], port.prg_next(seed, first, 16 * nblk)), (seed, first, nblk)
        assert not got[16 * nblk:].any()
    with pytest.raises(pkg.InvalidArgument):
        ctx.prg_expand_bitsliced_dev("k", 0, 40, torch.zeros(64, dtype=torch.uint8, device="cuda"))


def test_prg_counter_high_word(ctx, port):
    first = (1 << 32) - 3  # crosses the 32-bit boundary of the counter
    assert np.array_equal(ctx.prg_expand("k", first, 160), port.prg_next("k", first, 160))
    first = (1 << 63) - 4
    assert np.array_equal(ctx.prg_expand("k", first, 64), port.prg_next("k", first, 64))


# ------------------------------------------------------------------ random / read
def test_random_golden(ctx, golden):
    for c in golden["random"]:
        f = ctx.vector_random if c["kind"] == "vector" else ctx.ff_random
        got = f(c["field"], c["seed"], c["first_block"], c["n"])
        assert ints(ctx_o(), got, c["field"]) == [int(h, 16) for h in c["hex"]], c


class _O:
    @staticmethod
    def to_ints(arr, field):
        arr = np.asarray(arr, dtype=np.uint64)
        if field == 61:
            return np.array([int(v) for v in arr.reshape(-1)], dtype=object).reshape(arr.shape)
        flat = arr.reshape(-1, 2)
        return np.array([int(lo) | (int(hi) << 64) for lo, hi in flat], dtype=object).reshape(arr.shape[:-1])


def ctx_o():
    return _O


@pytest.mark.parametrize("field", [61, 127])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 1000, 1001, (1 << 20) + 1])
def test_random_vs_oracle(ctx, orc, field, n):
    assert np.array_equal(ctx.vector_random(field, "secrets", 5, n), orc.vector_random(field, "secrets", 5, n))
    m = min(n, 5000)
    assert np.array_equal(ctx.ff_random(field, "ff", 9, m), orc.ff_random(field, "ff", 9, m))


def test_from_bytes_golden_and_edges(ctx, port, golden):
    for c in golden["from_bytes"]:
        got = ctx.from_bytes(c["field"], bytes.fromhex(c["raw"]))
        assert ints(port, got, c["field"]) == [int(h, 16) for h in c["hex"]]
    raw = bytes(port.prg_next("raw", 0, 16 * 1000))
    for field in (61, 127):
        assert np.array_equal(ctx.from_bytes(field, raw), port.from_bytes(field, raw))


# ------------------------------------------------------------------ Shamir
def test_shamir_golden(ctx, port, golden):
    for c in golden["shamir"]:
        f = c["field"]
        secrets = unhex(port, c["secrets"], f)
        sh = ctx.shamir_share(f, secrets, c["t"], c["n"], c["seed"], c["first_block"])
        assert ints(port, sh, f) == [int(h, 16) for h in c["shares"]], (f, c["t"], c["n"])
        assert ints(port, ctx.recover_p(f, sh), f) == [int(h, 16) for h in c["recover_p"]]


def test_survey_sum_of_1023_calls(ctx, port, golden):
    s = port.from_ints([123 + j for j in range(1024)], 61)
    sh = ctx.shamir_share(61, s, 15, 32, "shamir bench")
    assert sum(ints(port, sh[1:], 61)) % P[61] == int(golden["survey_sum_1023"], 16) == 0x44672D90DD13206


@pytest.mark.parametrize("t,n,N", [(15, 32, 1 << 16), (15, 32, 1000), (2, 5, (1 << 15) + 77), (0, 1, 7), (1, 3, 33),
                                   (7, 16, 129), (15, 31, 4097), (6, 13, 100000), (16, 33, 257)])
def test_fused_share_recover_vs_oracle(ctx, orc, t, n, N):
    """sclgpu_fp61_shamir_share_recover_dev (k_share_recover61: share groups + reconstruction warps in one persistent
    launch): the share planes and the reconstructed secrets against the oracle, in the dependent mode (the tiles this
    launch stores are read back by the same CTA) and in the independent mode (another batch's planes)."""
    import torch
    secrets = orc.vector_random(61, "secrets", 3, N)
    first = 77
    want = orc.shamir_share(61, secrets, t, n, "shamir bench", first)          # [N][n]
    d_sec = torch.from_numpy(secrets.view(np.int64)).cuda()
    d_sh = torch.zeros((n, N), dtype=torch.int64, device="cuda")
    d_out = torch.zeros(N, dtype=torch.int64, device="cuda")
    ctx.shamir_share_recover_dev(d_sec, N, t, n, "shamir bench", first, d_sh, d_out)
    torch.cuda.synchronize()
    assert np.array_equal(d_sh.cpu().numpy().view(np.uint64).T, want)
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), orc.recover_p(61, want))
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), secrets)
    # independent mode: reconstruct ANOTHER batch (arbitrary canonical words, not a sharing) while sharing this one
    other = orc.vector_random(61, "other planes", 0, N * n).reshape(N, n)
    d_other = torch.from_numpy(np.ascontiguousarray(other.T).view(np.int64)).cuda()
    d_sh.zero_()
    d_out.zero_()
    ctx.shamir_share_recover_dev(d_sec, N, t, n, "shamir bench", first, d_sh, d_out, rec_shares=d_other)
    torch.cuda.synchronize()
    assert np.array_equal(d_sh.cpu().numpy().view(np.uint64).T, want)
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), orc.recover_p(61, other))
    # custom nodes / evaluation point
    alphas = orc.from_ints([3 * i + 2 for i in range(n)], 61)
    ctx.shamir_share_recover_dev(d_sec, N, t, n, "shamir bench", first, d_sh, d_out, rec_shares=d_other, alphas=alphas, x=5)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), orc.recover_p(61, other, alphas=alphas, x=5))


@pytest.mark.parametrize("field,t,n,N", [
    (61, 2, 5, 1 << 16), (61, 15, 32, 1 << 14), (61, 15, 32, 1000), (61, 0, 1, 7), (61, 1, 3, 33),
    (61, 16, 33, 257), (61, 17, 40, 129), (61, 31, 64, 65), (61, 3, 70000 // 1000, 100),
    (127, 7, 16, 1 << 13), (127, 2, 5, 1 << 14), (127, 0, 2, 5), (127, 8, 17, 300), (127, 9, 20, 123),
    (127, 15, 32, 64),
])
def test_share_recover_vs_oracle(ctx, orc, field, t, n, N):
    secrets = orc.vector_random(field, "secrets", 0, N)
    first = 1000
    got = ctx.shamir_share(field, secrets, t, n, "shamir bench", first)
    want = orc.shamir_share(field, secrets, t, n, "shamir bench", first)
    assert np.array_equal(got, want)
    rec = ctx.recover_p(field, got)
    assert np.array_equal(rec, orc.recover_p(field, want))
    assert np.array_equal(rec, secrets)


@pytest.mark.parametrize("first", [0, 1, 7, 8, 250, 255, 256, (1 << 32) - 3, (1 << 40) + 12])
def test_share_prg_offsets_aligned_and_not(ctx, pkg, port, first):
    """The fused share kernels draw two keystream blocks per thread and iteration when no secret of a warp crosses a
    256-counter group (first_block a multiple of the blocks per sharing) and one block at a time otherwise: both loops,
    with the crossing at every position inside a secret, for the share kernel (both fields, even and odd block counts)
    and for the single-launch step, against the oracle (shamir.h:52-68 on one PRG, prg.cc:124-146)."""
    import torch

    for field, t, n, N in [(61, 15, 32, 1500), (61, 7, 16, 700), (61, 4, 9, 300), (127, 7, 16, 700), (127, 4, 11, 260)]:
        sec = port.vector_random(field, "secrets", 0, N)
        assert np.array_equal(ctx.shamir_share(field, sec, t, n, "offsets", first), port.shamir_share(field, sec, t, n, "offsets", first)), \
            (field, t, n, first)
    ctx.use_torch_stream()
    N, t, n = 3000, 15, 32
    sec = port.vector_random(61, "secrets", 0, N)
    want = port.shamir_share(61, sec, t, n, "offsets", first)
    d_sec = torch.from_numpy(sec.view(np.int64)).cuda()
    d_prev = torch.from_numpy(np.ascontiguousarray(want.T).view(np.int64)).cuda()     # the batch to reconstruct: party-major planes
    d_sh = torch.zeros((n, N), dtype=torch.int64, device="cuda")
    d_out = torch.zeros(N, dtype=torch.int64, device="cuda")
    ctx.shamir_share_recover_dev(d_sec, N, t, n, "offsets", first, d_sh, d_out, rec_shares=d_prev)
    torch.cuda.synchronize()
    assert np.array_equal(d_sh.cpu().numpy().view(np.uint64).T, want) and np.array_equal(d_out.cpu().numpy().view(np.uint64), sec)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 8, 16, 31, 32])
def test_recover_p_secret_major_device_kernel(ctx, pkg, port, n):
    """sclgpu_fp61_recover_p_dev on SCL's own [N][n] layout (k_recover61_sm: the rows read where they lie, a warp per 32
    secrets): ragged batch sizes, default and custom nodes / evaluation point, arbitrary canonical words as shares --
    against the oracle's shamirRecoverP (shamir.h:82-104)."""
    import torch

    ctx.use_torch_stream()
    for N in (1, 31, 33, 1000, 4097):
        sh = port.vector_random(61, "any shares", 7, N * n).reshape(N, n)
        d_sh = torch.from_numpy(sh.view(np.int64)).cuda()
        d_out = torch.zeros(N + 2, dtype=torch.int64, device="cuda")
        ctx.recover_p_dev(61, d_sh, N, n, d_out, pkg.binding.SECRET_MAJOR)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy().view(np.uint64)
        assert np.array_equal(got[:N], port.recover_p(61, sh)) and not got[N:].any(), (n, N)
        alphas = port.from_ints([5 * i + 2 for i in range(n)], 61)
        ctx.recover_p_dev(61, d_sh, N, n, d_out, pkg.binding.SECRET_MAJOR, alphas=alphas, x=11)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy().view(np.uint64)[:N], port.recover_p(61, sh, alphas=alphas, x=11)), (n, N, "nodes")


def test_share_empty_and_degenerate(ctx, port):
    for field in (61, 127):
        e = port.from_ints([], field).reshape((0,) + (() if field == 61 else (2,)))
        assert ctx.shamir_share(field, e, 2, 5, "x").shape[0] == 0
        s = port.from_ints([9, 10], field)
        assert ctx.shamir_share(field, s, 2, 0, "x").shape[:2] == (2, 0)
        # t = 0: every share equals the secret
        sh = ctx.shamir_share(field, s, 0, 4, "x")
        assert ints(port, sh, field) == [9, 9, 9, 9, 10, 10, 10, 10]


def test_share_large_point_count(ctx, port):
    """n above the small-point Horner limit exercises the generic evaluation path."""
    s = port.from_ints([1, 2, 3], 61)
    n = 70000
    got = ctx.shamir_share(61, s, 2, n, "big n", 0)
    want = port.shamir_share(61, s, 2, n, "big n", 0)
    assert np.array_equal(got, want)


def test_recover_p_custom_golden(ctx, port, golden):
    for c in golden["recover_p_custom"]:
        f = c["field"]
        sh = unhex(port, c["shares"], f, (c["N"], c["n"]))
        out = ctx.recover_p(f, sh, unhex(port, c["alphas"], f), c["x"])
        assert ints(port, out, f) == [int(h, 16) for h in c["out"]]


def test_recover_d_golden(ctx, port, golden):
    for c in golden["recover_d"]:
        f = c["field"]
        sh = unhex(port, c["shares"], f, (c["N"], c["n"]))
        out, err, rc = ctx.recover_d(f, sh, c["t"])
        assert rc == c["rc"] and [int(e) for e in err] == c["err"], (f, c["t"])
        assert ints(port, out, f) == [int(h, 16) for h in c["out"]]
    c = golden["recover_d_not_enough"]
    sh = port.shamir_share(61, port.from_ints([5], 61), c["t"], c["n"], "few")
    assert ctx.recover_d(61, sh, c["t"])[2] == -1
    c = golden["recover_d_custom"]
    sh = unhex(port, c["shares"], 61, (2, c["n"]))
    out, err, rc = ctx.recover_d(61, sh, c["t"], alphas=unhex(port, c["alphas"], 61), d=c["d"], x=c["x"])
    assert rc == c["rc"] and ints(port, out, 61) == [int(h, 16) for h in c["out"]]


@pytest.mark.parametrize("field,t,n,N", [(127, 7, 16, 4096), (61, 7, 16, 4096), (61, 15, 32, 2048), (127, 2, 5, 999)])
def test_recover_d_tamper_vs_oracle(ctx, orc, field, t, n, N):
    """C3's tamper set: flip one share at idx in {0, 8, 2t-1, 2t, 2t+1}; the last two
    are never checked by the reference (shamir.h:129) and must stay undetected."""
    secrets = orc.vector_random(field, "secrets127", 0, N)
    sh = orc.shamir_share(field, secrets, t, n, "shamir bench", 0)
    flat = sh.reshape(N, n, -1)
    idxs = [0, min(8, n - 1), 2 * t - 1, 2 * t, min(2 * t + 1, n - 1)]
    for k, j in enumerate(range(0, N, 16)):
        flat[j, idxs[k % len(idxs)], k % flat.shape[2]] ^= np.uint64(1 << (k % 60))
    out, err, nd = ctx.recover_d(field, sh, t)
    w_out, w_err, w_nd = orc.recover_d(field, sh, t)
    assert nd == w_nd and nd > 0
    assert np.array_equal(err, w_err)
    assert np.array_equal(out, w_out)


def test_recover_d_raises_reference_strings(ctx, pkg, port):
    lib = ctx.lib
    sh = port.shamir_share(61, port.from_ints([5], 61), 3, 5, "few")
    import ctypes as C
    out = np.zeros(1, dtype=np.uint64)
    err = np.zeros(1, dtype=np.uint8)
    nd = C.c_uint64()
    rc = lib.sclgpu_fp61_recover_d(ctx._ctx, sh.ctypes.data_as(C.c_void_p), 1, 5, 3, None, 0, 3, None,
                                   out.ctypes.data_as(C.c_void_p), err.ctypes.data_as(C.c_void_p), C.byref(nd))
    assert rc == pkg.binding.ELOGIC
    assert lib.sclgpu_last_error(ctx._ctx) == b"not enough shares provided to detect errors"
    sh = port.shamir_share(61, port.from_ints([5], 61), 3, 7, "ok")
    sh[0, 2] = 4
    rc = lib.sclgpu_fp61_recover_d(ctx._ctx, sh.ctypes.data_as(C.c_void_p), 1, 7, 3, None, 0, 3, None,
                                   out.ctypes.data_as(C.c_void_p), err.ctypes.data_as(C.c_void_p), C.byref(nd))
    assert rc == pkg.binding.EDETECT and nd.value == 1 and err[0] == 1
    assert lib.sclgpu_last_error(ctx._ctx) == b"error detected during recovery"


def test_lagrange_golden_and_collision(ctx, pkg, port, golden):
    for c in golden["lagrange"]:
        f = c["field"]
        lb = ctx.lagrange(f, port.from_ints(c["nodes"], f), int(c["x"], 16))
        assert ints(port, lb, f) == [int(h, 16) for h in c["hex"]]
    with pytest.raises(pkg.LogicError, match="0 not invertible modulo prime"):
        ctx.lagrange(61, port.from_ints([1, 2, 2], 61), 0)


# ------------------------------------------------------------------ Vector / Matrix
def test_vec_golden(ctx, port, golden):
    for c in golden["vec"]:
        f = c["field"]
        a, b = port.vector_random(f, "a", 0, 37), port.vector_random(f, "b", 0, 37)
        if c["op"] == "beaver":
            e, d, cc = (port.vector_random(f, s, 0, 37) for s in ("e", "d", "c"))
            got = ctx.beaver(f, e, b, d, a, cc)
        else:
            got = ctx.vec_op(f, c["op"], a, b)
        assert ints(port, got, f) == [int(h, 16) for h in c["hex"]], c["op"]


@pytest.mark.parametrize("field", [61, 127])
@pytest.mark.parametrize("n", [1, 2, 31, 1000, 1001, (1 << 18) + 3])
def test_vec_ops_vs_oracle(ctx, orc, field, n):
    a, b = orc.vector_random(field, "va", 0, n), orc.vector_random(field, "vb", 9, n)
    for op in range(6):
        assert np.array_equal(ctx.vec_op(field, op, a, b), orc.vec_op(field, op, a, b)), op
    e, d, c = (orc.vector_random(field, s, 0, n) for s in ("e", "d", "c"))
    assert np.array_equal(ctx.beaver(field, e, b, d, a, c), orc.beaver(field, e, b, d, a, c))


@pytest.mark.parametrize("field", [61, 127])
def test_vec_equal(ctx, port, field):
    """Vector::equals (vector.h:358-375, test_vector.cc): equal, one element off (first / last / middle), sizes."""
    for n in (1, 2, 1000, (1 << 18) + 5):
        a = port.vector_random(field, "eq", 0, n)
        assert ctx.vec_equal(field, a, a.copy())
        for pos in {0, n - 1, n // 2}:
            b = a.copy()
            b.reshape(n, -1)[pos, -1] ^= np.uint64(1 << 40)
            assert not ctx.vec_equal(field, a, b)
    assert not ctx.vec_equal(field, port.vector_random(field, "eq", 0, 4), port.vector_random(field, "eq", 0, 5))
    e = port.from_ints([], field)
    assert ctx.vec_equal(field, e, e)


@pytest.mark.parametrize("field", [61, 127])
def test_field_edge_values(ctx, port, field):
    p = P[field]
    vals = [0, 1, 2, p - 1, p - 2, (p + 1) // 2, (1 << 60) + 5, p // 3, (1 << 32) - 1, 1 << 32]
    if field == 127:
        vals += [(1 << 64) - 1, 1 << 64, (1 << 126) + (1 << 64) - 1, (1 << 63), (1 << 127) - (1 << 64)]
    A = port.from_ints([a for a in vals for _ in vals], field)
    B = port.from_ints([b for _ in vals for b in vals], field)
    for op in (0, 1, 2, 4):
        assert np.array_equal(ctx.vec_op(field, op, A, B), port.vec_op(field, op, A, B)), op
    assert np.array_equal(ctx.vec_op(field, 5, A), port.vec_op(field, 5, A))


def test_matvec_golden_and_vs_oracle(ctx, orc, port, golden):
    for c in golden["matvec"]:
        f, rows, cols = c["field"], c["rows"], c["cols"]
        A = port.vector_random(f, "mat A", 0, rows * cols).reshape((rows, cols) + (() if f == 61 else (2,)))
        x = port.vector_random(f, "vec x", 0, cols)
        assert ints(port, ctx.matvec(f, A, x), f) == [int(h, 16) for h in c["hex"]]
    for field, rows, cols in [(61, 300, 1024), (61, 7, 4097), (61, 1, 1), (127, 64, 513), (61, 128, 8192), (61, 33, 258),
                              (61, 5, 1000), (61, 2049, 256)]:
        A = orc.vector_random(field, "mat A", 0, rows * cols).reshape((rows, cols) + (() if field == 61 else (2,)))
        x = orc.vector_random(field, "vec x", 0, cols)
        assert np.array_equal(ctx.matvec(field, A, x), orc.matvec(field, A, x)), (field, rows, cols)


@pytest.mark.parametrize("field,rows,inner,cols", [
    (61, 128, 16, 32), (61, 1, 2, 1), (61, 300, 520, 70), (61, 129, 4100, 33), (61, 64, 8200, 40), (61, 257, 130, 257),
    (61, 5, 7, 3), (61, 200, 333, 100), (127, 40, 50, 30), (127, 3, 1, 2), (127, 130, 520, 40), (127, 200, 2100, 20),
    (127, 257, 9, 33), (127, 64, 64, 16), (127, 129, 31, 17)])
def test_matmul_vs_oracle(ctx, pkg, port, field, rows, inner, cols):
    """Matrix::multiply(Matrix) (matrix.h:476-495).  Fp61 with even inner dimension and Fp127 run on the tensor
    cores (inner > 4096 / 2048: more than one accumulation round; ragged tiles in every dimension), tiny or
    odd-inner Fp61 products on the integer pipe; edge residues 0, 1, p-1 planted in both operands."""
    sh = () if field == 61 else (2,)
    A = port.vector_random(field, "mat A", 0, rows * inner).reshape((rows, inner) + sh).copy()
    Bm = port.vector_random(field, "mat B", 7, inner * cols).reshape((inner, cols) + sh).copy()
    pm1 = port.from_ints([P[field] - 1, 0, 1], field)
    A[0, 0], Bm[0, 0] = pm1[0], pm1[0]
    A[rows - 1, inner - 1], Bm[inner - 1, cols - 1] = pm1[0], pm1[2]
    A[rows // 2, inner // 2] = pm1[1]
    got = ctx.matmul(field, A, Bm)
    assert np.array_equal(got, port.matmul(field, A, Bm)), (field, rows, inner, cols)


def test_matmul_all_max_residues(ctx, port):
    """every product is (p-1)^2 = 1: the limb accumulators run at their maximum (255 * 255 per byte pair)"""
    rows, inner, cols = 128, 4096, 32
    A = np.full((rows, inner), P[61] - 1, dtype=np.uint64)
    Bm = np.full((inner, cols), P[61] - 1, dtype=np.uint64)
    got = ctx.matmul(61, A, Bm)
    assert np.all(got == np.uint64(inner % P[61]))


def test_matmul_all_max_residues_fp127(ctx, port):
    rows, inner, cols = 128, 2048, 16
    pm1 = port.from_ints([P[127] - 1], 127)[0]
    A = np.broadcast_to(pm1, (rows, inner, 2)).copy()
    Bm = np.broadcast_to(pm1, (inner, cols, 2)).copy()
    got = ctx.matmul(127, A, Bm)
    assert ints(port, got, 127) == [inner] * (rows * cols)


def test_matmul_errors_and_golden_identity(ctx, pkg, port):
    with pytest.raises(pkg.InvalidArgument, match="this->cols\\(\\) != that->rows\\(\\)"):
        ctx.matmul(61, np.zeros((2, 3), dtype=np.uint64), np.zeros((4, 2), dtype=np.uint64))
    # test_matrix.cc:342-365: Vandermonde * coefficients = polynomial evaluations = the shares
    V = ctx.vandermonde(61, 32, 16)                                   # 32 x 16
    coeffs = port.vector_random(61, "coeffs", 0, 16 * 40).reshape(16, 40)
    sh = ctx.matmul(61, V, coeffs)                                   # 32 x 40: column j = shares of polynomial j
    assert np.array_equal(sh, port.matmul(61, V, coeffs))


def test_matvec_and_vandermonde_errors(ctx, pkg, port, golden):
    with pytest.raises(pkg.InvalidArgument, match="n or m cannot be 0"):
        ctx.vandermonde(61, 0, 3)
    for c in golden["vandermonde"]:
        assert ints(port, ctx.vandermonde(c["field"], c["n"], c["m"]), c["field"]) == [int(h, 16) for h in c["hex"]]


# ------------------------------------------------------------------ device-pointer path
def test_dev_path_party_major_and_secret_major(ctx, pkg, orc):
    import torch

    ctx.use_torch_stream()
    B = pkg.binding
    for field, t, n, N in [(61, 15, 32, 5000), (127, 7, 16, 3000), (61, 2, 5, 777)]:
        w = 1 if field == 61 else 2
        secrets = orc.vector_random(field, "secrets", 0, N)
        want = orc.shamir_share(field, secrets, t, n, "shamir bench", 0)
        d_sec = torch.from_numpy(secrets.view(np.int64)).cuda()
        d_pm = torch.empty((n, N, w), dtype=torch.int64, device="cuda")
        d_sm = torch.empty((N, n, w), dtype=torch.int64, device="cuda")
        ctx.shamir_share_dev(field, d_sec, N, t, n, "shamir bench", 0, d_pm, B.PARTY_MAJOR)
        ctx.shamir_share_dev(field, d_sec, N, t, n, "shamir bench", 0, d_sm, B.SECRET_MAJOR)
        torch.cuda.synchronize()
        sm = d_sm.cpu().numpy().view(np.uint64).reshape(want.shape)
        pm = d_pm.cpu().numpy().view(np.uint64)
        assert np.array_equal(sm, want)
        assert np.array_equal(np.swapaxes(pm, 0, 1).reshape(want.shape), want)
        for layout, buf in ((B.PARTY_MAJOR, d_pm), (B.SECRET_MAJOR, d_sm)):
            d_out = torch.empty((N, w), dtype=torch.int64, device="cuda")
            ctx.recover_p_dev(field, buf, N, n, d_out, layout)
            torch.cuda.synchronize()
            assert np.array_equal(d_out.cpu().numpy().view(np.uint64).reshape(secrets.shape), secrets)
            if n >= 2 * t + 1:
                d_err = torch.empty(N, dtype=torch.uint8, device="cuda")
                nd = ctx.recover_d_dev(field, buf, N, n, t, d_out, d_err, layout)
                assert nd == 0 and int(d_err.sum()) == 0
                assert np.array_equal(d_out.cpu().numpy().view(np.uint64).reshape(secrets.shape), secrets)


@pytest.mark.parametrize("n,N", [(1, 64), (3, 1000), (8, 4096), (9, 4097), (32, 1 << 16), (33, 2050), (100, 513), (2048, 96)])
def test_recover_p_party_major_kernel_vs_oracle(ctx, pkg, orc, n, N):
    """k_recover61_pm (party-major planes, limb-split Lagrange coefficients) on ARBITRARY
    share words, i.e. the full linear map and not only consistent sharings, against
    shamirRecoverP of the oracle; default nodes and custom (alphas, x)."""
    import torch

    ctx.use_torch_stream()
    B = pkg.binding
    raw = orc.vector_random(61, "recover pm", 7, N * n)          # canonical residues
    raw[:3] = [0, (1 << 61) - 2, 1]                                # edge residues first
    sm = raw.reshape(N, n)                                        # SCL's [N][n]
    d_pm = torch.from_numpy(np.ascontiguousarray(sm.T).view(np.int64)).cuda()
    d_out = torch.empty(N, dtype=torch.int64, device="cuda")
    ctx.recover_p_dev(61, d_pm, N, n, d_out, B.PARTY_MAJOR)
    torch.cuda.synchronize()
    K = min(N, 2048 if n <= 100 else 4)                           # the real oracle recomputes its basis per secret
    assert np.array_equal(d_out[:K].cpu().numpy().view(np.uint64), orc.recover_p(61, sm[:K]))
    if n <= 100:
        alphas = orc.from_ints([3 * i + 2 for i in range(n)], 61)
        ctx.recover_p_dev(61, d_pm, N, n, d_out, B.PARTY_MAJOR, alphas=alphas, x=5)
        torch.cuda.synchronize()
        assert np.array_equal(d_out[:K].cpu().numpy().view(np.uint64), orc.recover_p(61, sm[:K], alphas, 5))
    # the generic kernel (secret-major input) must agree everywhere
    d_sm = torch.from_numpy(sm.view(np.int64)).cuda()
    d_out2 = torch.empty(N, dtype=torch.int64, device="cuda")
    ctx.recover_p_dev(61, d_pm, N, n, d_out, B.PARTY_MAJOR)
    ctx.recover_p_dev(61, d_sm, N, n, d_out2, B.SECRET_MAJOR)
    torch.cuda.synchronize()
    assert torch.equal(d_out, d_out2)


def test_sharded_gpu_engine(ctx, pkg, orc):
    """The N>1 driver logic with the GPU engine: slices shared with the offset counter
    concatenate to the one-PRG batch (what tests/dist_worker.py checks under gloo)."""
    sh = pkg.sharding
    field, t, n, N = 61, 15, 32, 4099
    secrets = orc.vector_random(field, "secrets", 0, N)
    full = ctx.shamir_share(field, secrets, t, n, "shamir bench", 3)
    assert np.array_equal(full, orc.shamir_share(field, secrets, t, n, "shamir bench", 3))
    parts = []
    for r in range(8):
        s = sh.shard_range(N, 8, r)
        parts.append(ctx.shamir_share(field, secrets[s.lo:s.hi], t, n, "shamir bench", sh.share_first_block(field, t, 3, s)))
    assert np.array_equal(np.concatenate(parts, axis=0), full)


# ------------------------------------------------------------------ full-size properties
def test_full_size_c2_properties(ctx, pkg, orc):
    """BASELINE config C2 at its full size (2^26 secrets, n=32, t=15: 16 GiB of shares per pass):
    (i) share -> recoverP round trip returns every secret, (ii) a prefix and strided
    samples equal the oracle bit for bit, (iii) linearity: share(a)+share(b) over the
    same coefficients... is checked through sum of shares == share of sums at x (party) level
    via recoverP(shares_a + shares_b) == a + b."""
    import torch

    ctx.use_torch_stream()
    B = pkg.binding
    field, t, n, N = 61, 15, 32, 1 << 26
    d_sec = torch.empty(N, dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "secrets", 0, N, d_sec)
    d_sh = torch.empty((n, N), dtype=torch.int64, device="cuda")
    ctx.shamir_share_dev(61, d_sec, N, t, n, "shamir bench", 0, d_sh, B.PARTY_MAJOR)
    d_out = torch.empty(N, dtype=torch.int64, device="cuda")
    ctx.recover_p_dev(61, d_sh, N, n, d_out, B.PARTY_MAJOR)
    torch.cuda.synchronize()
    assert torch.equal(d_out, d_sec)
    # prefix + strided samples against the oracle
    K = 2048
    sec_h = d_sec[:K].cpu().numpy().view(np.uint64)
    assert np.array_equal(sec_h, orc.vector_random(61, "secrets", 0, K))
    want = orc.shamir_share(61, sec_h, t, n, "shamir bench", 0)
    assert np.array_equal(d_sh[:, :K].t().contiguous().cpu().numpy().view(np.uint64), want)
    if orc.kind == "port":  # seekable oracle: check far-away slices too
        for lo in (N // 2 - 5, N - K):
            sec_h = d_sec[lo:lo + K].cpu().numpy().view(np.uint64)
            want = orc.shamir_share(61, sec_h, t, n, "shamir bench", lo * 8)
            assert np.array_equal(d_sh[:, lo:lo + K].t().contiguous().cpu().numpy().view(np.uint64), want)
    # linearity of reconstruction: recoverP(sh_a + sh_b) = a + b
    d_sh2 = torch.empty((n, N), dtype=torch.int64, device="cuda")
    ctx.shamir_share_dev(61, d_out, N, t, n, "other", 0, d_sh2, B.PARTY_MAJOR)
    ctx.vec_op_dev(61, 0, d_sh, d_sh2, n * N, d_sh2)
    d_sum = torch.empty(N, dtype=torch.int64, device="cuda")
    ctx.recover_p_dev(61, d_sh2, N, n, d_sum, B.PARTY_MAJOR)
    d_want = torch.empty(N, dtype=torch.int64, device="cuda")
    ctx.vec_op_dev(61, 0, d_sec, d_sec, N, d_want)
    torch.cuda.synchronize()
    assert torch.equal(d_sum, d_want)


def test_full_size_array_sharing_properties(ctx, pkg, port):
    """Pedersen's sharing step at C2's volume: 2^25 pairs (math::Array<Fp61, 2>), n=32, t=15 = 2^26 component
    polynomials, 16 GiB of shares.  (i) share -> recoverP returns every pair in both layouts, (ii) a prefix and two
    far-away slices equal the (seekable) oracle bit for bit, (iii) party-major and secret-major hold the same shares."""
    import torch

    ctx.use_torch_stream()
    B = pkg.binding
    W, t, n, N = 2, 15, 32, 1 << 25
    blocks = pkg.api.blocks_per_array_share_call(61, W, t)
    d_sec = torch.empty((N, W), dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "pairs", 0, N * W, d_sec)
    d_pm = torch.empty((n, N, W), dtype=torch.int64, device="cuda")
    d_out = torch.empty((N, W), dtype=torch.int64, device="cuda")
    ctx.shamir_share_array_dev(61, d_sec, N, W, t, n, "pedersen", 0, d_pm, B.PARTY_MAJOR)
    ctx.recover_p_array_dev(61, d_pm, N, W, n, d_out, B.PARTY_MAJOR)
    torch.cuda.synchronize()
    assert torch.equal(d_out, d_sec)
    K = 1024
    for lo in (0, N // 2 - 3, N - K):
        sec_h = d_sec[lo:lo + K].cpu().numpy().view(np.uint64)
        want = port.shamir_share_array(61, sec_h, t, n, "pedersen", lo * blocks)          # [K, n, W]
        got = d_pm[:, lo:lo + K, :].permute(1, 0, 2).contiguous().cpu().numpy().view(np.uint64)
        assert np.array_equal(got, want), lo
    # secret-major = SCL's own layout: same shares, and it reconstructs
    Nh = N // 2                                                                             # 8 GiB more
    d_sm = torch.empty((Nh, n, W), dtype=torch.int64, device="cuda")
    ctx.shamir_share_array_dev(61, d_sec, Nh, W, t, n, "pedersen", 0, d_sm, B.SECRET_MAJOR)
    d_out.zero_()
    ctx.recover_p_array_dev(61, d_sm, Nh, W, n, d_out, B.SECRET_MAJOR)
    torch.cuda.synchronize()
    assert torch.equal(d_out[:Nh], d_sec[:Nh])
    for lo in (0, Nh - 4096):
        assert torch.equal(d_sm[lo:lo + 4096], d_pm[:, lo:lo + 4096, :].permute(1, 0, 2))


# ------------------------------------------------------------------ shamirRecoverC
def test_recover_c_golden(ctx, port, golden):
    for c in golden["recover_c"]:
        f, n, N = c["field"], c["n"], c["N"]
        sh = unhex(port, c["shares"], f, (N, n))
        pf, pe, st, nf = ctx.recover_c(f, sh)
        assert [int(v) for v in st] == c["status"] and nf == c["n_failed"]
        assert ints(port, pf, f) == [int(h, 16) for h in c["f"]]
        assert ints(port, pe, f) == [int(h, 16) for h in c["err"]]


@pytest.mark.parametrize("field,n,N", [(61, 1, 50), (61, 4, 3000), (61, 7, 2000), (61, 16, 4097), (61, 22, 600), (61, 31, 300),
                                       (61, 33, 200), (127, 4, 1000), (127, 16, 1500), (127, 31, 150)])
def test_recover_c_vs_oracle(ctx, pkg, orc, field, n, N):
    """0..t+1 corrupted shares per sharing; f, err, status and the count against shamirRecoverC of the
    oracle; default and custom nodes; device-pointer path in both layouts."""
    import torch

    rng = np.random.default_rng(n)
    t = (n - 1) // 3
    sec = orc.vector_random(field, "secrets", 0, N)
    sh = orc.shamir_share(field, sec, t, n, "rc", 3).copy()
    flat = sh.reshape(N, n, -1)
    for j in range(N):
        k = j % (t + 2)
        for i in (rng.choice(3 * t + 1, size=min(k, 3 * t + 1), replace=False) if k else []):
            flat[j, i, rng.integers(flat.shape[2])] ^= np.uint64(1 + j)
    K = N if orc.kind == "port" or n <= 16 else min(N, 64)
    want = orc.recover_c(field, sh[:K])
    got = ctx.recover_c(field, sh)
    for g, w, name in zip(got[:3], want[:3], ("f", "err", "status")):
        assert np.array_equal(g[:K], w), (field, n, name)
    ok = np.array([j % (t + 2) <= t for j in range(N)])   # at most t corrupted shares: must be corrected
    assert not got[2][ok].any()                            # (beyond the radius the outcome depends on the data)
    f0 = got[0][:, 0]
    assert np.array_equal(f0[ok], sec[ok])
    assert got[3] == int((got[2] != 0).sum())
    alphas = orc.from_ints([7 * i + 2 for i in range(n)], field)
    w2, g2 = orc.recover_c(field, sh[:min(K, 256)], alphas), ctx.recover_c(field, sh[:min(K, 256)], alphas)
    assert all(np.array_equal(a, b) for a, b in zip(g2[:3], w2[:3])) and g2[3] == w2[3]
    # device pointers
    ctx.use_torch_stream()
    w = 1 if field == 61 else 2
    d_sm = torch.from_numpy(sh.view(np.int64)).cuda()
    d_pm = torch.from_numpy(np.ascontiguousarray(np.swapaxes(sh.reshape(N, n, w), 0, 1)).view(np.int64)).cuda()
    for layout, buf in ((pkg.binding.SECRET_MAJOR, d_sm), (pkg.binding.PARTY_MAJOR, d_pm)):
        d_f = torch.empty((N, 3 * t + 1, w), dtype=torch.int64, device="cuda")
        d_e = torch.empty((N, t + 1, w), dtype=torch.int64, device="cuda")
        d_st = torch.empty(N, dtype=torch.uint8, device="cuda")
        nf = ctx.recover_c_dev(field, buf, N, n, d_f, d_e, d_st, layout)
        assert nf == got[3]
        assert np.array_equal(d_f.cpu().numpy().view(np.uint64).reshape(got[0].shape), got[0])
        assert np.array_equal(d_e.cpu().numpy().view(np.uint64).reshape(got[1].shape), got[1])
        assert np.array_equal(d_st.cpu().numpy(), got[2])


@pytest.mark.parametrize("field,n,N", [(61, 34, 160), (61, 40, 130), (61, 64, 96), (61, 100, 48), (61, 166, 12), (127, 34, 80),
                                       (127, 64, 40), (127, 118, 8)])
def test_recover_c_more_than_32_points(ctx, pkg, orc, port, field, n, N):
    """3t+1 > 32: one CTA per sharing (k_recover_c_cta), error-free sharings settled by k_recover_c_clean_any.  The
    reference takes any size (shamir.h:203-246, matrix.h:598-828); f, err, status and the count against its
    shamirRecoverC with 0..t+1 corrupted shares, default and custom nodes, both layouts on the device."""
    import torch

    rng = np.random.default_rng(n)
    t = (n - 1) // 3
    sec = port.vector_random(field, "secrets", 0, N)
    sh = port.shamir_share(field, sec, t, n, "rc", 3).copy()
    flat = sh.reshape(N, n, -1)
    nerr = [0, 1, t, t + 1, t // 2, 2]
    for j in range(N):
        k = nerr[j % len(nerr)]
        for i in (rng.choice(3 * t + 1, size=min(k, 3 * t + 1), replace=False) if k else []):
            flat[j, i, rng.integers(flat.shape[2])] ^= np.uint64(1 + j)
    K = min(N, 24)                                           # the unmodified reference on a prefix, the port on everything
    want_ref = orc.recover_c(field, sh[:K])
    want = port.recover_c(field, sh)
    got = ctx.recover_c(field, sh)
    for g, w, r, name in zip(got[:3], want[:3], want_ref[:3], ("f", "err", "status")):
        assert np.array_equal(g, w), (field, n, name)
        assert np.array_equal(g[:K], r), (field, n, name, "reference")
    assert got[3] == want[3] == int((got[2] != 0).sum())
    ok = np.array([nerr[j % len(nerr)] <= t for j in range(N)])
    assert not got[2][ok].any() and np.array_equal(got[0][:, 0][ok], sec[ok])
    alphas = port.from_ints([5 * i + 3 for i in range(n)], field)
    w2, g2 = port.recover_c(field, sh[:16], alphas), ctx.recover_c(field, sh[:16], alphas)
    assert all(np.array_equal(a, b) for a, b in zip(g2[:3], w2[:3])) and g2[3] == w2[3]
    ctx.use_torch_stream()
    w = 1 if field == 61 else 2
    d_pm = torch.from_numpy(np.ascontiguousarray(np.swapaxes(sh.reshape(N, n, w), 0, 1)).view(np.int64)).cuda()
    d_f = torch.empty((N, 3 * t + 1, w), dtype=torch.int64, device="cuda")
    d_e = torch.empty((N, t + 1, w), dtype=torch.int64, device="cuda")
    d_st = torch.empty(N, dtype=torch.uint8, device="cuda")
    assert ctx.recover_c_dev(field, d_pm, N, n, d_f, d_e, d_st, pkg.binding.PARTY_MAJOR) == got[3]
    assert np.array_equal(d_f.cpu().numpy().view(np.uint64).reshape(got[0].shape), got[0])
    assert np.array_equal(d_st.cpu().numpy(), got[2])


def test_recover_c_errors(ctx, pkg, port):
    with pytest.raises(pkg.InvalidArgument):
        ctx.recover_c(61, port.from_ints(list(range(400)), 61).reshape(1, 400))   # the system no longer fits shared memory


# ------------------------------------------------------------------ Polynomial::evaluate from coefficient planes
@pytest.mark.parametrize("field,t,n,N", [(61, 15, 32, 5000), (61, 2, 5, 129), (61, 0, 3, 7), (61, 9, 20, 1 << 16),
                                         (127, 7, 16, 3001), (127, 3, 9, 640), (61, 20, 40, 300), (127, 9, 20, 100)])
def test_share_from_coefficient_planes(ctx, pkg, orc, field, t, n, N):
    """shamir_share_coeffs_dev (poly.h:56-64): coefficients supplied by the caller as [t+1][N] planes
    (arbitrary residues, plane 0 = the secrets).  Oracle: the same polynomial written as a Vandermonde
    product, shares[j] = V(n, t+1) * c_j (the reference's test_matrix.cc:342-365 identity), through
    Matrix::multiply(Vector) of the oracle."""
    import torch

    ctx.use_torch_stream()
    w = 1 if field == 61 else 2
    coeffs = orc.vector_random(field, "coeff planes", 3, (t + 1) * N).reshape((t + 1, N) + (() if w == 1 else (2,)))
    d_c = torch.from_numpy(coeffs.view(np.int64)).cuda()
    d_pm = torch.empty((n, N, w), dtype=torch.int64, device="cuda")
    d_sm = torch.empty((N, n, w), dtype=torch.int64, device="cuda")
    ctx.shamir_share_coeffs_dev(field, d_c, N, t, n, d_pm, pkg.binding.PARTY_MAJOR)
    ctx.shamir_share_coeffs_dev(field, d_c, N, t, n, d_sm, pkg.binding.SECRET_MAJOR)
    torch.cuda.synchronize()
    pm = d_pm.cpu().numpy().view(np.uint64)
    sm = d_sm.cpu().numpy().view(np.uint64)
    assert np.array_equal(np.swapaxes(pm, 0, 1), sm)
    V = orc.vandermonde(field, n, t + 1)
    K = min(N, 64)
    for j in list(range(K)) + [N - 1]:
        cj = np.ascontiguousarray(coeffs[:, j])
        assert np.array_equal(sm[j].reshape(V.shape[:1] + V.shape[2:]), orc.matvec(field, V, cj)), (field, j)


@pytest.mark.parametrize("field,t,n", [(61, 15, 32), (61, 1, 5), (127, 7, 16), (127, 1, 4)])
def test_share_limb_recombination_edges(ctx, pkg, port, field, t, n):
    """Crafted coefficient planes that drive the tensor-core epilogue through its corner cases: every
    coefficient p-1 (largest limb sums), c0 = p-1 and c1 = 1 (the share at x = 1 is exactly p before
    canonicalisation and must come out as 0), all zero, and single-bit coefficients."""
    import torch

    ctx.use_torch_stream()
    w = 1 if field == 61 else 2
    p = P[field]
    cols = []
    cols.append([p - 1] * (t + 1))
    cols.append([p - 1, 1] + [0] * (t - 1))
    cols.append([0] * (t + 1))
    cols.append([1] * (t + 1))
    for b in (0, 7, 8, 60, field - 1):
        cols.append([(1 << b) % p] * (t + 1))
        cols.append([(p - (1 << b)) % p] + [(1 << b) % p] * t)
    N = len(cols)
    planes = port.from_ints([cols[j][k] for k in range(t + 1) for j in range(N)], field).reshape((t + 1, N) + (() if w == 1 else (2,)))
    d_c = torch.from_numpy(planes.view(np.int64)).cuda()
    d_sm = torch.empty((N, n, w), dtype=torch.int64, device="cuda")
    ctx.shamir_share_coeffs_dev(field, d_c, N, t, n, d_sm, pkg.binding.SECRET_MAJOR)
    torch.cuda.synchronize()
    got = d_sm.cpu().numpy().view(np.uint64)
    for j in range(N):
        want = [sum(c * pow(i, k, p) for k, c in enumerate(cols[j])) % p for i in range(1, n + 1)]
        assert ints(port, got[j], field) == want, (field, j)


# ------------------------------------------------------------------ per-party packets
@pytest.mark.parametrize("field,t,n,N", [(61, 15, 32, 5000), (61, 2, 5, 1), (61, 3, 7, 4097), (127, 7, 16, 3001),
                                         (61, 20, 40, 300), (61, 1, 3, (1 << 21) + 7)])
def test_share_packets_wire_layout(ctx, pkg, orc, field, t, n, N):
    """packet i == Serializer<Vector<FF>>::write(column i of SCL's share matrix): u32 count, then
    FF::write bytes (vector.h:596-629, serializer.h:160-176); recover_p_packets inverts it."""
    import struct

    secrets = orc.vector_random(field, "secrets", 0, N)
    want = orc.shamir_share(field, secrets, t, n, "packets", 9)          # [N][n](,2)
    packets = ctx.shamir_share_packets(field, secrets, t, n, "packets", 9)
    assert len(packets) == n
    for i, p in enumerate(packets):
        col = np.ascontiguousarray(want[:, i])
        assert bytes(p[:4]) == struct.pack("<I", N)
        assert p[4:].tobytes() == col.tobytes(), (field, i)
    rec = ctx.recover_p_packets(field, packets, N)
    assert np.array_equal(rec, secrets)
    bad = [p.copy() for p in packets]
    bad[n - 1][:4] = np.frombuffer(struct.pack("<I", N + 1), dtype=np.uint8)
    with pytest.raises(pkg.InvalidArgument, match="Vec sizes mismatch"):
        ctx.recover_p_packets(field, bad, N)


@pytest.mark.parametrize("field,t,n,N", [(61, 15, 32, 4098), (61, 2, 5, 1001), (127, 7, 16, 2050), (127, 2, 20, 515), (61, 3, 2100, 40)])
def test_recover_p_packets_non_canonical_words(ctx, orc, field, t, n, N):
    """Wire packets come from OTHER parties: any byte string is a legal element, and SCL canonicalises on receive
    (Serializer<Vector<FF>>::read -> FF::read -> `% p`; vector.h:623-626, ff.h:63-67, mersenne61.cc:87-90).  Words in
    [p, 2^64) / [p, 2^128) must therefore reconstruct to what the reference reconstructs from the same bytes:
    FF::read on every word (oracle from_bytes) followed by shamirRecoverP -- on the plane kernel, the tensor-core
    kernel (Fp127, n <= 16) and the generic kernel (Fp127 n > 16, Fp61 n > 2048)."""
    import struct
    rng = np.random.default_rng(field * 1000 + n)
    w = 1 if field == 61 else 2
    secrets = orc.vector_random(field, "secrets", 0, N)
    packets = ctx.shamir_share_packets(field, secrets, min(t, n - 1), n, "wire", 0)
    p_int = (1 << field) - 1
    raw = []
    for i, pk in enumerate(packets):
        words = pk[4:].view(np.uint64).reshape(N, w).copy()
        for j in range(i % 3, N, 3):                      # a third of the words of every packet
            v = int(words[j, 0]) | ((int(words[j, 1]) << 64) if w == 2 else 0)
            k = int(rng.integers(1, 8 if field == 61 else 2))
            nv = v + k * p_int if j % 2 else (1 << (64 * w)) - 1 - int(rng.integers(0, 1 << 20))
            if nv >= 1 << (64 * w):
                nv = v + p_int
            words[j, 0] = nv & 0xFFFFFFFFFFFFFFFF
            if w == 2:
                words[j, 1] = nv >> 64
        raw.append(words)
        pk[4:] = words.reshape(-1).view(np.uint8)
        assert bytes(pk[:4]) == struct.pack("<I", N)
    canon = np.stack([orc.from_bytes(field, r.tobytes()).reshape((N,) + ((2,) if w == 2 else ())) for r in raw], axis=1)
    want = orc.recover_p(field, np.ascontiguousarray(canon))
    got = ctx.recover_p_packets(field, packets, N)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("field,t", [(61, 7), (61, 15), (61, 20), (61, 31), (127, 7), (127, 15)])
def test_reconstruction_extreme_limbs(ctx, port, field, t):
    """Largest byte-limb sums through the tensor-core reconstruction kernels: every share p - 1 (the sharing of the
    constant polynomial p - 1: all bytes 0xff but the top one), and sharings of random secrets with a zero polynomial
    tail scaled to p - 1.  With more than 16 Fp61 shares per row the accumulators reach 24 bits (two K tiles); the
    recombination must still be exact."""
    N, n = 257, 2 * t + 1
    pm1 = (1 << field) - 2
    sh = port.from_ints([pm1] * (N * n), field).reshape((N, n) + (() if field == 61 else (2,)))
    out, err, nd = ctx.recover_d(field, sh, t)
    w_out, w_err, w_nd = port.recover_d(field, sh, t)
    assert nd == w_nd == 0 and np.array_equal(out, w_out) and np.array_equal(err, w_err)
    assert np.array_equal(ctx.recover_p(field, sh), port.recover_p(field, sh))
    assert np.array_equal(ctx.recover_p(field, sh[:, : min(n, 32 if field == 61 else 16)]),
                          port.recover_p(field, sh[:, : min(n, 32 if field == 61 else 16)]))
    # one share of every sharing replaced by 0 / by 1: flagged (or not) exactly as the oracle says
    sh2 = sh.copy()
    sh2.reshape(N, n, -1)[::2, t + 1] = 0
    sh2.reshape(N, n, -1)[1::2, 0, 0] = 1
    out, err, nd = ctx.recover_d(field, sh2, t)
    w_out, w_err, w_nd = port.recover_d(field, sh2, t)
    assert nd == w_nd and np.array_equal(out, w_out) and np.array_equal(err, w_err)


@pytest.mark.parametrize("field", [61, 127])
def test_recover_d_degenerate_thresholds(ctx, pkg, port, field):
    """t = 0: the reference's own size check (n_given >= d + t) lets d + 1 > n_given through and then reads one share
    past the end (shamir.h:125-127); this ABI returns the reference's logic_error instead of reading out of bounds or
    dividing by n_given = 0."""
    s = port.vector_random(field, "secrets", 0, 10)
    empty_rows = pkg.api.empty(field, 10, 0)
    out, err, rc = ctx.recover_d(field, empty_rows, 0)       # n_given = 0, t = 0, d = 0
    assert rc == -1
    sh = port.shamir_share(field, s, 2, 2, "deg", 0)          # two shares per secret
    out, err, rc = ctx.recover_d(field, sh, 0, alphas=port.from_ints([1, 2], field), d=2, x=0)   # needs shares 0..2
    assert rc == -1
    sh = port.shamir_share(field, s, 0, 3, "deg", 0)          # degree 0, t = 0: one share interpolates, no checks
    out, err, rc = ctx.recover_d(field, sh, 0)
    assert rc == 0 and np.array_equal(out, s)


# ------------------------------------------------------------------ array-valued secrets, hyperInvertible
def test_share_array_and_him_golden(ctx, port, golden):
    for c in golden["share_array"]:
        f, W, N, n = c["field"], c["W"], c["N"], c["n"]
        secrets = unhex(port, c["secrets"], f, (N, W))
        sh = ctx.shamir_share_array(f, secrets, c["t"], n, c["seed"], c["first_block"])
        assert ints(port, sh, f) == [int(h, 16) for h in c["shares"]], c
        assert ints(port, ctx.recover_p_array(f, sh), f) == [int(h, 16) for h in c["recover"]]
    for c in golden["hyper_invertible"]:
        f = c["field"]
        assert ints(port, ctx.hyper_invertible(f, c["n"], c["m"]), f) == [int(h, 16) for h in c["him"]]


@pytest.mark.parametrize("field,W,t,n,N", [(61, 2, 15, 32, 1 << 14), (61, 2, 2, 5, 4099), (61, 3, 7, 16, 3001), (61, 1, 4, 9, 777),
                                           (61, 5, 1, 3, 130), (61, 2, 20, 40, 515), (61, 4, 0, 2, 33), (61, 9, 3, 7, 260),
                                           (61, 2, 14, 31, 70001), (61, 6, 1, 32, 1111), (61, 2, 0, 1, 129), (61, 4, 9, 17, 6400),
                                           (61, 2, 7, 16, 1 << 17),
                                           (127, 2, 7, 16, 1 << 12), (127, 3, 2, 5, 1025), (127, 1, 3, 8, 300), (127, 5, 9, 12, 129),
                                           (127, 2, 0, 3, 257), (127, 4, 6, 13, 20000)])
def test_share_array_vs_oracle(ctx, pkg, port, orc, field, W, t, n, N):
    """shamirSecretShare on math::Array<FF, W> + shamirRecoverP, host and device-pointer paths, both layouts."""
    import torch

    o = orc if (orc.kind == "port" or W <= 5) else port      # the reference driver instantiates W = 1..5
    es = () if field == 61 else (2,)
    secrets = port.vector_random(field, "secrets", 0, N * W).reshape((N, W) + es)
    first = (1 << 33) - 77 if o.kind == "port" else 1234
    want = o.shamir_share_array(field, secrets, t, n, "shamir bench", first)
    got = ctx.shamir_share_array(field, secrets, t, n, "shamir bench", first)
    assert np.array_equal(got, want)
    rec = ctx.recover_p_array(field, got)
    assert np.array_equal(rec, o.recover_p_array(field, want))
    if t < n:
        assert np.array_equal(rec, secrets)
    if W == 1:
        assert np.array_equal(got.reshape(-1), ctx.shamir_share(field, secrets.reshape((N,) + es), t, n, "shamir bench", first).reshape(-1))
    assert pkg.binding.load().sclgpu_share_array_blocks(8 if field == 61 else 16, W, t) == pkg.api.blocks_per_array_share_call(field, W, t)
    # device pointers: secret-major [N][n][W] and party-major [n][N][W]
    ctx.use_torch_stream()
    w = 1 if field == 61 else 2
    d_sec = torch.from_numpy(secrets.view(np.int64)).cuda()
    for layout in (pkg.binding.SECRET_MAJOR, pkg.binding.PARTY_MAJOR):
        shape = (N, n, W, w) if layout == pkg.binding.SECRET_MAJOR else (n, N, W, w)
        d_sh = torch.zeros(shape, dtype=torch.int64, device="cuda")
        d_out = torch.zeros((N, W, w), dtype=torch.int64, device="cuda")
        ctx.shamir_share_array_dev(field, d_sec, N, W, t, n, "shamir bench", first, d_sh, layout)
        ctx.recover_p_array_dev(field, d_sh, N, W, n, d_out, layout)
        torch.cuda.synchronize()
        h = d_sh.cpu().numpy().view(np.uint64)
        if layout == pkg.binding.PARTY_MAJOR:
            h = np.swapaxes(h, 0, 1)
        assert np.array_equal(h.reshape(want.shape), want), layout
        assert np.array_equal(d_out.cpu().numpy().view(np.uint64).reshape(rec.shape), rec), layout


@pytest.mark.parametrize("field", [61, 127])
def test_hyper_invertible_vs_oracle(ctx, pkg, orc, field):
    for n, m in ((4, 5), (1, 1), (7, 3), (16, 16), (33, 20), (2, 64)):
        assert np.array_equal(ctx.hyper_invertible(field, n, m), orc.hyper_invertible(field, n, m)), (n, m)
    with pytest.raises(pkg.InvalidArgument):
        ctx.hyper_invertible(field, 0, 3)
    with pytest.raises(pkg.InvalidArgument):
        ctx.hyper_invertible(field, 3, 0)


def test_share_array_errors_and_empty(ctx, pkg, port):
    with pytest.raises(pkg.InvalidArgument):
        ctx.shamir_share_array(61, np.zeros((3, 0), dtype=np.uint64), 1, 3, "x")
    out = ctx.shamir_share_array(61, np.zeros((0, 2), dtype=np.uint64), 1, 3, "x")
    assert out.shape == (0, 3, 2)
    assert ctx.recover_p_array(61, out).shape == (0, 2)


# ------------------------------------------------------------------ pageable host buffers (std::vector-backed SCL containers)
def test_pageable_host_buffers_staged(ctx, pkg, orc):
    """Host entry points on ordinary (pageable) numpy memory large enough for the pinned-ring stager
    (csrc/host_stage.h), against the oracle and against the same calls on pinned buffers."""
    N, t, n = 1 << 19, 3, 9                       # 36 MiB of shares: several ring pieces, ragged tail
    N -= 12345
    secrets = orc.vector_random(61, "secrets", 0, N)
    want = orc.shamir_share(61, secrets, t, n, "pageable", 99)
    got = ctx.shamir_share(61, secrets, t, n, "pageable", 99)          # numpy in, numpy out: pageable both ways
    assert np.array_equal(got, want)
    assert np.array_equal(ctx.recover_p(61, got), secrets)
    out, err, nd = ctx.recover_d(61, got, t)
    assert nd == 0 and np.array_equal(out, secrets) and not err.any()
    pinned = ctx.host_alloc(got.nbytes).view(np.uint64).reshape(got.shape)
    lib = pkg.binding.load()
    assert lib.sclgpu_fp61_shamir_share(ctx._ctx, secrets.ctypes.data, N, t, n, pkg.api.seed16("pageable"), 99, pinned.ctypes.data) == 0
    assert np.array_equal(pinned, want)
    ctx.host_free(pinned)
    # Fp127, packets and the vector ops take the same route
    s127 = orc.vector_random(127, "secrets127", 0, 1 << 18)
    g127 = ctx.shamir_share(127, s127, 2, 5, "pageable", 7)
    assert np.array_equal(g127, orc.shamir_share(127, s127, 2, 5, "pageable", 7))
    assert np.array_equal(ctx.recover_p(127, g127), s127)
    a = orc.vector_random(61, "a", 0, (1 << 21) + 77)
    b = orc.vector_random(61, "b", 0, (1 << 21) + 77)
    assert np.array_equal(ctx.beaver(61, a, b, a, b, a), orc.beaver(61, a, b, a, b, a))
    assert np.array_equal(ctx.random(61, "pageable", 5, (1 << 22) + 3), orc.vector_random(61, "pageable", 5, (1 << 22) + 3))
    assert np.array_equal(ctx.prg_expand("pageable", 11, (40 << 20) + 5), orc.prg_next("pageable", 11, (40 << 20) + 5))


# ------------------------------------------------------------------ additive sharing
def test_additive_golden(ctx, port, golden):
    for c in golden["additive"]:
        f = c["field"]
        secrets = unhex(port, c["secrets"], f)
        sh = ctx.additive_share(f, secrets, c["n"], c["seed"], c["first_block"])
        assert ints(port, sh, f) == [int(h, 16) for h in c["shares"]], c
        assert ints(port, ctx.additive_recover(f, sh), f) == [int(h, 16) for h in c["recover"]]


@pytest.mark.parametrize("field,n,N", [(61, 3, 1 << 16), (61, 1, 100), (61, 2, 4097), (61, 40, 3000), (61, 300, 77),
                                       (127, 3, 1 << 14), (127, 1, 5), (127, 17, 2222)])
def test_additive_vs_oracle(ctx, pkg, orc, field, n, N):
    import torch

    secrets = orc.vector_random(field, "secrets", 0, N)
    first = (1 << 33) - 1000 if orc.kind == "port" else 4321
    got = ctx.additive_share(field, secrets, n, "additive", first)
    want = orc.additive_share(field, secrets, n, "additive", first)
    assert np.array_equal(got, want)
    assert np.array_equal(ctx.additive_recover(field, got), secrets)
    assert np.array_equal(ctx.additive_recover(field, got), orc.additive_recover(field, want))
    # device-pointer path, party-major planes
    ctx.use_torch_stream()
    w = 1 if field == 61 else 2
    d_sec = torch.from_numpy(secrets.view(np.int64)).cuda()
    d_pm = torch.empty((n, N, w), dtype=torch.int64, device="cuda")
    d_out = torch.empty((N, w), dtype=torch.int64, device="cuda")
    ctx.additive_share_dev(field, d_sec, N, n, "additive", first, d_pm, pkg.binding.PARTY_MAJOR)
    ctx.additive_recover_dev(field, d_pm, N, n, d_out, pkg.binding.PARTY_MAJOR)
    torch.cuda.synchronize()
    assert np.array_equal(np.swapaxes(d_pm.cpu().numpy().view(np.uint64), 0, 1).reshape(want.shape), want)
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64).reshape(secrets.shape), secrets)


def test_additive_errors(ctx, pkg, port):
    with pytest.raises(pkg.InvalidArgument):
        ctx.additive_share(61, port.from_ints([1], 61), 0, "x")
    e = port.from_ints([], 61)
    assert ctx.additive_share(61, e, 3, "x").shape[0] == 0


# ------------------------------------------------------------------ both Fp61 share kernels
@pytest.mark.parametrize("tc", ["3", "4", "2", "1", "0"])
def test_share_kernel_paths_vs_oracle(tc):
    """tests/tc_check.py sweeps (t, n, N) on the device-pointer path against the plain-C oracle;
    SCLGPU_SHARE_TC selects the tcgen05 kernels (3 = default: A operand in tensor memory, 5 groups;
    4 = warp-specialised producers / consumers; 2 = 4 groups; 1 = A operand in shared memory) or the
    integer-pipe kernel (0).
    A separate process because the library reads the knob once."""
    import os
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SCLGPU_SHARE_TC=tc)
    r = subprocess.run([sys.executable, os.path.join(repo, "tests", "tc_check.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "TC_CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("knob", ["SCLGPU_MATMUL_V1", "SCLGPU_MATMUL_GENERIC", "SCLGPU_RECOVER_GENERIC", "SCLGPU_SHARE_GENERIC",
                                  "SCLGPU_RECOVER_C_FULL", "SCLGPU_RECOVER_C_NOSYN", "SCLGPU_MATVEC_WARP", "SCLGPU_MATVEC_VARIANT=0",
                                  "SCLGPU_PRG_BITSLICED", "SCLGPU_TRANSPOSE_TILES", "SCLGPU_SHARE_SM_TRANSPOSE", "SCLGPU_SR_WARPS=4", "SCLGPU_SR_WARPS=8", "SCLGPU_SR_WARPS=108",
                                  "SCLGPU_NO_FUSED_STEP", "SCLGPU_HOST_CHUNK_MB=1,SCLGPU_HOST_PIPES=4",
                                  "SCLGPU_HOST_CHUNK_MB=1,SCLGPU_HOST_PIPES=1",
                                  "SCLGPU_NO_KNOB"])
def test_selectable_kernels_vs_oracle(knob):
    """tests/knob_check.py with one kernel-selection knob set: the first GEMM form, the
    integer-pipe GEMM, the integer-pipe reconstruction kernels, the staged share path, Berlekamp-Welch without the
    error-free fast path / without the syndrome decoder, the one-warp-per-row and the first chunked mat-vec;
    the bitsliced keystream kernel, the other forms of the single-launch step and the two-kernel step, the host pipelines with four / one chunk(s) of 1 MiB in flight; SCLGPU_NO_KNOB is
    the same sweep on the defaults."""
    import os
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    for item in knob.split(","):
        name, _, value = item.partition("=")
        env[name] = value or "1"
    r = subprocess.run([sys.executable, os.path.join(repo, "tests", "knob_check.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "KNOB_CHECK PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


# ------------------------------------------------------------------ full-size properties, configs C3 / C4 / C5
def test_full_size_c3_fp127_recover_d_tamper_set(ctx, pkg, orc, port):
    """BASELINE config C3 at its full size: Fp127 n=16 t=7, 2^24 secrets, share -> recoverD with the tamper
    set of SURVEY 8d (one flipped share at idx in {0, 8, 13, 14, 15} for 1/1024 of the secrets): idx 14 and 15
    are never checked by the reference (shamir.h:129) and must stay undetected, the others must be flagged."""
    import torch

    ctx.use_torch_stream()
    B = pkg.binding
    N, n, t = 1 << 24, 16, 7
    d_sec = torch.empty((N, 2), dtype=torch.int64, device="cuda")
    ctx.random_dev(127, "secrets127", 0, N, d_sec)
    d_sh = torch.empty((n, N, 2), dtype=torch.int64, device="cuda")
    ctx.shamir_share_dev(127, d_sec, N, t, n, "shamir bench", 0, d_sh, B.PARTY_MAJOR)
    # prefix + far slice against the oracle
    K = 512
    sec_h = d_sec[:K].cpu().numpy().view(np.uint64)
    want = orc.shamir_share(127, sec_h, t, n, "shamir bench", 0)
    assert np.array_equal(d_sh[:, :K].permute(1, 0, 2).contiguous().cpu().numpy().view(np.uint64), want)
    lo = N - K
    sec_h = d_sec[lo:].cpu().numpy().view(np.uint64)
    want = port.shamir_share(127, sec_h, t, n, "shamir bench", lo * 8)
    assert np.array_equal(d_sh[:, lo:].permute(1, 0, 2).contiguous().cpu().numpy().view(np.uint64), want)
    # tamper: secret j = 1024*k gets share idxs[k % 5] flipped
    idxs = [0, 8, 13, 14, 15]
    js = torch.arange(0, N, 1024, device="cuda")
    which = torch.tensor(idxs, device="cuda")[torch.arange(js.numel(), device="cuda") % 5]
    d_sh[which, js, 0] ^= 1
    d_out = torch.empty((N, 2), dtype=torch.int64, device="cuda")
    d_err = torch.empty(N, dtype=torch.uint8, device="cuda")
    nd = ctx.recover_d_dev(127, d_sh, N, n, t, d_out, d_err, B.PARTY_MAJOR)
    detected = which < 14                                   # idx 0, 8, 13 are read and checked
    assert nd == int(detected.sum())
    flags = torch.zeros(N, dtype=torch.uint8, device="cuda")
    flags[js[detected]] = 1
    assert torch.equal(d_err, flags)
    good = flags == 0
    assert torch.equal(d_out[good], d_sec[good])            # untouched and UNDETECTED-tampered sharings recover
    assert int(d_out[~good].abs().sum()) == 0               # flagged ones are zeroed
    # the same 64 tampered sharings through the oracle
    pick = js[:64].cpu().numpy()
    sm = d_sh[:, js[:64]].permute(1, 0, 2).contiguous().cpu().numpy().view(np.uint64)
    o_out, o_err, o_nd = orc.recover_d(127, sm, t)
    assert np.array_equal(o_err, d_err[js[:64]].cpu().numpy()) and np.array_equal(o_out, d_out[js[:64]].cpu().numpy().view(np.uint64))


def test_full_size_c4_prg_expansion(ctx, pkg, orc, port):
    """C4: Vector<Fp61>::random of 2^28 elements (2 GiB of keystream): prefix against the real oracle, far
    slices against the seekable port, and the raw-keystream kernel against the element kernel on every element
    (from_bytes of the former must equal the latter)."""
    import torch

    ctx.use_torch_stream()
    n = 1 << 28
    d_el = torch.empty(n, dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "prg bench", 0, n, d_el)
    K = 1 << 16
    assert np.array_equal(d_el[:K].cpu().numpy().view(np.uint64), orc.vector_random(61, "prg bench", 0, K))
    for lo in (n // 2 - 6, n - K):
        lo -= lo % 2
        assert np.array_equal(d_el[lo:lo + K].cpu().numpy().view(np.uint64), port.vector_random(61, "prg bench", lo // 2, K))
    d_raw = torch.empty(n, dtype=torch.int64, device="cuda")
    ctx.prg_expand_dev("prg bench", 0, 8 * n, d_raw)
    d_chk = torch.empty(n, dtype=torch.int64, device="cuda")
    ctx.lib.sclgpu_fp61_from_bytes_dev(ctx._ctx, d_raw.data_ptr(), n, d_chk.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(d_chk, d_el)


def test_full_size_c5_matvec_and_muladd(ctx, pkg, port):
    """C5: Fp61 mat-vec 8192 x 8192 against the oracle on every row, and the Beaver combination on 2^26
    elements against the oracle on a prefix / far slice plus an algebraic identity on everything:
    z - c = e*(b+d) + d*a."""
    import torch

    ctx.use_torch_stream()
    rows = cols = 8192
    d_A = torch.empty((rows, cols), dtype=torch.int64, device="cuda")
    d_x = torch.empty(cols, dtype=torch.int64, device="cuda")
    d_y = torch.empty(rows, dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "mat A", 0, rows * cols, d_A)
    ctx.random_dev(61, "vec x", 0, cols, d_x)
    ctx.matvec_dev(61, d_A, rows, cols, d_x, d_y)
    torch.cuda.synchronize()
    want = port.matvec(61, d_A.cpu().numpy().view(np.uint64), d_x.cpu().numpy().view(np.uint64))
    assert np.array_equal(d_y.cpu().numpy().view(np.uint64), want)
    del d_A
    n = 1 << 26
    v = {k: torch.empty(n, dtype=torch.int64, device="cuda") for k in "ebdacz"}
    for k in "ebdac":
        ctx.random_dev(61, k, 0, n, v[k])
    ctx.beaver_dev(61, v["e"], v["b"], v["d"], v["a"], v["c"], n, v["z"])
    torch.cuda.synchronize()
    K = 1 << 16
    for lo in (0, n - K):
        h = {k: v[k][lo:lo + K].cpu().numpy().view(np.uint64) for k in "ebdacz"}
        assert np.array_equal(h["z"], port.beaver(61, h["e"], h["b"], h["d"], h["a"], h["c"]))
    t1, t2, lhs = (torch.empty(n, dtype=torch.int64, device="cuda") for _ in range(3))
    ctx.vec_op_dev(61, 0, v["b"], v["d"], n, t1)        # b + d
    ctx.vec_op_dev(61, 2, v["e"], t1, n, t1)            # e * (b + d)
    ctx.vec_op_dev(61, 2, v["d"], v["a"], n, t2)        # d * a
    ctx.vec_op_dev(61, 0, t1, t2, n, t1)
    ctx.vec_op_dev(61, 1, v["z"], v["c"], n, lhs)       # z - c
    torch.cuda.synchronize()
    assert torch.equal(lhs, t1)


# ------------------------------------------------------------------ Matrix::vandermonde(xs), Polynomial::evaluate, transpose
@pytest.mark.parametrize("field", [61, 127])
def test_vandermonde_xs_poly_evaluate_transpose(ctx, pkg, orc, port, field):
    """Matrix::vandermonde(n, m, xs) with caller nodes (matrix.h:445-460, "|xs| != number of rows"),
    Polynomial::evaluate at caller points (poly.h:56-64) in the host and device forms, Matrix::transpose
    (matrix.h:344-355) -- against the reference's own functions (ref_driver.cc calls them)."""
    import torch
    if orc.kind != "reference":
        pytest.skip("needs oracle/_ref (the unmodified reference)")
    p_int = (1 << field) - 1
    for n, m in ((1, 1), (5, 3), (32, 16), (100, 7)):
        xs = port.from_ints([(i * i * 7919 + 3) % p_int if i % 5 else p_int - i - 1 for i in range(n)], field)
        assert np.array_equal(ctx.vandermonde_xs(field, n, m, xs), orc.vandermonde_xs(field, n, m, xs))
    with pytest.raises(pkg.InvalidArgument, match=r"\|xs\| != number of rows"):
        ctx.vandermonde_xs(field, 4, 3, port.from_ints([1, 2, 3], field))
    with pytest.raises(ValueError):
        orc.vandermonde_xs(field, 4, 3, port.from_ints([1, 2, 3], field))
    w = 1 if field == 61 else 2
    for N, t, n in ((1, 0, 1), (7, 3, 5), (3000, 15, 32), (1025, 40, 70), (70000, 2, 3)):
        coeffs = port.vector_random(field, "poly", 0, N * (t + 1)).reshape((N, t + 1) + ((2,) if w == 2 else ()))
        if N > 5:
            coeffs[3, t] = 0       # a zero leading coefficient: Polynomial::create strips it, the value is the same
            coeffs[4] = 0          # the zero polynomial
        xs = port.from_ints([0, 1, p_int - 1] + [(j * 104729 + 17) % p_int for j in range(n)], field)[:n]
        want = orc.poly_evaluate(field, coeffs, xs)
        assert np.array_equal(ctx.poly_evaluate(field, coeffs, xs), want)
        # device form: coefficient planes [t+1][N], both layouts
        planes = np.ascontiguousarray(np.swapaxes(coeffs.reshape(N, t + 1, w), 0, 1))
        d_pl = torch.from_numpy(planes.view(np.int64)).cuda()
        d_pm = torch.empty((n, N, w), dtype=torch.int64, device="cuda")
        d_sm = torch.empty((N, n, w), dtype=torch.int64, device="cuda")
        ctx.poly_evaluate_dev(field, d_pl, N, t, xs, d_pm, pkg.binding.PARTY_MAJOR)
        ctx.poly_evaluate_dev(field, d_pl, N, t, xs, d_sm, pkg.binding.SECRET_MAJOR)
        torch.cuda.synchronize()
        assert np.array_equal(d_sm.cpu().numpy().view(np.uint64).reshape(want.shape), want)
        assert np.array_equal(np.swapaxes(d_pm.cpu().numpy().view(np.uint64), 0, 1).reshape(want.shape), want)
    # the last eight: one side narrow and the other long -- k_transpose_narrow, both directions, ragged ends
    for rows, cols in ((1, 1), (3, 5), (64, 33), (1000, 17), (5000, 5), (7, 4097), (32, 1025), (3001, 32), (2049, 1), (1, 5000),
                       (8, 2000), (16, 3000)):
        A = port.vector_random(field, "tr", 0, rows * cols).reshape((rows, cols) + ((2,) if w == 2 else ()))
        assert np.array_equal(ctx.transpose(field, A), orc.transpose(field, A))


# ------------------------------------------------------------------ gather / multi-device / async on whatever is visible
@pytest.mark.parametrize("N,n", [(4096, 32), (1001, 5), (2, 3), (100000, 16)])
def test_recover_p_gather_destinations(ctx, orc, N, n):
    """sclgpu_fp61_recover_p_gather_dev: the reconstructed secrets land in every destination buffer at the given offset
    (here: several buffers of the same device; the peer-memory case runs in tests/dist_gpu_worker.py)."""
    import torch
    t = (n - 1) // 2
    secrets = orc.vector_random(61, "secrets", 0, N)
    want = orc.shamir_share(61, secrets, t, n, "g", 4)
    d_sh = torch.from_numpy(np.ascontiguousarray(want.T).view(np.int64)).cuda()
    off = 6
    bufs = [torch.zeros(N + 16, dtype=torch.int64, device="cuda") for _ in range(3)]
    ctx.recover_p_gather_dev(d_sh, N, n, [b_.data_ptr() for b_ in bufs], off)
    torch.cuda.synchronize()
    for b_ in bufs:
        h = b_.cpu().numpy().view(np.uint64)
        assert np.array_equal(h[off:off + N], secrets)
        assert not h[:off].any() and not h[off + N:].any()


def test_fused_launch_with_gather_destinations(ctx, orc):
    """sclgpu_fp61_shamir_share_recover_gather_dev: share planes as usual, the reconstructed secrets of ANOTHER batch
    stored into every destination at the offset (single GPU: several local buffers; peers: tests/dist_gpu_worker.py)."""
    import torch
    N, t, n = 40000, 15, 32
    secrets = orc.vector_random(61, "secrets", 0, N)
    prev = orc.shamir_share(61, secrets, t, n, "previous batch", 0)
    d_prev = torch.from_numpy(np.ascontiguousarray(prev.T).view(np.int64)).cuda()
    d_sec = torch.from_numpy(secrets.view(np.int64)).cuda()
    d_sh = torch.zeros((n, N), dtype=torch.int64, device="cuda")
    off = 10
    bufs = [torch.zeros(N + 32, dtype=torch.int64, device="cuda") for _ in range(4)]
    ctx.shamir_share_recover_gather_dev(d_sec, N, t, n, "shamir bench", 5, d_sh, [b_.data_ptr() for b_ in bufs], off, rec_shares=d_prev)
    torch.cuda.synchronize()
    assert np.array_equal(d_sh.cpu().numpy().view(np.uint64).T, orc.shamir_share(61, secrets, t, n, "shamir bench", 5))
    for b_ in bufs:
        h = b_.cpu().numpy().view(np.uint64)
        assert np.array_equal(h[off:off + N], secrets) and not h[:off].any() and not h[off + N:].any()
    # same batch (dependent mode), odd offset: 8-byte aligned destinations fall back to the two-kernel path
    for b_ in bufs:
        b_.zero_()
    ctx.shamir_share_recover_gather_dev(d_sec, N, t, n, "shamir bench", 5, d_sh, [b_.data_ptr() for b_ in bufs], 7)
    torch.cuda.synchronize()
    for b_ in bufs:
        assert np.array_equal(b_.cpu().numpy().view(np.uint64)[7:7 + N], secrets)


def test_multi_context_vs_oracle(pkg, orc):
    """sclgpu_mctx over every visible GPU (one is enough: the slicing and the PRG offsets are the same code): share,
    recoverP, recoverD and Vector::random equal the one-PRG batch of the oracle."""
    import torch
    g = max(1, min(torch.cuda.device_count(), 8))
    devs = list(range(g)) if g > 1 else [0, 0, 0]   # one GPU: three slices on the same device
    m = pkg.MultiContext(devs)
    try:
        for N in (0, 1, 2, 5, 1000, 77777):
            secrets = orc.vector_random(61, "secrets", 0, N)
            assert np.array_equal(m.random("secrets", 0, N), secrets)
            sh = m.shamir_share(61, secrets, 15, 32, "shamir bench", 11)
            assert np.array_equal(sh, orc.shamir_share(61, secrets, 15, 32, "shamir bench", 11))
            assert np.array_equal(m.recover_p(61, sh), secrets)
        s127 = orc.vector_random(127, "secrets127", 0, 2049)
        sh127 = m.shamir_share(127, s127, 7, 16, "m127", 1)
        assert np.array_equal(sh127, orc.shamir_share(127, s127, 7, 16, "m127", 1))
        sh127[5, 2, 1] ^= np.uint64(4)
        sh127[2048, 13, 0] ^= np.uint64(1)
        sh127[1000, 15, 0] ^= np.uint64(1)   # unchecked index (shamir.h:129)
        out, err, nd = m.recover_d(127, sh127, 7)
        w_out, w_err, w_nd = orc.recover_d(127, sh127, 7)
        assert nd == w_nd == 2 and np.array_equal(err, w_err) and np.array_equal(out, w_out)
    finally:
        m.close()


def test_async_host_calls(ctx, orc):
    """share_async(batch k) overlapping recover_p(batch k-1) on one context; results equal the synchronous calls."""
    N, t, n = 50000, 15, 32
    secrets = orc.vector_random(61, "secrets", 0, N)
    want = [orc.shamir_share(61, secrets, t, n, "shamir bench", 100 * k) for k in range(3)]
    bufs = [np.zeros((N, n), dtype=np.uint64) for _ in range(2)]
    out = np.zeros(N, dtype=np.uint64)
    ctx.shamir_share_async(61, secrets, t, n, "shamir bench", 0, bufs[0])
    ctx.wait()
    assert np.array_equal(bufs[0], want[0])
    for k in (1, 2):
        ctx.shamir_share_async(61, secrets, t, n, "shamir bench", 100 * k, bufs[k & 1])
        rec = ctx.recover_p(61, bufs[(k - 1) & 1])       # synchronous call on the same context, other batch
        ctx.wait()
        assert np.array_equal(rec, secrets)
        assert np.array_equal(bufs[k & 1], want[k])
    ctx.recover_p_async(61, bufs[0], N, n, out)
    ctx.sync()                                            # sync completes the pending asynchronous call too
    assert np.array_equal(out, secrets)
    # an error inside the asynchronous call surfaces at wait()
    with pytest.raises(Exception):
        ctx.recover_p_async(61, bufs[0], N, n, out, alphas=np.zeros(n, dtype=np.uint64), x=0)   # equal nodes: not invertible
        ctx.wait()


# ------------------------------------------------------------------ several GPUs, NCCL
def test_multi_gpu_nccl():
    """tests/dist_gpu_worker.py under torchrun: batch-sharded share / reconstruct with per-rank PRG offsets, error counts
    summed over ranks, C5's row-sharded mat-vec with its all-gather, both peer-memory gathers, the multi-device handle --
    all against the oracle.  One process per visible GPU over NCCL when there are at least two; on a one-GPU box two
    ranks share the GPU (collectives over gloo, peer memory through cudaIpc between the two processes), so the N > 1
    path is exercised on every box."""
    import os
    import subprocess
    import sys

    import torch

    g = torch.cuda.device_count()
    world = min(g, 8) if g >= 2 else 2
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(repo, "tests", "dist_gpu_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
