"""Pins the oracle: the plain-C restatement (oracle/scl_oracle.c) against
(a) vectors recorded from the unmodified reference (tests/golden/scl_golden.json,
made by tests/golden/make_golden.py), (b) the SURVEY's
hard-coded known answers, (c) the reference's own behavioural tests restated
(test/scl/ss/test_shamir.cc, test_poly.cc, test_vector.cc, test_matrix.cc), and
(d) the compiled reference itself on seeded random inputs when its .so is here."""
import numpy as np
import pytest

P = {61: (1 << 61) - 1, 127: (1 << 127) - 1}


def ints(o, arr, field):
    return [int(v) for v in o.to_ints(arr, field).reshape(-1)]


def unhex(o, hexes, field, shape=None):
    a = o.from_ints([int(h, 16) for h in hexes], field)
    if shape is not None:
        a = a.reshape(tuple(shape) + (() if field == 61 else (2,)))
    return a


# ------------------------------------------------------------- SURVEY 8c KATs
def test_survey_prg_kat(port):
    assert bytes(port.prg_next(b"", 0, 48)).hex() == (
        "7727a8004ea0c9708441893d2808ca94570feebdca7b0c8ef044a2dc19fd880350e50584d2a1c30fa50deb669e963045")
    assert bytes(port.prg_next("shamir passive", 0, 32)).hex() == (
        "965ef33d3c2cdc3467662ff22077cede048cd9a369c0f7ec8b1e76f36baec71a")


def test_prg_matches_independent_aes(port):
    """block i = AES128_seed(LE64(i) || LE64(PRG_NONCE)) checked against `cryptography`."""
    import struct
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

    key = b"shamir bench" + b"\0" * 4
    enc = Cipher(algorithms.AES(key), modes.ECB()).encryptor()
    want = b"".join(enc.update(struct.pack("<QQ", c, 0x0123456789ABCDEF)) for c in range(1000, 1010))
    assert bytes(port.prg_next("shamir bench", 1000, 160)) == want


def test_survey_random_consumption_pattern(port):
    assert [hex(v) for v in ints(port, port.ff_random(61, b"", 0, 2), 61)] == ["0x10c9a04e00a8277a", "0xe0c7bcabdee0f5b"]
    assert [hex(v) for v in ints(port, port.vector_random(61, b"", 0, 4), 61)] == [
        "0x10c9a04e00a8277a", "0x14ca08283d894188", "0xe0c7bcabdee0f5b", "0x388fd19dca244f0"]


def test_survey_shamir_kat(port):
    s = port.from_ints([123, 124], 61)
    sh = port.shamir_share(61, s, 2, 5, "shamir passive")
    assert [hex(v) for v in ints(port, sh[0], 61)] == [
        "0xbc6378a9608f2f4", "0x117befe873c4fd84", "0x112129199934202b", "0xab5e31e06565ae9", "0x1e3a1df5bb2badbd"]
    assert [hex(v) for v in ints(port, sh[1], 61)] == [
        "0x37ab993dfb3a789", "0x36a638150d15476", "0x1fcefdc853590742", "0x18a88868e74abfef", "0xdf703630ca67e7c"]
    s = port.from_ints([123 + j for j in range(1024)], 61)
    sh = port.shamir_share(61, s, 15, 32, "shamir bench")
    v = ints(port, sh[0], 61)
    assert (hex(v[0]), hex(v[1]), hex(v[31])) == ("0x18a322ca09e6e84c", "0x9e2135ee856c5ee", "0x12734b17e48df10")
    assert ints(port, port.recover_p(61, sh[:1]), 61) == [0x7B]
    assert sum(ints(port, sh[1:], 61)) % P[61] == 0x44672D90DD13206


def test_survey_lagrange_and_edges(port):
    lb = ints(port, port.lagrange(61, port.from_ints([1, 2, 3, 4, 5], 61), 0), 61)
    assert lb == [5, 0x1FFFFFFFFFFFFFF5, 0xA, 0x1FFFFFFFFFFFFFFA, 1]
    lb = ints(port, port.lagrange(61, port.from_ints(list(range(1, 33)), 61), 0), 61)
    assert (lb[0], lb[1], lb[31], sum(lb) % P[61]) == (0x20, 0x1FFFFFFFFFFFFE0F, 0x1FFFFFFFFFFFFFFE, 1)
    p = P[61]
    assert port.scalar_op(61, 2, p - 1, p - 1) == (0, 1)
    assert port.scalar_op(61, 0, p - 1, p - 1) == (0, 0x1FFFFFFFFFFFFFFD)
    assert port.scalar_op(61, 1, 0, 1) == (0, 0x1FFFFFFFFFFFFFFE)
    assert port.scalar_op(61, 4, 2) == (0, 0x1000000000000000)
    assert port.scalar_op(61, 4, 0)[0] == -2
    assert ints(port, port.from_bytes(61, b"\xff" * 8 + p.to_bytes(8, "little")), 61) == [7, 0]
    assert ints(port, port.from_bytes(127, b"\xff" * 16), 127) == [1]
    assert port.scalar_op(127, 2, P[127] - 1, P[127] - 1) == (0, 1)


def test_survey_matvec_kat(port):
    A = port.vector_random(61, "mat A", 0, 64 * 64).reshape(64, 64)
    x = port.vector_random(61, "vec x", 0, 64)
    y = ints(port, port.matvec(61, A, x), 61)
    assert (y[0], y[63], sum(y) % P[61]) == (0x1172BF06CC5D2E8B, 0x5AD2E6879DFACC8, 0xF8FDCF52691A572)


# ---------------------------------------------- golden JSON from the reference
def test_golden_prg(port, golden):
    for c in golden["prg"]:
        assert bytes(port.prg_next(c["seed"], c["first_block"], c["n_bytes"])).hex() == c["hex"], c


def test_golden_random(port, golden):
    for c in golden["random"]:
        f = port.vector_random if c["kind"] == "vector" else port.ff_random
        got = f(c["field"], c["seed"], c["first_block"], c["n"])
        assert ints(port, got, c["field"]) == [int(h, 16) for h in c["hex"]], c


def test_golden_from_bytes(port, golden):
    for c in golden["from_bytes"]:
        got = port.from_bytes(c["field"], bytes.fromhex(c["raw"]))
        assert ints(port, got, c["field"]) == [int(h, 16) for h in c["hex"]]


def test_golden_scalar(port, golden):
    for field, op, a, b, rc, v in golden["scalar"]:
        got = port.scalar_op(field, op, int(a, 16), int(b, 16))
        assert got[0] == rc, (field, op, a, b)
        if rc == 0:
            assert got[1] == int(v, 16), (field, op, a, b)


def test_golden_shamir(port, golden):
    for c in golden["shamir"]:
        f = c["field"]
        secrets = unhex(port, c["secrets"], f)
        sh = port.shamir_share(f, secrets, c["t"], c["n"], c["seed"], c["first_block"])
        assert ints(port, sh, f) == [int(h, 16) for h in c["shares"]], (f, c["t"], c["n"])
        assert ints(port, port.recover_p(f, sh), f) == [int(h, 16) for h in c["recover_p"]]
    assert golden["survey_sum_1023"] == "44672d90dd13206"


def test_golden_recover_c(port, golden):
    """shamirRecoverC (Berlekamp-Welch, shamir.h:203-258): vectors recorded from the reference with
    0..t+2 corrupted shares per sharing (the last two beyond the correction radius)."""
    for c in golden["recover_c"]:
        f, n, N = c["field"], c["n"], c["N"]
        sh = unhex(port, c["shares"], f, (N, n))
        pf, pe, st, nf = port.recover_c(f, sh)
        assert [int(v) for v in st] == c["status"] and nf == c["n_failed"], c["n"]
        assert ints(port, pf, f) == [int(h, 16) for h in c["f"]]
        assert ints(port, pe, f) == [int(h, 16) for h in c["err"]]
        t = (n - 1) // 3
        for j in range(min(N, t + 1)):               # within the radius: the secret is f(0)
            assert ints(port, pf[j], f)[0] == 1000 + j


def test_port_vs_reference_recover_c(port, ref):
    rng = np.random.default_rng(7)
    for field in (61, 127):
        for n in (1, 4, 6, 10, 16, 22, 31):
            t, N = (n - 1) // 3, 40
            sec = port.vector_random(field, "secrets", 0, N)
            sh = port.shamir_share(field, sec, t, n, "rc", 3).copy()
            flat = sh.reshape(N, n, -1)
            for j in range(N):
                k = j % (t + 2)
                for i in (rng.choice(3 * t + 1, size=min(k, 3 * t + 1), replace=False) if k else []):
                    flat[j, i, 0] ^= np.uint64(1 + j)
            a, b = port.recover_c(field, sh), ref.recover_c(field, sh)
            assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3] == b[3], (field, n)
            alphas = port.from_ints([5 * i + 3 for i in range(n)], field)
            sh2 = sh.copy()                                   # custom nodes: garbage in, same answer out
            a, b = port.recover_c(field, sh2, alphas), ref.recover_c(field, sh2, alphas)
            assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3] == b[3], (field, n, "alphas")


def test_golden_share_array_and_him(port, golden):
    """shamirSecretShare / shamirRecoverP on math::Array<FF, W> (pedersen.h:137-138) and
    Matrix::hyperInvertible (matrix.h:462-475): vectors recorded from the reference."""
    for c in golden["share_array"]:
        f, W, N, n = c["field"], c["W"], c["N"], c["n"]
        secrets = unhex(port, c["secrets"], f, (N, W))
        sh = port.shamir_share_array(f, secrets, c["t"], n, c["seed"], c["first_block"])
        assert ints(port, sh, f) == [int(h, 16) for h in c["shares"]], c
        assert ints(port, port.recover_p_array(f, sh), f) == [int(h, 16) for h in c["recover"]] == ints(port, secrets, f)
    for c in golden["hyper_invertible"]:
        f = c["field"]
        assert ints(port, port.hyper_invertible(f, c["n"], c["m"]), f) == [int(h, 16) for h in c["him"]]
    # W = 1 is the plain shamirSecretShare; component w of coefficient k is stream element k*W + w
    sec = port.from_ints([[5, 6], [7, 8]], 61)
    sh = port.shamir_share_array(61, sec, 1, 3, "pattern", 4)           # t = 1: share_i = s + c_1 * i
    c1 = port.vector_random(61, "pattern", 4, 8).reshape(2, 4)[:, 2:]   # 2 blocks per sharing, elements 2, 3
    p = (1 << 61) - 1
    want = [[[(int(sec[j, w]) + int(c1[j, w]) * i) % p for w in range(2)] for i in (1, 2, 3)] for j in range(2)]
    assert port.to_ints(sh, 61).tolist() == want


def test_port_vs_reference_share_array_and_him(port, ref):
    for field in (61, 127):
        for W in (1, 2, 3, 5):
            for t, n in ((2, 5), (15, 32), (0, 3), (7, 16), (4, 4), (20, 40)):
                sec = port.vector_random(field, "secrets", 0, 9 * W).reshape((9, W) + (() if field == 61 else (2,)))
                a = port.shamir_share_array(field, sec, t, n, "shamir bench", 3)
                assert np.array_equal(a, ref.shamir_share_array(field, sec, t, n, "shamir bench", 3)), (field, W, t, n)
                ra = port.recover_p_array(field, a)
                assert np.array_equal(ra, ref.recover_p_array(field, a))
                if t < n:
                    assert np.array_equal(ra, sec)
                if W == 1:
                    assert np.array_equal(a.reshape(-1), port.shamir_share(field, sec.reshape((9,) + (() if field == 61 else (2,))), t, n, "shamir bench", 3).reshape(-1))
        for n, m in ((4, 5), (1, 1), (7, 3), (16, 16), (33, 20)):
            assert np.array_equal(port.hyper_invertible(field, n, m), ref.hyper_invertible(field, n, m))


def test_golden_additive(port, golden):
    """additiveShare (additive.h:42-53): vectors recorded from the reference; reconstruction = sum."""
    for c in golden["additive"]:
        f = c["field"]
        secrets = unhex(port, c["secrets"], f)
        sh = port.additive_share(f, secrets, c["n"], c["seed"], c["first_block"])
        assert ints(port, sh, f) == [int(h, 16) for h in c["shares"]], c
        assert ints(port, port.additive_recover(f, sh), f) == [int(h, 16) for h in c["recover"]] == ints(port, secrets, f)
    # the consumption pattern: share i of secret j is FF::random at block first + j(n-1) + i
    sec = port.from_ints([5, 6, 7], 61)
    sh = port.additive_share(61, sec, 4, "pattern", 10)
    assert np.array_equal(sh[:, :3].reshape(-1), port.ff_random(61, "pattern", 10, 9))


def test_port_vs_reference_additive(port, ref):
    for field in (61, 127):
        sec = port.vector_random(field, "secrets", 0, 300)
        for n in (1, 2, 7, 40):
            a = port.additive_share(field, sec, n, "additive", 77)
            assert np.array_equal(a, ref.additive_share(field, sec, n, "additive", 77))
            assert np.array_equal(ref.additive_recover(field, a), sec)


def test_golden_recover_p_custom(port, golden):
    for c in golden["recover_p_custom"]:
        f = c["field"]
        sh = unhex(port, c["shares"], f, (c["N"], c["n"]))
        out = port.recover_p(f, sh, unhex(port, c["alphas"], f), c["x"])
        assert ints(port, out, f) == [int(h, 16) for h in c["out"]]


def test_golden_recover_d(port, golden):
    for c in golden["recover_d"]:
        f = c["field"]
        sh = unhex(port, c["shares"], f, (c["N"], c["n"]))
        out, err, rc = port.recover_d(f, sh, c["t"])
        assert rc == c["rc"] and [int(e) for e in err] == c["err"]
        assert ints(port, out, f) == [int(h, 16) for h in c["out"]]
    c = golden["recover_d_not_enough"]
    sh = port.shamir_share(61, port.from_ints([5], 61), c["t"], c["n"], "few")
    assert port.recover_d(61, sh, c["t"])[2] == -1 == c["rc"]
    c = golden["recover_d_custom"]
    sh = unhex(port, c["shares"], 61, (2, c["n"]))
    out, err, rc = port.recover_d(61, sh, c["t"], alphas=unhex(port, c["alphas"], 61), d=c["d"], x=c["x"])
    assert rc == c["rc"] and ints(port, out, 61) == [int(h, 16) for h in c["out"]]


def test_golden_lagrange_vec_mat(port, golden):
    for c in golden["lagrange"]:
        f = c["field"]
        lb = port.lagrange(f, port.from_ints(c["nodes"], f), int(c["x"], 16))
        assert ints(port, lb, f) == [int(h, 16) for h in c["hex"]]
    for c in golden["vec"]:
        f = c["field"]
        a, b = port.vector_random(f, "a", 0, 37), port.vector_random(f, "b", 0, 37)
        if c["op"] == "beaver":
            e, d, cc = (port.vector_random(f, s, 0, 37) for s in ("e", "d", "c"))
            got = port.beaver(f, e, b, d, a, cc)
        else:
            got = port.vec_op(f, c["op"], a, b)
        assert ints(port, got, f) == [int(h, 16) for h in c["hex"]], c["op"]
    for c in golden["matvec"]:
        f, rows, cols = c["field"], c["rows"], c["cols"]
        A = port.vector_random(f, "mat A", 0, rows * cols).reshape((rows, cols) + (() if f == 61 else (2,)))
        x = port.vector_random(f, "vec x", 0, cols)
        assert ints(port, port.matvec(f, A, x), f) == [int(h, 16) for h in c["hex"]]
    for c in golden["vandermonde"]:
        assert ints(port, port.vandermonde(c["field"], c["n"], c["m"]), c["field"]) == [int(h, 16) for h in c["hex"]]


# ---------------------- the reference's own behavioural tests, restated
@pytest.mark.parametrize("field", [61, 127])
def test_ref_test_shamir_behaviour(port, field):
    # test_shamir.cc:34-40
    s = port.from_ints([123], field)
    sh = port.shamir_share(field, s, 3, 4, "shamir")
    assert ints(port, port.recover_p(field, sh), field) == [123]
    # test_shamir.cc:68-79: recoverD ok, then shares[2] = 4 -> "error detected during recovery"
    sh = port.shamir_share(field, s, 3, 7, "shamir")
    out, err, rc = port.recover_d(field, sh, 3)
    assert rc == 0 and ints(port, out, field) == [123]
    bad = sh.copy()
    bad.reshape(1, 7, -1)[0, 2, :] = 0
    bad.reshape(1, 7, -1)[0, 2, 0] = 4
    out, err, rc = port.recover_d(field, bad, 3)
    assert rc == 1 and list(err) == [1]


def test_ref_small_integer_identities(port):
    # test_poly.cc:64-71: 4 + 5x + x^2 at 5 = 54 == the share of party 5 with fixed coefficients
    # (checked through vandermonde x coefficient dot product, test_matrix.cc:342-365)
    V = ints(port, port.vandermonde(61, 5, 3), 61)
    row5 = V[4 * 3: 5 * 3]
    assert row5 == [1, 5, 25] and sum(c * v for c, v in zip([4, 5, 1], row5)) == 54
    # test_vector.cc: dot of small vectors
    a, b = port.from_ints([1, 2, 3, 4], 61), port.from_ints([5, 6, 7, 8], 61)
    assert ints(port, port.vec_op(61, 4, a, b), 61) == [70]
    assert ints(port, port.vec_op(61, 5, a), 61) == [10]


# ------------------------------ port vs compiled reference, seeded random
@pytest.mark.parametrize("field", [61, 127])
def test_port_vs_reference_random(port, ref, field):
    rng = np.random.default_rng(field)
    for t, n, N in [(2, 5, 300), (15, 32, 50), (7, 16, 64), (0, 1, 5), (20, 41, 6)]:
        seed = f"seed {t} {n}"
        first = int(rng.integers(0, 1 << 16))  # the reference PRG cannot seek: skipping is O(first)
        secrets = ref.vector_random(field, "secrets", 3, N)
        assert np.array_equal(secrets, port.vector_random(field, "secrets", 3, N))
        a = ref.shamir_share(field, secrets, t, n, seed, first)
        b = port.shamir_share(field, secrets, t, n, seed, first)
        assert np.array_equal(a, b)
        assert np.array_equal(ref.recover_p(field, a), port.recover_p(field, b))
        if n >= 2 * t + 1 and t >= 1:
            tam = a.copy()
            flat = tam.reshape(N, n, -1)
            for j in range(0, N, 3):
                flat[j, int(rng.integers(0, n)), 0] ^= np.uint64(rng.integers(1, 1 << 30))
            ra, pa = ref.recover_d(field, tam, t), port.recover_d(field, tam, t)
            assert ra[2] == pa[2] and np.array_equal(ra[1], pa[1]) and np.array_equal(ra[0], pa[0])
    n = 1001
    a, b = ref.vector_random(field, "va", 0, n), ref.vector_random(field, "vb", 9, n)
    for op in range(6):
        assert np.array_equal(ref.vec_op(field, op, a, b), port.vec_op(field, op, a, b)), op
    assert np.array_equal(ref.ff_random(field, "ff", 77, 33), port.ff_random(field, "ff", 77, 33))
    assert bytes(ref.prg_next("zz", 12345, 1000)) == bytes(port.prg_next("zz", 12345, 1000))
