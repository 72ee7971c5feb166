"""Generates tests/golden/scl_golden.json from the UNMODIFIED reference.

Run in the build container (needs /root/reference to build oracle/_ref/libsclref.so):

    python tests/golden/make_golden.py

Every value is what SCL 0.1.0's own code returns (through oracle/ref_driver.cc);
elements are recorded as hex of the integer value, keystream as hex bytes.  The
JSON is small and committed; the GPU box and the CPU-only test run read it back
(tests/test_oracle_golden.py, tests/test_gpu_parity.py) -- /root/reference is
never needed at test time.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import PRIME, RefOracle, from_ints, to_ints  # noqa: E402


def hx(arr, field):
    return [format(int(v), "x") for v in to_ints(arr, field).reshape(-1)]


def main():
    r = RefOracle()
    g = {"generator": "tests/golden/make_golden.py", "reference": "scl-0.1.0-unmodified (oracle/_ref/libsclref.so)"}

    # --- util::PRG keystream (prg.cc:82-84,124-146)
    g["prg"] = []
    for seed, first, nbytes in [("", 0, 48), ("shamir passive", 0, 32), ("shamir bench", 0, 64),
                                ("shamir bench", 1 << 20, 40), ("a seed longer than sixteen bytes", 7, 33),
                                ("prg bench", (1 << 27) - 2, 32), ("x", 3, 1), ("x", 3, 17)]:
        g["prg"].append({"seed": seed, "first_block": first, "n_bytes": nbytes,
                         "hex": bytes(r.prg_next(seed, first, nbytes)).hex()})

    # --- Vector::random / FF::random
    g["random"] = []
    for field in (61, 127):
        for seed, first, n in [("", 0, 4), ("secrets", 0, 9), ("prg bench", 12345, 7)]:
            g["random"].append({"field": field, "kind": "vector", "seed": seed, "first_block": first, "n": n,
                                "hex": hx(r.vector_random(field, seed, first, n), field)})
            g["random"].append({"field": field, "kind": "ff", "seed": seed, "first_block": first, "n": n,
                                "hex": hx(r.ff_random(field, seed, first, n), field)})

    # --- FF::read edge cases
    g["from_bytes"] = []
    for field in (61, 127):
        bs = 8 if field == 61 else 16
        p = PRIME[field]
        raws = [b"\xff" * bs, p.to_bytes(bs, "little"), (p + 1).to_bytes(bs, "little"), (p - 1).to_bytes(bs, "little"),
                b"\x00" * bs, bytes(range(1, bs + 1)), (2 * p + 1).to_bytes(bs, "little")]
        raw = b"".join(raws)
        g["from_bytes"].append({"field": field, "raw": raw.hex(), "hex": hx(r.from_bytes(field, raw), field)})

    # --- scalar ops incl. edges
    g["scalar"] = []
    for field in (61, 127):
        p = PRIME[field]
        vals = [0, 1, 2, 3, p - 1, p - 2, (p + 1) // 2, 0x123456789ABCDEF % p, (1 << 60) + 12345, p // 3]
        if field == 127:
            vals += [(1 << 126) + (1 << 64) - 1, (1 << 64), (1 << 64) - 1, (1 << 100) + 77]
        for op in range(6):
            for a in vals:
                for b in (vals if op in (0, 1, 2, 5) else [0]):
                    rc, v = r.scalar_op(field, op, a, b)
                    g["scalar"].append([field, op, format(a, "x"), format(b, "x"), rc, format(v, "x")])

    # --- shamirSecretShare / recoverP / recoverD
    g["shamir"] = []
    cases = [(61, 2, 5, 6, "shamir passive", 0), (61, 15, 32, 4, "shamir bench", 0), (61, 15, 32, 3, "shamir bench", 8 * 1000),
             (61, 0, 3, 2, "t0", 0), (61, 1, 1, 2, "n1", 5), (61, 16, 40, 2, "t16", 0), (61, 20, 48, 2, "t20", 3),
             (61, 3, 4, 3, "shamir", 0),
             (127, 7, 16, 4, "m127", 0), (127, 2, 5, 3, "shamir passive", 0), (127, 8, 20, 2, "t8", 11), (127, 10, 24, 2, "t10", 0),
             (127, 0, 2, 2, "t0", 0)]
    for field, t, n, N, seed, first in cases:
        secrets = from_ints([123 + j for j in range(N)], field)
        sh = r.shamir_share(field, secrets, t, n, seed, first)
        rec = r.recover_p(field, sh)
        g["shamir"].append({"field": field, "t": t, "n": n, "N": N, "seed": seed, "first_block": first,
                            "secrets": hx(secrets, field), "shares": hx(sh, field), "recover_p": hx(rec, field)})

    # --- shamirRecoverC (shamir.h:203-258): 0..t+1 corrupted shares per sharing
    g["recover_c"] = []
    for field, n, seed in [(61, 4, "rc4"), (61, 7, "rc7"), (61, 16, "rc16"), (61, 5, "rc5"), (127, 7, "rc127"), (127, 10, "rc10")]:
        t = (n - 1) // 3
        N = t + 3
        secrets = from_ints([1000 + j for j in range(N)], field)
        sh = r.shamir_share(field, secrets, t, n, seed, 0).copy()
        flat = sh.reshape(N, n, -1)
        for j in range(N):                      # sharing j gets j corrupted shares (j = t+1, t+2: beyond the radius)
            for i in range(min(j, 3 * t + 1)):
                flat[j, (2 * i + j) % (3 * t + 1), 0] ^= np.uint64(0x10001 + i)
        f, e, st, nf = r.recover_c(field, sh)
        g["recover_c"].append({"field": field, "n": n, "N": N, "shares": hx(sh, field), "f": hx(f, field),
                               "err": hx(e, field), "status": [int(v) for v in st], "n_failed": nf})

    # --- additiveShare (additive.h:42-53): n-1 FF::random (one block each) + secret - sum
    g["additive"] = []
    for field, n, N, seed, first in [(61, 3, 4, "additive", 0), (61, 1, 2, "additive", 3), (61, 5, 3, "shamir bench", 1 << 20),
                                     (127, 4, 3, "additive", 0), (127, 2, 2, "a127", 9)]:
        secrets = from_ints([123 + j for j in range(N)], field)
        sh = r.additive_share(field, secrets, n, seed, first)
        g["additive"].append({"field": field, "n": n, "N": N, "seed": seed, "first_block": first,
                              "secrets": hx(secrets, field), "shares": hx(sh, field),
                              "recover": hx(r.additive_recover(field, sh), field)})

    # --- shamirSecretShare / shamirRecoverP on math::Array<FF, W> (pedersen.h:137-138 uses W = 2), and
    #     Matrix::hyperInvertible (matrix.h:462-475)
    g["share_array"] = []
    for field, W, t, n, N, seed, first in [(61, 2, 2, 5, 3, "pedersen", 0), (61, 2, 15, 32, 2, "shamir bench", 7),
                                           (61, 3, 1, 4, 2, "array", 0), (61, 1, 2, 5, 2, "shamir passive", 0),
                                           (127, 2, 2, 5, 2, "pedersen", 0), (127, 3, 7, 16, 2, "m127", 11)]:
        secrets = from_ints([[100 * j + w + 1 for w in range(W)] for j in range(N)], field)
        sh = r.shamir_share_array(field, secrets, t, n, seed, first)
        g["share_array"].append({"field": field, "W": W, "t": t, "n": n, "N": N, "seed": seed, "first_block": first,
                                 "secrets": hx(secrets, field), "shares": hx(sh, field),
                                 "recover": hx(r.recover_p_array(field, sh), field)})
    g["hyper_invertible"] = []
    for field, n, m in [(61, 4, 5), (61, 1, 1), (61, 6, 3), (127, 4, 5), (127, 3, 3)]:
        g["hyper_invertible"].append({"field": field, "n": n, "m": m, "him": hx(r.hyper_invertible(field, n, m), field)})

    # sum over all shares of 1024 calls (secrets 123..1146), t=15 n=32, PRG("shamir bench")
    secrets = from_ints([123 + j for j in range(1024)], 61)
    sh = r.shamir_share(61, secrets, 15, 32, "shamir bench", 0)
    tot = 0
    for v in to_ints(sh[1:], 61).reshape(-1):
        tot = (tot + int(v)) % PRIME[61]
    g["survey_sum_1023"] = format(tot, "x")

    # custom alphas / x (test_shamir.cc:42-66, 100-109)
    g["recover_p_custom"] = []
    for field in (61, 127):
        secrets = from_ints([555, 7], field)
        sh = r.shamir_share(field, secrets, 3, 8, "custom", 0)
        for lo, x in [(0, 0), (2, 0), (2, 27)]:
            al = from_ints([i + 1 for i in range(lo, lo + 6)], field)
            win = np.ascontiguousarray(sh[:, lo:lo + 6])
            rec = r.recover_p(field, win, al, x)
            g["recover_p_custom"].append({"field": field, "shares": hx(win, field), "alphas": hx(al, field), "x": x,
                                          "N": 2, "n": 6, "out": hx(rec, field)})

    # recoverD incl. the unchecked-index quirk (shamir.h:129)
    g["recover_d"] = []
    for field, t, n in [(127, 7, 16), (61, 7, 16), (61, 3, 7), (127, 2, 5), (61, 1, 3)]:
        N = 8
        secrets = from_ints([1000 + j for j in range(N)], field)
        sh = r.shamir_share(field, secrets, t, n, "m127", 0)
        tam = sh.copy()
        idxs = [0, min(n - 1, 2 * t - 1), min(n - 1, 2 * t), n - 1, t, t + 1, None, None]
        for j, ix in enumerate(idxs):
            if ix is not None:
                tam.reshape(N, n, -1)[j, ix, 0] ^= np.uint64(1)
        out, err, rc = r.recover_d(field, tam, t)
        g["recover_d"].append({"field": field, "t": t, "n": n, "N": N, "shares": hx(tam, field), "tampered_idx": idxs,
                               "out": hx(out, field), "err": [int(e) for e in err], "rc": rc})
    # not enough shares
    sh = r.shamir_share(61, from_ints([5], 61), 3, 5, "few", 0)
    out, err, rc = r.recover_d(61, sh, 3)
    g["recover_d_not_enough"] = {"field": 61, "t": 3, "n": 5, "rc": rc}
    # five-argument form
    sh = r.shamir_share(61, from_ints([42, 43], 61), 2, 9, "five", 0)
    al = from_ints([i + 1 for i in range(9)], 61)
    out, err, rc = r.recover_d(61, sh, 2, alphas=al, d=4, x=0)
    g["recover_d_custom"] = {"field": 61, "t": 2, "d": 4, "n": 9, "shares": hx(sh, 61), "alphas": hx(al, 61), "x": 0,
                             "out": hx(out, 61), "err": [int(e) for e in err], "rc": rc}

    # --- Lagrange basis
    g["lagrange"] = []
    for field in (61, 127):
        for nodes, x in [(list(range(1, 6)), 0), (list(range(1, 33)), 0), ([42, 43, 44, 45, 46, 47, 48, 49], 0),
                         (list(range(1, 9)), 9), ([3, 1, 4, 15, 9, 2, 6], PRIME[field] - 5)]:
            lb = r.lagrange(field, from_ints(nodes, field), x)
            g["lagrange"].append({"field": field, "nodes": nodes, "x": format(x, "x"), "hex": hx(lb, field)})

    # --- Vector ops / beaver / matvec / vandermonde on PRG-drawn inputs
    g["vec"] = []
    for field in (61, 127):
        a = r.vector_random(field, "a", 0, 37)
        b = r.vector_random(field, "b", 0, 37)
        for op in range(6):
            out = r.vec_op(field, op, a, b)
            g["vec"].append({"field": field, "op": op, "n": 37, "hex": hx(out, field)})
        e, d, c = (r.vector_random(field, s, 0, 37) for s in ("e", "d", "c"))
        g["vec"].append({"field": field, "op": "beaver", "n": 37, "hex": hx(r.beaver(field, e, b, d, a, c), field)})
    g["matvec"] = []
    for field, rows, cols in [(61, 64, 64), (61, 5, 33), (127, 9, 20)]:
        A = r.vector_random(field, "mat A", 0, rows * cols).reshape((rows, cols) + (() if field == 61 else (2,)))
        x = r.vector_random(field, "vec x", 0, cols)
        g["matvec"].append({"field": field, "rows": rows, "cols": cols, "hex": hx(r.matvec(field, A, x), field)})
    g["vandermonde"] = [{"field": f, "n": n, "m": m, "hex": hx(r.vandermonde(f, n, m), f)}
                        for f, n, m in [(61, 5, 3), (61, 32, 16), (127, 16, 8)]]

    path = os.path.join(HERE, "scl_golden.json")
    with open(path, "w") as fh:
        json.dump(g, fh, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
