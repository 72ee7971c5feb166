#!/usr/bin/env python
"""Small workload that touches every hot kernel once or twice (for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry

pkg = entry.load_package()
o = entry.load_oracle(); port = o.PortOracle()
ctx = pkg.Context(0)
for field, t, n, N in [(61, 15, 32, 700), (61, 2, 5, 130), (127, 7, 16, 300), (61, 16, 40, 64)]:
    sec = port.vector_random(field, "secrets", 0, N)
    sh = ctx.shamir_share(field, sec, t, n, "shamir bench", 5)
    assert np.array_equal(sh, port.shamir_share(field, sec, t, n, "shamir bench", 5))
    assert np.array_equal(ctx.recover_p(field, sh), sec)
    if n >= 2 * t + 1:
        out, err, nd = ctx.recover_d(field, sh, t)
        assert nd == 0 and np.array_equal(out, sec)
    pk = ctx.shamir_share_packets(field, sec, t, n, "shamir bench", 5)
    assert np.array_equal(ctx.recover_p_packets(field, pk, N), sec)
    ad = ctx.additive_share(field, sec, 4, "additive", 1)
    assert np.array_equal(ctx.additive_recover(field, ad), sec)
assert np.array_equal(ctx.prg_expand("k", 250, 5000), port.prg_next("k", 250, 5000))
assert np.array_equal(ctx.vector_random(61, "v", 3, 1001), port.vector_random(61, "v", 3, 1001))
A = port.vector_random(61, "mat A", 0, 64 * 512).reshape(64, 512); x = port.vector_random(61, "vec x", 0, 512)
assert np.array_equal(ctx.matvec(61, A, x), port.matvec(61, A, x))
a = port.vector_random(61, "a", 0, 1001); b = port.vector_random(61, "b", 0, 1001)
assert np.array_equal(ctx.beaver(61, a, b, a, b, a), port.beaver(61, a, b, a, b, a))
# error-correcting reconstruction, GEMM on tensor cores (two accumulation rounds, ragged tiles), Fp127 GEMM
sec = port.vector_random(61, "secrets", 0, 200)
sh = ctx.shamir_share(61, sec, 3, 10, "rc", 0).copy(); sh[::3, 4] ^= np.uint64(9)
f, e, st, nf = ctx.recover_c(61, sh)
assert nf == 0 and np.array_equal(f[:, 0], sec)
A = port.vector_random(61, "mat A", 0, 130 * 4100).reshape(130, 4100); Bm = port.vector_random(61, "mat B", 0, 4100 * 33).reshape(4100, 33)
assert np.array_equal(ctx.matmul(61, A, Bm), port.matmul(61, A, Bm))
A7 = port.vector_random(127, "mat A", 0, 9 * 6).reshape(9, 6, 2); B7 = port.vector_random(127, "mat B", 0, 6 * 5).reshape(6, 5, 2)
assert np.array_equal(ctx.matmul(127, A7, B7), port.matmul(127, A7, B7))
# array-valued secrets: fused pair exchange (Fp61 even W), strided counters (Fp127), staged path (odd W), wide transposes
for field, W, t, n, N in [(61, 2, 15, 32, 300), (61, 4, 2, 5, 77), (61, 3, 3, 7, 100), (127, 2, 7, 16, 150), (127, 3, 9, 12, 40)]:
    sec = port.vector_random(field, "pairs", 0, N * W).reshape((N, W) + (() if field == 61 else (2,)))
    sh = ctx.shamir_share_array(field, sec, t, n, "pedersen", 3)
    assert np.array_equal(sh, port.shamir_share_array(field, sec, t, n, "pedersen", 3))
    assert np.array_equal(ctx.recover_p_array(field, sh), sec)
assert np.array_equal(ctx.hyper_invertible(61, 4, 5), port.hyper_invertible(61, 4, 5))
ctx.close()
print("SANITIZE_DRIVER_OK")
