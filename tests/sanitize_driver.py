#!/usr/bin/env python
"""Small workload that touches every hot kernel once or twice (for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry

pkg = entry.load_package()
o = entry.load_oracle(); port = o.PortOracle()
ctx = pkg.Context(0)
B = pkg.binding
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
# round 2: the single-launch step (both modes, with gather destinations), reconstruction with the gather fused in,
# Berlekamp-Welch with more than 32 points (CTA per sharing), chunked mat-vec, caller-node Vandermonde / evaluation
import torch
N, t, n = 1500, 15, 32
sec = port.vector_random(61, "secrets", 0, N)
want = port.shamir_share(61, sec, t, n, "shamir bench", 9)
d_sec = torch.from_numpy(sec.view(np.int64)).cuda()
d_sh = torch.zeros((n, N), dtype=torch.int64, device="cuda"); d_sh2 = torch.zeros((n, N), dtype=torch.int64, device="cuda")
d_out = torch.zeros(N, dtype=torch.int64, device="cuda")
ctx.shamir_share_recover_dev(d_sec, N, t, n, "shamir bench", 9, d_sh, d_out)
torch.cuda.synchronize()
assert np.array_equal(d_sh.cpu().numpy().view(np.uint64).T, want) and np.array_equal(d_out.cpu().numpy().view(np.uint64), sec)
d_out.zero_()
ctx.shamir_share_recover_dev(d_sec, N, t, n, "shamir bench", 9, d_sh2, d_out, rec_shares=d_sh)
bufs = [torch.zeros(N + 8, dtype=torch.int64, device="cuda") for _ in range(3)]
ctx.shamir_share_recover_gather_dev(d_sec, N, t, n, "shamir bench", 9, d_sh2, [b_.data_ptr() for b_ in bufs], 4, rec_shares=d_sh)
ctx.recover_p_gather_dev(d_sh, N, n, [b_.data_ptr() for b_ in bufs[:2]], 2)
torch.cuda.synchronize()
assert np.array_equal(d_out.cpu().numpy().view(np.uint64), sec)
assert np.array_equal(bufs[2].cpu().numpy().view(np.uint64)[4:4 + N], sec) and np.array_equal(bufs[0].cpu().numpy().view(np.uint64)[2:2 + N], sec)
for field, nn in ((61, 40), (127, 34)):
    tt = (nn - 1) // 3
    sec = port.vector_random(field, "secrets", 0, 12)
    sh = port.shamir_share(field, sec, tt, nn, "rc", 0).copy()
    sh.reshape(12, nn, -1)[::2, 5, 0] ^= np.uint64(3)
    sh.reshape(12, nn, -1)[1, :tt + 1, 0] ^= np.uint64(7)
    got, want = ctx.recover_c(field, sh), port.recover_c(field, sh)
    assert all(np.array_equal(g, w) for g, w in zip(got[:3], want[:3]))
A = port.vector_random(61, "mat A", 0, 1024 * 4096).reshape(1024, 4096); x = port.vector_random(61, "vec x", 0, 4096)
assert np.array_equal(ctx.matvec(61, A[:64], x), port.matvec(61, A[:64], x))
assert np.array_equal(ctx.matvec(61, A, x)[:8], port.matvec(61, A[:8], x))
xs = port.vector_random(61, "xs", 0, 9)
cf = port.vector_random(61, "cf", 0, 50 * 6).reshape(50, 6)
assert ctx.poly_evaluate(61, cf, xs).shape == (50, 9) and ctx.vandermonde_xs(61, 9, 4, xs).shape == (9, 4)
# round 2, second session: the bitsliced keystream kernel (ranges inside / across 32-counter groups)
d_ks = torch.zeros(16 * 300, dtype=torch.uint8, device="cuda")
ctx.prg_expand_bitsliced_dev("k", 250, 16 * 300, d_ks)
torch.cuda.synchronize()
assert np.array_equal(d_ks.cpu().numpy(), port.prg_next("k", 250, 16 * 300))
# narrow transpositions (k_transpose_narrow, both directions, ragged ends) and shamirRecoverP on SCL's layout on the device
for rows, cols in ((3001, 5), (7, 2049), (32, 1025), (1100, 32)):
    A = port.vector_random(61, "tr", 0, rows * cols).reshape(rows, cols)
    assert np.array_equal(ctx.transpose(61, A), np.ascontiguousarray(A.T))
for n_, N_ in ((32, 1001), (5, 333), (16, 64)):
    sh_ = port.vector_random(61, "any shares", 7, N_ * n_).reshape(N_, n_)
    d_o = torch.zeros(N_, dtype=torch.int64, device="cuda")
    ctx.recover_p_dev(61, torch.from_numpy(sh_.view(np.int64)).cuda(), N_, n_, d_o, B.SECRET_MAJOR)
    torch.cuda.synchronize()
    assert np.array_equal(d_o.cpu().numpy().view(np.uint64), port.recover_p(61, sh_))
ctx.close()
print("SANITIZE_DRIVER_OK")
