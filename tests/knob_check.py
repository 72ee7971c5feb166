#!/usr/bin/env python
"""TEST HELPER (run by tests/test_gpu_parity.py in subprocesses with different SCLGPU_* knobs set): a compact sweep
of share / recoverP / recoverD / mat-mul against the plain-C oracle, so that every selectable kernel stays verified."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry

pkg = entry.load_package()
o = entry.load_oracle(); port = o.PortOracle()
ctx = pkg.Context(0)
for field, t, n, N in [(61, 15, 32, 3000), (61, 2, 5, 777), (127, 7, 16, 1500), (127, 2, 7, 300), (61, 7, 16, 2048),
                       (61, 15, 32, 21001)]:   # the last one: several chunks in the host pipelines when SCLGPU_HOST_CHUNK_MB=1
    sec = port.vector_random(field, "secrets", 0, N)
    sh = ctx.shamir_share(field, sec, t, n, "knobs", 11)
    assert np.array_equal(sh, port.shamir_share(field, sec, t, n, "knobs", 11)), ("share", field, t, n)
    assert np.array_equal(ctx.recover_p(field, sh), sec), ("recover_p", field, n)
    if n >= 2 * t + 1:
        bad = sh.copy()
        bad.reshape(N, n, -1)[::7, t + 1, 0] ^= np.uint64(3)
        g, w = ctx.recover_d(field, bad, t), port.recover_d(field, bad, t)
        assert np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]) and g[2] == w[2], ("recover_d", field, t)
for field, W, t, n, N in [(61, 2, 15, 32, 1500), (61, 4, 2, 5, 301), (127, 2, 7, 16, 700), (127, 3, 1, 4, 129)]:   # array-valued secrets
    sec = port.vector_random(field, "pairs", 0, N * W).reshape((N, W) + (() if field == 61 else (2,)))
    sh = ctx.shamir_share_array(field, sec, t, n, "knobs", 5)
    assert np.array_equal(sh, port.shamir_share_array(field, sec, t, n, "knobs", 5)), ("share_array", field, W, t, n)
    assert np.array_equal(ctx.recover_p_array(field, sh), sec), ("recover_p_array", field, W, n)
for field, n, N in [(61, 16, 600), (127, 7, 200)]:   # shamirRecoverC: 0..t corrupted shares per sharing
    t = (n - 1) // 3
    sec = port.vector_random(field, "secrets", 0, N)
    sh = port.shamir_share(field, sec, t, n, "rc", 3).copy()
    flat = sh.reshape(N, n, -1)
    for j in range(N):
        for i in range(j % (t + 1)):
            flat[j, (5 * i + j) % (3 * t + 1), 0] ^= np.uint64(1 + j)
    g, w = ctx.recover_c(field, sh), port.recover_c(field, sh)
    assert all(np.array_equal(a, b) for a, b in zip(g[:3], w[:3])) and g[3] == w[3] == 0, ("recover_c", field, n)
for field, rows, inner, cols in [(61, 300, 520, 70), (61, 129, 4100, 33), (61, 257, 130, 257), (127, 130, 520, 40), (127, 64, 2100, 20)]:
    shp = () if field == 61 else (2,)
    A = port.vector_random(field, "mat A", 0, rows * inner).reshape((rows, inner) + shp)
    Bm = port.vector_random(field, "mat B", 7, inner * cols).reshape((inner, cols) + shp)
    assert np.array_equal(ctx.matmul(field, A, Bm), port.matmul(field, A, Bm)), ("matmul", field, rows, inner, cols)
for rows, cols in [(512, 8192), (1024, 4608), (700, 6144)]:   # long rows: the chunked sweep (8 KiB and 4 KiB chunks) + finish
    A = port.vector_random(61, "mat A", 0, rows * cols).reshape(rows, cols)
    x = port.vector_random(61, "vec x", 0, cols)
    assert np.array_equal(ctx.matvec(61, A, x), port.matvec(61, A, x)), ("matvec", rows, cols)
for seed, first, nbytes in [("prg bench", 0, 1 << 16), ("", 5, 16 * 1000), ("shamir bench", (1 << 32) - 40, 16 * 200), ("k", 31, 48),
                            ("k", 7, 1001)]:   # util::PRG keystream (with SCLGPU_PRG_BITSLICED: the bitsliced kernel)
    assert np.array_equal(ctx.prg_expand(seed, first, nbytes), port.prg_next(seed, first, nbytes)), ("prg", seed, first, nbytes)
# the single-launch step (k_share_recover61; form selected by SCLGPU_SR_WARPS): same-batch and pipelined mode
import torch
ctx.use_torch_stream()
for field, N, t, n, first in [(61, 4100, 15, 32, 8), (61, 999, 7, 16, 5), (61, 300, 2, 5, 0), (127, 700, 7, 16, 3)]:   # shares in SCL's layout on the device
    sec = port.vector_random(field, "secrets", 0, N)
    want = port.shamir_share(field, sec, t, n, "sm", first)
    w = 1 if field == 61 else 2
    d_sm = torch.zeros((N, n, w), dtype=torch.int64, device="cuda")
    ctx.shamir_share_dev(field, torch.from_numpy(sec.view(np.int64)).cuda(), N, t, n, "sm", first, d_sm, pkg.binding.SECRET_MAJOR)
    torch.cuda.synchronize()
    assert np.array_equal(d_sm.cpu().numpy().view(np.uint64).reshape(want.shape), want), ("share_dev secret-major", field, t, n)
for N, t, n, first in [(5000, 15, 32, 16), (777, 7, 16, 3), (130, 2, 5, 0)]:
    sec = port.vector_random(61, "secrets", 0, N)
    want = port.shamir_share(61, sec, t, n, "step", first)
    d_sec = torch.from_numpy(sec.view(np.int64)).cuda()
    d_sh = torch.zeros((n, N), dtype=torch.int64, device="cuda")
    d_sh2 = torch.zeros((n, N), dtype=torch.int64, device="cuda")
    d_out = torch.zeros(N, dtype=torch.int64, device="cuda")
    ctx.shamir_share_recover_dev(d_sec, N, t, n, "step", first, d_sh, d_out)                       # reconstructs its own tiles
    torch.cuda.synchronize()
    assert np.array_equal(d_sh.cpu().numpy().view(np.uint64).T, want) and np.array_equal(d_out.cpu().numpy().view(np.uint64), sec), ("step", N, t, n)
    d_out.zero_()
    ctx.shamir_share_recover_dev(d_sec, N, t, n, "step", first, d_sh2, d_out, rec_shares=d_sh)     # reconstructs another batch
    torch.cuda.synchronize()
    assert np.array_equal(d_sh2.cpu().numpy().view(np.uint64).T, want) and np.array_equal(d_out.cpu().numpy().view(np.uint64), sec), ("step2", N, t, n)
ctx.close()
print("KNOB_CHECK PASSED", {k: v for k, v in os.environ.items() if k.startswith("SCLGPU_")})
