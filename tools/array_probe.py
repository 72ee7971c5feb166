#!/usr/bin/env python
"""Plain share (2^26 secrets) against the array share (2^25 pairs, W = 2) back to back: same number of component
polynomials, same keystream volume.  Development tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

pkg = entry.load_package(); B = pkg.binding
ctx = pkg.Context(0); ctx.use_torch_stream()
N, t, n = 1 << 26, 15, 32
sec = torch.empty((N,), dtype=torch.int64, device="cuda")
sh = torch.empty((n, N), dtype=torch.int64, device="cuda")
ctx.random_dev(61, "secrets", 0, N, sec)

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for rnd in range(2):
    print("plain  ", timeit(lambda: ctx.shamir_share_dev(61, sec, N, t, n, "x", 0, sh, B.PARTY_MAJOR)))
    for W in (2, 4, 8):
        print("array W=%d" % W, timeit(lambda: ctx.shamir_share_array_dev(61, sec, N // W, W, t, n, "x", 0, sh, B.PARTY_MAJOR)))

# secret-major ([N][n][W], SCL's layout): chunked share + wide transposition
W = 2
sm = torch.empty((N // W) * n * W, dtype=torch.int64, device="cuda")
out = torch.empty(N, dtype=torch.int64, device="cuda")
for lg in (22, 25):
    Np = 1 << lg
    a = timeit(lambda: ctx.shamir_share_array_dev(61, sec, Np, W, t, n, "x", 0, sm, B.SECRET_MAJOR), reps=3)
    b = timeit(lambda: ctx.recover_p_array_dev(61, sm, Np, W, n, out, B.SECRET_MAJOR), reps=3)
    print(f"secret-major 2^{lg} pairs: share {a:.2f} ms, recover {b:.2f} ms, ok={bool(torch.equal(out[:Np * W], sec[:Np * W]))}")
