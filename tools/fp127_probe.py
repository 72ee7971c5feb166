#!/usr/bin/env python
"""C3 (Fp127, n=16, t=7, 2^24 secrets): share and recoverD timed a few rounds in a row.  Development tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

pkg = entry.load_package(); B = pkg.binding
ctx = pkg.Context(0); ctx.use_torch_stream()
N, t, n = 1 << 24, 7, 16
sec = torch.empty((N, 2), dtype=torch.int64, device="cuda")
sh = torch.empty((n, N, 2), dtype=torch.int64, device="cuda")
out = torch.empty((N, 2), dtype=torch.int64, device="cuda")
err = torch.empty(N, dtype=torch.uint8, device="cuda")
ctx.random_dev(127, "secrets127", 0, N, sec)

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for rnd in range(3):
    a = timeit(lambda: ctx.shamir_share_dev(127, sec, N, t, n, "shamir bench", 0, sh, B.PARTY_MAJOR))
    b = timeit(lambda: ctx.recover_d_dev(127, sh, N, n, t, out, err, B.PARTY_MAJOR))
    print("share127 %.3f ms  recoverD127 %.3f ms  ok=%s" % (a, b, bool(torch.equal(out, sec))))
