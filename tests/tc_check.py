#!/usr/bin/env python
"""TEST HELPER (run by tests/test_gpu_parity.py in a subprocess): the active share kernels against the plain-C oracle on a
sweep of (t, n, N), device-pointer path, both layouts.  Usage: tc_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as entry

pkg = entry.load_package(); B = pkg.binding
o = entry.load_oracle(); port = o.PortOracle()
ctx = pkg.Context(0); ctx.use_torch_stream()
bad = 0
CASES = [(61, t, n, N) for t, n, N in [(15, 32, 128), (15, 32, 1), (15, 32, 5000), (2, 5, 1000), (0, 1, 7), (1, 3, 129),
                                       (7, 16, 4097), (15, 32, 1 << 17), (14, 31, 12345), (3, 9, 384), (15, 17, 999),
                                       (8, 24, 100000), (15, 32, 1 << 20)]]
CASES += [(127, t, n, N) for t, n, N in [(7, 16, 128), (7, 16, 1), (7, 16, 5000), (2, 5, 1000), (0, 1, 7), (1, 3, 129),
                                         (7, 16, 1 << 17), (6, 15, 12345), (3, 9, 384), (7, 9, 999), (4, 12, 100000)]]
for field, t, n, N in CASES:
    w = 1 if field == 61 else 2
    secrets = port.vector_random(field, "secrets", 0, N)
    first = 1000 + 250 * t
    want = port.shamir_share(field, secrets, t, n, "shamir bench", first).reshape(N, n * w)
    d_sec = torch.from_numpy(secrets.view(np.int64)).cuda()
    d_pm = torch.zeros((n, N, w), dtype=torch.int64, device="cuda")
    ctx.shamir_share_dev(field, d_sec, N, t, n, "shamir bench", first, d_pm, B.PARTY_MAJOR)
    torch.cuda.synchronize()
    got = np.ascontiguousarray(np.swapaxes(d_pm.cpu().numpy().view(np.uint64), 0, 1)).reshape(N, n * w)
    ok = np.array_equal(got, want)
    if not ok:
        bad += 1
        diff = np.argwhere(got != want)
        print(f"MISMATCH field={field} t={t} n={n} N={N}: {len(diff)} of {got.size} differ; first at (secret,party)={diff[0]}"
              f" got={int(got[tuple(diff[0])]):x} want={int(want[tuple(diff[0])]):x}")
        js = sorted(set(int(d[0]) for d in diff))[:10]; ps = sorted(set(int(d[1]) for d in diff))[:40]
        print("   secrets:", js, " parties:", ps)
    else:
        print(f"ok field={field} t={t} n={n} N={N}")
print("TC_CHECK", "FAILED" if bad else "PASSED")
sys.exit(1 if bad else 0)
