#!/usr/bin/env python
"""Development check of the active Fp61 share kernel against the plain-C oracle on a
sweep of (t, n, N), device-pointer path, both layouts.  Usage: tc_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as entry

pkg = entry.load_package(); B = pkg.binding
o = entry.load_oracle(); port = o.PortOracle()
ctx = pkg.Context(0); ctx.use_torch_stream()
bad = 0
for t, n, N in [(15, 32, 128), (15, 32, 1), (15, 32, 5000), (2, 5, 1000), (0, 1, 7), (1, 3, 129), (7, 16, 4097),
                (15, 32, 1 << 17), (14, 31, 12345), (3, 9, 384), (15, 17, 999), (8, 24, 100000), (15, 32, 1 << 20)]:
    secrets = port.vector_random(61, "secrets", 0, N)
    first = 1000 + 250 * t
    want = port.shamir_share(61, secrets, t, n, "shamir bench", first)
    d_sec = torch.from_numpy(secrets.view(np.int64)).cuda()
    d_pm = torch.zeros((n, N), dtype=torch.int64, device="cuda")
    ctx.shamir_share_dev(61, d_sec, N, t, n, "shamir bench", first, d_pm, B.PARTY_MAJOR)
    torch.cuda.synchronize()
    got = d_pm.cpu().numpy().view(np.uint64).T
    ok = np.array_equal(got, want)
    if not ok:
        bad += 1
        diff = np.argwhere(got != want)
        print(f"MISMATCH t={t} n={n} N={N}: {len(diff)} of {got.size} differ; first at (secret,party)={diff[0]}"
              f" got={int(got[tuple(diff[0])]):x} want={int(want[tuple(diff[0])]):x}")
        js = sorted(set(int(d[0]) for d in diff))[:10]; ps = sorted(set(int(d[1]) for d in diff))[:40]
        print("   secrets:", js, " parties:", ps)
    else:
        print(f"ok t={t} n={n} N={N}")
print("TC_CHECK", "FAILED" if bad else "PASSED")
sys.exit(1 if bad else 0)
