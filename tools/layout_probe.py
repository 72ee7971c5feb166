#!/usr/bin/env python
"""Device-pointer entry points in SCL's secret-major layout [N][n] against party-major planes (development tool):
share and recoverP at 2^log2N secrets, n=32, t=15."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 25
pkg = entry.load_package(); B = pkg.binding
ctx = pkg.Context(0); ctx.use_torch_stream()
N, n, t = 1 << lg, 32, 15
sec = torch.empty(N, dtype=torch.int64, device="cuda")
pm = torch.empty((n, N), dtype=torch.int64, device="cuda")
sm = torch.empty((N, n), dtype=torch.int64, device="cuda")
out = torch.empty(N, dtype=torch.int64, device="cuda")
ctx.random_dev(61, "secrets", 0, N, sec)

def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for name, buf, lay in (("party-major", pm, B.PARTY_MAJOR), ("secret-major", sm, B.SECRET_MAJOR)):
    a = timeit(lambda: ctx.shamir_share_dev(61, sec, N, t, n, "shamir bench", 0, buf, lay))
    b = timeit(lambda: ctx.recover_p_dev(61, buf, N, n, out, lay))
    print(f"{name}: share {a:.3f} ms  recoverP {b:.3f} ms  ok={bool(torch.equal(out, sec))}")
print("layouts agree:", bool(torch.equal(pm.t().contiguous(), sm)))
