#!/usr/bin/env python
"""shamirRecoverC alone (Fp61, n=16, t=5, 2^20 sharings; every second / every 64th / no sharing with a corrupted
share), for timing and for ncu.  Development tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

pkg = entry.load_package(); B = pkg.binding
ctx = pkg.Context(0); ctx.use_torch_stream()
N, n, t = 1 << 20, 16, 5
sec = torch.empty(N, dtype=torch.int64, device="cuda")
sh = torch.empty((n, N), dtype=torch.int64, device="cuda")
ctx.random_dev(61, "secrets", 0, N, sec)
ctx.shamir_share_dev(61, sec, N, t, n, "rc", 0, sh, B.PARTY_MAJOR)
fo = torch.empty((N, 3 * t + 1), dtype=torch.int64, device="cuda")
eo = torch.empty((N, t + 1), dtype=torch.int64, device="cuda")
st = torch.empty(N, dtype=torch.uint8, device="cuda")

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for every in (2, 64, 0):
    if every: sh[3, ::every] ^= 5
    ms = timeit(lambda: ctx.recover_c_dev(61, sh, N, n, fo, eo, st, B.PARTY_MAJOR))
    ok = bool(torch.equal(fo[:, 0], sec)) and int(st.sum().item()) == 0
    print(f"corrupted share in every {every or 'no'} sharing: {ms:.3f} ms  {N / ms / 1e3:.1f} M sharings/s  ok={ok}")
    if every: sh[3, ::every] ^= 5

# several corrupted shares per sharing (up to t), and larger t
for n2, k_err in ((16, 5), (31, 10), (31, 3)):
    t2 = (n2 - 1) // 3
    N2 = 1 << 19
    sec2 = torch.empty(N2, dtype=torch.int64, device="cuda")
    sh2 = torch.empty((n2, N2), dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "secrets", 0, N2, sec2)
    ctx.shamir_share_dev(61, sec2, N2, t2, n2, "rc", 0, sh2, B.PARTY_MAJOR)
    for i in range(k_err):
        sh2[(3 * i + 1) % (3 * t2 + 1)] ^= (7 + i)
    fo2 = torch.empty((N2, 3 * t2 + 1), dtype=torch.int64, device="cuda")
    eo2 = torch.empty((N2, t2 + 1), dtype=torch.int64, device="cuda")
    st2 = torch.empty(N2, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: ctx.recover_c_dev(61, sh2, N2, n2, fo2, eo2, st2, B.PARTY_MAJOR))
    ok = bool(torch.equal(fo2[:, 0], sec2)) and int(st2.sum().item()) == 0
    print(f"n={n2} t={t2}: {k_err} corrupted shares in EVERY sharing: {ms:.3f} ms  {N2 / ms / 1e3:.1f} M sharings/s  ok={ok}")

# more than 32 points: the CTA-per-sharing kernel's range (run once with SCLGPU_RECOVER_C_NOSYN=1 for the elimination alone)
for n3, k_err, lgN in ((64, 21, 13), (64, 3, 13), (166, 55, 11)):
    t3 = (n3 - 1) // 3
    N3 = 1 << lgN
    sec3 = torch.empty(N3, dtype=torch.int64, device="cuda")
    sh3 = torch.empty((n3, N3), dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "secrets", 0, N3, sec3)
    ctx.shamir_share_dev(61, sec3, N3, t3, n3, "rc", 0, sh3, B.PARTY_MAJOR)
    for i in range(k_err):
        sh3[(3 * i + 1) % (3 * t3 + 1)] ^= (7 + i)
    fo3 = torch.empty((N3, 3 * t3 + 1), dtype=torch.int64, device="cuda")
    eo3 = torch.empty((N3, t3 + 1), dtype=torch.int64, device="cuda")
    st3 = torch.empty(N3, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: ctx.recover_c_dev(61, sh3, N3, n3, fo3, eo3, st3, B.PARTY_MAJOR), reps=2)
    ok = bool(torch.equal(fo3[:, 0], sec3)) and int(st3.sum().item()) == 0
    print(f"n={n3} t={t3}: {k_err} corrupted shares in EVERY one of 2^{lgN} sharings: {ms:.3f} ms  {N3 / ms:.1f} k sharings/s  ok={ok}"
          f"  nosyn={os.environ.get('SCLGPU_RECOVER_C_NOSYN', '0')}")
