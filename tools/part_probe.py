#!/usr/bin/env python
"""SM-partition probe (development tool): share on S SMs and reconstruction on R SMs, alone and concurrently on two
streams.  Env: SCLGPU_SHARE_SMS, SCLGPU_RECOVER_SMS, SCLGPU_RECOVER61_TC."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    reps = 5
    pkg = entry.load_package(); B = pkg.binding
    ctx = pkg.Context(0); ctx2 = pkg.Context(0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ctx.set_stream(s1.cuda_stream); ctx2.set_stream(s2.cuda_stream)
    N, n, t = 1 << lg, 32, 15
    d_sec = torch.empty(N, dtype=torch.int64, device="cuda")
    d_sh = torch.empty((n, N), dtype=torch.int64, device="cuda")
    d_sh2 = torch.empty((n, N), dtype=torch.int64, device="cuda")
    d_out = torch.empty(N, dtype=torch.int64, device="cuda")
    with torch.cuda.stream(s1):
        ctx.random_dev(61, "secrets", 0, N, d_sec)
        ctx.shamir_share_dev(61, d_sec, N, t, n, "shamir bench", 0, d_sh2, B.PARTY_MAJOR)
    torch.cuda.synchronize()
    res = {"env": {k: v for k, v in os.environ.items() if k.startswith("SCLGPU_")}}
    def timed(fn):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s1)
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        e1.record(s1); torch.cuda.synchronize()
        return None
    import time
    def wall(fn):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0) / reps
    res["share_ms"] = wall(lambda: ctx.shamir_share_dev(61, d_sec, N, t, n, "shamir bench", 0, d_sh, B.PARTY_MAJOR))
    res["recover_ms"] = wall(lambda: ctx2.recover_p_dev(61, d_sh2, N, n, d_out, B.PARTY_MAJOR))
    res["recover_ok"] = bool(torch.equal(d_out, d_sec))
    def both():
        ctx.shamir_share_dev(61, d_sec, N, t, n, "shamir bench", 0, d_sh, B.PARTY_MAJOR)
        ctx2.recover_p_dev(61, d_sh2, N, n, d_out, B.PARTY_MAJOR)
    res["concurrent_ms"] = wall(both)
    print(json.dumps(res))
main()
