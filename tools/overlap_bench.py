#!/usr/bin/env python
"""Development experiment: C2 step as a two-stream chunk pipeline -- recoverP of chunk c
(HBM-bound) runs under shamirSecretShare of chunk c+1 (LSU/ALU-bound).  Measured on B200 at 2^26,
8 and 16 chunks, with 128-thread reconstruction CTAs so that they fit beside the share kernel's CTA:
14.27 ms serial -> 13.58 ms overlapped (5 %), far from the 2.9 ms the reconstruction costs alone,
so bench.py keeps the plain two-call step.  Usage: overlap_bench.py [log2N] [chunks]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
C = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pkg = entry.load_package(); B = pkg.binding
ctx = pkg.Context(0)
N, n, t = 1 << lg, 32, 15
Nc = N // C
lo_pri, hi_pri = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
s_share = torch.cuda.Stream(priority=-1)
s_rec = torch.cuda.Stream(priority=0)
d_sec = torch.empty(N, dtype=torch.int64, device="cuda")
ctx.use_torch_stream(); ctx.random_dev(61, "secrets", 0, N, d_sec); torch.cuda.synchronize()
d_sh = [torch.empty((n, Nc), dtype=torch.int64, device="cuda") for _ in range(C)]
d_out = torch.empty(N, dtype=torch.int64, device="cuda")
evs = [torch.cuda.Event() for _ in range(C)]

def step_serial():
    ctx.set_stream(s_share.cuda_stream)
    for c in range(C):
        ctx.shamir_share_dev(61, d_sec[c * Nc:(c + 1) * Nc], Nc, t, n, "shamir bench", 8 * c * Nc, d_sh[c], B.PARTY_MAJOR)
        ctx.recover_p_dev(61, d_sh[c], Nc, n, d_out[c * Nc:(c + 1) * Nc], B.PARTY_MAJOR)

def step_overlap():
    for c in range(C):
        ctx.set_stream(s_share.cuda_stream)
        ctx.shamir_share_dev(61, d_sec[c * Nc:(c + 1) * Nc], Nc, t, n, "shamir bench", 8 * c * Nc, d_sh[c], B.PARTY_MAJOR)
        evs[c].record(s_share)
        if c >= 1:
            ctx.set_stream(s_rec.cuda_stream)
            s_rec.wait_event(evs[c - 1])
            ctx.recover_p_dev(61, d_sh[c - 1], Nc, n, d_out[(c - 1) * Nc:c * Nc], B.PARTY_MAJOR)
    ctx.set_stream(s_rec.cuda_stream)
    s_rec.wait_event(evs[C - 1])
    ctx.recover_p_dev(61, d_sh[C - 1], Nc, n, d_out[(C - 1) * Nc:], B.PARTY_MAJOR)

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    s_share.wait_stream(torch.cuda.current_stream()); s_rec.wait_stream(torch.cuda.current_stream())
    for _ in range(reps): fn()
    torch.cuda.current_stream().wait_stream(s_share); torch.cuda.current_stream().wait_stream(s_rec)
    e1.record(torch.cuda.current_stream()); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = {"log2N": lg, "chunks": C}
res["serial_ms"] = timeit(step_serial)
d_out.zero_()
res["overlap_ms"] = timeit(step_overlap)
torch.cuda.synchronize()
res["ok"] = bool(torch.equal(d_out, d_sec))
print(json.dumps(res))
