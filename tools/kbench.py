#!/usr/bin/env python
"""Kernel-level timing helper (development tool): times the device-pointer entry
points with CUDA events.  Usage: kbench.py [log2N] [reps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    pkg = entry.load_package(); B = pkg.binding
    ctx = pkg.Context(0); ctx.use_torch_stream()
    N, n, t = 1 << lg, 32, 15
    d_sec = torch.empty(N, dtype=torch.int64, device="cuda")
    d_sh = torch.empty((n, N), dtype=torch.int64, device="cuda")
    d_out = torch.empty(N, dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "secrets", 0, N, d_sec)
    def timeit(fn, reps=reps):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    res = {"log2N": lg, "env": {k: v for k, v in os.environ.items() if k.startswith("SCLGPU_")}}
    res["share_ms"] = timeit(lambda: ctx.shamir_share_dev(61, d_sec, N, t, n, "shamir bench", 0, d_sh, B.PARTY_MAJOR))
    res["recover_ms"] = timeit(lambda: ctx.recover_p_dev(61, d_sh, N, n, d_out, B.PARTY_MAJOR))
    res["ok"] = bool(torch.equal(d_out, d_sec))
    d_out.zero_()
    res["fused_step_ms"] = timeit(lambda: ctx.shamir_share_recover_dev(d_sec, N, t, n, "shamir bench", 0, d_sh, d_out))
    res["fused_ok"] = bool(torch.equal(d_out, d_sec))
    d_sh2 = d_sh.clone()
    d_out.zero_()
    res["fused_indep_ms"] = timeit(lambda: ctx.shamir_share_recover_dev(d_sec, N, t, n, "shamir bench", 0, d_sh, d_out, rec_shares=d_sh2))
    res["fused_indep_ok"] = bool(torch.equal(d_out, d_sec))
    del d_sh2
    res["random_ms"] = timeit(lambda: ctx.random_dev(61, "prg bench", 0, N, d_out))
    print(json.dumps(res))
main()
