#!/usr/bin/env python
"""Step-kernel timing helper (development tool): the pipelined one-launch step (share batch k, reconstruct batch k-1)
for the variant in SCLGPU_SR_WARPS, CUDA events, verified against the secrets.  Usage: kbench2.py [log2N] [reps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # PRG offset: a multiple of 8 keeps every warp on the two-block loop
    pkg = entry.load_package(); B = pkg.binding
    ctx = pkg.Context(0); ctx.use_torch_stream()
    N, n, t = 1 << lg, 32, 15
    d_sec = torch.empty(N, dtype=torch.int64, device="cuda")
    d_sh = [torch.empty((n, N), dtype=torch.int64, device="cuda") for _ in range(2)]
    d_out = torch.empty(N, dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "secrets", 0, N, d_sec)
    ctx.shamir_share_dev(61, d_sec, N, t, n, "shamir bench", first, d_sh[1], B.PARTY_MAJOR)
    ref = d_sh[1].clone()
    step = lambda k: ctx.shamir_share_recover_dev(d_sec, N, t, n, "shamir bench", first, d_sh[k & 1], d_out, rec_shares=d_sh[(k - 1) & 1])
    for k in range(3): step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_out.zero_()
    e0.record()
    for k in range(3, 3 + reps): step(k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ok = bool(torch.equal(d_out, d_sec)) and bool(torch.equal(d_sh[0], ref)) and bool(torch.equal(d_sh[1], ref))
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("SCLGPU_")}, "step_ms": ms, "ok": ok, "first_block": first,
                      "frac_of_18.54T": 2048 * N / (ms * 1e-3) / 18.54e12}))
main()
