#!/usr/bin/env python
"""End-to-end duplex step (development tool): sclgpu_fp61_shamir_share_async(batch k) || sclgpu_fp61_recover_p(batch k-1)
through the host C ABI on pinned buffers, as bench.py's `e2e`, at 2^lg secrets.  The pipeline depth / chunk size come
from SCLGPU_HOST_PIPES / SCLGPU_HOST_CHUNK_MB (read once per process).  Usage: e2e_probe.py [log2N] [steps]"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as entry


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    pkg = entry.load_package()
    ctx = pkg.Context(0)
    lib = ctx.lib
    N, n, t = 1 << lg, 32, 15
    h_sec = ctx.host_alloc(8 * N).view(np.uint64)
    h_sh = [ctx.host_alloc(8 * N * n).view(np.uint64) for _ in range(2)]
    h_out = ctx.host_alloc(8 * N).view(np.uint64)
    d_sec = torch.empty(N, dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "secrets", 0, N, d_sec)
    h_sec[:] = d_sec.cpu().numpy().view(np.uint64)
    seed = pkg.api.seed16("shamir bench")
    p = lambda a: a.ctypes.data_as(C.c_void_p)

    def seq():
        ctx._check(lib.sclgpu_fp61_shamir_share(ctx._ctx, p(h_sec), N, t, n, seed, 0, p(h_sh[0])))
        ctx._check(lib.sclgpu_fp61_recover_p(ctx._ctx, p(h_sh[0]), N, n, None, None, p(h_out)))

    def dup(k):
        ctx._check(lib.sclgpu_fp61_shamir_share_async(ctx._ctx, p(h_sec), N, t, n, seed, 0, p(h_sh[k & 1])))
        ctx._check(lib.sclgpu_fp61_recover_p(ctx._ctx, p(h_sh[(k - 1) & 1]), N, n, None, None, p(h_out)))
        ctx._check(lib.sclgpu_wait(ctx._ctx))

    res = {"log2N": lg, "env": {k: v for k, v in os.environ.items() if k.startswith("SCLGPU_")}}
    seq()
    t0 = time.perf_counter()
    for _ in range(steps):
        seq()
    res["seq_ms"] = 1e3 * (time.perf_counter() - t0) / steps
    ok = bool(np.array_equal(h_out, h_sec))
    ctx._check(lib.sclgpu_fp61_shamir_share(ctx._ctx, p(h_sec), N, t, n, seed, 0, p(h_sh[1])))
    h_out[:] = 0
    dup(0)
    t0 = time.perf_counter()
    for k in range(1, 1 + steps):
        dup(k)
    res["dup_ms"] = 1e3 * (time.perf_counter() - t0) / steps
    ok = ok and bool(np.array_equal(h_out, h_sec))
    gb = 8 * N * (n + 1) / 1e9
    res["seq_GBps"] = 2 * gb / (res["seq_ms"] / 1e3)
    res["dup_GBps_each_way"] = gb / (res["dup_ms"] / 1e3)
    res["ok"] = ok
    print(json.dumps(res))


main()
