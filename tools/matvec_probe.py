#!/usr/bin/env python
"""Mat-vec / dot timing at equal bytes (development tool): 8192 x 8192 Fp61 mat-vec (512 MiB read once) beside a dot
product of two 2^25-element vectors (512 MiB read once) and the 2^26 one.  Usage: matvec_probe.py [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    pkg = entry.load_package()
    ctx = pkg.Context(0)
    ctx.use_torch_stream()
    R = Cc = 8192
    A = torch.empty(R * Cc, dtype=torch.int64, device="cuda")
    x = torch.empty(Cc, dtype=torch.int64, device="cuda")
    y = torch.empty(R, dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "mat A", 0, R * Cc, A)
    ctx.random_dev(61, "vec x", 0, Cc, x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    res = {"env": {k: v for k, v in os.environ.items() if k.startswith("SCLGPU_")}}
    med, best = timeit(lambda: ctx.matvec_dev(61, A, R, Cc, x, y))
    res["matvec_ms"], res["matvec_best_ms"] = med, best
    res["matvec_GBps"] = R * Cc * 8 / med / 1e6
    half = R * Cc // 2
    d1 = torch.empty(1, dtype=torch.int64, device="cuda")
    med, best = timeit(lambda: ctx.vec_op_dev(61, 4, A[:half], A[half:], half, d1))
    res["dot_2^25_ms"], res["dot_2^25_GBps"] = med, R * Cc * 8 / med / 1e6
    print(json.dumps(res))


main()
