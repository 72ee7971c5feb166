#!/usr/bin/env python
"""Matrix::multiply(Matrix) timing sweep (device-resident, CUDA events).
Usage: matmul_bench.py [--fp127] dims..."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
pkg = entry.load_package(); ctx = pkg.Context(0); ctx.use_torch_stream()
field = 127 if "--fp127" in sys.argv else 61
w = 1 if field == 61 else 2
dims = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [2048, 3072, 4096, 4160, 5120, 6144, 8192]
for dim in dims:
    A = torch.empty((dim, dim, w), dtype=torch.int64, device="cuda"); B = torch.empty_like(A); C = torch.empty_like(A)
    ctx.random_dev(field, "mat A", 0, dim * dim, A); ctx.random_dev(field, "mat B", 0, dim * dim, B)
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.matmul_dev(field, A, dim, dim, B, dim, C); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    print(json.dumps({"field": field, "dim": dim, "ms": ts, "T_mults_per_s": round(dim ** 3 / min(ts) / 1e9, 2)}))
    del A, B, C; torch.cuda.empty_cache()
