#!/usr/bin/env python
"""Transposition kernels timed alone (the secret-major <-> party-major step of the host paths).  Development tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

pkg = entry.load_package(); B = pkg.binding
ctx = pkg.Context(0); ctx.use_torch_stream()

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for field, w in ((61, 1), (127, 2)):
    for rows, cols in ((32, 1 << 22), (1 << 22, 32), (16, 1 << 22), (1 << 22, 16), (5, 1 << 22)):
        a = torch.randint(0, 1 << 60, (rows * cols * w,), dtype=torch.int64, device="cuda")
        b = torch.empty_like(a)
        ms = timeit(lambda: ctx.transpose_dev(field, a, rows, cols, b))
        print(f"Fp{field} {rows} x {cols}: {ms:.3f} ms  {2 * a.numel() * 8 / ms / 1e6:.0f} GB/s")
