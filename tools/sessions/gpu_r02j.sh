#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "vandermonde or matvec or cpp_shim or poly" > gpurun_out/r02j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -15 gpurun_out/r02j_pytest.log
python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import __graft_entry__ as entry
pkg = entry.load_package(); ctx = pkg.Context(0); ctx.use_torch_stream()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for rows, cols in ((8192, 8192), (4096, 16384), (16384, 4096), (1024, 8192)):
    A = torch.empty((rows, cols), dtype=torch.int64, device="cuda"); x = torch.empty(cols, dtype=torch.int64, device="cuda"); y = torch.empty(rows, dtype=torch.int64, device="cuda")
    ctx.random_dev(61, "mat A", 0, rows * cols, A); ctx.random_dev(61, "vec x", 0, cols, x)
    ms = timeit(lambda: ctx.matvec_dev(61, A, rows, cols, x, y))
    print(json.dumps({"rows": rows, "cols": cols, "ms": ms, "GBps": 8 * rows * cols / ms / 1e6}))
PY
