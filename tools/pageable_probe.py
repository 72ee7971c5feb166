#!/usr/bin/env python
"""Host entry points with pageable buffers (what an SCL user's std::vector is) against pinned ones.
SCLGPU_HOST_STAGING=0 in the environment leaves pageable buffers to the driver.  Development tool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as entry

pkg = entry.load_package()
ctx = pkg.Context(0)
N, t, n = 1 << 24, 15, 32
lib = pkg.binding.load()
def run(sec, sh, out, tag):
    for rnd in range(3):
        t0 = time.perf_counter()
        rc = lib.sclgpu_fp61_shamir_share(ctx._ctx, sec.ctypes.data, N, t, n, pkg.api.seed16("x"), 0, sh.ctypes.data)
        t1 = time.perf_counter()
        rc |= lib.sclgpu_fp61_recover_p(ctx._ctx, sh.ctypes.data, N, n, None, None, out.ctypes.data)
        t2 = time.perf_counter()
        assert rc == 0
    gb = 8 * N * n / 1e9
    print(f"{tag} staging={os.environ.get('SCLGPU_HOST_STAGING', '1')}: share {t1 - t0:.3f} s ({gb / (t1 - t0):.1f} GB/s)  "
          f"recover {t2 - t1:.3f} s ({gb / (t2 - t1):.1f} GB/s)  ok={np.array_equal(out, sec)}")
    return sh

sec_p = ctx.host_alloc(8 * N).view(np.uint64); sh_p = ctx.host_alloc(8 * N * n).view(np.uint64); out_p = ctx.host_alloc(8 * N).view(np.uint64)
sec_p[:] = np.arange(N, dtype=np.uint64) * 977 + 5
run(sec_p, sh_p, out_p, "pinned  ")
sec = np.array(sec_p); sh = np.zeros(N * n, dtype=np.uint64); out = np.zeros(N, dtype=np.uint64)
run(sec, sh, out, "pageable")
print("same shares:", np.array_equal(sh, sh_p))
