#!/usr/bin/env python
"""Keystream kernels side by side (development tool): T-table k_prg_bytes vs bitsliced k_prg_bitsliced, 2 GiB (C4)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

pkg = entry.load_package()
ctx = pkg.Context(0); ctx.use_torch_stream()
nb = 1 << 31
buf = torch.empty(nb, dtype=torch.uint8, device="cuda")
ref = torch.empty(nb, dtype=torch.uint8, device="cuda")

def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

t_tab = timeit(lambda: ctx.prg_expand_dev("prg bench", 0, nb, ref))
t_bs = timeit(lambda: ctx.prg_expand_bitsliced_dev("prg bench", 0, nb, buf))
print(json.dumps({"t_table_ms": t_tab, "bitsliced_ms": t_bs, "ratio": t_bs / t_tab, "equal": bool(torch.equal(buf, ref)),
                  "t_table_Gblocks_s": nb / 16 / t_tab / 1e6, "bitsliced_Gblocks_s": nb / 16 / t_bs / 1e6}))
