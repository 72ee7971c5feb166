#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp
instructions, share, and stall samples.  Usage: ncu_opmix.py file.csv [units]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); smp = collections.Counter(); stalls = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDS", "STG", "LDG", "SHF", "IADD3", "LOP3")) else op.split(".")[0]
    n = int(float(r[ix["Instructions Executed"]] or 0))
    ops[op] += n; tot += n
    smp[op] += int(float(r[ix["# Samples"]] or 0))
    for c in stall_cols:
        stalls[c] += int(float(r[ix[c]] or 0))
print(f"total warp instr {tot:,}" + (f"  = {tot*32/units:.1f} thread-instr per unit" if units else ""))
for op, n in ops.most_common(25):
    print(f"{op:18s} {n:14,d} {100*n/tot:5.1f}%  samples {smp[op]:7d}" + (f"  {n*32/units:8.1f}/unit" if units else ""))
ts = sum(stalls.values())
print("stalls:", ", ".join(f"{k[6:]} {100*v/ts:.1f}%" for k, v in stalls.most_common(8)))
