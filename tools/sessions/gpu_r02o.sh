#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/r02o_kbench.json
for w in 4 104 108; do
SCLGPU_SR_WARPS=$w timeout 300 python tools/kbench.py 26 5 >> gpurun_out/r02o_kbench.json 2>> gpurun_out/r02o_kbench.err
done
SCLGPU_SHARE_TC=2 timeout 300 python tools/kbench.py 26 5 >> gpurun_out/r02o_kbench.json 2>> gpurun_out/r02o_kbench.err
python - <<'PY'
import json
for l in open('gpurun_out/r02o_kbench.json'):
    d=json.loads(l); print(d['env'], {k:round(v,3) for k,v in d.items() if k.endswith('_ms')}, d['fused_ok'], d['fused_indep_ok'])
PY
