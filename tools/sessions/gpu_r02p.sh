#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or full_size_c2" > gpurun_out/r02p_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r02p_pytest.log
: > gpurun_out/r02p_kbench.json
for w in 204 4; do
SCLGPU_SR_WARPS=$w timeout 200 python tools/kbench.py 26 5 >> gpurun_out/r02p_kbench.json 2>> gpurun_out/r02p_kbench.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02p_kbench.json'):
    d=json.loads(l); print(d['env'], {k:round(v,3) for k,v in d.items() if k.endswith('_ms')}, d['fused_ok'], d['fused_indep_ok'])
PY
tail -3 gpurun_out/r02p_kbench.err
