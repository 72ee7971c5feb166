#!/bin/bash
set -u
mkdir -p gpurun_out
SO=secure-computation-library_b200/csrc/libsclgpu.so
cp $SO /tmp/new.so
: > gpurun_out/r03e_kbench.json
for v in head new head new; do
  if [ $v = new ]; then cp /tmp/new.so $SO; else cp ab/libsclgpu_$v.so $SO; fi
  timeout 300 python tools/kbench.py 26 8 >> gpurun_out/r03e_kbench.json 2>> gpurun_out/r03e_kbench.err
done
cp /tmp/new.so $SO
python - <<'PY'
import json
for l in open('gpurun_out/r03e_kbench.json'):
    d=json.loads(l); print({k:round(v,3) for k,v in d.items() if k.endswith('_ms')}, d['ok'], d['fused_ok'], d['fused_indep_ok'])
PY
timeout 300 python tools/fp127_probe.py | tail -2
timeout 1200 python -m pytest tests -x -q -m gpu -k "share or prg or random or additive or array or full_size or selectable or path or smoke or multi" > gpurun_out/r03e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r03e_pytest.log
