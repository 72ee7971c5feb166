#!/bin/bash
# A/B of the step kernel on ONE box: libraries under ab/ (git-ignored builds of variants) against the current build
set -u
mkdir -p gpurun_out
SO=secure-computation-library_b200/csrc/libsclgpu.so
T=$1; shift
: > gpurun_out/${T}_kbench.json
cp $SO /tmp/new.so
for v in "$@"; do
  if [ $v = new ]; then cp /tmp/new.so $SO; else cp ab/libsclgpu_$v.so $SO; fi
  echo "{\"variant\": \"$v\"}" >> gpurun_out/${T}_kbench.json
  SCLGPU_SR_WARPS=204 timeout 200 python tools/kbench2.py 26 10 >> gpurun_out/${T}_kbench.json 2>> gpurun_out/${T}_kbench.err
done
cp /tmp/new.so $SO
python - <<PY
import json
v=None
for l in open('gpurun_out/${T}_kbench.json'):
    d=json.loads(l)
    if 'variant' in d: v=d['variant']; continue
    print(v, round(d['step_ms'],3), d['ok'])
PY
tail -3 gpurun_out/${T}_kbench.err
