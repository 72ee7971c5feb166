#!/bin/bash
# A/B of the step kernel on ONE box: the library of the round's start (ab/libsclgpu_base.so) against the current build
set -u
mkdir -p gpurun_out
SO=secure-computation-library_b200/csrc/libsclgpu.so
T=${1:-r02t}
: > gpurun_out/${T}_kbench.json
cp $SO /tmp/new.so
for v in base new base new new; do
  if [ $v = base ]; then cp ab/libsclgpu_base.so $SO; else cp /tmp/new.so $SO; fi
  SCLGPU_SR_WARPS=204 timeout 200 python tools/kbench2.py 26 10 >> gpurun_out/${T}_kbench.json 2>> gpurun_out/${T}_kbench.err
done
cp /tmp/new.so $SO
cat gpurun_out/${T}_kbench.json; tail -5 gpurun_out/${T}_kbench.err
