#!/bin/bash
set -u
mkdir -p gpurun_out
SO=secure-computation-library_b200/csrc/libsclgpu.so
: > gpurun_out/r02t_kbench.json
cp $SO /tmp/new.so
cp ab/libsclgpu_base.so $SO
SCLGPU_SR_WARPS=204 timeout 200 python tools/kbench2.py 26 10 >> gpurun_out/r02t_kbench.json 2>> gpurun_out/r02t_kbench.err
cp /tmp/new.so $SO
for w in 204 305 315 304 204 305; do
SCLGPU_SR_WARPS=$w timeout 200 python tools/kbench2.py 26 10 >> gpurun_out/r02t_kbench.json 2>> gpurun_out/r02t_kbench.err
done
cat gpurun_out/r02t_kbench.json; tail -5 gpurun_out/r02t_kbench.err
