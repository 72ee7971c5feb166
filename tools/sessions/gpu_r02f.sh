#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/r02f_kbench.json
for w in 8 4; do
SCLGPU_SR_WARPS=$w timeout 300 python tools/kbench.py 26 5 >> gpurun_out/r02f_kbench.json 2>> gpurun_out/r02f_kbench.err
done
cat gpurun_out/r02f_kbench.json
