#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/r02c_kbench.json
for d in 0 1 2; do
SCLGPU_SR_DBG=$d timeout 300 python tools/kbench.py 26 5 >> gpurun_out/r02c_kbench.json 2>> gpurun_out/r02c_kbench.err
done
cat gpurun_out/r02c_kbench.json
