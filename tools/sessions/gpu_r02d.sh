#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/r02d_part.json
run() { env "$@" timeout 300 python tools/part_probe.py 26 >> gpurun_out/r02d_part.json 2>> gpurun_out/r02d_part.err; }
run SCLGPU_NOOP=1
run SCLGPU_RECOVER61_TC=1
run SCLGPU_SHARE_SMS=134 SCLGPU_RECOVER_SMS=14 SCLGPU_RECOVER61_TC=1
run SCLGPU_SHARE_SMS=132 SCLGPU_RECOVER_SMS=16 SCLGPU_RECOVER61_TC=1
run SCLGPU_SHARE_SMS=128 SCLGPU_RECOVER_SMS=20 SCLGPU_RECOVER61_TC=1
run SCLGPU_SHARE_SMS=128 SCLGPU_RECOVER_SMS=20
run SCLGPU_SHARE_SMS=124 SCLGPU_RECOVER_SMS=24
cat gpurun_out/r02d_part.json; tail -3 gpurun_out/r02d_part.err
