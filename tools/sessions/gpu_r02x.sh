#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/r02x_matvec.json
for v in 1 2 3 2 3; do
SCLGPU_MATVEC_VARIANT=$v timeout 120 python tools/matvec_probe.py 30 >> gpurun_out/r02x_matvec.json 2>> gpurun_out/r02x_matvec.err
done
cut -c1-200 gpurun_out/r02x_matvec.json; tail -3 gpurun_out/r02x_matvec.err
SCLGPU_MATVEC_VARIANT=3 timeout 600 python -m pytest tests -x -q -m gpu -k "matvec or c5" > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02x_pytest.log
