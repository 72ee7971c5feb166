#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/r02w_matvec.json
for v in 0 1 2 0 1 2; do
SCLGPU_MATVEC_VARIANT=$v timeout 120 python tools/matvec_probe.py 30 >> gpurun_out/r02w_matvec.json 2>> gpurun_out/r02w_matvec.err
done
cat gpurun_out/r02w_matvec.json; tail -3 gpurun_out/r02w_matvec.err
timeout 600 python -m pytest tests -x -q -m gpu -k "matvec or c5 or vec_ops" > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02w_pytest.log
SCLGPU_MATVEC_VARIANT=2 timeout 600 python -m pytest tests -x -q -m gpu -k "matvec or c5" >> gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02w_pytest.log
