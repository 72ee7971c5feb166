#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "recover_c" > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02v_pytest.log
timeout 600 python tools/recover_c_probe.py > gpurun_out/r02v_recover_c.txt 2>&1; tail -8 gpurun_out/r02v_recover_c.txt
SCLGPU_RECOVER_C_NOSYN=1 timeout 900 python tools/recover_c_probe.py > gpurun_out/r02v_recover_c_nosyn.txt 2>&1; tail -3 gpurun_out/r02v_recover_c_nosyn.txt
