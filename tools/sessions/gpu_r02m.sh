#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "recover_c or cpp_shim or plain_c" > gpurun_out/r02m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -12 gpurun_out/r02m_pytest.log
echo "--- syndrome decoder on"; timeout 600 python tools/recover_c_probe.py 2>&1 | tail -8
echo "--- syndrome decoder off"; SCLGPU_RECOVER_C_NOSYN=1 timeout 600 python tools/recover_c_probe.py 2>&1 | tail -8
