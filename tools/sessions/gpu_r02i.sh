#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "recover_c" > gpurun_out/r02i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -25 gpurun_out/r02i_pytest.log
