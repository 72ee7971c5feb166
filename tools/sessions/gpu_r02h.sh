#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "non_canonical or degenerate or cpp_shim or packets or recover_d or async or multi_context or gather" > gpurun_out/r02h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -25 gpurun_out/r02h_pytest.log
