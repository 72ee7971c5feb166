#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02_full_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_full_pytest.log
tail -12 gpurun_out/r02_full_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
