#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_share_recover" > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
timeout 300 python tools/kbench.py 26 5 > gpurun_out/r02b_kbench.json 2> gpurun_out/r02b_kbench.err
timeout 300 python tools/kbench.py 23 20 >> gpurun_out/r02b_kbench.json 2>> gpurun_out/r02b_kbench.err
if [ "${1:-}" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_share_recover61 -s 2 -c 1 -f -o gpurun_out/prof_r02b_fused python tools/kbench.py 26 1 > gpurun_out/r02b_ncu.log 2>&1
fi
tail -3 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_kbench.json; tail -3 gpurun_out/r02b_kbench.err
