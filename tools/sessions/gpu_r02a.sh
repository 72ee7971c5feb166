#!/bin/bash
# round-2 GPU session A: fused share+recover kernel -- parity, timing, ncu
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_share_recover or share_recover_vs_oracle or smoke" > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 300 python tools/kbench.py 26 5 > gpurun_out/r02a_kbench.json 2> gpurun_out/r02a_kbench.err
timeout 300 python tools/kbench.py 23 20 >> gpurun_out/r02a_kbench.json 2>> gpurun_out/r02a_kbench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_share_recover61 -s 2 -c 1 -f -o gpurun_out/prof_r02a_fused python tools/kbench.py 26 1 > gpurun_out/r02a_ncu.log 2>&1
tail -5 gpurun_out/r02a_pytest.log; cat gpurun_out/r02a_kbench.json; tail -3 gpurun_out/r02a_kbench.err
