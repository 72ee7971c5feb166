#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_recover_c_syndrome -c 1 -f -o gpurun_out/prof_recover_c_syn_r02 python tools/recover_c_probe.py > gpurun_out/ncu_recover_c_syn.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_recover_c_clean -c 1 -f -o gpurun_out/prof_recover_c_clean_r02 python tools/recover_c_probe.py > gpurun_out/ncu_recover_c_clean.log 2>&1
ls -la gpurun_out/prof_recover_c_*r02*
