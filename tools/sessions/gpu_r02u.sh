#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/r02u_kbench.json
for w in 204 204; do
SCLGPU_SR_WARPS=$w timeout 200 python tools/kbench2.py 26 10 >> gpurun_out/r02u_kbench.json 2>> gpurun_out/r02u_kbench.err
done
timeout 200 python tools/kbench.py 26 8 >> gpurun_out/r02u_kbench.json 2>> gpurun_out/r02u_kbench.err
cat gpurun_out/r02u_kbench.json; tail -5 gpurun_out/r02u_kbench.err
timeout 900 python -m pytest tests -x -q -m gpu -k "share or prg or random or additive or array or full_size or knob or path" > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02u_pytest.log
SCLGPU_HOST_CHUNK_MB=512 timeout 300 python tools/e2e_probe.py 25 3 > gpurun_out/r02u_e2e.json 2>> gpurun_out/r02u_e2e.err
timeout 300 python tools/e2e_probe.py 25 3 >> gpurun_out/r02u_e2e.json 2>> gpurun_out/r02u_e2e.err
cat gpurun_out/r02u_e2e.json
timeout 300 ncu --set full --clock-control none -k regex:"k_matvec61" -c 4 --launch-skip 6 -o gpurun_out/r02u_matvec python tools/matvec_probe.py 2 > gpurun_out/r02u_ncu.log 2>&1
ncu -i gpurun_out/r02u_matvec.ncu-rep --page details --csv 2>/dev/null | grep -E "Duration|DRAM Throughput|Registers Per|Achieved Occupancy" | cut -d, -f5,13- | head -20
