#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/prg_probe.py > gpurun_out/r02y_prg.json 2> gpurun_out/r02y_prg.err; cat gpurun_out/r02y_prg.json; tail -3 gpurun_out/r02y_prg.err
timeout 1200 python -m pytest tests -x -q -m gpu -k "prg or selectable or matvec" > gpurun_out/r02y_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02y_pytest.log
timeout 300 ncu --set full --clock-control none -k regex:"k_prg_bitsliced" -c 1 --launch-skip 2 -o gpurun_out/r02y_bitsliced python tools/prg_probe.py > gpurun_out/r02y_ncu.log 2>&1
ncu -i gpurun_out/r02y_bitsliced.ncu-rep --page details --csv 2>/dev/null | grep -E "Duration|Registers Per|Achieved Occupancy|Issue Slots Busy|ALU|Executed Ipc|Local" | cut -d, -f12- | head -20
