#!/bin/bash
# Round-2 profiling recipe (run under gpurun, 1 GPU).  Outputs land in gpurun_out/.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-staged --no-configs --no-gathered"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
# the step's kernel (one persistent launch: share groups + reconstruction warps), full set; -s skips the warm-up launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_share_recover61 -s 4 -c 1 \
    -f -o gpurun_out/prof_fused_${TAG} $BENCH > gpurun_out/ncu_fused_${TAG}.log 2>&1
# the two kernels of the back-to-back schedule
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_share_tcm -s 3 -c 1 \
    -f -o gpurun_out/prof_share_${TAG} $BENCH > gpurun_out/ncu_share_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_recover61_pm -s 3 -c 1 \
    -f -o gpurun_out/prof_recover_${TAG} $BENCH > gpurun_out/ncu_recover_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
