#!/bin/bash
# Profiling recipe (run under gpurun, 1 GPU).  Outputs land in gpurun_out/.
#   $1 = tag (e.g. r01c)   $2 = kernel regex for the full capture (default: both hot kernels)
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-staged"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
# the two hot kernels, full set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_share -s 3 -c 1 \
    -f -o gpurun_out/prof_share_${TAG} $BENCH > gpurun_out/ncu_share_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_recover -s 3 -c 1 \
    -f -o gpurun_out/prof_recover_${TAG} $BENCH > gpurun_out/ncu_recover_${TAG}.log 2>&1
ls -la gpurun_out
