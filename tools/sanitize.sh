#!/bin/bash
# compute-sanitizer passes over the hot kernels (run under gpurun, 1 GPU): memcheck,
# racecheck (shared-memory hazards), synccheck (barrier misuse) on a small sweep of every
# share / recover kernel.  Output: gpurun_out/sanitize_<tool>.log
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for tc in ${SANITIZE_TCS:-3 1 0}; do
    SCLGPU_SANITIZE_SMALL=1 SCLGPU_SHARE_TC=$tc timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 \
      python tests/sanitize_driver.py > gpurun_out/sanitize_${tool}_tc${tc}.log 2>&1
    echo "$tool tc=$tc rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_tc${tc}.log | tail -1)"
  done
done
