#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r02e_bench.err
head -c 6000 gpurun_out/r02e_bench.json
