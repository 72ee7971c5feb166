#!/bin/bash
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-staged --no-configs --no-gathered"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_share_recover61 -s 4 -c 1 -f -o gpurun_out/prof_fused_r02c $BENCH > gpurun_out/ncu_fused_r02c.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02c.csv $BENCH > gpurun_out/ncu_launches_r02c.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02q_bench.json'))
print(d['value'], d['ms_per_step'], d['int_roofline']['frac'], d['verified_vs_oracle_all_ranks'])
print({k:round(v['ms_per_step'],3) for k,v in d['schedules'].items() if isinstance(v,dict)})
print({k:round(v['ms_per_step'],3) for k,v in d['gathered'].items() if isinstance(v,dict)})
PY
