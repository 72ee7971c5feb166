#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02l_bench_reference.json 2> gpurun_out/r02l_ref.err
timeout 900 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02l_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','verified_vs_oracle_all_ranks','clocks')})
print({k:round(v['ms_per_step'],3) for k,v in d['schedules'].items() if isinstance(v,dict)})
print(d['int_roofline']['frac'], d['roofline']['frac'], d['roofline']['achieved'])
print({k:round(v['ms_per_step'],3) for k,v in d['gathered'].items() if isinstance(v,dict)})
e=d['e2e']; print(e['value'], e['ms_per_step'], e['frac_of_host_ceiling'])
for k,v in d['configs'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('note','checked','collective','bound')})
print(d['cpu_baseline'])
r=json.load(open('gpurun_out/r02l_bench_reference.json')); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
