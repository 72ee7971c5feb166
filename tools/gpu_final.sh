#!/bin/bash
# End-of-session check (1 GPU): sanitizer pass, full GPU test tier, smoke, both bench arms as the driver runs them
set -u
mkdir -p gpurun_out
SANITIZE_TCS=3 bash tools/sanitize.sh
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','verified_vs_oracle_all_ranks','clocks')})
print(d['int_roofline']['frac'], d['roofline']['frac'])
print({k:round(v['ms_per_step'],3) for k,v in d['schedules'].items() if isinstance(v,dict)})
e=d['e2e']; print(e['value'], e['ms_per_step'], e.get('frac_of_host_ceiling'), e['verified'])
for k,v in d['configs'].items(): print(k, round(v['value']/1e9,3), v['verified_vs_oracle'], v.get('hbm_frac'))
r=json.load(open('gpurun_out/final_bench_reference.json')); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
