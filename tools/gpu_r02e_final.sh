#!/bin/bash
# Final-form measurements of round 2 (1 GPU): full GPU test tier, both bench arms as the driver runs them, launch list
# and the ncu capture of the step kernel (profiles/run_ncu_r02.sh).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log; tail -3 gpurun_out/r02e_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02e_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/r02e_bench_reference.json 2> gpurun_out/r02e_ref.err; echo "ref rc=$?"
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','verified_vs_oracle_all_ranks','clocks')})
print(d['int_roofline']['frac'], d['roofline']['frac'], d['roofline']['kernel'])
print({k:round(v['ms_per_step'],3) for k,v in d['schedules'].items() if isinstance(v,dict)})
e=d['e2e']; print(e['value'], e['ms_per_step'], e.get('frac_of_host_ceiling'))
for k,v in d['configs'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('note','checked','collective','bound')})
r=json.load(open('gpurun_out/r02e_bench_reference.json')); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
bash profiles/run_ncu_r02.sh r02e > gpurun_out/run_ncu_r02e.log 2>&1; tail -4 gpurun_out/run_ncu_r02e.log
