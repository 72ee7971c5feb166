#!/bin/bash
# Round-2 re-entry check from HEAD: full GPU test tier, smoke, both bench arms as the driver runs them.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
tail -4 gpurun_out/r02r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/r02r_smoke.log
S=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/r02r_bench_reference.json 2> gpurun_out/r02r_ref.err
echo "ref rc=$? wall=$(( $(date +%s) - S ))"
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err
echo "bench rc=$? wall=$(( $(date +%s) - S ))"; tail -c 400 gpurun_out/r02r_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02r_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','verified_vs_oracle_all_ranks','clocks')})
print(d['int_roofline']['frac'], d['roofline']['frac'], d['roofline']['kernel'])
e=d['e2e']; print(e['value'], e['ms_per_step'], e.get('frac_of_host_ceiling'))
for k,v in d['configs'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('note','checked','collective','bound')})
r=json.load(open('gpurun_out/r02r_bench_reference.json')); print('ref', r['value'], r['cpu_baseline']['cores'])
PY
