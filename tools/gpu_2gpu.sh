#!/bin/bash
# 2-GPU check of the default bench line (with the host legs), as the driver's scaling run launches it
set -u
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench2.json 2> gpurun_out/r02g_bench2.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02g_bench2.json"))
print(d["n_gpus"], d["ms_per_step"], d["value"], d["int_roofline"]["frac"], d["verified_vs_oracle_all_ranks"])
print(d["strong"]["ms_per_step"], d["strong"]["value"], {k:round(v["ms_per_step"],3) for k,v in d["strong"]["gathered"].items() if isinstance(v,dict)})
e=d["e2e"]; print(e["value"], e["ms_per_step"], e.get("frac_of_host_ceiling"), e["verified"], (e.get("single_process_all_gpus") or {}).get("value"))
print(d["configs"]["C5_fp61_matvec_8192_muladd_2^26"]["matvec_ms"])
PY
