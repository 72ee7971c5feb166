#!/bin/bash
# 8-GPU session: multi-GPU parity worker + bench under torchrun
set -u
mkdir -p gpurun_out
G=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29551 tests/dist_gpu_worker.py > gpurun_out/r02k_dist_gpu_${G}.log 2>&1
echo "dist worker rc=$?"; grep -E "DIST_GPU_OK|Error|assert" gpurun_out/r02k_dist_gpu_${G}.log | head -5
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r02k_bench${G}.json 2> gpurun_out/r02k_bench${G}.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r02k_bench${G}.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02k_bench${G}.json'))
    keep={k:d.get(k) for k in ('value','ms_per_step','n_gpus','verified_vs_oracle_all_ranks','strong','gathered','gpu_launches')}
    print(json.dumps(keep,indent=1)[:4500])
    e=d.get('e2e') or {}
    print(json.dumps({k:e.get(k) for k in ('value','ms_per_step','one_direction_at_a_time','host_ceiling','frac_of_host_ceiling','single_process_all_gpus','secrets_per_gpu')},indent=1))
    print(json.dumps(d['configs'].get('C5_fp61_matvec_8192_muladd_2^26'),indent=1))
except Exception as ex:
    print('parse failed',ex)
PY
