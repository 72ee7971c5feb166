#!/bin/bash
# 8-GPU session, final kernels of round 2: multi-GPU parity worker + the bench line without the host legs
set -u
mkdir -p gpurun_out
G=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29551 tests/dist_gpu_worker.py > gpurun_out/r02e_dist_gpu_${G}.log 2>&1
echo "dist worker rc=$?"; grep -E "DIST_GPU_OK|Error|assert" gpurun_out/r02e_dist_gpu_${G}.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $G --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02e_bench${G}.json 2> gpurun_out/r02e_bench${G}.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r02e_bench${G}.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02e_bench${G}.json'))
    keep={k:d.get(k) for k in ('value','ms_per_step','n_gpus','verified_vs_oracle_all_ranks','gpu_launches')}
    print(json.dumps(keep))
    print(json.dumps(d.get('strong'))[:1800])
    print(json.dumps(d.get('gathered'))[:1500])
    print(json.dumps(d['configs'].get('C5_fp61_matvec_8192_muladd_2^26')))
except Exception as ex:
    print('parse failed',ex)
PY
