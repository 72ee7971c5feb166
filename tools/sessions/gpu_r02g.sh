#!/bin/bash
# 2-GPU session: multi-GPU tests + bench under torchrun
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_gpu_nccl or gather or multi_context or fused" > gpurun_out/r02g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
tail -15 gpurun_out/r02g_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02g_bench2.json 2> gpurun_out/r02g_bench2.err
echo "bench rc=$?"
tail -c 2500 gpurun_out/r02g_bench2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02g_bench2.json'))
    keep={k:d[k] for k in ('value','ms_per_step','n_gpus','verified_vs_oracle_all_ranks','schedules','strong','gathered','gpu_launches')}
    print(json.dumps(keep,indent=1)[:5000])
    print(json.dumps(d.get('e2e'),indent=1)[:3000])
    print(json.dumps(d['configs'].get('C5_fp61_matvec_8192_muladd_2^26'),indent=1))
except Exception as e:
    print('parse failed',e)
PY
