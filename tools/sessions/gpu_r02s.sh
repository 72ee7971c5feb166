#!/bin/bash
# A/B of the AES final-round change on the step kernel; e2e pipeline depth sweep; mat-vec vs dot + ncu
set -u
mkdir -p gpurun_out
SO=secure-computation-library_b200/csrc/libsclgpu.so
cp $SO /tmp/new.so
: > gpurun_out/r02s_kbench.json
for v in base new base new; do
  if [ $v = base ]; then cp ab/libsclgpu_base.so $SO; else cp /tmp/new.so $SO; fi
  echo "{\"variant\": \"$v\"}" >> gpurun_out/r02s_kbench.json
  timeout 300 python tools/kbench.py 26 8 >> gpurun_out/r02s_kbench.json 2>> gpurun_out/r02s_kbench.err
done
cp /tmp/new.so $SO
python - <<'PY'
import json
for l in open('gpurun_out/r02s_kbench.json'):
    d=json.loads(l)
    if 'variant' in d: print(d); continue
    print({k:round(v,3) for k,v in d.items() if k.endswith('_ms')}, d['ok'], d['fused_ok'], d['fused_indep_ok'])
PY
timeout 600 python -m pytest tests -x -q -m gpu -k "share or prg or random or additive or host or async or pageable" > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s_pytest.log
: > gpurun_out/r02s_e2e.json
for cfg in "2 256" "4 64" "3 128" "4 128" "4 32"; do
  set -- $cfg
  SCLGPU_HOST_PIPES=$1 SCLGPU_HOST_CHUNK_MB=$2 timeout 300 python tools/e2e_probe.py 25 3 >> gpurun_out/r02s_e2e.json 2>> gpurun_out/r02s_e2e.err
done
cat gpurun_out/r02s_e2e.json
timeout 120 python tools/matvec_probe.py > gpurun_out/r02s_matvec.json 2> gpurun_out/r02s_matvec.err; cat gpurun_out/r02s_matvec.json
SCLGPU_MATVEC_WARP=1 timeout 120 python tools/matvec_probe.py >> gpurun_out/r02s_matvec.json 2>> gpurun_out/r02s_matvec.err; tail -1 gpurun_out/r02s_matvec.json
timeout 300 ncu --set full --clock-control none -k regex:"k_matvec61_chunks|k_dot_partial" -c 4 --launch-skip 8 -o gpurun_out/r02s_matvec python tools/matvec_probe.py 2 > gpurun_out/r02s_ncu.log 2>&1
ncu -i gpurun_out/r02s_matvec.ncu-rep --page details --csv 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|Registers|Achieved Occupancy|L2 Cache Throughput|Theoretical Occ" | head -40
