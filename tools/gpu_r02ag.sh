#!/bin/bash
# run-to-run: does the number of hardware queues (CUDA_DEVICE_MAX_CONNECTIONS) explain bench.py's slower state
mkdir -p gpurun_out
for c in 32 8 32 8 32 32 16; do
  echo "== CUDA_DEVICE_MAX_CONNECTIONS=$c" >> gpurun_out/conn_r02ag.txt
  CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_per_step_min_median_max'])" >> gpurun_out/conn_r02ag.txt
done
for c in 32 8; do
  echo "== off, CUDA_DEVICE_MAX_CONNECTIONS=$c" >> gpurun_out/conn_r02ag.txt
  CRYOGPU_ZP_EARLY_CTAS=0 CRYOGPU_ZP_JOBS=0 CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_per_step_min_median_max'])" >> gpurun_out/conn_r02ag.txt
done
cat gpurun_out/conn_r02ag.txt
