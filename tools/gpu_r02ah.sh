#!/bin/bash
# run-to-run: the early pass with at most one working CTA per SM (CRYOGPU_ZP_EARLY_ONE_PER_SM)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py -x -q > gpurun_out/pytest_r02ah.log 2>&1; tail -2 gpurun_out/pytest_r02ah.log
for c in 1 0 1 0 1 0 1 0 1 1; do
  echo "== CRYOGPU_ZP_EARLY_ONE_PER_SM=$c" >> gpurun_out/one_r02ah.txt
  CRYOGPU_ZP_EARLY_ONE_PER_SM=$c timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_per_step_min_median_max'])" >> gpurun_out/one_r02ah.txt
done
cat gpurun_out/one_r02ah.txt
