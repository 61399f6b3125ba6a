#!/bin/bash
# round 2, session 3: raw / RLE stage -- edge stores and their fence off the issuing warp, one fence per batch of publications
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_pages.py -x -q > gpurun_out/pytest_r02ai.log 2>&1; tail -2 gpurun_out/pytest_r02ai.log
for rep in 1 2 3; do
for l in tools/_prof/libcryogpu_prev.so pg_cryogen_b200/libcryogpu.so; do
  echo "== $l" >> gpurun_out/ab_r02ai.txt
  CRYOGPU_LIB=$l timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_per_step_min_median_max'])" >> gpurun_out/ab_r02ai.txt
done; done
cat gpurun_out/ab_r02ai.txt
timeout 200 python tools/zp_timeline.py 2>&1 | tail -13 | grep -v "lz4_cta\|execute_cta\|seq_large\|parse"
timeout 300 python tools/gpu_probe.py 1024 1:1:S:hex 1:1:M:hex 1:1:D:hex 2>&1 | grep method
