#!/bin/bash
# round 2, session 3: the executor's long runs as jobs for the raw / RLE stage (CRYOGPU_ZP_JOBS), with the early pass
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py tests/test_gpu_pages.py -x -q > gpurun_out/pytest_r02y.log 2>&1; tail -3 gpurun_out/pytest_r02y.log
ab() { echo "== $1" >> gpurun_out/ab_r02y.txt; env $1 timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02y.txt; }
ab CRYOGPU_ZP_JOBS=0
ab CRYOGPU_ZP_JOBS=1
ab "CRYOGPU_ZP_JOBS=1 CRYOGPU_ZP_PF_INFLIGHT=1"
ab "CRYOGPU_ZP_JOBS=1 CRYOGPU_ZP_PF_INFLIGHT=4"
ab "CRYOGPU_ZP_JOBS=1 CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=40"
ab "CRYOGPU_ZP_JOBS=1 CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=60"
ab "CRYOGPU_ZP_JOBS=1 CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=40 CRYOGPU_ZP_PF_INFLIGHT=1"
ab "CRYOGPU_ZP_JOBS=1 CRYOGPU_ZP_EARLY_CTAS=74 CRYOGPU_ZP_EARLY_PCT=40 CRYOGPU_ZP_PF_INFLIGHT=1"
ab CRYOGPU_ZP_JOBS=0
cat gpurun_out/ab_r02y.txt
timeout 200 python tools/zp_timeline.py > gpurun_out/timeline_r02y.txt 2>&1; tail -12 gpurun_out/timeline_r02y.txt
CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=40 timeout 200 python tools/zp_timeline.py > gpurun_out/timeline_r02y_e.txt 2>&1; tail -13 gpurun_out/timeline_r02y_e.txt
