#!/bin/bash
# round 2, session 3: raw / RLE stage with units (blocks and jobs), jobs claimed when they exist and served between frames;
# executor waits only for sources inside what stage 0 owes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py tests/test_gpu_pages.py -x -q > gpurun_out/pytest_r02ab.log 2>&1; tail -3 gpurun_out/pytest_r02ab.log
ab() { echo "== $1" >> gpurun_out/ab_r02ab.txt; env $1 timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02ab.txt; }
ab "CRYOGPU_LIB=tools/_prof/libcryogpu_prev.so CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=55"
ab CRYOGPU_X=1
ab CRYOGPU_ZP_EARLY_PCT=45
ab CRYOGPU_ZP_EARLY_PCT=65
ab CRYOGPU_ZP_PF_INFLIGHT=4
ab "CRYOGPU_ZP_EARLY_CTAS=0 CRYOGPU_ZP_JOBS=0"
ab CRYOGPU_X=1
cat gpurun_out/ab_r02ab.txt
timeout 200 python tools/zp_timeline.py > gpurun_out/timeline_r02ab.txt 2>&1; tail -13 gpurun_out/timeline_r02ab.txt
timeout 600 python tools/gpu_probe.py 1024 1:1:S:hex 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard 1:1:D:hex 1:3:M:lowcard > gpurun_out/probe_r02ab.log 2>&1; cat gpurun_out/probe_r02ab.log
CRYOGPU_ZP_EARLY_CTAS=0 CRYOGPU_ZP_JOBS=0 timeout 600 python tools/gpu_probe.py 1024 1:1:S:hex 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard > gpurun_out/probe_r02ab_off.log 2>&1; cat gpurun_out/probe_r02ab_off.log
