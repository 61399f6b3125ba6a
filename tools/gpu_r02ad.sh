#!/bin/bash
# round 2, session 3: matches that run out of the known range split; jobs after the frames (1) or also between them (2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py tests/test_gpu_pages.py -x -q > gpurun_out/pytest_r02ad.log 2>&1; tail -3 gpurun_out/pytest_r02ad.log
ab() { echo "== $1" >> gpurun_out/ab_r02ad.txt; env $1 timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02ad.txt; }
for rep in 1 2 3; do
ab "CRYOGPU_ZP_EARLY_CTAS=0 CRYOGPU_ZP_JOBS=0"
ab CRYOGPU_ZP_JOBS=1
ab CRYOGPU_ZP_JOBS=2
done
ab "CRYOGPU_LIB=tools/_prof/libcryogpu_prev.so CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=55"
cat gpurun_out/ab_r02ad.txt
timeout 200 python tools/zp_timeline.py 296960 M hex 2>&1 | tail -14 > gpurun_out/mhex_r02ad.txt; grep -v "lz4_cta\|execute_cta\|seq_large\|parse" gpurun_out/mhex_r02ad.txt
timeout 600 python tools/gpu_probe.py 1024 1:1:S:hex 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard 1:1:D:hex 1:3:M:lowcard > gpurun_out/probe_r02ad.log 2>&1; cat gpurun_out/probe_r02ad.log
