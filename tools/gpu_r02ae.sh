#!/bin/bash
# round 2, session 3: four tickets per CTA in the job phase of the raw / RLE stage
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py tests/test_gpu_pages.py -x -q > gpurun_out/pytest_r02ae.log 2>&1; tail -3 gpurun_out/pytest_r02ae.log
ab() { echo "== $1" >> gpurun_out/ab_r02ae.txt; env $1 timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02ae.txt; }
for rep in 1 2 3; do
ab CRYOGPU_LIB=tools/_prof/libcryogpu_prev.so
ab CRYOGPU_X=1
done
ab CRYOGPU_ZP_EARLY_PCT=65
ab CRYOGPU_ZP_EARLY_PCT=45
ab "CRYOGPU_ZP_EARLY_CTAS=0 CRYOGPU_ZP_JOBS=0"
cat gpurun_out/ab_r02ae.txt
for i in 1 2; do timeout 200 python tools/zp_timeline.py 2>&1 | tail -13 >> gpurun_out/timeline_r02ae.txt; done; grep -v "lz4_cta\|execute_cta\|seq_large\|parse" gpurun_out/timeline_r02ae.txt
timeout 600 python tools/gpu_probe.py 1024 1:1:S:hex 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard > gpurun_out/probe_r02ae.log 2>&1; cat gpurun_out/probe_r02ae.log
