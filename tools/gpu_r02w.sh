#!/bin/bash
# round 2, session 3: bulk groups of the raw / RLE stage in flight per CTA (CRYOGPU_ZP_PF_INFLIGHT), with and without the early pass
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py tests/test_gpu_pages.py -x -q > gpurun_out/pytest_r02w.log 2>&1; tail -3 gpurun_out/pytest_r02w.log
ab() { echo "== $1" >> gpurun_out/ab_r02w.txt; env $1 timeout 300 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02w.txt; }
ab CRYOGPU_LIB=tools/_prof/libcryogpu_prev.so
for g in 1 2 3 4 8; do ab CRYOGPU_ZP_PF_INFLIGHT=$g; done
ab "CRYOGPU_ZP_PF_INFLIGHT=1 CRYOGPU_ZP_PREFILL_CTAS=2"
ab "CRYOGPU_ZP_PF_INFLIGHT=1 CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=40"
ab "CRYOGPU_ZP_PF_INFLIGHT=2 CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=40"
ab "CRYOGPU_ZP_PF_INFLIGHT=1 CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=60"
ab CRYOGPU_LIB=tools/_prof/libcryogpu_prev.so
cat gpurun_out/ab_r02w.txt
CRYOGPU_ZP_PF_INFLIGHT=1 timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02w_1.txt 2>&1; tail -12 gpurun_out/timeline_r02w_1.txt
CRYOGPU_ZP_PF_INFLIGHT=2 timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02w_2.txt 2>&1; tail -12 gpurun_out/timeline_r02w_2.txt
CRYOGPU_ZP_PF_INFLIGHT=1 CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=40 timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02w_e.txt 2>&1; tail -13 gpurun_out/timeline_r02w_e.txt
