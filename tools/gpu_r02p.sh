#!/bin/bash
# round 2, session 3: raw / RLE stage with warp-parallel positions; partial early pass (CTAs x share of the frames)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py tests/test_gpu_pages.py -x -q > gpurun_out/pytest_r02p.log 2>&1; tail -3 gpurun_out/pytest_r02p.log
ab() { echo "== $1" >> gpurun_out/ab_r02p.txt; env $1 timeout 300 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02p.txt; }
ab CRYOGPU_ZP_EARLY_CTAS=0
ab "CRYOGPU_ZP_EARLY_CTAS=32 CRYOGPU_ZP_EARLY_PCT=30"
ab "CRYOGPU_ZP_EARLY_CTAS=64 CRYOGPU_ZP_EARLY_PCT=30"
ab "CRYOGPU_ZP_EARLY_CTAS=64 CRYOGPU_ZP_EARLY_PCT=45"
ab "CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=45"
ab "CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=60"
ab "CRYOGPU_ZP_EARLY_CTAS=96 CRYOGPU_ZP_EARLY_PCT=50"
ab "CRYOGPU_ZP_EARLY_CTAS=32 CRYOGPU_ZP_EARLY_PCT=20"
ab CRYOGPU_ZP_EARLY_CTAS=0
cat gpurun_out/ab_r02p.txt
CRYOGPU_ZP_EARLY_CTAS=0 timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02p_0.txt 2>&1; tail -12 gpurun_out/timeline_r02p_0.txt
CRYOGPU_ZP_EARLY_CTAS=64 CRYOGPU_ZP_EARLY_PCT=45 timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02p_64_45.txt 2>&1; tail -13 gpurun_out/timeline_r02p_64_45.txt
CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=45 timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02p_148_45.txt 2>&1; tail -13 gpurun_out/timeline_r02p_148_45.txt
