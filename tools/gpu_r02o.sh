#!/bin/bash
# round 2, session 3: early pass of the raw / RLE stage beside the entropy stages (guessed positions), source prefetch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py tests/test_gpu_pages.py -x -q > gpurun_out/pytest_r02o.log 2>&1; tail -3 gpurun_out/pytest_r02o.log
ab() { echo "== $1" >> gpurun_out/ab_r02o.txt; env $1 timeout 300 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02o.txt; }
ab CRYOGPU_ZP_EARLY=0
ab CRYOGPU_ZP_EARLY=1
ab "CRYOGPU_ZP_EARLY=1 CRYOGPU_ZP_SRC_PREFETCH=0"
ab CRYOGPU_ZP_EARLY=2
ab CRYOGPU_ZP_EARLY=0
ab CRYOGPU_ZP_EARLY=1
cat gpurun_out/ab_r02o.txt
for e in 1 2; do
CRYOGPU_ZP_EARLY=$e timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02o_$e.txt 2>&1; tail -13 gpurun_out/timeline_r02o_$e.txt
done
CRYOGPU_ZP_EARLY=1 CRYOGPU_ZP_SRC_PREFETCH=0 timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02o_nopf.txt 2>&1; tail -13 gpurun_out/timeline_r02o_nopf.txt
timeout 600 python tools/gpu_probe.py 1024 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard 1:1:D:hex 1:1:D:lowcard > gpurun_out/probe_r02o.log 2>&1; cat gpurun_out/probe_r02o.log
CRYOGPU_ZP_EARLY=0 timeout 600 python tools/gpu_probe.py 1024 1:1:M:hex 1:1:D:hex > gpurun_out/probe_r02o_e0.log 2>&1; cat gpurun_out/probe_r02o_e0.log
timeout 300 python tools/zp_timeline.py 294912 D hex > gpurun_out/timeline_r02o_Dhex.txt 2>&1; tail -13 gpurun_out/timeline_r02o_Dhex.txt
