#!/bin/bash
# round 2, session 3: share of the frames for the early pass, with the executor's long runs as jobs
mkdir -p gpurun_out
ab() { echo "== $1" >> gpurun_out/ab_r02z.txt; env $1 timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02z.txt; }
ab CRYOGPU_ZP_JOBS=0
for p in 50 60 70 80 100; do ab "CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=$p"; done
ab "CRYOGPU_ZP_EARLY_CTAS=296 CRYOGPU_ZP_EARLY_PCT=60"
ab "CRYOGPU_ZP_EARLY_CTAS=296 CRYOGPU_ZP_EARLY_PCT=80"
ab "CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=60 CRYOGPU_ZP_PF_INFLIGHT=4"
ab "CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=60 CRYOGPU_ZP_JOBS=0"
cat gpurun_out/ab_r02z.txt
CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=70 timeout 200 python tools/zp_timeline.py > gpurun_out/timeline_r02z_70.txt 2>&1; tail -13 gpurun_out/timeline_r02z_70.txt
CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=100 timeout 200 python tools/zp_timeline.py > gpurun_out/timeline_r02z_100.txt 2>&1; tail -13 gpurun_out/timeline_r02z_100.txt
