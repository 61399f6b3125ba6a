#!/bin/bash
# round 2, session 3: defaults (early pass one CTA per SM x 55 %, jobs, two tickets per round); full GPU suite; other kinds
mkdir -p gpurun_out
ab() { echo "== $1" >> gpurun_out/ab_r02aa.txt; env $1 timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02aa.txt; }
ab "CRYOGPU_LIB=tools/_prof/libcryogpu_prev.so CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=55"
ab CRYOGPU_X=1
ab CRYOGPU_ZP_EARLY_PCT=50
ab CRYOGPU_ZP_EARLY_PCT=60
ab CRYOGPU_ZP_EARLY_PCT=65
ab "CRYOGPU_ZP_EARLY_CTAS=0 CRYOGPU_ZP_JOBS=0"
ab CRYOGPU_X=1
cat gpurun_out/ab_r02aa.txt
timeout 200 python tools/zp_timeline.py > gpurun_out/timeline_r02aa.txt 2>&1; tail -13 gpurun_out/timeline_r02aa.txt
timeout 600 python tools/gpu_probe.py 1024 1:1:S:hex 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard 1:1:D:hex 1:1:D:lowcard 1:3:M:lowcard > gpurun_out/probe_r02aa.log 2>&1; cat gpurun_out/probe_r02aa.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02aa.log 2>&1; tail -3 gpurun_out/pytest_r02aa.log
