#!/bin/bash
# round 2, session 3: sequence windows for the working lanes only, source prefetch inside the parse kernel; full GPU suite
mkdir -p gpurun_out
ab() { echo "== $1" >> gpurun_out/ab_r02s.txt; env $1 timeout 300 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02s.txt; }
ab CRYOGPU_LIB=tools/_prof/libcryogpu_old.so
ab CRYOGPU_X=1
ab CRYOGPU_ZP_SRC_PREFETCH=0
ab CRYOGPU_X=1
cat gpurun_out/ab_r02s.txt
timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02s.txt 2>&1; tail -12 gpurun_out/timeline_r02s.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02s.log 2>&1; tail -3 gpurun_out/pytest_r02s.log
