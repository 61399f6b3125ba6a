#!/bin/bash
# round 2, session 3: executor descriptors in shared memory, table CTAs of 4 warps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py tests/test_gpu_pages.py tests/test_gpu_batch.py -x -q > gpurun_out/pytest_r02r.log 2>&1; tail -3 gpurun_out/pytest_r02r.log
ab() { echo "== $1" >> gpurun_out/ab_r02r.txt; env $1 timeout 300 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02r.txt; }
ab CRYOGPU_LIB=tools/_prof/libcryogpu_old.so
ab CRYOGPU_X=1
ab CRYOGPU_LIB=tools/_prof/libcryogpu_old.so
ab CRYOGPU_X=1
ab CRYOGPU_ZP_PREFILL_CTAS=2
cat gpurun_out/ab_r02r.txt
timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02r.txt 2>&1; tail -12 gpurun_out/timeline_r02r.txt
timeout 600 python tools/gpu_probe.py 1024 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard 1:1:D:hex 1:1:D:lowcard 1:3:M:lowcard > gpurun_out/probe_r02r.log 2>&1; cat gpurun_out/probe_r02r.log
