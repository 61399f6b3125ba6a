#!/bin/bash
# round 2, session 3: table stages per 8 frames, vectorised table loads / window refills, source prefetch on its own,
# repeat-offset prefix, leaner step loop; executor CTAs of 4 warps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py tests/test_gpu_pages.py -x -q > gpurun_out/pytest_r02q.log 2>&1; tail -3 gpurun_out/pytest_r02q.log
ab() { echo "== $1" >> gpurun_out/ab_r02q.txt; env $1 timeout 300 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02q.txt; }
ab CRYOGPU_LIB=tools/_prof/libcryogpu_old.so
ab CRYOGPU_ZP_SRC_PREFETCH=1
ab CRYOGPU_ZP_SRC_PREFETCH=0
ab CRYOGPU_LIB=tools/_prof/libcryogpu_w4.so
ab CRYOGPU_ZP_SRC_PREFETCH=1
cat gpurun_out/ab_r02q.txt
timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02q.txt 2>&1; tail -12 gpurun_out/timeline_r02q.txt
CRYOGPU_ZP_SRC_PREFETCH=0 timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02q_nopf.txt 2>&1; tail -12 gpurun_out/timeline_r02q_nopf.txt
timeout 600 python tools/gpu_probe.py 1024 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard 1:1:D:hex 1:1:D:lowcard 1:3:M:lowcard > gpurun_out/probe_r02q.log 2>&1; cat gpurun_out/probe_r02q.log
CRYOGPU_ZP_SRC_PREFETCH=0 timeout 600 python tools/gpu_probe.py 1024 1:1:M:hex 1:1:D:hex > gpurun_out/probe_r02q_nopf.log 2>&1; cat gpurun_out/probe_r02q_nopf.log
mkdir -p /tmp/nr
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_zp_execute$' -s 3 -c 1 -o /tmp/nr/zp_exec -f python tools/gpu_probe.py 3449 1:1:S:hex > /tmp/nr/zp_exec.log 2>&1
python tools/ncu_kernel_summary.py /tmp/nr/zp_exec.ncu-rep headline_execute > gpurun_out/r02q_zp_execute_ncu.txt 2>&1
python tools/ncu_lines.py /tmp/nr/zp_exec.ncu-rep 50 | cut -c1-220 >> gpurun_out/r02q_zp_execute_ncu.txt 2>&1
head -3 gpurun_out/r02q_zp_execute_ncu.txt
