#!/bin/bash
# round 2, session 3: executor -- literal window filled whole and kept across runs, fills without read-back
mkdir -p gpurun_out /tmp/nr
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py tests/test_gpu_pages.py -x -q > gpurun_out/pytest_r02v.log 2>&1; tail -3 gpurun_out/pytest_r02v.log
ab() { echo "== $1" >> gpurun_out/ab_r02v.txt; env $1 timeout 300 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02v.txt; }
ab CRYOGPU_LIB=tools/_prof/libcryogpu_prev.so
ab CRYOGPU_X=1
ab CRYOGPU_LIB=tools/_prof/libcryogpu_prev.so
ab CRYOGPU_X=1
ab "CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=35"
ab "CRYOGPU_ZP_EARLY_CTAS=148 CRYOGPU_ZP_EARLY_PCT=50"
cat gpurun_out/ab_r02v.txt
timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02v.txt 2>&1; tail -12 gpurun_out/timeline_r02v.txt
timeout 600 python tools/gpu_probe.py 1024 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard 1:1:D:hex 1:1:D:lowcard 1:3:M:lowcard > gpurun_out/probe_r02v.log 2>&1; cat gpurun_out/probe_r02v.log
k=k_zp_execute
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 3 -c 1 -o /tmp/nr/$k -f python tools/gpu_probe.py 3449 1:1:S:hex > /tmp/nr/$k.log 2>&1
python tools/ncu_kernel_summary.py /tmp/nr/$k.ncu-rep headline > gpurun_out/r02v_execute_ncu.txt 2>&1
python tools/ncu_lines.py /tmp/nr/$k.ncu-rep 40 | cut -c1-200 >> gpurun_out/r02v_execute_ncu.txt 2>&1
head -12 gpurun_out/r02v_execute_ncu.txt
