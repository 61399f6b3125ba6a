#!/bin/bash
# round 2, session 3: lane-serial stages in one wave (warps walk their groups' block indices), sequence walk by bit position
mkdir -p gpurun_out /tmp/nr
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_shim.py -x -q > gpurun_out/pytest_r02n.log 2>&1; tail -3 gpurun_out/pytest_r02n.log
ab() { echo "== $1" >> gpurun_out/ab_r02n.txt; CRYOGPU_LIB=$2 timeout 300 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02n.txt; }
ab old tools/_prof/libcryogpu_old.so
ab new pg_cryogen_b200/libcryogpu.so
ab old tools/_prof/libcryogpu_old.so
ab new pg_cryogen_b200/libcryogpu.so
cat gpurun_out/ab_r02n.txt
timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02n.txt 2>&1; tail -12 gpurun_out/timeline_r02n.txt
for k in k_zp_huftab k_zp_fsetab k_zp_sequences_small k_zp_literals; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 3 -c 1 -o /tmp/nr/$k -f python tools/gpu_probe.py 3449 1:1:S:hex > /tmp/nr/$k.log 2>&1
  tail -1 /tmp/nr/$k.log
  python tools/ncu_kernel_summary.py /tmp/nr/$k.ncu-rep headline_$k > gpurun_out/r02n_${k}_ncu.txt 2>&1
  python tools/ncu_lines.py /tmp/nr/$k.ncu-rep 45 | cut -c1-220 >> gpurun_out/r02n_${k}_ncu.txt 2>&1
done
timeout 600 python tools/gpu_probe.py 1024 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard 1:1:D:hex 1:1:D:lowcard 1:3:M:lowcard > gpurun_out/probe_r02n.log 2>&1; cat gpurun_out/probe_r02n.log
