#!/bin/bash
# round 2, session 3: byte-mapped runs in the warp executor -- parity, A/B against the previous build, timeline, ncu of the
# entropy stages and the executor with source lines
mkdir -p gpurun_out /tmp/nr
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py -x -q > gpurun_out/pytest_r02m.log 2>&1; tail -3 gpurun_out/pytest_r02m.log
ab() { echo "== $1" >> gpurun_out/ab_r02m.txt; CRYOGPU_LIB=$2 timeout 300 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" >> gpurun_out/ab_r02m.txt; }
ab old tools/_prof/libcryogpu_old.so
ab new72 pg_cryogen_b200/libcryogpu.so
ab new64 tools/_prof/libcryogpu_r64.so
ab old tools/_prof/libcryogpu_old.so
ab new72 pg_cryogen_b200/libcryogpu.so
cat gpurun_out/ab_r02m.txt
timeout 300 python tools/zp_timeline.py > gpurun_out/timeline_r02m.txt 2>&1; tail -24 gpurun_out/timeline_r02m.txt
for k in k_zp_execute k_zp_literals k_zp_sequences_small; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 3 -c 1 -o /tmp/nr/$k -f python tools/gpu_probe.py 3449 1:1:S:hex > /tmp/nr/$k.log 2>&1
  tail -2 /tmp/nr/$k.log
  python tools/ncu_kernel_summary.py /tmp/nr/$k.ncu-rep headline_$k > gpurun_out/r02m_${k}_ncu.txt 2>&1
  python tools/ncu_lines.py /tmp/nr/$k.ncu-rep 45 | cut -c1-220 >> gpurun_out/r02m_${k}_ncu.txt 2>&1
done
timeout 600 python tools/gpu_probe.py 1024 1:1:S:lowcard 1:1:M:hex 1:1:M:lowcard 1:1:D:hex 1:3:M:lowcard > gpurun_out/probe_r02m.log 2>&1; cat gpurun_out/probe_r02m.log
