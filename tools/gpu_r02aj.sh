#!/bin/bash
# round 2, session 3: stage 4 queues jobs only while stage 0's late pass is running (kernels one after the other under ncu: it then
# writes its runs itself); parity, A/B against the previous build, launch list and full captures of the two kernels under ncu
mkdir -p gpurun_out /tmp/nr
timeout 600 python -m pytest tests/test_gpu_zstd_decode.py tests/test_gpu_zstd_encode.py tests/test_gpu_pages.py tests/test_gpu_shim.py -x -q > gpurun_out/pytest_r02aj.log 2>&1; tail -2 gpurun_out/pytest_r02aj.log
for rep in 1 2; do
for l in tools/_prof/libcryogpu_prev.so pg_cryogen_b200/libcryogpu.so; do
  echo "== $l" >> gpurun_out/ab_r02aj.txt
  CRYOGPU_LIB=$l timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_per_step_min_median_max'])" >> gpurun_out/ab_r02aj.txt
done; done
cat gpurun_out/ab_r02aj.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02c_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
grep "k_zp_" gpurun_out/r02c_launches_bench.csv | awk -F'","' '{print $5, $NF}' | tail -12
for k in k_zp_prefill k_zp_execute; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 3 -c 1 -o /tmp/nr/$k -f python tools/gpu_probe.py 3449 1:1:S:hex > /tmp/nr/$k.log 2>&1
  python tools/ncu_kernel_summary.py /tmp/nr/$k.ncu-rep headline >> gpurun_out/r02c_under_ncu_two_kernels.txt 2>&1
done
cat gpurun_out/r02c_under_ncu_two_kernels.txt
