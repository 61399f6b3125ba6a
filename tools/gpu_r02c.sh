#!/bin/bash
# round-2 session: parity, small batches, cx phase profile, kinds probe, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_small_batches.py > gpurun_out/small_batches_r02c.txt 2>&1; cat gpurun_out/small_batches_r02c.txt
timeout 600 python tools/gpu_one_block.py > gpurun_out/one_block_r02c.txt 2>&1; cat gpurun_out/one_block_r02c.txt
ONE_REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_one_block.csv python tools/gpu_one_block.py > /dev/null 2>&1
CXPROF_ZSTD=1 CRYOGPU_LIB=tools/_prof/libcryogpu_prof.so timeout 300 python tools/gpu_cxprof.py > gpurun_out/cxprof_r02c.txt 2>&1; cat gpurun_out/cxprof_r02c.txt
timeout 600 python tools/gpu_probe.py 1024 > gpurun_out/probe_r02c.log 2>&1; cat gpurun_out/probe_r02c.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_r02c.json 2> gpurun_out/bench.err; cat gpurun_out/bench_r02c.json; tail -3 gpurun_out/bench.err
