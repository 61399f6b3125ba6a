#!/bin/bash
# round-2 session d: parity, one-block latency, cx phase profile, kinds probe, small batches, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_one_block.py > gpurun_out/one_block_r02d.txt 2>&1; cat gpurun_out/one_block_r02d.txt
CXPROF_ZSTD=1 CRYOGPU_LIB=tools/_prof/libcryogpu_prof.so timeout 300 python tools/gpu_cxprof.py > gpurun_out/cxprof_r02d.txt 2>&1; cat gpurun_out/cxprof_r02d.txt
timeout 600 python tools/gpu_probe.py 1024 > gpurun_out/probe_r02d.log 2>&1; cat gpurun_out/probe_r02d.log
timeout 600 python tools/gpu_small_batches.py > gpurun_out/small_batches_r02d.txt 2>&1; cat gpurun_out/small_batches_r02d.txt
timeout 600 python bench.py --no-cpu > gpurun_out/bench_r02d.json 2> gpurun_out/bench.err; cat gpurun_out/bench_r02d.json; tail -3 gpurun_out/bench.err
