#!/bin/bash
# round-2 session e
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/bench_r02e.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_r02e.json; tail -3 gpurun_out/bench.err
python tools/zp_timeline.py > gpurun_out/timeline_r02e.txt 2>&1; tail -24 gpurun_out/timeline_r02e.txt
timeout 600 python tools/gpu_one_block.py > gpurun_out/one_block_r02e.txt 2>&1; cat gpurun_out/one_block_r02e.txt
CXPROF_ZSTD=1 CRYOGPU_LIB=tools/_prof/libcryogpu_prof.so timeout 300 python tools/gpu_cxprof.py > gpurun_out/cxprof_r02e.txt 2>&1; grep -v Warn gpurun_out/cxprof_r02e.txt
timeout 600 python tools/gpu_probe.py 1024 > gpurun_out/probe_r02e.log 2>&1; cat gpurun_out/probe_r02e.log
timeout 600 python tools/gpu_small_batches.py > gpurun_out/small_batches_r02e.txt 2>&1; cat gpurun_out/small_batches_r02e.txt
