#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_one_block.py > gpurun_out/one_block_r02l.txt 2>&1; grep -v Warn gpurun_out/one_block_r02l.txt
CXPROF_ZSTD=0 CRYOGPU_LIB=tools/_prof/libcryogpu_prof.so timeout 300 python tools/gpu_cxprof.py > gpurun_out/cxprof_r02l.txt 2>&1; grep -v "Warn\|d_me" gpurun_out/cxprof_r02l.txt
timeout 600 python tools/gpu_probe.py 1024 0:1:S:hex 0:1:M:hex 0:1:M:lowcard 0:1:D:hex 0:1:D:lowcard > gpurun_out/probe_r02l.log 2>&1; cat gpurun_out/probe_r02l.log
