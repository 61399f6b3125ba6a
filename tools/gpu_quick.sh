#!/bin/bash
# quick GPU check: zstd parity tests, probe of block kinds, headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_probe.py 1024 1:1:S:hex 1:1:M:hex 1:1:D:hex 1:1:D:lowcard 1:-5:D:lowcard 1:3:M:lowcard > gpurun_out/probe.log 2>&1; cat gpurun_out/probe.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
grep -v "at::" gpurun_out/launches.csv | awk -F'","' '{print $5, $NF}' | tail -16
