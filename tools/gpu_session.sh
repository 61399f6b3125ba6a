#!/bin/bash
# one GPU-box session: parity tests, smoke, bench (both arms), probes, launch list, full ncu capture of the pipeline, timeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 python tools/gpu_probe.py 3449 0:1:S:hex 1:1:S:hex 0:1:M:hex 1:1:M:hex > gpurun_out/probe.log 2>&1
timeout 600 python tools/gpu_probe.py 1024 >> gpurun_out/probe.log 2>&1
timeout 600 python tools/gpu_probe_enc.py 1024 >> gpurun_out/probe.log 2>&1; cat gpurun_out/probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
# one full capture of one step's pipeline kernels (8 k_zp_ launches per step; 3 warm-up steps + the gate)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_zp_ -s 24 -c 8 -o gpurun_out/zp_full -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python tools/zp_timeline.py > gpurun_out/timeline.txt 2>&1; tail -9 gpurun_out/timeline.txt
