#!/bin/bash
# full ncu capture of one launch of kernel $1 (regex) from the headline bench; report -> gpurun_out/$2.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-3} -c 1 -o gpurun_out/$2 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log
