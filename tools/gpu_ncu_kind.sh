#!/bin/bash
# full ncu capture of kernel $1 on a table of kind $3/$4 with $2 rows (through tools/zp_probe_once.py)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -o gpurun_out/$5 -f python tools/gpu_probe.py 512 1:1:$3:$4 > gpurun_out/ncu_$5.log 2>&1
tail -2 gpurun_out/ncu_$5.log
