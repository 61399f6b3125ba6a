#!/bin/bash
# full capture of the warp executor of the headline batch with source lines
mkdir -p gpurun_out /tmp/nr
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_zp_execute$' -s 3 -c 1 -o /tmp/nr/zp_exec -f python tools/gpu_probe.py 3449 1:1:S:hex > /tmp/nr/zp_exec.log 2>&1
tail -3 /tmp/nr/zp_exec.log
python tools/ncu_kernel_summary.py /tmp/nr/zp_exec.ncu-rep headline_execute > gpurun_out/r02_zp_execute_ncu.txt 2>&1
python tools/ncu_lines.py /tmp/nr/zp_exec.ncu-rep 40 | cut -c1-200 >> gpurun_out/r02_zp_execute_ncu.txt 2>&1
cat gpurun_out/r02_zp_execute_ncu.txt
