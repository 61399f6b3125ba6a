#!/bin/bash
# round-2 profiling session: launch list of the bench command; full captures of the pipeline kernels of one headline step
# and of the CTA-per-block decoders on dense low-cardinality blocks, summarised on the box
mkdir -p gpurun_out /tmp/nr
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
grep -v "at::" gpurun_out/r02_launches_bench.csv | awk -F'","' '{print $5, $NF}' | tail -14
# one full capture of one step's pipeline kernels (9 k_zp_ launches per step; 3 warm-up steps + the gate = 4 steps before)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_zp_ -s 36 -c 9 -o /tmp/nr/zp_full -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python tools/ncu_kernel_summary.py /tmp/nr/zp_full.ncu-rep r02_headline_step > gpurun_out/r02_pipeline_ncu_summary.txt 2>&1; cat gpurun_out/r02_pipeline_ncu_summary.txt | cut -c1-260
cap() { # label kernel-regex skip command...
  local label=$1 k=$2 s=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o /tmp/nr/$label -f "$@" > /tmp/nr/$label.log 2>&1
  python tools/ncu_kernel_summary.py /tmp/nr/$label.ncu-rep "$label" >> gpurun_out/r02_cx_ncu_summary.txt
  python tools/ncu_lines.py /tmp/nr/$label.ncu-rep 8 | cut -c1-170 >> gpurun_out/r02_cx_ncu_summary.txt
  echo >> gpurun_out/r02_cx_ncu_summary.txt
}
rm -f gpurun_out/r02_cx_ncu_summary.txt
cap lz4_cta_decode_Dlowcard_1024   k_lz4_decode_c 3 python tools/gpu_probe.py 1024 0:1:D:lowcard
cap zstd_cta_execute_Dlowcard_1024 k_zp_execute_c 3 python tools/gpu_probe.py 1024 1:1:D:lowcard
cap lz4_cta_decode_one_block_S     k_lz4_decode_c 3 python tools/gpu_probe.py 1 0:1:S:hex
cat gpurun_out/r02_cx_ncu_summary.txt | cut -c1-200
