#!/bin/bash
# round 2, session 3: ncu --set full of every pipeline kernel of one headline step (one capture per kernel: a capture
# of all of them at once saw none of the side streams' launches in this round, profiles/README.md), with source lines
mkdir -p gpurun_out /tmp/nr
out=gpurun_out/r02_pipeline_ncu_summary.txt
echo "# ncu --set full --clock-control none --import-source on, one capture per kernel (-k regex:^NAME\$ -s 3 -c 1) of tools/gpu_probe.py 3449 1:1:S:hex" > $out
echo "# (the headline batch: 3449 S/hex zstd-1 frames); per kernel: duration (cold, serialised), DRAM bytes, warp-instructions, issue-active, resident warps, registers, grid x block, dynamic smem" >> $out
for k in k_zp_parse k_zp_huftab k_zp_literals k_zp_fsetab k_zp_sequences_small k_zp_sequences_large k_zp_prefill k_zp_execute; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 3 -c 1 -o /tmp/nr/$k -f python tools/gpu_probe.py 3449 1:1:S:hex > /tmp/nr/$k.log 2>&1
  python tools/ncu_kernel_summary.py /tmp/nr/$k.ncu-rep headline >> $out 2>&1
  python tools/ncu_lines.py /tmp/nr/$k.ncu-rep 30 | cut -c1-200 > gpurun_out/r02t_${k}_lines.txt 2>&1
done
cat $out
ncu -i /tmp/nr/k_zp_prefill.ncu-rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for n in ('dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','dram__cycles_active.avg.pct_of_peak_sustained_elapsed'):
    if n in h: print(n, r[h.index(n)])
" > gpurun_out/r02t_prefill_metrics.txt 2>&1; cat gpurun_out/r02t_prefill_metrics.txt
head -14 gpurun_out/r02t_k_zp_prefill_lines.txt
