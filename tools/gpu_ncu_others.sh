#!/bin/bash
# full ncu captures of the kernels outside the headline bench, summarised on the box (the reports are too big to bring back together)
mkdir -p gpurun_out /tmp/nr
cap() { # label kernel-regex skip command...
  local label=$1 k=$2 s=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o /tmp/nr/$label -f "$@" > /tmp/nr/$label.log 2>&1
  python tools/ncu_kernel_summary.py /tmp/nr/$label.ncu-rep "$label" >> gpurun_out/others_summary.txt
  python tools/ncu_lines.py /tmp/nr/$label.ncu-rep 6 | cut -c1-170 >> gpurun_out/others_summary.txt
  echo >> gpurun_out/others_summary.txt
}
rm -f gpurun_out/others_summary.txt
cap lz4_decode_S_3449blocks   k_lz4_decode_w 3 python tools/gpu_probe.py 3449 0:1:S:hex
cap lz4_decode_Dlowcard_1024  k_lz4_decode_w 3 python tools/gpu_probe.py 1024 0:1:D:lowcard
cap lz4_encode_S_512          k_lz4_encode   1 python tools/gpu_probe_enc.py 512 0:1:S:hex
cap zstd_encode_S_512         k_zstd_encode  1 python tools/gpu_probe_enc.py 512 1:1:S:hex
cap zstd_encode_Dhex_512      k_zstd_encode  1 python tools/gpu_probe_enc.py 512 1:1:D:hex
cat gpurun_out/others_summary.txt | cut -c1-200
