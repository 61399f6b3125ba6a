#!/bin/bash
# round 2, session 3: what holds the executor's phase -- timelines with the executor's long fills and / or the raw / RLE stage left out
# (timing experiments: the outputs of these builds are wrong on purpose)
mkdir -p gpurun_out
for v in tl tl_nofill tl_nopf tl_none; do
  echo "== $v" >> gpurun_out/ablate_r02x.txt
  TL_LIB=tools/_prof/libcryogpu_$v.so timeout 300 python tools/zp_timeline.py 2>&1 | tail -12 >> gpurun_out/ablate_r02x.txt
done
cat gpurun_out/ablate_r02x.txt
