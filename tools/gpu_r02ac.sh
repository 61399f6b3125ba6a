#!/bin/bash
# M/hex regression: which of jobs / early pass, and where the time goes
mkdir -p gpurun_out
for cfg in "CRYOGPU_ZP_EARLY_CTAS=0 CRYOGPU_ZP_JOBS=0" "CRYOGPU_ZP_EARLY_CTAS=0 CRYOGPU_ZP_JOBS=1" "CRYOGPU_ZP_JOBS=0" "CRYOGPU_X=1"; do
  echo "== $cfg" >> gpurun_out/mhex_r02ac.txt
  env $cfg timeout 200 python tools/zp_timeline.py 296960 M hex 2>&1 | tail -14 >> gpurun_out/mhex_r02ac.txt
done
cat gpurun_out/mhex_r02ac.txt
