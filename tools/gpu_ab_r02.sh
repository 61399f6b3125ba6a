#!/bin/bash
# A/B: which of the round-2 additions slowed the headline step
mkdir -p gpurun_out
for cfg in "default" "CRYOGPU_ZSTD_EXEC=warp" "CRYOGPU_LZ4_KERNEL=warp" "CRYOGPU_ZSTD_EXEC=warp CRYOGPU_LZ4_KERNEL=warp"; do
  echo "== $cfg" >> gpurun_out/ab_r02.txt
  if [ "$cfg" = "default" ]; then timeout 300 python bench.py --no-cpu --no-e2e --steps 10 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/ab_r02.txt
  else env $cfg timeout 300 python bench.py --no-cpu --no-e2e --steps 10 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/ab_r02.txt; fi
done
cat gpurun_out/ab_r02.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
grep -v "at::" gpurun_out/launches_r02c.csv | awk -F'","' '{print $5, $NF}' | tail -36
