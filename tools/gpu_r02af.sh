#!/bin/bash
# run-to-run: eight fresh processes of the timeline tool and of the bench, what differs in the slow ones
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  echo "== process $i" >> gpurun_out/states_r02af.txt
  timeout 200 python tools/zp_timeline.py 2>&1 | grep -A12 "iteration 2\|iteration 3" | grep -v "lz4_cta\|execute_cta\|seq_large\|parse\|huftab\|fsetab" >> gpurun_out/states_r02af.txt
done
for i in 1 2 3 4 5 6; do
  timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_per_step_min_median_max'])" >> gpurun_out/states_r02af.txt
done
cat gpurun_out/states_r02af.txt
