#!/bin/bash
# last measurements of the round with the final build: parity tests, smoke, the bench line, the reference arm, timelines, repeats
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 900 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02c_bench_ref.json 2>> gpurun_out/bench.err
rm -f gpurun_out/r02c_timeline.txt gpurun_out/r02c_repeat.txt
for i in 1 2 3; do timeout 200 python tools/zp_timeline.py 2>&1 | tail -13 >> gpurun_out/r02c_timeline.txt; done
for i in 1 2 3 4; do timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_per_step_min_median_max'], d['roofline']['frac'])" >> gpurun_out/r02c_repeat.txt; done; cat gpurun_out/r02c_repeat.txt
CRYOGPU_ZP_JOBS=0 CRYOGPU_ZP_EARLY_CTAS=0 timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('jobs and early pass off', d['ms_per_step'], d['ms_per_step_min_median_max'], d['roofline']['frac'])" >> gpurun_out/r02c_repeat.txt; tail -1 gpurun_out/r02c_repeat.txt
timeout 600 python tools/gpu_probe.py 1024 > gpurun_out/r02c_probe.txt 2>&1
timeout 600 python tools/gpu_probe.py 3449 0:1:S:hex 1:1:S:hex 0:1:M:hex 1:1:M:hex >> gpurun_out/r02c_probe.txt 2>&1
grep -v Warn gpurun_out/r02c_probe.txt | grep "method=1"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['ms_per_step_min_median_max'], d['value'], d['roofline']['frac'], {k:v for k,v in d['e2e'].items() if k in ('value','value_every_byte_written')}, d['cpu_baseline']['value'])
for c in d['secondary']:
    print(' ', c['op'], c['codec'], c['blocks'], round(c['value'],1), round(c['roofline_frac'],4), c.get('bit_exact_all_blocks', c.get('roundtrip_through_reference_decompressor')), round(c['cpu_reference']['all_cores'],1))
r=json.loads(open('gpurun_out/r02c_bench_ref.json').read().strip().splitlines()[-1]); print('reference arm', r['value'], r['ms_per_step'])
PY
