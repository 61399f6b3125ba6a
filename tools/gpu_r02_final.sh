#!/bin/bash
# final round-2 session: parity tests, smoke, bench (default line with secondary + next rows, reference arm, configs 3/4/5),
# probes, launch list, timeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --config 3 > gpurun_out/r02_bench_cfg3.json 2>> gpurun_out/bench.err
timeout 900 python bench.py --config 4 > gpurun_out/r02_bench_cfg4.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --config 5 > gpurun_out/r02_bench_cfg5.json 2>> gpurun_out/bench.err
timeout 600 python tools/gpu_probe.py 1024 > gpurun_out/r02_probe.txt 2>&1
timeout 600 python tools/gpu_probe.py 3449 0:1:S:hex 1:1:S:hex 0:1:M:hex 1:1:M:hex >> gpurun_out/r02_probe.txt 2>&1
timeout 600 python tools/gpu_probe_enc.py 1024 >> gpurun_out/r02_probe.txt 2>&1; grep -v Warn gpurun_out/r02_probe.txt | grep method
ONE_REPS=100 timeout 600 python tools/gpu_one_block.py 2>&1 | grep "one block" > gpurun_out/r02_one_block.txt
CXPROF_ZSTD=1 CRYOGPU_LIB=tools/_prof/libcryogpu_prof.so timeout 300 python tools/gpu_cxprof.py 2>&1 | grep -v "Warn\|d_me" > gpurun_out/r02_cxprof.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['value'], d['roofline']['frac'], {k:v for k,v in d['e2e'].items() if k in ('value','value_every_byte_written')}, d['cpu_baseline']['value'])
for c in d['secondary']:
    print(' ', c['op'], c['codec'], c['blocks'], round(c['value'],1), round(c['roofline_frac'],4), c.get('bit_exact_all_blocks', c.get('roundtrip_through_reference_decompressor')), c.get('ratio_vs_reference'), round(c['cpu_reference']['all_cores'],1))
for c in d['next_rows']:
    print(' ', {k:(round(v,2) if isinstance(v,float) else v) for k,v in c.items() if k!='api'})
r=json.loads(open('gpurun_out/r02_bench_ref.json').read().strip().splitlines()[-1]); print('reference arm', r['value'], r['ms_per_step'])
PY
