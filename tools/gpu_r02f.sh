#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
CXPROF_ZSTD=1 CRYOGPU_LIB=tools/_prof/libcryogpu_prof.so timeout 300 python tools/gpu_cxprof.py > gpurun_out/cxprof_r02f.txt 2>&1; grep -v Warn gpurun_out/cxprof_r02f.txt
timeout 600 python tools/gpu_probe.py 1024 0:1:M:lowcard 0:1:D:lowcard 1:1:M:lowcard 1:1:D:lowcard > gpurun_out/probe_r02f.log 2>&1; cat gpurun_out/probe_r02f.log
( time timeout 900 python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench.err ) 2>&1 | grep real; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02f.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'], d['cpu_baseline'])
for c in d['secondary'] or []:
    print(c['op'], c['codec'], c['blocks'], round(c['value'],1), round(c['roofline_frac'],4), c.get('bit_exact_all_blocks', c.get('roundtrip_through_reference_decompressor')), c.get('ratio_vs_reference'), {k: (round(v,2) if isinstance(v,float) else v) for k,v in c['cpu_reference'].items()})
PY
