#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_probe.py 1024 > gpurun_out/probe_r02g.log 2>&1; cat gpurun_out/probe_r02g.log
timeout 600 python tools/gpu_probe.py 3449 0:1:S:hex 1:1:S:hex 0:1:M:hex 1:1:M:hex >> gpurun_out/probe_r02g.log 2>&1; tail -4 gpurun_out/probe_r02g.log
( time timeout 900 python bench.py --config 5 > gpurun_out/bench_cfg5_r02g.json 2> gpurun_out/bench.err ) 2>&1 | grep real; tail -3 gpurun_out/bench.err
( time timeout 900 python bench.py --config 3 > gpurun_out/bench_cfg3_r02g.json 2>> gpurun_out/bench.err ) 2>&1 | grep real; tail -3 gpurun_out/bench.err
( time timeout 1200 python bench.py --config 4 > gpurun_out/bench_cfg4_r02g.json 2>> gpurun_out/bench.err ) 2>&1 | grep real; tail -3 gpurun_out/bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02g.json 2>> gpurun_out/bench.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/bench_ref_r02g.json
python - <<'PY'
import json
for f in ('bench_cfg5_r02g','bench_cfg3_r02g','bench_cfg4_r02g'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'failed', e); continue
    print(f, d['metric'], d['value'])
    for r in d.get('batches', []) + d.get('sweep', []) + d.get('levels', []):
        print('  ', {k: (round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
PY
