#!/bin/bash
# final measurements of the third session of round 2 (after the last change to the raw / RLE stage)
mkdir -p gpurun_out /tmp/nr
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 900 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02c_bench_ref.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02c_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
CRYOGPU_ZP_JOBS=0 CRYOGPU_ZP_EARLY_CTAS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02c_launches_bench_nojobs.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/bench_under_ncu2.log 2>&1
rm -f gpurun_out/r02c_timeline.txt; for i in 1 2 3; do timeout 200 python tools/zp_timeline.py 2>&1 | tail -13 >> gpurun_out/r02c_timeline.txt; done
out=gpurun_out/r02c_pipeline_ncu_summary.txt
echo "# ncu --set full --clock-control none --import-source on, one capture per kernel (-k regex:^NAME\$ -s 3 -c 1) of tools/gpu_probe.py 3449 1:1:S:hex" > $out
echo "# (the headline batch: 3449 S/hex zstd-1 frames); per kernel: duration (cold, serialised), DRAM bytes, warp-instructions, issue-active, resident warps, registers, grid x block, dynamic smem" >> $out
for k in k_zp_parse k_zp_prefill_early k_zp_huftab k_zp_literals k_zp_fsetab k_zp_sequences_small k_zp_sequences_large k_zp_prefill k_zp_execute k_zp_check; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 3 -c 1 -o /tmp/nr/$k -f python tools/gpu_probe.py 3449 1:1:S:hex > /tmp/nr/$k.log 2>&1
  python tools/ncu_kernel_summary.py /tmp/nr/$k.ncu-rep headline >> $out 2>&1
  python tools/ncu_lines.py /tmp/nr/$k.ncu-rep 25 | cut -c1-200 > gpurun_out/r02c_${k}_lines.txt 2>&1
done
cat $out
for i in 1 2 3 4; do timeout 200 python bench.py --no-cpu --no-e2e --no-secondary --steps 20 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_per_step_min_median_max'], d['roofline']['frac'])" >> gpurun_out/r02c_repeat.txt; done; cat gpurun_out/r02c_repeat.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['ms_per_step_min_median_max'], d['value'], d['roofline']['frac'], {k:v for k,v in d['e2e'].items() if k in ('value','value_every_byte_written')}, d['cpu_baseline']['value'])
for c in d['secondary']:
    print(' ', c['op'], c['codec'], c['blocks'], round(c['value'],1), round(c['roofline_frac'],4), c.get('bit_exact_all_blocks', c.get('roundtrip_through_reference_decompressor')), c.get('ratio_vs_reference'), round(c['cpu_reference']['all_cores'],1))
for c in d['next_rows']:
    print(' ', {k:(round(v,2) if isinstance(v,float) else v) for k,v in c.items() if k!='api'})
r=json.loads(open('gpurun_out/r02c_bench_ref.json').read().strip().splitlines()[-1]); print('reference arm', r['value'], r['ms_per_step'])
PY
