#!/bin/bash
# final run of the third session of round 2: parity tests, smoke, the bench line with secondary cells and next rows, the reference
# arm, configs 4/5, probes, launch list, timeline, ncu --set full of every pipeline kernel, memcheck of the pipeline's tests
mkdir -p gpurun_out /tmp/nr
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 900 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02c_bench_ref.json 2>> gpurun_out/bench.err
timeout 900 python bench.py --config 4 > gpurun_out/r02c_bench_cfg4.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --config 5 > gpurun_out/r02c_bench_cfg5.json 2>> gpurun_out/bench.err
timeout 600 python tools/gpu_probe.py 1024 > gpurun_out/r02c_probe.txt 2>&1
timeout 600 python tools/gpu_probe.py 3449 0:1:S:hex 1:1:S:hex 0:1:M:hex 1:1:M:hex >> gpurun_out/r02c_probe.txt 2>&1
grep -v Warn gpurun_out/r02c_probe.txt | grep method
ONE_REPS=100 timeout 600 python tools/gpu_one_block.py 2>&1 | grep "one block" > gpurun_out/r02c_one_block.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/r02c_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
for i in 1 2 3; do timeout 200 python tools/zp_timeline.py 2>&1 | tail -13 >> gpurun_out/r02c_timeline.txt; done
out=gpurun_out/r02c_pipeline_ncu_summary.txt
echo "# ncu --set full --clock-control none --import-source on, one capture per kernel (-k regex:^NAME\$ -s 3 -c 1) of tools/gpu_probe.py 3449 1:1:S:hex" > $out
echo "# (the headline batch: 3449 S/hex zstd-1 frames); per kernel: duration (cold, serialised), DRAM bytes, warp-instructions, issue-active, resident warps, registers, grid x block, dynamic smem" >> $out
for k in k_zp_parse k_zp_prefill_early k_zp_huftab k_zp_literals k_zp_fsetab k_zp_sequences_small k_zp_sequences_large k_zp_prefill k_zp_execute k_zp_check; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k\$" -s 3 -c 1 -o /tmp/nr/$k -f python tools/gpu_probe.py 3449 1:1:S:hex > /tmp/nr/$k.log 2>&1
  python tools/ncu_kernel_summary.py /tmp/nr/$k.ncu-rep headline >> $out 2>&1
  python tools/ncu_lines.py /tmp/nr/$k.ncu-rep 25 | cut -c1-200 > gpurun_out/r02c_${k}_lines.txt 2>&1
done
cat $out
OUT=gpurun_out/r02c_sanitizer.txt
echo "# compute-sanitizer --tool memcheck on one B200 (gpurun), code of the third session of round 2 (zstd pipeline: early pass, jobs, units)" > $OUT
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_zstd_decode.py -k "bit_exact or mixed or malformed or large_batch" -x -q > /tmp/san.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" /tmp/san.log | tail -3 | sed 's/^/    /' >> $OUT
grep -E "Invalid|hazard" /tmp/san.log | head -5 | sed 's/^/    /' >> $OUT
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_zstd_encode.py -k "roundtrip" -x -q > /tmp/san.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" /tmp/san.log | tail -3 | sed 's/^/    /' >> $OUT
cat $OUT
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print('headline', d['ms_per_step'], d['value'], d['roofline']['frac'], {k:v for k,v in d['e2e'].items() if k in ('value','value_every_byte_written')}, d['cpu_baseline']['value'])
for c in d['secondary']:
    print(' ', c['op'], c['codec'], c['blocks'], round(c['value'],1), round(c['roofline_frac'],4), c.get('bit_exact_all_blocks', c.get('roundtrip_through_reference_decompressor')), c.get('ratio_vs_reference'), round(c['cpu_reference']['all_cores'],1))
r=json.loads(open('gpurun_out/r02c_bench_ref.json').read().strip().splitlines()[-1]); print('reference arm', r['value'], r['ms_per_step'])
PY
