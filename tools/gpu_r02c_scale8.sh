#!/bin/bash
# eight ranks on one box with the final code of round 2 (weak scaling of the headline; the driver runs 1/2/4/8 itself at round end)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-secondary 2>gpurun_out/scale8_err.txt | grep '^{' | tail -1 > gpurun_out/r02c_scale_8.json
python -c "
import json; j=json.load(open('gpurun_out/r02c_scale_8.json')); print('headline 8 GPUs', round(j['value'],1), round(j['ms_per_step'],4), j['ms_per_step_min_median_max'], round(j['e2e']['value'],1) if j['e2e'] else None, round(j['e2e']['value_every_byte_written'],1) if j['e2e'] else None)"
tail -2 gpurun_out/scale8_err.txt | cut -c1-200
