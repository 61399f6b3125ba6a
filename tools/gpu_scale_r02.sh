#!/bin/bash
# scaling on one multi-GPU box: headline (weak) at 1/2/4/8, config 4 (zstd levels, 10 GB, strong) at 1/2/4/8
mkdir -p gpurun_out
NG=${1:-8}
for n in 1 2 4 8; do
  [ $n -gt $NG ] && continue
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-secondary 2>gpurun_out/scale_err_$n.txt | tail -1 > gpurun_out/r02_scale_$n.json;
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/scale_err_$n.txt | grep '^{' | tail -1 > gpurun_out/r02_scale_$n.json; fi
  python -c "
import json; j=json.load(open('gpurun_out/r02_scale_$n.json')); print('headline', $n, round(j['value'],1), round(j['ms_per_step'],4), round(j['e2e']['value'],1) if j['e2e'] else None, round(j['e2e']['value_every_byte_written'],1) if j['e2e'] else None)"
done
for n in 1 2 4 8; do
  [ $n -gt $NG ] && continue
  if [ $n = 1 ]; then python bench.py --config 4 --gpus 1 2>>gpurun_out/scale_err_$n.txt | tail -1 > gpurun_out/r02_cfg4_$n.json;
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --config 4 --gpus $n 2>>gpurun_out/scale_err_$n.txt | grep '^{' | tail -1 > gpurun_out/r02_cfg4_$n.json; fi
  python -c "
import json; j=json.load(open('gpurun_out/r02_cfg4_$n.json'))
l=[r for r in j['levels'] if r['zstd_compression_level']==1][0]
print('config4', $n, 'level 1 compress', round(l['compress_GBps_in'],1), 'decompress', round(l['decompress_GBps_out'],1), 'level -5', round(j['levels'][0]['compress_GBps_in'],1), round(j['levels'][0]['decompress_GBps_out'],1))"
done
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-200
nproc
