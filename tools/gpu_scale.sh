for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/scale_$n.json;
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/scale_$n.json; fi
  python -c "
import json; j=json.load(open('gpurun_out/scale_$n.json')); print($n, j['value'], j['ms_per_step'], j['e2e']['value'] if j['e2e'] else None)"
done
