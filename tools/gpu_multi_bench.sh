#!/bin/bash
mkdir -p gpurun_out
for cfg in "4 fused" "8 fused" "8 nccl"; do set -- $cfg
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 295$1$1 bench.py --gpus $1 --steps 20 --warmup 3 --exchange $2 2>&1 | tail -1 > gpurun_out/bench_n$1_$2.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_n$1_$2.json')); print($1, d['exchange'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'clocks', d['clocks'])"
done
