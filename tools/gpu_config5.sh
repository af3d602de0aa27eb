#!/bin/bash
# BASELINE config 5: 8 shards x 12.5 M x 768, ef=200, k=10, batches of 65 536 queries.
# Not run in round 1 (an 8-GPU call is charged 8x).  ~38 GB of vectors and a ~100 s build per GPU.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_config5.sh'
mkdir -p gpurun_out
N=${N:-8}
timeout 1400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --gpus $N --rows 12500000 --batch 65536 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_config5_n$N.json
