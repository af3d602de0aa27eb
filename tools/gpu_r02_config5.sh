#!/bin/bash
# BASELINE config 5 (8 x 12.5M x 768, ef=200, k=10, B=1M tiled by 65536) + the strong-scaling point at N=8.
# Run with: gpurun --gpus 8 --timeout 2400 -- 'bash tools/gpu_r02_config5.sh'
mkdir -p gpurun_out/c5
N=${N:-8}
run() { # name, extra args
  name=$1; shift
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 \
    bench.py --gpus $N --watchdog-s 1450 "$@" > gpurun_out/c5/$name.json 2> gpurun_out/c5/$name.err
  echo "== $name rc=$?"; tail -c 3000 gpurun_out/c5/$name.json; tail -3 gpurun_out/c5/$name.err
}
free -g | head -2
# pre-flight: the same code path at toy size (device-side generation, borrowed vectors, 3 tiles, both exchanges, parity)
run preflight_n$N --workload config5 --rows 150000 --batch 150000 --tile 65536 --device-gen 1 --steps 2 --warmup 1 --also-exchange nccl --parity-sample 32
grep -q '"merged_equals_numpy_merge_of_gpu_lists": true' gpurun_out/c5/preflight_n$N.json || { echo "PREFLIGHT FAILED"; exit 1; }
run strong10m_n$N --workload strong10m --steps 10 --warmup 3 --also-exchange nccl --parity-sample 128
run config5_n$N --workload config5 --steps 3 --warmup 1 --also-exchange nccl --parity-sample 64
