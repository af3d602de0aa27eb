#!/bin/bash
# round-2 call 2 (2 GPUs): sharded C ABI at world 1 and 2, bench at N=1 (both arms) and N=2 (both exchanges)
mkdir -p gpurun_out/c2
python -m pytest tests/test_sharded_gpu.py -x -q -s > gpurun_out/c2/pytest_sharded.txt 2>&1
tail -30 gpurun_out/c2/pytest_sharded.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c2/bench_ref_n1.json 2> gpurun_out/c2/bench_ref_n1.err
tail -c 1500 gpurun_out/c2/bench_ref_n1.json; tail -5 gpurun_out/c2/bench_ref_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c2/bench_n1.json 2> gpurun_out/c2/bench_n1.err
tail -c 3000 gpurun_out/c2/bench_n1.json; tail -5 gpurun_out/c2/bench_n1.err
for x in fused nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus 2 --steps 20 --warmup 5 --exchange $x > gpurun_out/c2/bench_n2_$x.json 2> gpurun_out/c2/bench_n2_$x.err
  tail -c 2500 gpurun_out/c2/bench_n2_$x.json; tail -5 gpurun_out/c2/bench_n2_$x.err
done
