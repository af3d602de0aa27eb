#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_1m_v3.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['step_ms'])"
ncu --set full --clock-control none -k regex:pr_iter_kernel -s 2 -c 1 -o gpurun_out/prof_pagerank24 python tools/bench_pagerank.py --scale 24 --reps 1 --no-cpu > gpurun_out/ncu_pr24.log 2>&1
tail -2 gpurun_out/ncu_pr24.log
