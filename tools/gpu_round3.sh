#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_host_gpu.py -m gpu -q 2>&1 | tail -8
timeout 900 python tools/bench_pagerank.py --scale 24 2>&1 | tail -1 | tee gpurun_out/pagerank_rmat24_v1.json
ncu --set full --clock-control none --import-source on -k regex:pr_iter_kernel -s 2 -c 1 -o gpurun_out/prof_pagerank python tools/bench_pagerank.py --scale 22 --reps 1 --no-cpu > gpurun_out/ncu_pr.log 2>&1
