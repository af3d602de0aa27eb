#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 600 python tools/sweep.py --grid "hnsw.min_blocks=4,7;hnsw.stages=2,3,4" 2>&1 | tee gpurun_out/sweep2.txt | tail -8
timeout 900 python tools/bench_pagerank.py --scale 24 2>&1 | tail -2 | tee gpurun_out/pagerank_rmat24.json
