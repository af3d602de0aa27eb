#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 900 python tools/bench_pagerank.py --scale 24 2>&1 | tail -1 | tee gpurun_out/pagerank_rmat24_v2.json
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_1m_v2.json
