#!/bin/bash
# round-2 call 3 (1 GPU): PageRank propagation-blocking engine, filter mask, bench N=1
mkdir -p gpurun_out/c3
python -m pytest tests/test_graph_gpu.py -x -q 2>&1 | tail -40 > gpurun_out/c3/pytest_pagerank.txt
cat gpurun_out/c3/pytest_pagerank.txt
python -m pytest tests/test_hnsw_gpu.py tests/test_host_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/c3/pytest_hnsw.txt
cat gpurun_out/c3/pytest_hnsw.txt
for mode in 1 0; do
  timeout 900 python bench.py --workload pagerank --steps 5 --warmup 3 --opt pagerank.mode=$mode > gpurun_out/c3/bench_pagerank_mode$mode.json 2> gpurun_out/c3/bench_pagerank_mode$mode.err
  tail -c 2500 gpurun_out/c3/bench_pagerank_mode$mode.json; tail -5 gpurun_out/c3/bench_pagerank_mode$mode.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c3/bench_n1.json 2> gpurun_out/c3/bench_n1.err
tail -c 3500 gpurun_out/c3/bench_n1.json; tail -5 gpurun_out/c3/bench_n1.err
python -m pytest tests/test_fullsize_gpu.py -x -q -s -k pagerank 2>&1 | tail -15 > gpurun_out/c3/pytest_fullsize_pr.txt
cat gpurun_out/c3/pytest_fullsize_pr.txt
