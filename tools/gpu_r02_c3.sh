#!/bin/bash
# round-2 call 3 (1 GPU): PageRank propagation-blocking engine, filter mask, F64, graph streams, bench N=1
mkdir -p gpurun_out/c3
python -m pytest tests/test_graph_gpu.py -q 2>&1 | tail -40 > gpurun_out/c3/pytest_graph.txt
cat gpurun_out/c3/pytest_graph.txt
python -m pytest tests/test_hnsw_gpu.py tests/test_host_gpu.py tests/test_sharded_gpu.py -q 2>&1 | tail -30 > gpurun_out/c3/pytest_hnsw.txt
cat gpurun_out/c3/pytest_hnsw.txt
for mode in 1 0; do
  timeout 900 python bench.py --workload pagerank --steps 5 --warmup 3 --opt pagerank.mode=$mode > gpurun_out/c3/bench_pagerank_mode$mode.json 2> gpurun_out/c3/bench_pagerank_mode$mode.err
  tail -c 2500 gpurun_out/c3/bench_pagerank_mode$mode.json; tail -5 gpurun_out/c3/bench_pagerank_mode$mode.err
done
# geometry sweep of the blocking (no CPU leg)
for o in "pagerank.window=49152" "pagerank.window=12288" "pagerank.hub_slots=8192" "pagerank.hub_slots=32768" \
         "pagerank.group_slots=16384" "pagerank.group_slots=49152" "pagerank.chunk=1048576"; do
  echo "== sweep $o"
  timeout 300 python bench.py --workload pagerank --steps 3 --warmup 3 --no-cpu --opt $o 2>&1 | tail -c 700
done > gpurun_out/c3/pagerank_sweep.txt 2>&1
grep -E "sweep|ms_per_iteration" gpurun_out/c3/pagerank_sweep.txt | sed 's/.*"ms_per_iteration": \([0-9.]*\).*/  ms_per_iteration \1/'
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c3/bench_n1.json 2> gpurun_out/c3/bench_n1.err
tail -c 3500 gpurun_out/c3/bench_n1.json; tail -5 gpurun_out/c3/bench_n1.err
python -m pytest tests/test_fullsize_gpu.py -q -s -k pagerank 2>&1 | tail -15 > gpurun_out/c3/pytest_fullsize_pr.txt
cat gpurun_out/c3/pytest_fullsize_pr.txt
