#!/bin/bash
# round-1 close-out: full GPU suite, ncu capture of the PageRank pull kernel, default bench, smoke
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pr_pull_kernel -s 3 -c 1 \
    -o gpurun_out/prof_pagerank_v6 python tools/bench_pagerank.py --no-cpu --reps 1 > gpurun_out/ncu_pr.log 2>&1
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_final2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_final2.json')); print({k:d[k] for k in ['value','ms_per_step','recall_at_k_vs_oracle','e2e','clocks','gpu_launches']}); print(d['roofline']); print(d['cpu_baseline'])"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
