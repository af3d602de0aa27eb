#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hnsw_search|topk_merge" --csv \
    --log-file gpurun_out/launches_final.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 3 -c 1 \
    -o gpurun_out/prof_search_final python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_final.json
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print({k:d[k] for k in ['value','ms_per_step','recall_at_k_vs_oracle','e2e','clocks']}); print(d['roofline']); print(d['cpu_baseline'])"
python __graft_entry__.py smoke 2>&1 | tail -1
