#!/bin/bash
# round close-out: full GPU suite, compute-sanitizer memcheck over every kernel, graph bench, default bench, smoke
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|sanitize workload done" gpurun_out/memcheck.log
timeout 300 python tools/bench_graph.py 2>&1 | tail -1 > gpurun_out/graph_air_routes.json; cat gpurun_out/graph_air_routes.json
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_closeout.json
python -c "
import json; d=json.load(open('gpurun_out/bench_closeout.json')); print({k:d[k] for k in ['value','ms_per_step','recall_at_k_vs_oracle','e2e','clocks','gpu_launches']}); print(d['roofline']); print(d['cpu_baseline'])"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
