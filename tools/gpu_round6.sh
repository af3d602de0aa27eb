#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python tools/bench_graph.py 2>&1 | tail -1 | tee gpurun_out/graph_air_routes.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.json; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print({k:d[k] for k in ['value','ms_per_step','recall_at_k_vs_oracle','cpu_baseline','e2e','clocks','gpu_launches']}); print(d['roofline'])"
