#!/bin/bash
# profile + sweep batch for the HNSW search kernel (run under gpurun from the repo root)
set -x
mkdir -p gpurun_out
# 1. launch list of the search / merge kernels for the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hnsw_search|topk_merge" --csv \
    --log-file gpurun_out/launches_search.csv python bench.py --steps 3 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
# 2. full capture of the search kernel (2 launches after the first)
ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 1 -c 2 \
    -o gpurun_out/prof_search python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
# 3. option sweep at 1M
for st in 2 3 4 6 8; do for w in 2 4 8; do
  echo "== stages=$st wpc=$w"; timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu --opt hnsw.stages=$st --opt hnsw.warps_per_cta=$w 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['ms_per_step'])"
done; done > gpurun_out/sweep.txt 2>&1
# 4. config 3: 10M x 768, ef=200, k=100, batch=65536
timeout 1500 python bench.py --rows 10000000 --batch 65536 --k 100 --steps 3 --warmup 1 --cpu-sample 256 > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err
tail -c 3000 gpurun_out/bench_10m.json
cat gpurun_out/sweep.txt
