#!/bin/bash
# round-2 call 4 (1 GPU): everything new under test + memcheck + PageRank benches
mkdir -p gpurun_out/c4
python -m pytest tests/test_graph_gpu.py -q 2>&1 | tail -40 > gpurun_out/c4/pytest_graph.txt
cat gpurun_out/c4/pytest_graph.txt
python -m pytest tests/test_hnsw_gpu.py tests/test_host_gpu.py tests/test_sharded_gpu.py -q 2>&1 | tail -40 > gpurun_out/c4/pytest_hnsw.txt
cat gpurun_out/c4/pytest_hnsw.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py > gpurun_out/c4/memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -8 gpurun_out/c4/memcheck.txt
for mode in 1; do
  timeout 900 python bench.py --workload pagerank --steps 5 --warmup 3 --opt pagerank.mode=$mode > gpurun_out/c4/bench_pagerank_mode$mode.json 2> gpurun_out/c4/bench_pagerank_mode$mode.err
  tail -c 2500 gpurun_out/c4/bench_pagerank_mode$mode.json; tail -5 gpurun_out/c4/bench_pagerank_mode$mode.err
done
for o in "pagerank.window=49152" "pagerank.window=12288" "pagerank.hub_slots=8192" "pagerank.hub_slots=32768" \
         "pagerank.group_slots=16384" "pagerank.group_slots=49152" "pagerank.chunk=1048576"; do
  echo "== sweep $o"
  timeout 300 python bench.py --workload pagerank --steps 3 --warmup 3 --no-cpu --opt $o 2>&1 | tail -c 900
done > gpurun_out/c4/pagerank_sweep.txt 2>&1
grep -E "sweep|ms_per_iteration" gpurun_out/c4/pagerank_sweep.txt | sed 's/.*"ms_per_iteration": \([0-9.]*\).*/  ms_per_iteration \1/'
# pre-flight of the config-5 code path at toy size on one GPU (device-side generation, borrowed vectors, B > tile)
timeout 600 python bench.py --workload config5 --rows 300000 --batch 200000 --device-gen 1 --steps 2 --warmup 1 > gpurun_out/c4/bench_config5_preflight.json 2> gpurun_out/c4/bench_config5_preflight.err
tail -c 1200 gpurun_out/c4/bench_config5_preflight.json; tail -3 gpurun_out/c4/bench_config5_preflight.err
# per-kernel times of one PageRank call
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pb_ --csv --log-file gpurun_out/c4/pagerank_launches.csv \
  python bench.py --workload pagerank --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/c4/pagerank_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
for k, v in agg.items(): print(k, len(v), "launches, mean us", sum(v) / len(v) / 1e3)
PY
python -m pytest tests/test_fullsize_gpu.py -q -s -k pagerank 2>&1 | tail -15 > gpurun_out/c4/pytest_fullsize_pr.txt
cat gpurun_out/c4/pytest_fullsize_pr.txt
