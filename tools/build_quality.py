"""Is the batched device builder as good as the reference's sequential insert on HIGH-dimensional data?
Builds the same 20k x 768 corpus with the oracle (faithful hnsw_put_vector order) and on the device and
compares recall@10 against brute force for both graphs, searched by the same oracle code."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import gen_vectors  # noqa: E402
from cozo_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402

capi.init(0)
n, dim, m, efc = (int(os.environ.get("BQ_N", 20000)), int(os.environ.get("BQ_DIM", 768)), 16,
                  int(os.environ.get("BQ_EFC", 100)))
X = gen_vectors(n, dim, 7)
Q = gen_vectors(500, dim, 8)
t0 = time.perf_counter()
seq = O.OracleHnsw.new(n, dim, m=m, ef_construction=efc, level_seed=3)
seq.insert_all(X)
t_seq = time.perf_counter() - t0
t0 = time.perf_counter()
g = capi.HnswIndex.build(X, m=m, ef_construction=efc, level_seed=3)
t_dev = time.perf_counter() - t0
dev = O.OracleHnsw.from_levels(X, O.HnswLevels(*g.export_levels()))
bi, _ = O.bruteforce_knn(X, Q, 10, n_threads=os.cpu_count())
out = {"n": n, "dim": dim, "m": m, "ef_construction": efc, "build_s": {"oracle_sequential": t_seq, "device": t_dev}}
for ef in (50, 200):
    r = {}
    for name, ix in (("oracle_sequential", seq), ("device", dev)):
        ids, _, _, st = ix.search(Q, 10, ef, n_threads=os.cpu_count())
        r[name] = {"recall_at_10_vs_bruteforce": float(np.mean([len(set(a) & set(b)) / 10 for a, b in zip(ids, bi)])),
                   "dist_evals_per_query": float(st[:, 0].mean())}
    out[f"ef={ef}"] = r
lv = seq.levels()
out["mean_degree_layer0"] = {"oracle_sequential": float(len(lv.col_idx[0]) / n),
                             "device": float(len(g.export_levels()[2][0]) / n)}
print(json.dumps(out))
