"""Freezes small oracle outputs as regression fixtures (tests/golden/oracle_small.npz).
These are NOT reference outputs (the reference cannot run here; parity is unpinned);
they pin the oracle against accidental change and give the GPU tests a frozen expectation."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(0x5EED0001)
X = rng.random((600, 24), dtype=np.float32)
Q = np.random.default_rng(0x5EED0002).random((40, 24), dtype=np.float32)
ix = O.OracleHnsw.new(600, 24, m=6, ef_construction=40, level_seed=0x5EED0003)
ix.insert_all(X)
lv = ix.levels()
ids, dist, cnt, st = ix.search(Q, 8, 32)
src = rng.integers(0, 200, 1500).astype(np.uint32)
dst = rng.integers(0, 200, 1500).astype(np.uint32)
w = (rng.integers(1, 40, 1500) / 4.0).astype(np.float32)
g = O.OracleGraph(200, src, dst, w)
pr, it, err = O.OracleGraph(200, src, dst).pagerank(0.85, 1e-4, 10)
sd, _ = g.sssp(np.arange(0, 200, 25, dtype=np.uint32))
out = dict(X=X, Q=Q, ids=ids, dist=dist, cnt=cnt, stats=st, entry=np.int64(lv.entry), n_levels=np.int64(lv.n_levels),
           deg0=np.diff(lv.row_ptr[0]).astype(np.uint16), col0=lv.col_idx[0], src=src, dst=dst, w=w, pr=pr,
           pr_iters=np.int64(it), sssp=sd, closeness=g.closeness(), betweenness=g.betweenness())
np.savez_compressed(os.path.join(os.path.dirname(__file__), "oracle_small.npz"), **out)
print("written", {k: getattr(v, "shape", None) for k, v in out.items()})
