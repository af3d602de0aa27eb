"""Tiny end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cozo_b200 import capi  # noqa: E402

capi.init(0)
rng = np.random.default_rng(0)
for dim, metric in [(64, capi.L2), (100, capi.COSINE)]:
    X = rng.random((1500, dim), dtype=np.float32) - 0.5
    g = capi.HnswIndex.build(X, metric=metric, m=8, ef_construction=40, keep_pruned_connections=(metric == capi.COSINE))
    Q = rng.random((64, dim), dtype=np.float32) - 0.5
    for mode in (1, 0, 2):                        # TMA ring, ld.global.nc, cooperative CTA-per-query
        capi.set_option("hnsw.mode", mode)
        ids, dist, cnt, st = g.search(Q, 5, 40)
        assert (cnt == 5).all()
    capi.set_option("hnsw.mode", -1)
    ni, rp, ci, ep = g.export_levels()
    g2 = capi.HnswIndex.stage(X, ni, rp, ci, ep, metric=metric, m_max0=16, m_max=8)
    ids2, _, _, _ = g2.search(Q, 5, 40, radius=float(np.median(dist)))
    # maintenance: append, change vectors under existing ids (non-tail ranges), remove, search again
    first = g.insert(rng.random((200, dim), dtype=np.float32) - 0.5)
    assert first == 1500
    g.update(np.array([3, 4, 5, 900, 1699], np.uint32), rng.random((5, dim), dtype=np.float32) - 0.5)
    g.remove(np.arange(100, 160, dtype=np.uint32))
    ids3, _, cnt3, _ = g.search(Q, 5, 40)
    assert (cnt3 == 5).all() and not np.isin(ids3, np.arange(100, 160)).any()
    g.export_levels(); g.export_dists(); g.export_live()
src = rng.integers(0, 300, 3000).astype(np.uint32)
dst = rng.integers(0, 300, 3000).astype(np.uint32)
w = (rng.random(3000) + 0.1).astype(np.float32)
gg = capi.Graph(300, src, dst, w)
gg.pagerank(0.85, 1e-4, 5)
gg.sssp(np.arange(0, 300, 50, dtype=np.uint32))
gg.closeness()
gg.betweenness()
gg.clustering()
gg.sssp_paths(np.array([0, 5], np.uint32), np.array([7, 9], np.uint32), forb_nodes=[[3], []], forb_edges=[[], [(5, 9)]])
# PageRank work items of every kind: a hub row (> 4096 in-edges), a medium row (257..4096), mini-blocks
hsrc = np.concatenate([np.arange(1, 6000), np.arange(2, 900), rng.integers(0, 6000, 4000)]).astype(np.uint32)
hdst = np.concatenate([np.zeros(5999), np.ones(898), rng.integers(0, 6000, 4000)]).astype(np.uint32)
hub = capi.Graph(6000, hsrc, hdst)
for warps in (32, 8):
    capi.set_option("pagerank.warps", warps)
    for dyn in (1, 0):
        capi.set_option("pagerank.dynamic", dyn)
        hub.pagerank(0.85, 0.0, 2)
print("sanitize workload done")
