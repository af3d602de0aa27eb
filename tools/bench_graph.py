#!/usr/bin/env python
"""SSSP family on the reference's air-routes graph and a larger random graph: device time vs oracle time."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cozo_b200 import capi  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.test_air_routes_cpu import load_routes  # noqa: E402

capi.init(0)
cores = os.cpu_count()
out = {}
n, src, dst, w, _, _ = load_routes()
for name, (nn, s, d, ww) in {"air_routes(3476n,50637e)": (n, src, dst, w)}.items():
    g = capi.Graph(nn, s, d, ww)
    o = O.OracleGraph(nn, s, d, ww)
    g.closeness()  # warm
    t0 = time.perf_counter(); gc, ms_c = g.closeness(); wall_c = time.perf_counter() - t0
    t0 = time.perf_counter(); gb, ms_b = g.betweenness(); wall_b = time.perf_counter() - t0
    t0 = time.perf_counter(); oc = o.closeness(n_threads=cores); cpu_c = time.perf_counter() - t0
    t0 = time.perf_counter(); ob = o.betweenness(n_threads=cores); cpu_b = time.perf_counter() - t0
    srcs = np.arange(0, nn, 7, dtype=np.uint32)
    t0 = time.perf_counter(); gd, gp, ms_s = g.sssp(srcs); wall_s = time.perf_counter() - t0
    t0 = time.perf_counter(); od, _ = o.sssp(srcs, n_threads=cores); cpu_s = time.perf_counter() - t0
    out[name] = {"closeness": {"gpu_kernel_ms": ms_c, "gpu_wall_ms": 1e3 * wall_c, "cpu_ms": 1e3 * cpu_c},
                 "betweenness": {"gpu_kernel_ms": ms_b, "gpu_wall_ms": 1e3 * wall_b, "cpu_ms": 1e3 * cpu_b,
                                 "max_abs_diff": float(np.max(np.abs(gb - ob)))},
                 "sssp_%d_sources" % srcs.size: {"gpu_kernel_ms": ms_s, "gpu_wall_ms": 1e3 * wall_s, "cpu_ms": 1e3 * cpu_s,
                                                 "dist_identical": bool(np.array_equal(gd, od))},
                 "cpu_threads": cores}
# multi-source SSSP on R-MAT graphs that do not fit shared memory (compacted frontier queues): time per source next
# to the 8E + 12N bytes-per-relaxation-pass model (a lower bound of ONE pass: a label-correcting search
# makes several) and the oracle's Dijkstra on the host threads
from tests.util import rmat_edges  # noqa: E402
for scale in [int(x) for x in os.environ.get("COZO_BENCH_RMAT_SCALES", "18,20").split(",") if x]:
    nn, s, d = rmat_edges(scale, 16, 0x5EED0004)
    ww = (np.random.default_rng(scale).integers(1, 256, s.size) / 16.0).astype(np.float32)
    g = capi.Graph(nn, s, d, ww)
    o = O.OracleGraph(nn, s, d, ww)
    srcs = np.arange(0, nn, nn // 64, dtype=np.uint32)[:64]
    g.sssp(srcs[:2], want_pred=False)
    t0 = time.perf_counter(); gd, _, ms_s = g.sssp(srcs, want_pred=False); wall_s = time.perf_counter() - t0
    t0 = time.perf_counter(); od, _ = o.sssp(srcs, n_threads=cores); cpu_s = time.perf_counter() - t0
    model = 8 * s.size + 12 * nn
    out[f"rmat{scale}(n={nn},m={s.size})"] = {
        "sssp_64_sources": {"gpu_kernel_ms": ms_s, "gpu_wall_ms": 1e3 * wall_s, "cpu_ms": 1e3 * cpu_s,
                            "dist_identical": bool(np.array_equal(gd, od)), "ms_per_source": ms_s / 64,
                            "one_pass_bytes_model": model,
                            "one_pass_model_GBs_per_source": model / (ms_s / 64 / 1e3) / 1e9},
        "cpu_threads": cores}
print(json.dumps(out))
