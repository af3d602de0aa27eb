#!/usr/bin/env python
"""Build one index, then time the search kernel under several option sets (same graph, same queries)."""
import argparse
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import gen_vectors  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--ef", type=int, default=200)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--grid", default="hnsw.min_blocks=4,7;hnsw.stages=2,3,4")
    a = ap.parse_args()
    import torch
    from cozo_b200 import capi
    capi.init(0)
    X = gen_vectors(a.n, a.dim, 0x5EED0001)
    g = capi.HnswIndex.build(X, m=16, ef_construction=200, level_seed=0x5EED0003)
    nsteps = a.steps + 2
    Q = torch.from_numpy(gen_vectors(a.batch * nsteps, a.dim, 0x5EED0002).reshape(nsteps, a.batch, a.dim)).cuda()
    ids = torch.empty((a.batch, a.k), dtype=torch.int32, device="cuda")
    dd = torch.empty((a.batch, a.k), dtype=torch.float32, device="cuda")
    qs = torch.zeros((a.batch, 4), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    axes = []
    for part in a.grid.split(";"):
        name, vals = part.split("=")
        axes.append([(name, int(v)) for v in vals.split(",")])
    ref_ids = None
    for combo in itertools.product(*axes):
        for name, v in combo:
            capi.set_option(name, v)
        ts = []
        for s in range(nsteps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.search_dev(Q[s].data_ptr(), a.batch, a.k, a.ef, ids.data_ptr(), dd.data_ptr(), None, qs.data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            if s >= 2:
                ts.append(e0.elapsed_time(e1))
        st = qs.to(torch.int64).sum(0).cpu().numpy()
        cur = ids.cpu().numpy().copy()
        same = None if ref_ids is None else float((cur == ref_ids).mean())
        ref_ids = cur if ref_ids is None else ref_ids
        bytes_ = st[0] * a.dim * 4 + st[1] * 8 + st[2] * 4 + a.batch * a.dim * 4
        ms = float(np.mean(ts))
        print(json.dumps({"opts": dict(combo), "ms": ms, "qps": a.batch / ms * 1e3, "frac": bytes_ / (ms / 1e3) / 1e9 / 6574.5,
                          "dist_evals_q": float(st[0]) / a.batch, "ids_same_as_first": same}), flush=True)


if __name__ == "__main__":
    main()
