#!/usr/bin/env python
"""Latency / throughput of the batched search as a function of the batch size (1M x 768, ef=200, k=10)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import gen_vectors  # noqa: E402


def main():
    import torch
    from cozo_b200 import capi
    capi.init(0)
    n, dim, k, ef = 1_000_000, 768, 10, 200
    X = gen_vectors(n, dim, 0x5EED0001)
    g = capi.HnswIndex.build(X, m=16, ef_construction=200, level_seed=0x5EED0003)
    Qall = torch.from_numpy(gen_vectors(65536, dim, 0x5EED0002)).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    modes = [int(x) for x in os.environ.get("SWEEP_MODES", "1,2").split(",")]
    for mode in modes:
      capi.set_option("hnsw.mode", mode)
      for B in (1, 8, 64, 512, 2368, 4096, 4736, 16384, 65536):
        if mode == 2 and B > 16384:
            continue
        ids = torch.empty((B, k), dtype=torch.int32, device="cuda")
        dd = torch.empty((B, k), dtype=torch.float32, device="cuda")
        ts = []
        for rep in range(7):
            off = (rep * B) % (65536 - B + 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.search_dev(Qall[off:off + B].data_ptr(), B, k, ef, ids.data_ptr(), dd.data_ptr(), None, None, stream)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(json.dumps({"mode": mode, "batch": B, "ms": ms, "qps": B / ms * 1e3, "us_per_query": ms / B * 1e3,
                          "checksum": int(ids.long().sum().item())}), flush=True)


if __name__ == "__main__":
    main()
