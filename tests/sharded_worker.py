"""Worker of tests/test_sharded_gpu.py: one process per GPU (torchrun), drives the sharded operator
through the C ABI only (cozo_gpu_shards_* / cozo_gpu_hnsw_*_sharded) and checks it against
  (a) a numpy merge of the per-shard lists the single-shard call returns, and
  (b) the CPU oracle run shard by shard with the same sharding (that is what parity means).
torch.distributed (gloo) is only the host channel that carries the 128-byte unique id and the checks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


from cozo_b200.sharded import merge_lists as numpy_merge  # noqa: E402  (the merge specification)


def main():
    import torch
    import torch.distributed as dist
    from cozo_b200 import capi
    from oracle import oracle as O
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    capi.init(local)
    n = 3000 + 500 * rank                      # ragged shards
    dim, m, k, ef, B = 48, 8, 10, 64, 333
    X = np.random.default_rng(100 + rank).random((n, dim), dtype=np.float32)
    Q = np.random.default_rng(7).random((B, dim), dtype=np.float32)
    g = capi.HnswIndex.build(X, m=m, ef_construction=60, level_seed=5 + rank)
    uid = [capi.ShardGroup.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    grp = capi.ShardGroup(uid[0], rank, world)
    off, total = grp.attach(g)
    rows = [None] * world
    dist.all_gather_object(rows, n)
    assert off == sum(rows[:rank]) and total == sum(rows)
    # reference lists: the single-shard call + the oracle on the exported graph
    li, ld, lc, lst = g.search(Q, k, ef)
    ni, rp, ci, ep = g.export_levels()
    oi, od, oc, _ = O.OracleHnsw.from_levels(X, O.HnswLevels(ni, rp, ci, ep)).search(Q, k, ef, n_threads=4)
    both = [None] * world
    dist.all_gather_object(both, (li, ld, oi, od))
    offsets = np.cumsum([0] + rows[:-1])
    exp_i, exp_d = numpy_merge(np.stack([b[0] for b in both]), np.stack([b[1] for b in both]), offsets, k)
    orc_i, orc_d = numpy_merge(np.stack([b[2] for b in both]), np.stack([b[3] for b in both]), offsets, k)
    seen = set()
    for exchange in (1, 0):
        for tile in (65536, 100):
            capi.set_option("shard.exchange", exchange)
            capi.set_option("shard.tile", tile)
            if tile == 65536:      # a fresh communicator re-negotiates the exchange
                grp.close()
                dist.broadcast_object_list(uid := [capi.ShardGroup.unique_id() if rank == 0 else None], src=0)
                grp = capi.ShardGroup(uid[0], rank, world)
                grp.attach(g)
            for root in (-1, 0, world - 1):
                ids, dd, cnt, st = grp.search(Q if root in (-1, rank) else None, k, ef, root=root, B=B)
                assert np.array_equal(ids, exp_i), (exchange, tile, root)
                assert np.array_equal(dd, exp_d)
                assert np.all(cnt == k) and st.dist_evals == lst.dist_evals
            seen.add(grp.info()["exchange"])
            # device form on a caller stream, twice back to back (double-buffered sets)
            qd = torch.from_numpy(Q).cuda()
            oi_d = torch.empty((B, k), dtype=torch.int64, device="cuda")
            od_d = torch.empty((B, k), dtype=torch.float32, device="cuda")
            for _ in range(3):
                grp.search_dev(qd.data_ptr(), B, k, ef, oi_d.data_ptr(), od_d.data_ptr(), None,
                               torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(oi_d.cpu().numpy().view(np.uint64), exp_i)
    # parity with the oracle under the same sharding
    rec = np.mean([len(set(a) & set(b)) / k for a, b in zip(exp_i, orc_i)])
    assert rec >= 0.999, rec
    assert np.allclose(exp_d, orc_d, rtol=1e-5, atol=1e-6)
    # radius and k > shard results
    ids, dd, cnt, _ = grp.search(Q, k, ef, radius=float(np.median(exp_d[:, 3])))
    assert np.array_equal(cnt, (exp_d <= np.float32(np.median(exp_d[:, 3]))).sum(1))
    dist.barrier()
    grp.close()
    if rank == 0:
        print(f"SHARDED_OK world={world} exchanges={sorted(seen)} recall_vs_oracle={rec:.4f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
