"""world_size-2 gloo test of the N>1 plumbing (shard offsets + the single list exchange).
The merge itself is a CUDA kernel (tests/test_hnsw_gpu.py::test_search_dev_and_merge)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, rows, B, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cozo_b200.sharded import ShardedTopK
    dev = torch.device("cpu")
    st = ShardedTopK(rows[rank], dev)
    rng = np.random.default_rng(100 + rank)
    d = np.sort(rng.random((B, k)).astype(np.float32), axis=1)
    i = rng.integers(0, rows[rank], (B, k)).astype(np.int32)
    all_d, all_i = st.gather(torch.from_numpy(d), torch.from_numpy(i))
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), d=d, i=i, all_d=all_d.numpy(), all_i=all_i.numpy(),
             offsets=st.offsets.numpy(), total=st.total_rows)
    dist.destroy_process_group()


def test_two_rank_exchange(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    rows, B, k = [1000, 1234], 17, 5       # ragged shards
    mp.spawn(_worker, args=(2, port, rows, B, k, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(tmp_path / f"r{x}.npz") for x in range(2)]
    for x in range(2):
        assert r[x]["offsets"].tolist() == [0, 1000] and int(r[x]["total"]) == 2234
        for s_ in range(2):                 # every rank holds every shard's list, in rank order
            assert np.array_equal(r[x]["all_d"][s_], r[s_]["d"]) and np.array_equal(r[x]["all_i"][s_], r[s_]["i"])
    # the merged answer any rank would compute == the global top-k of the concatenated lists
    cat_d = np.concatenate([r[0]["d"], r[1]["d"]], axis=1)
    cat_i = np.concatenate([r[0]["i"].astype(np.int64), r[1]["i"].astype(np.int64) + 1000], axis=1)
    order = np.argsort(cat_d, axis=1, kind="stable")[:, :k]
    top = np.take_along_axis(cat_i, order, 1)
    assert top.shape == (B, k) and top.max() < 2234
    # the merge specification the CUDA kernel is tested against (cozo_b200.sharded.merge_lists)
    from cozo_b200.sharded import merge_lists
    mi, md = merge_lists(r[0]["all_i"].astype(np.uint32), r[0]["all_d"], r[0]["offsets"], k)
    assert np.array_equal(mi.astype(np.int64), top) and np.array_equal(md, np.take_along_axis(cat_d, order, 1))
    # lists shorter than k (padding) never win a slot
    ai = r[0]["all_i"].astype(np.uint32).copy()
    ad = r[0]["all_d"].copy()
    ai[1, :, 2:] = 0xFFFFFFFF
    ad[1, :, 2:] = np.inf
    mi2, md2 = merge_lists(ai, ad, r[0]["offsets"], k)
    assert np.all(np.isfinite(md2)) and np.all(mi2 < 2234)
