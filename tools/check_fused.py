"""2+ GPU check (torchrun): fused peer-store exchange == NCCL all-gather exchange, bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import gen_vectors  # noqa: E402
from cozo_b200 import capi  # noqa: E402
from cozo_b200.sharded import ShardedHnswSearch  # noqa: E402

rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
capi.init(lr)
n, dim, B, k = 50000, 128, 1000, 10
X = gen_vectors(n, dim, 1 + rank)
g = capi.HnswIndex.build(X, m=16, ef_construction=100, level_seed=rank)
Q = torch.from_numpy(gen_vectors(B * 3, dim, 99).reshape(3, B, dim)).to(dev)
a = ShardedHnswSearch(g, n, dev, exchange="nccl")
b = ShardedHnswSearch(g, n, dev, exchange="fused")
ok = True
for s in range(3):
    ia, da = a.search(Q[s], k, 64)
    ib, db = b.search(Q[s], k, 64)
    torch.cuda.synchronize()
    ok &= bool(torch.equal(ia, ib) and torch.equal(da, db))
    # global ids must span both shards
    assert int(ia.max()) >= n or dist.get_world_size() == 1
t = torch.tensor([int(ok)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("FUSED_EQUALS_NCCL", bool(t.item()))
dist.destroy_process_group()
