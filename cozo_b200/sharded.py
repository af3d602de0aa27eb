"""Row-sharded corpus: one process per GPU, one HNSW graph per shard.

The query batch is replicated; every rank searches its own shard, the per-shard top-k
lists are exchanged with ONE all-gather (NCCL over NVLink on the GPU box; gloo in the CPU
tests of the plumbing) and merged on the device by cozo_gpu_topk_merge_dev.  The result is
the k-NN over the union of the shards *as searched shard by shard* — parity is defined
against the oracle run with the same sharding, not against a single big graph.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class ShardedTopK:
    """Collective plumbing shared by bench.py and the tests: shard offsets + list exchange."""

    def __init__(self, local_rows: int, device: torch.device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device
        rows = torch.tensor([local_rows], dtype=torch.int64, device=device)
        if self.world > 1:
            allr = torch.empty(self.world, dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(allr, rows, group=group)
        else:
            allr = rows
        # global id of shard s' row 0 = rows of the shards before it (row-contiguous partition)
        self.offsets = torch.cumsum(allr, 0) - allr
        self.total_rows = int(allr.sum().item())

    def gather(self, local_dist: torch.Tensor, local_ids: torch.Tensor):
        """[B,k] per rank -> ([S,B,k] dist, [S,B,k] ids); exactly one all-gather per tensor."""
        B, k = local_dist.shape
        if self.world == 1:
            return local_dist.view(1, B, k), local_ids.view(1, B, k)
        all_d = torch.empty((self.world, B, k), dtype=local_dist.dtype, device=self.device)
        all_i = torch.empty((self.world, B, k), dtype=local_ids.dtype, device=self.device)
        dist.all_gather_into_tensor(all_d.view(-1), local_dist.contiguous().view(-1), group=self.group)
        dist.all_gather_into_tensor(all_i.view(-1), local_ids.contiguous().view(-1), group=self.group)
        return all_d, all_i


class _PeerBuffers:
    """[S,B,k] gather buffers in symmetric memory, double-buffered: every rank's search kernel
    stores its lists straight into all peers' buffers over NVLink (fused exchange)."""

    def __init__(self, world, B, k, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        gname = group.group_name if group is not None else dist.group.WORLD.group_name
        self.d, self.i, self.hd, self.hi = [], [], [], []
        for _ in range(2):
            td = symm_mem.empty((world, B, k), dtype=torch.float32, device=device)
            ti = symm_mem.empty((world, B, k), dtype=torch.int32, device=device)
            self.d.append(td)
            self.i.append(ti)
            self.hd.append(symm_mem.rendezvous(td, gname))
            self.hi.append(symm_mem.rendezvous(ti, gname))
        self.step = 0


class ShardedHnswSearch:
    """Device path (CUDA only, no CPU fallback).
    exchange="nccl":  local search -> ONE NCCL all-gather per list -> merge kernel
    exchange="fused": the search kernel's epilogue stores the lists into every peer's gather buffer
                      (symmetric memory over NVLink), one device-side barrier, merge kernel."""

    def __init__(self, index, local_rows: int, device: torch.device, group=None, exchange: str = "nccl"):
        from . import capi
        if device.type != "cuda":
            raise capi.CozoGpuError(capi.E_NODEV, "ShardedHnswSearch needs a CUDA device (no CPU fallback)")
        self.capi = capi
        self.index = index
        self.group = group
        self.plumb = ShardedTopK(local_rows, device, group)
        self.device = device
        self.exchange = exchange if self.plumb.world > 1 else "nccl"
        self._peer = {}

    def _search_fused(self, q_dev, k, ef, qstats):
        B = q_dev.shape[0]
        key = (B, k)
        if key not in self._peer:
            # symmetric memory needs every peer to be NVLink/P2P reachable; agree on the outcome
            try:
                pb = _PeerBuffers(self.plumb.world, B, k, self.device, self.group)
                ok = 1
            except Exception as e:  # pragma: no cover - depends on the box
                pb, ok = None, 0
                self.fallback_reason = repr(e)
            t = torch.tensor([ok], device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            if int(t.item()) == 0:
                self.exchange = "nccl"      # plain NCCL all-gather (the north-star exchange)
                return self.search(q_dev, k, ef, qstats)
            self._peer[key] = pb
        pb = self._peer[key]
        s = pb.step & 1
        pb.step += 1
        stream = torch.cuda.current_stream().cuda_stream
        self.index.search_scatter_dev(q_dev.data_ptr(), B, k, ef, pb.hi[s].buffer_ptrs, pb.hd[s].buffer_ptrs,
                                      self.plumb.rank, None if qstats is None else qstats.data_ptr(), stream)
        pb.hd[s].barrier(channel=s)      # all ranks' stores have landed (and step-2's merge is long done)
        out_i = torch.empty((B, k), dtype=torch.int64, device=self.device)
        out_d = torch.empty((B, k), dtype=torch.float32, device=self.device)
        self.capi.topk_merge_dev(pb.d[s].data_ptr(), pb.i[s].data_ptr(), self.plumb.world, B, k,
                                 self.plumb.offsets.data_ptr(), out_i.data_ptr(), out_d.data_ptr(), stream)
        return out_i, out_d

    def search(self, q_dev: torch.Tensor, k: int, ef: int, qstats: torch.Tensor | None = None):
        if self.exchange == "fused":
            return self._search_fused(q_dev, k, ef, qstats)
        B = q_dev.shape[0]
        stream = torch.cuda.current_stream().cuda_stream
        ids = torch.empty((B, k), dtype=torch.int32, device=self.device)
        dd = torch.empty((B, k), dtype=torch.float32, device=self.device)
        self.index.search_dev(q_dev.data_ptr(), B, k, ef, ids.data_ptr(), dd.data_ptr(), None,
                              None if qstats is None else qstats.data_ptr(), stream)
        all_d, all_i = self.plumb.gather(dd, ids)
        out_i = torch.empty((B, k), dtype=torch.int64, device=self.device)
        out_d = torch.empty((B, k), dtype=torch.float32, device=self.device)
        self.capi.topk_merge_dev(all_d.data_ptr(), all_i.data_ptr(), self.plumb.world, B, k,
                                 self.plumb.offsets.data_ptr(), out_i.data_ptr(), out_d.data_ptr(), stream)
        return out_i, out_d
