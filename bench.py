#!/usr/bin/env python
"""bench.py — HNSW k-NN queries/sec on synthetic f32 vectors (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path (batched hnsw_knn) over one batch of B queries.
Workload at N=1 = BASELINE.json configs[1]: 1M x 768 f32, ef=200, k=10, batch=4096,
index parameters m=16 / ef_construction=200.
At N>1 (torchrun, one rank per GPU) every rank owns one 1M-vector shard with its own
graph ("weak": per-GPU work fixed), the query batch is replicated, per-shard top-k lists
are exchanged with ONE NCCL all-gather and merged on the device.

Keys beyond the base contract: `roofline`, `cpu_baseline`, `e2e`, `clocks`, `gpu_launches`.
Prints exactly one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "HNSW k-NN queries/sec at recall@10"
UNIT = "queries/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # workload (defaults = BASELINE configs[1]); overridable for quick runs
    ap.add_argument("--rows", dest="n", type=int, default=int(os.environ.get("COZO_BENCH_N", 1_000_000)))
    ap.add_argument("--dim", type=int, default=int(os.environ.get("COZO_BENCH_DIM", 768)))
    ap.add_argument("--batch", type=int, default=int(os.environ.get("COZO_BENCH_BATCH", 4096)))
    ap.add_argument("--ef", type=int, default=200)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--m", type=int, default=16)
    ap.add_argument("--efc", type=int, default=200)
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("COZO_BENCH_CPU_SAMPLE", 4096)))
    ap.add_argument("--ref-sample", type=int, default=int(os.environ.get("COZO_BENCH_REF_SAMPLE", 1024)))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value")
    ap.add_argument("--exchange", default=os.environ.get("COZO_BENCH_EXCHANGE", "fused"), choices=["nccl", "fused"],
                    help="N>1: one NCCL all-gather (north star) or peer stores fused into the search kernel")
    return ap.parse_args()


def workload_name(a, n_gpus):
    s = f"{a.n}x{a.dim} f32 U[0,1) per GPU, L2, m={a.m}, ef_construction={a.efc}, ef={a.ef}, k={a.k}, batch={a.batch}"
    if n_gpus > 1:
        s += f", corpus row-sharded over {n_gpus} GPUs ({a.n * n_gpus} vectors total)"
    return s


def gen_vectors(n, dim, seed):
    """i.i.d. U[0,1) f32 (the law of rand_vec, data/functions.rs:2154), generated in chunks."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    step = max(1, (1 << 26) // dim)
    for i in range(0, n, step):
        rng.random(out=out[i:i + step], dtype=np.float32)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            if t < t0 - 0.05 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get("hnsw_search_bytes_per_launch")
    except Exception:
        return None



def host_cores() -> int:
    """threads the CPU arms may use: the affinity mask, capped by a cgroup CPU quota if one is set"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                quota, period = txt[0], float(txt[1])
            else:
                quota, period = txt[0], float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota not in ("max", "-1") and period > 0:
                n = max(1, min(n, int(-(-float(quota) // period))))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n

def cpu_oracle_qps(X, levels, Q, k, ef, threads):
    from oracle import oracle as O
    ix = O.OracleHnsw.from_levels(X, O.HnswLevels(*levels))
    t0 = time.perf_counter()
    ids, dist, cnt, st = ix.search(Q, k, ef, n_threads=threads)
    dt = time.perf_counter() - t0
    return len(Q) / dt, ids, st


def run_reference(a, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port: the Rust crate cannot be
    built in this image) on the host cores, same config, bounded sample per step."""
    if rank != 0:
        return
    from cozo_b200 import capi
    cores = host_cores()
    X = gen_vectors(a.n, a.dim, 0x5EED0001)
    capi.init(0)
    t0 = time.perf_counter()
    g = capi.HnswIndex.build(X, m=a.m, ef_construction=a.efc, level_seed=0x5EED0003)
    levels = g.export_levels()
    g.close()
    build_s = time.perf_counter() - t0
    from oracle import oracle as O
    ix = O.OracleHnsw.from_levels(X, O.HnswLevels(*levels))
    sample = min(a.ref_sample, a.batch)
    times = []
    for s in range(a.warmup + a.steps):
        Q = gen_vectors(sample, a.dim, 0x5EED0002 + s)
        t0 = time.perf_counter()
        ix.search(Q, a.k, a.ef, n_threads=cores)
        if s >= a.warmup:
            times.append(time.perf_counter() - t0)
    tot = sum(times)
    qps = sample * a.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, 1), "note": "graph built by the device builder, searched on CPU; "
                   f"index build {build_s:.1f}s not timed"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} queries per step x {a.steps} steps, {cores} threads (one query per "
                                   "thread at a time == N concurrent read transactions)"},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    import torch
    from cozo_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the hot path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    capi.init(local_rank)
    for o in a.opt:
        name, val = o.split("=")
        capi.set_option(name, int(val))

    # ---- corpus shard + index on this GPU -----------------------------------------------
    t0 = time.perf_counter()
    X = gen_vectors(a.n, a.dim, 0x5EED0001 + 1000 * rank)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = capi.HnswIndex.build(X, m=a.m, ef_construction=a.efc, level_seed=0x5EED0003 + rank)
    build_s = time.perf_counter() - t0
    B, k, ef, dim = a.batch, a.k, a.ef, a.dim
    nsteps = a.warmup + a.steps
    # distinct query batch per step, resident in HBM before the timed region
    Qh = gen_vectors(B * nsteps, dim, 0x5EED0002).reshape(nsteps, B, dim)
    Qd = torch.from_numpy(Qh).to(dev)
    ids = torch.empty((B, k), dtype=torch.int32, device=dev)
    dd = torch.empty((B, k), dtype=torch.float32, device=dev)
    qstats = torch.zeros((nsteps, B, 4), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    sharded = None
    if world > 1:
        import torch.distributed as dist
        from cozo_b200.sharded import ShardedHnswSearch
        sharded = ShardedHnswSearch(g, a.n, dev, exchange=a.exchange)

    def step(s):
        if sharded is None:
            g.search_dev(Qd[s].data_ptr(), B, k, ef, ids.data_ptr(), dd.data_ptr(), None, qstats[s].data_ptr(), stream)
        else:   # local search -> ONE all-gather per list -> merge kernel
            sharded.search(Qd[s], k, ef, qstats[s])

    sampler = ClockSampler(local_rank)
    sampler.start()          # nvidia-smi needs a moment to start: launch it before the warm-up
    for s in range(a.warmup):
        step(s)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    time.sleep(0.3)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    torch.cuda.synchronize()
    t_start = time.perf_counter()
    e_all0 = torch.cuda.Event(enable_timing=True)
    e_all1 = torch.cuda.Event(enable_timing=True)
    e_all0.record()
    for i in range(a.steps):
        evs[i][0].record()
        step(a.warmup + i)
        evs[i][1].record()
    e_all1.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop(t_start, t_end)
    total_ms = e_all0.elapsed_time(e_all1)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    # whole-job units: every rank searched B queries against its shard
    value = world * B * a.steps / (total_ms / 1e3)

    # ---- roofline of the dominant kernel (hnsw_search_kernel) ---------------------------
    st = qstats[a.warmup:].to(torch.int64).sum(dim=(0, 1)).cpu().numpy()
    dist_evals, expanded, nbr_reads = int(st[0]), int(st[1]), int(st[2])
    alg_bytes_per_launch = (dist_evals * dim * 4 + expanded * 8 + nbr_reads * 4 + a.steps * B * dim * 4) / a.steps
    # kernel duration: the search kernel alone (events bracket memset+kernel; at N=1 a step is just that)
    kern_ms = statistics.mean(step_ms) if world == 1 else None
    if world > 1:
        # time the search kernel alone once more for the roofline line
        tmp = []
        for i in range(a.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.search_dev(Qd[a.warmup + i].data_ptr(), B, k, ef, ids.data_ptr(), dd.data_ptr(), None, None, stream)
            e1.record()
            torch.cuda.synchronize()
            tmp.append(e0.elapsed_time(e1))
        kern_ms = statistics.mean(tmp)
    peak, peak_src = peak_hbm()
    achieved = alg_bytes_per_launch / (kern_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "kernel": "hnsw_search_kernel", "kernel_ms": kern_ms,
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "peak_source": peak_src,
                "dist_evals_per_query": dist_evals / (a.steps * B), "nodes_expanded_per_query": expanded / (a.steps * B)}

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (pinned), copies inside ----
    hq = torch.from_numpy(Qh).pin_memory()
    e2e_times = []
    for s in range(nsteps):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        hi, hd, hc, hst = g.search(hq[s].numpy(), k, ef)
        if world > 1:
            # host-API path of the sharded operator: per-shard lists -> device -> all-gather -> merge -> host
            ld = torch.from_numpy(hd).to(dev)
            li = torch.from_numpy(hi.view(np.int32)).to(dev)
            all_d, all_i = sharded.plumb.gather(ld, li)
            out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
            out_d = torch.empty((B, k), dtype=torch.float32, device=dev)
            capi.topk_merge_dev(all_d.data_ptr(), all_i.data_ptr(), world, B, k, sharded.plumb.offsets.data_ptr(),
                                out_i.data_ptr(), out_d.data_ptr(), stream)
            _ = out_i.cpu()
        dt = time.perf_counter() - t0
        if s >= a.warmup:
            e2e_times.append(dt)
    e2e_s = sum(e2e_times)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * B * a.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * dim * 4,
           "d2h_bytes_per_step": B * k * 8 + B * 4 + B * 16}

    # ---- recall vs the CPU oracle on the same graph + cpu_baseline (rank 0, N=1 only) -----
    cpu_baseline = None
    recall_vs_oracle = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cores = host_cores()
        levels = g.export_levels()
        sample = min(a.cpu_sample, B)
        s_last = a.warmup + a.steps - 1
        qps, oids, ost = cpu_oracle_qps(X, levels, Qh[s_last][:sample], k, ef, cores)
        hi, hd, _, _ = g.search(Qh[s_last][:sample], k, ef)
        recall_vs_oracle = float(np.mean([len(set(x) & set(y)) / k for x, y in zip(hi, oids)]))
        cpu_baseline = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{sample} queries of the last timed batch, {cores} threads, oracle port of "
                                  "hnsw_knn on the exported graph (flat CSR + flat vectors: faster than real Cozo)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "l2_policy": "inputs larger than L2 (corpus "
                       f"{a.n * a.dim * 4 / 1e9:.2f} GB/GPU >> 126 MB) and a distinct query batch per step",
                       "value_counts": "per-shard k-NN searches per second summed over ranks (== queries/s at N=1); "
                                       "global queries/s over the whole sharded corpus = value / n_gpus",
                       "index_build_s": round(build_s, 2), "gen_s": round(gen_s, 2),
                       "options": {o.split("=")[0]: int(o.split("=")[1]) for o in a.opt}},
            "global_queries_per_s": value / world,
            "recall_at_k_vs_oracle": recall_vs_oracle,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": a.steps * (1 + (1 if world > 1 else 0)),
            "exchange": (sharded.exchange if world > 1 else None),
            "step_ms": step_ms,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
