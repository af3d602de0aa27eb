"""ctypes front-end of the CPU oracle (oracle/cozo_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by cozo_b200.
PARITY UNPINNED (see header of cozo_oracle.cpp and DESIGN.md).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcozo_oracle.so")

L2, COSINE, IP = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cozo_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u32, u64, i64, f32p, f64 = C.c_uint32, C.c_uint64, C.c_int64, C.POINTER(C.c_float), C.c_double
        vp = C.c_void_p
        L.orc_vec_dist.restype = f64
        L.orc_vec_dist.argtypes = [C.c_int, vp, vp, u32]
        L.orc_hnsw_new.restype = vp
        L.orc_hnsw_new.argtypes = [u32, u32, C.c_int, u32, u32, C.c_int, C.c_int, u64]
        L.orc_hnsw_insert.restype = C.c_int
        L.orc_hnsw_insert.argtypes = [vp, u32, vp, i64]
        L.orc_hnsw_remove.restype = C.c_int
        L.orc_hnsw_remove.argtypes = [vp, u32]
        L.orc_hnsw_build_dist_evals.restype = u64
        L.orc_hnsw_build_dist_evals.argtypes = [vp]
        L.orc_hnsw_relation_rows.restype = u64
        L.orc_hnsw_relation_rows.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_hnsw_from_csr.restype = vp
        L.orc_hnsw_from_csr.argtypes = [u32, u32, C.c_int, vp, C.c_int, u32, vp, vp, vp, vp, u32]
        L.orc_hnsw_free.argtypes = [vp]
        L.orc_hnsw_n_levels.restype = u32
        L.orc_hnsw_n_levels.argtypes = [vp]
        L.orc_hnsw_entry.restype = u32
        L.orc_hnsw_entry.argtypes = [vp]
        L.orc_hnsw_level_size.argtypes = [vp, u32, vp, vp]
        L.orc_hnsw_export_level.argtypes = [vp, u32, vp, vp, vp]
        L.orc_hnsw_search_batch.restype = C.c_int
        L.orc_hnsw_search_batch.argtypes = [vp, vp, u32, u32, u32, f64, vp, vp, vp, vp, u32]
        L.orc_hnsw_search_batch_f64.restype = C.c_int
        L.orc_hnsw_search_batch_f64.argtypes = [vp, vp, vp, u32, u32, u32, f64, vp, vp, vp, vp, u32]
        L.orc_bruteforce_knn.restype = C.c_int
        L.orc_bruteforce_knn.argtypes = [vp, u32, u32, C.c_int, vp, u32, u32, vp, vp, u32]
        L.orc_graph_new.restype = vp
        L.orc_graph_new.argtypes = [u32, u64, vp, vp, vp]
        L.orc_graph_free.argtypes = [vp]
        L.orc_graph_export.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_pagerank.restype = u32
        L.orc_pagerank.argtypes = [vp, C.c_float, f64, u32, C.c_int, vp, vp, u32]
        L.orc_sssp.restype = C.c_int
        L.orc_sssp.argtypes = [vp, vp, u32, vp, u32, vp, vp, u32]
        L.orc_sssp_keep_ties.restype = i64
        L.orc_sssp_keep_ties.argtypes = [vp, u32, vp, u32, vp, vp, vp]
        L.orc_closeness.restype = C.c_int
        L.orc_closeness.argtypes = [vp, vp, u32]
        L.orc_betweenness.restype = C.c_int
        L.orc_betweenness.argtypes = [vp, vp, u32, u64]
        L.orc_yen.restype = C.c_int
        L.orc_yen.argtypes = [vp, u32, u32, u32, vp, vp, vp, vp]
        L.orc_clustering.restype = C.c_int
        L.orc_clustering.argtypes = [vp, vp, vp, vp, u32]
        L.orc_random_level.restype = i64
        L.orc_random_level.argtypes = [vp, u32]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def vec_dist(metric: int, a: np.ndarray, b: np.ndarray) -> float:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return lib().orc_vec_dist(metric, _p(a), _p(b), a.size)


class HnswLevels:
    """Flat per-level CSR (level 0 = bottom = reference layer 0, level L = layer -L)."""

    def __init__(self, node_ids, row_ptr, col_idx, entry):
        self.node_ids = node_ids  # list; node_ids[0] is None (identity)
        self.row_ptr = row_ptr
        self.col_idx = col_idx
        self.entry = entry

    @property
    def n_levels(self):
        return len(self.row_ptr)


class OracleHnsw:
    def __init__(self, handle, dim, keep=None):
        self._h = handle
        self.dim = dim
        self._keep = keep

    @classmethod
    def new(cls, n_max, dim, metric=L2, m=16, ef_construction=200, extend_candidates=False,
            keep_pruned_connections=False, level_seed=0x5EED0003):
        h = lib().orc_hnsw_new(n_max, dim, metric, m, ef_construction, int(extend_candidates),
                               int(keep_pruned_connections), level_seed)
        return cls(h, dim)

    @classmethod
    def from_levels(cls, vectors: np.ndarray, levels: HnswLevels, metric=L2, copy_vectors=False):
        vectors = np.ascontiguousarray(vectors, np.float32)
        n, dim = vectors.shape
        nl = levels.n_levels
        lvl_nodes = np.array([len(rp) - 1 for rp in levels.row_ptr], np.uint32)
        keep = [vectors, lvl_nodes]
        nid = (C.c_void_p * nl)()
        rpp = (C.c_void_p * nl)()
        cip = (C.c_void_p * nl)()
        for L in range(nl):
            rp = np.ascontiguousarray(levels.row_ptr[L], np.uint64)
            ci = np.ascontiguousarray(levels.col_idx[L], np.uint32)
            keep += [rp, ci]
            if L > 0:
                ni = np.ascontiguousarray(levels.node_ids[L], np.uint32)
                keep.append(ni)
                nid[L] = ni.ctypes.data
            rpp[L] = rp.ctypes.data
            cip[L] = ci.ctypes.data if ci.size else None
        entry = 0xFFFFFFFF if levels.entry is None else int(levels.entry)
        h = lib().orc_hnsw_from_csr(n, dim, metric, _p(vectors), int(copy_vectors), nl, _p(lvl_nodes), nid, rpp, cip,
                                    entry)
        return cls(h, dim, keep if not copy_vectors else None)

    def insert(self, id_: int, v: np.ndarray, forced_level: int = 1):
        v = np.ascontiguousarray(v, np.float32)
        assert v.size == self.dim
        rc = lib().orc_hnsw_insert(self._h, id_, _p(v), forced_level)
        if rc != 0:
            raise RuntimeError(f"oracle insert failed rc={rc}")

    def insert_all(self, vectors: np.ndarray):
        vectors = np.ascontiguousarray(vectors, np.float32)
        for i in range(vectors.shape[0]):
            self.insert(i, vectors[i])

    def remove(self, id_: int):
        lib().orc_hnsw_remove(self._h, id_)

    def build_dist_evals(self) -> int:
        return lib().orc_hnsw_build_dist_evals(self._h)

    def relation_rows(self):
        """Raw index-relation rows in key order: (layer i64, fr u32, to u32, dist f64, ignore_link u8);
        self-loop rows have fr == to and carry the degree in `dist`."""
        n = lib().orc_hnsw_relation_rows(self._h, None, None, None, None, None)
        layer = np.zeros(n, np.int64)
        fr = np.zeros(n, np.uint32)
        to = np.zeros(n, np.uint32)
        dist = np.zeros(n, np.float64)
        ign = np.zeros(n, np.uint8)
        lib().orc_hnsw_relation_rows(self._h, _p(layer), _p(fr), _p(to), _p(dist), _p(ign))
        return layer, fr, to, dist, ign

    def levels(self) -> HnswLevels:
        L = lib()
        nl = L.orc_hnsw_n_levels(self._h)
        node_ids, row_ptr, col_idx = [], [], []
        for lv in range(nl):
            nr, ne = C.c_uint32(), C.c_uint64()
            L.orc_hnsw_level_size(self._h, lv, C.byref(nr), C.byref(ne))
            ni = np.zeros(nr.value, np.uint32)
            rp = np.zeros(nr.value + 1, np.uint64)
            ci = np.zeros(ne.value, np.uint32)
            L.orc_hnsw_export_level(self._h, lv, _p(ni), _p(rp), _p(ci))
            node_ids.append(None if lv == 0 else ni)
            row_ptr.append(rp)
            col_idx.append(ci)
        entry = L.orc_hnsw_entry(self._h)
        return HnswLevels(node_ids, row_ptr, col_idx, None if entry == 0xFFFFFFFF else entry)

    def search(self, queries: np.ndarray, k: int, ef: int, radius: float | None = None, n_threads: int = 1):
        """Returns (ids[B,k] u32 padded 0xFFFFFFFF, dist[B,k] f64 padded inf, count[B], stats[B,3])."""
        queries = np.ascontiguousarray(queries, np.float32).reshape(-1, self.dim)
        B = queries.shape[0]
        ids = np.empty((B, k), np.uint32)
        dist = np.empty((B, k), np.float64)
        cnt = np.zeros(B, np.uint32)
        stats = np.zeros((B, 3), np.uint64)
        rc = lib().orc_hnsw_search_batch(self._h, _p(queries), B, k, ef, -1.0 if radius is None else float(radius),
                                         _p(ids), _p(dist), _p(cnt), _p(stats), n_threads)
        assert rc == 0
        return ids, dist, cnt, stats

    def search_f64(self, vectors64: np.ndarray, queries: np.ndarray, k: int, ef: int, radius: float | None = None,
                   n_threads: int = 1):
        """the same graph searched as an F64 index (manifest.dtype == F64): f64 payloads, f64 queries, f64 arithmetic"""
        vectors64 = np.ascontiguousarray(vectors64, np.float64).reshape(-1, self.dim)
        queries = np.ascontiguousarray(queries, np.float64).reshape(-1, self.dim)
        B = queries.shape[0]
        ids = np.empty((B, k), np.uint32)
        dist = np.empty((B, k), np.float64)
        cnt = np.zeros(B, np.uint32)
        stats = np.zeros((B, 3), np.uint64)
        rc = lib().orc_hnsw_search_batch_f64(self._h, _p(vectors64), _p(queries), B, k, ef,
                                             -1.0 if radius is None else float(radius), _p(ids), _p(dist), _p(cnt),
                                             _p(stats), n_threads)
        assert rc == 0
        return ids, dist, cnt, stats

    def close(self):
        if self._h:
            lib().orc_hnsw_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bruteforce_knn(vectors, queries, k, metric=L2, n_threads=1):
    vectors = np.ascontiguousarray(vectors, np.float32)
    queries = np.ascontiguousarray(queries, np.float32).reshape(-1, vectors.shape[1])
    B = queries.shape[0]
    ids = np.empty((B, k), np.uint32)
    dist = np.empty((B, k), np.float64)
    lib().orc_bruteforce_knn(_p(vectors), vectors.shape[0], vectors.shape[1], metric, _p(queries), B, k, _p(ids),
                             _p(dist), n_threads)
    return ids, dist


class OracleGraph:
    def __init__(self, n: int, src, dst, w=None):
        src = np.ascontiguousarray(src, np.uint32)
        dst = np.ascontiguousarray(dst, np.uint32)
        w = None if w is None else np.ascontiguousarray(w, np.float32)
        self.n, self.m = int(n), int(src.size)
        self.weighted = w is not None
        self._h = lib().orc_graph_new(n, src.size, _p(src), _p(dst), _p(w))

    def export(self):
        out_ptr = np.zeros(self.n + 1, np.uint64)
        in_ptr = np.zeros(self.n + 1, np.uint64)
        out_idx = np.zeros(self.m, np.uint32)
        in_idx = np.zeros(self.m, np.uint32)
        out_w = np.zeros(self.m, np.float32) if self.weighted else None
        lib().orc_graph_export(self._h, _p(out_ptr), _p(out_idx), _p(out_w), _p(in_ptr), _p(in_idx))
        return out_ptr, out_idx, out_w, in_ptr, in_idx

    def pagerank(self, damping=0.85, tol=1e-4, max_iter=10, variant="jacobi", n_threads=1):
        """theta/epsilon are cast to f32 first, epsilon then widened (pagerank.rs:37-38,49)."""
        scores = np.zeros(self.n, np.float32)
        err = C.c_double()
        tol64 = float(np.float32(tol))
        it = lib().orc_pagerank(self._h, float(np.float32(damping)), tol64, max_iter,
                                {"jacobi": 0, "gs": 1}[variant], _p(scores), C.byref(err), n_threads)
        return scores, it, err.value

    def sssp(self, sources, goals=None, n_threads=1, want_back=True):
        sources = np.ascontiguousarray(sources, np.uint32)
        goals_a = None if goals is None else np.ascontiguousarray(goals, np.uint32)
        dist = np.zeros((sources.size, self.n), np.float32)
        back = np.zeros((sources.size, self.n), np.uint32) if want_back else None
        lib().orc_sssp(self._h, _p(sources), sources.size, _p(goals_a), 0 if goals is None else goals_a.size,
                       _p(dist), _p(back), n_threads)
        return dist, back

    def sssp_keep_ties(self, source, goals=None):
        goals_a = None if goals is None else np.ascontiguousarray(goals, np.uint32)
        ng = 0 if goals is None else goals_a.size
        dist = np.zeros(self.n, np.float32)
        tot = lib().orc_sssp_keep_ties(self._h, source, _p(goals_a), ng, _p(dist), None, None)
        if tot < 0:
            raise RuntimeError("keep_ties did not terminate (zero-weight cycle)")
        bp = np.zeros(self.n + 1, np.uint64)
        bi = np.zeros(max(tot, 1), np.uint32)
        lib().orc_sssp_keep_ties(self._h, source, _p(goals_a), ng, _p(dist), _p(bp), _p(bi))
        return dist, bp, bi[:tot]

    def closeness(self, n_threads=1):
        out = np.zeros(self.n, np.float32)
        lib().orc_closeness(self._h, _p(out), n_threads)
        return out

    def betweenness(self, n_threads=1, path_cap=1 << 20):
        out = np.zeros(self.n, np.float32)
        rc = lib().orc_betweenness(self._h, _p(out), n_threads, path_cap)
        if rc != 0:
            raise RuntimeError("betweenness failed (zero-weight cycle)")
        return out

    def yen(self, start, goal, k):
        """k_shortest_path_yen (yen.rs:120-211) -> list of (cost f32, path list)"""
        tot = C.c_uint64()
        cnt = lib().orc_yen(self._h, start, goal, k, None, None, None, C.byref(tot))
        cost = np.zeros(max(cnt, 1), np.float32)
        ptr = np.zeros(cnt + 1, np.uint64)
        buf = np.zeros(max(tot.value, 1), np.uint32)
        lib().orc_yen(self._h, start, goal, k, _p(cost), _p(ptr), _p(buf), C.byref(tot))
        return [(float(cost[i]), buf[int(ptr[i]):int(ptr[i + 1])].tolist()) for i in range(cnt)]

    def clustering(self, n_threads=1):
        """(cc f64, n_triangles u64, degree u64) per node; the graph must hold the mirrored edge stream"""
        cc = np.zeros(self.n, np.float64)
        nt = np.zeros(self.n, np.uint64)
        deg = np.zeros(self.n, np.uint64)
        lib().orc_clustering(self._h, _p(cc), _p(nt), _p(deg), n_threads)
        return cc, nt, deg

    def close(self):
        if self._h:
            lib().orc_graph_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def random_levels(n: int, m: int, seed: int) -> np.ndarray:
    st = C.c_uint64(seed)
    out = np.empty(n, np.int64)
    L = lib()
    for i in range(n):
        out[i] = L.orc_random_level(C.byref(st), m)
    return out
