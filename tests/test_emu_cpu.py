"""Device code under a CPU SIMT emulator (tests/emu/cuda_emu.hpp; test infrastructure, like oracle/).
(tests/test_emu_lib_cpu.py goes one step further and runs the whole library, host code included, the same way.)

The kernels that have not run on a GPU yet are compiled from the product's own headers
(cozo_b200/csrc/pagerank_pb.cuh, graph_kernels.cuh) into host programs: one OS thread per CUDA thread, real barriers,
real atomics, a watchdog for hangs.  The harnesses check results against independent host references.
This is how the lost-update race of the compacted-frontier SSSP kernel was found (and fixed) without a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


# COZO_EMU_SANITIZE=1: the same harnesses under AddressSanitizer + UBSan (the buffers a kernel touches are host
# allocations of the size the product's host code gives them, so an out-of-bounds access in device code is reported)
_SANITIZE = ["-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"] if os.environ.get("COZO_EMU_SANITIZE") == "1" else []


def _build(name, tmp):
    exe = os.path.join(tmp, name)
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", *_SANITIZE, "-o", exe, os.path.join(EMU, name + ".cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.fixture(scope="module")
def exes(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("emu"))
    return {n: _build(n, tmp) for n in ("pagerank_pb_emu", "graph_emu")}


_ENV = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")


def _run(exe, *args, timeout=1800):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=timeout, env=_ENV)
    assert r.returncode == 0 and "EMU_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
    return r.stdout


# n, m, NH, GS, WIN, CHUNK, iterations, seed [, star]
@pytest.mark.parametrize("args", [
    (500, 4000, 64, 256, 512, 1024, 3, 1),            # several groups, a few bins
    (400, 3000, 0, 64, 64, 1024, 2, 2),               # no hub table, tiny tiles and windows
    (300, 1500, 8, 32, 48, 1024, 3, 6, "star"),       # one row straddling dozens of windows
    (400, 3000, 16384, 32768, 24576, 262144, 2, 5),   # the default geometry: every source in the hub table
    (200, 0, 64, 256, 512, 1024, 2, 7),               # no edges at all
])
def test_pagerank_blocking_device_code(exes, args):
    """staging kernels + gather / accumulate / straddle / final passes of pagerank.cu (mode 1) vs an f64 Jacobi iteration"""
    _run(exes["pagerank_pb_emu"], *args)


@pytest.mark.parametrize("args", [(40, 160, 9), (36, 200, 21)])
def test_graph_device_code(exes, args):
    """SSSP in all frontier forms (with and without forbidden sets), closeness, betweenness (+ ordered reduction, run
    twice), zero-weight-cycle flag, clustering vs host Dijkstra / Brandes / brute force"""
    _run(exes["graph_emu"], *args)


def _emu_build(exe, tmp, n, dim, m, efc, keep, extend, seed, max_batch=1):
    """run the harness; returns (X, level, n_levels, entry, [ {node: set(neighbours)} per level ])"""
    import struct

    import numpy as np
    rng = np.random.default_rng(seed)
    X = rng.random((n, dim), dtype=np.float32)
    level = np.minimum(np.floor(-np.log(rng.random(n)) / np.log(m)), 15).astype(np.uint8)     # the level law, hnsw.rs:46-52
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    open(fin, "wb").write(X.tobytes() + level.tobytes())
    r = subprocess.run([exe, fin, fout, str(n), str(dim), str(m), str(efc), str(int(keep)), str(int(extend)), str(max_batch)],
                       capture_output=True, text=True, timeout=1800, env=_ENV)
    assert r.returncode == 0 and "EMU_OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]
    buf = open(fout, "rb").read()
    nl, entry = struct.unpack_from("II", buf, 0)
    off, dev = 8, []
    for _ in range(nl):
        rows, = struct.unpack_from("I", buf, off)
        off += 4
        d = {}
        for _ in range(rows):
            i, deg = struct.unpack_from("II", buf, off)
            off += 8
            nb = struct.unpack_from(f"{deg}I", buf, off)
            assert len(set(nb)) == deg, "duplicate neighbour in a row"
            d[i] = set(nb)
            off += 4 * deg
        dev.append(d)
    return X, level, nl, entry, dev


def _oracle_build(X, level, m, efc, keep, extend):
    from oracle import oracle as O
    n, dim = X.shape
    ix = O.OracleHnsw.new(n, dim, m=m, ef_construction=efc, keep_pruned_connections=keep, extend_candidates=extend)
    for i in range(n):
        ix.insert(i, X[i], forced_level=-int(level[i]))
    return ix


def _build_graphs(exe, tmp, n, dim, m, efc, keep, extend, seed):
    import numpy as np
    X, level, nl, entry, dev = _emu_build(exe, tmp, n, dim, m, efc, keep, extend, seed)
    lv = _oracle_build(X, level, m, efc, keep, extend).levels()
    assert lv.n_levels == nl and lv.entry == entry
    same = total = 0
    for L in range(nl):
        nodes = np.arange(n) if L == 0 else lv.node_ids[L]
        assert sorted(dev[L]) == [int(x) for x in nodes]
        for r_, node in enumerate(nodes):
            ref = {int(x) for x in lv.col_idx[L][int(lv.row_ptr[L][r_]):int(lv.row_ptr[L][r_ + 1])]}
            same += ref == dev[L][int(node)]
            total += 1
    return same, total


_LONG = True      # every case: seconds since the emulator schedules lanes as fibers (was minutes with a thread per lane)


@pytest.mark.parametrize("n,efc,keep,extend", [
    (40, 10, False, False),
    pytest.param(40, 10, True, True, marks=pytest.mark.skipif(not _LONG, reason="minutes under the emulator; COZO_EMU_LONG=1 "
                                                                               "(tests/test_emu_lib_cpu.py covers the same through the C ABI)")),
    pytest.param(120, 16, False, False, marks=pytest.mark.skipif(not _LONG, reason="minutes under the emulator; COZO_EMU_LONG=1")),
    pytest.param(100, 16, True, False, marks=pytest.mark.skipif(not _LONG, reason="minutes under the emulator; COZO_EMU_LONG=1")),
    pytest.param(90, 12, False, True, marks=pytest.mark.skipif(not _LONG, reason="minutes under the emulator; COZO_EMU_LONG=1")),
    pytest.param(90, 12, True, True, marks=pytest.mark.skipif(not _LONG, reason="minutes under the emulator; COZO_EMU_LONG=1")),
])
def test_builder_fidelity_device_code(tmp_path, n, efc, keep, extend):
    """the index builder's kernels (batch search with the TMA ring, heuristic selection with / without candidate
    extension, link / shrink) in fidelity mode == the oracle's faithful hnsw_put_vector, edge for edge, given the same
    levels.  Measured with COZO_EMU_LONG=1: 156/156, 127/127 (keep_pruned), 117/118 (extend), 114/114 (both) rows."""
    exe = _build("hnsw_build_emu", str(tmp_path))
    same, total = _build_graphs(exe, str(tmp_path), n, 16, 4, efc, keep, extend, seed=n + efc)
    assert same >= total - max(1, total // 100), (same, total)    # a strict f32 comparison may flip once in a while


def test_builder_default_batched_device_code(tmp_path):
    """the DEFAULT builder mode (batches of linked/16 nodes that do not see each other, K1 with 4 warps per CTA and one TMA
    ring per warp, in-edges sorted by (layer, target), one warp per target row): row invariants, and the graph is as good
    as the oracle's sequential build of the same vectors and levels (searched by the same oracle code)"""
    import numpy as np

    from oracle import oracle as O
    n, dim, m, efc = 160, 16, 4, 16
    exe = _build("hnsw_build_emu", str(tmp_path))
    X, level, nl, entry, dev = _emu_build(exe, str(tmp_path), n, dim, m, efc, False, False, seed=n, max_batch=8192)
    node_ids, row_ptr, col_idx = [], [], []
    for L in range(nl):
        ids, rp, ci = [], [0], []
        for i in sorted(dev[L]):
            nb = dev[L][i]
            assert level[i] >= L and len(nb) <= (2 * m if L == 0 else m)
            assert all(level[x] >= L for x in nb), "neighbour that does not exist on this layer"
            ids.append(i)
            ci += sorted(nb)
            rp.append(len(ci))
        node_ids.append(np.array(ids, np.uint32))
        row_ptr.append(np.array(rp, np.uint32))
        col_idx.append(np.array(ci, np.uint32))
    assert level[entry] == nl - 1
    dev_ix = O.OracleHnsw.from_levels(X, O.HnswLevels(node_ids, row_ptr, col_idx, entry))
    seq_ix = _oracle_build(X, level, m, efc, False, False)
    Q = np.random.default_rng(5).random((100, dim), dtype=np.float32)
    bi, _ = O.bruteforce_knn(X, Q, 5, n_threads=4)
    rec = {}
    for name, ix in (("device", dev_ix), ("sequential", seq_ix)):
        ids, _, _, _ = ix.search(Q, 5, 20, n_threads=4)
        rec[name] = float(np.mean([len(set(a) & set(b)) / 5 for a, b in zip(ids, bi)]))
    assert rec["device"] >= rec["sequential"] - 0.03, rec       # measured: 0.962 vs 0.970
