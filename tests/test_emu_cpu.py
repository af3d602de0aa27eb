"""Device code under a CPU SIMT emulator (tests/emu/cuda_emu.hpp; test infrastructure, like oracle/).

The kernels that have not run on a GPU yet are compiled from the product's own headers
(cozo_b200/csrc/pagerank_pb.cuh, graph_kernels.cuh) into host programs: one OS thread per CUDA thread, real barriers,
real atomics, a watchdog for hangs.  The harnesses check results against independent host references.
This is how the lost-update race of the compacted-frontier SSSP kernel was found (and fixed) without a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


def _build(name, tmp):
    exe = os.path.join(tmp, name)
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-o", exe, os.path.join(EMU, name + ".cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.fixture(scope="module")
def exes(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("emu"))
    return {n: _build(n, tmp) for n in ("pagerank_pb_emu", "graph_emu")}


def _run(exe, *args, timeout=600):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=timeout)
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
