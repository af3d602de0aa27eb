"""The numpy model of the propagation-blocking PageRank layout (tools/pagerank_pb_model.py, which
cozo_b200/csrc/pagerank.cu follows pass by pass) against the oracle, over blocking geometries that exercise every
path: several groups, many bins, rows straddling windows, empty rows, no hub table, everything in the hub table."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import rmat_edges
from tools import pagerank_pb_model as M


@pytest.mark.parametrize("NH,GS,WIN", [(16384, 32768, 24576), (64, 256, 512), (0, 64, 64), (8, 32, 48), (1024, 4096, 3001)])
def test_model_matches_oracle(NH, GS, WIN):
    n, src, dst = rmat_edges(9, 8, 77)
    os_, oit, oerr = O.OracleGraph(n, src, dst).pagerank(0.85, 0.0, 4, variant="jacobi")
    ms, mit, merr = M.pagerank(n, src, dst, 0.85, 0.0, 4, NH, GS, WIN)
    assert mit == oit and np.max(np.abs(ms - os_) / os_) <= 1e-5
    assert abs(merr - oerr) <= 0.02 * oerr + 1e-6


def test_model_star_and_path():
    n = 600                                    # a row of 599 in-edges straddles many 48-entry windows
    src = np.concatenate([np.arange(1, n), np.zeros(n - 1)]).astype(np.uint32)
    dst = np.concatenate([np.zeros(n - 1), np.arange(1, n)]).astype(np.uint32)
    os_, _, _ = O.OracleGraph(n, src, dst).pagerank(0.85, 0.0, 3)
    for geo in ((8, 32, 48), (0, 64, 64), (64, 256, 512)):
        ms, _, _ = M.pagerank(n, src, dst, 0.85, 0.0, 3, *geo)
        assert np.max(np.abs(ms - os_) / os_) <= 2e-5
    src = np.arange(0, 100, dtype=np.uint32)
    dst = src + 1
    os_, _, _ = O.OracleGraph(300, src, dst).pagerank(0.85, 0.0, 5)
    ms, _, _ = M.pagerank(300, src, dst, 0.85, 0.0, 5, 8, 32, 48)
    assert np.allclose(ms, os_, rtol=1e-6)
