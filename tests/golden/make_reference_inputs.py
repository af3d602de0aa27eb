"""Inputs of tools/run_reference.sh: the config-1 vectors and queries as JSON rows."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.util import SEED_DATA, SEED_QUERY, uniform_vectors  # noqa: E402

out = sys.argv[1]
X = uniform_vectors(10_000, 128, SEED_DATA)
Q = uniform_vectors(1_000, 128, SEED_QUERY)
json.dump([[i, [float(x) for x in v]] for i, v in enumerate(X)], open(os.path.join(out, "hnsw_vectors.json"), "w"))
json.dump([[float(x) for x in q] for q in Q], open(os.path.join(out, "hnsw_queries.json"), "w"))
