"""Builds tests/golden/air_routes.npz from the reference's own test fixture
(cozo-core/tests/air-routes-latest-{nodes,edges}.csv), following the relation the reference's
integration tests create from it (cozo-core/tests/air_routes.rs:96-127):

    :replace route { fr: String, to: String => dist: Float }     rows with label == 'route'

Run in the build container (needs /root/reference); the GPU box only sees the committed .npz.
The reference pins the row count of this relation: 50637 (air_routes.rs:189-209)."""
import csv
import os

import numpy as np

REF = "/root/reference/cozo-core/tests"
code = {}
with open(os.path.join(REF, "air-routes-latest-nodes.csv"), newline="") as f:
    r = csv.reader(f)
    next(r)
    for row in r:
        code[int(row[0])] = row[3]
routes = {}
with open(os.path.join(REF, "air-routes-latest-edges.csv"), newline="") as f:
    r = csv.reader(f)
    next(r)
    for row in r:
        if row[3] != "route":
            continue
        routes[(code[int(row[1])], code[int(row[2])])] = float(row[4])      # keyed relation: last write wins
keys = sorted(routes)                                                        # scan order of a stored relation
assert len(keys) == 50637, len(keys)
codes = sorted({k for pair in keys for k in pair})
idx = {c: i for i, c in enumerate(codes)}
np.savez_compressed(os.path.join(os.path.dirname(__file__), "air_routes.npz"),
                    codes=np.array(codes), fr=np.array([idx[a] for a, _ in keys], np.uint16),
                    to=np.array([idx[b] for _, b in keys], np.uint16),
                    dist=np.array([routes[k] for k in keys], np.float32))
print(len(keys), "routes,", len(codes), "airports")
