"""Writes tests/golden/oracle_small.json: the oracle's outputs on a fixed 40-gene x 60-cell input.

The reference (Julia) cannot run in the build image and ships no fixtures, so these vectors pin the
ORACLE (regression guard), not the reference: parity stays "unpinned" (see oracle header).
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import cellscopes_oracle as oracle  # noqa: E402

rng = np.random.default_rng(20261019)
lat = rng.normal(size=(60, 3))
load = rng.normal(size=(40, 3)) * 0.8
lam = np.exp(-0.5 + load @ lat.T + 0.3 * rng.normal(size=(1, 60)))
counts = rng.poisson(lam)
counts[counts.sum(axis=1) == 0, 0] = 1
counts[0, counts.sum(axis=0) == 0] = 1
res = oracle.run_pipeline(sp.csc_matrix(counts), nFeatures=12, n_pcs=4, k=5)
out = {
    "counts": counts.tolist(), "nFeatures": 12, "n_pcs": 4, "k": 5,
    "norm": res["norm"].toarray().tolist(),
    "features": [int(v) for v in res["hvg"]["features"]],
    "variance_standardized": res["hvg"]["variance_standardized"].tolist(),
    "scaled": res["scaled"].tolist(),
    "principalvars": res["pca"]["principalvars"].tolist(),
    "embedding": res["pca"]["embedding"].tolist(),
    "knn_idx": res["knn_idx"].tolist(),
    "adj": res["adj"].toarray().tolist(),
}
with open(os.path.join(HERE, "oracle_small.json"), "w") as f:
    json.dump(out, f)
print("wrote oracle_small.json")
