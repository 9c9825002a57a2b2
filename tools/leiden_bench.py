"""Community detection timing on the graph of a benchmark workload: python tools/leiden_bench.py [C2|C3] [resolution]."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import csb200 as cs

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
res = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4           # the scRNA tutorial's value for 400k cells
n_genes, n_cells, density, n_features, n_pcs, k = bench.WORKLOADS[wl]
ctx = cs.Context(0)
raw = ctx.generate_counts(cs.synth.make_params(n_genes, n_cells, density=density, seed=1234), 0, n_cells)
out = cs.run_path(ctx, cs.LocalComm(), raw, n_cells, 0, cs.PathParams(n_features=n_features, n_pcs=n_pcs, k=30, graph_mode="sum"))
g = out["graph"]
for rep in range(2):
    ctx.synchronize()
    t0 = time.perf_counter()
    r = ctx.leiden(g, res)
    dt = time.perf_counter() - t0
print(json.dumps({"workload": wl, "n_cells": n_cells, "k": 30, "graph_nnz": g.info()[2], "resolution": res, "leiden_s": dt,
                  "n_communities": r["n_communities"], "levels": r["levels"], "quality": r["quality"],
                  "largest": int(np.bincount(r["labels"]).max()),
                  "reference": "run_clustering(res=1e-4, n_neighbors=30) incl. Leiden: 590 s at 400k cells (docs/scRNA_tutorial/README.md:280-282)"}))
