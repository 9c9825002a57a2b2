"""Writes the input matrix for julia/make_fixtures.jl: a small synthetic count matrix (the same generator as every other
test) as a MatrixMarket file, which is what the reference's own reader takes (src/fileio.jl:1-37).

    python tests/golden/make_ref_inputs.py [n_genes=6000] [n_cells=3000] [out=tests/golden/ref_in]
"""
import json
import os
import sys

import scipy.io

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import csb200 as cs  # noqa: E402

n_genes = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "tests", "golden", "ref_in")
os.makedirs(out, exist_ok=True)
p = cs.synth.make_params(n_genes, n_cells, density=0.06, seed=1234)
m = cs.synth.sample_counts(p)
scipy.io.mmwrite(os.path.join(out, "matrix.mtx"), m.astype("int64"), field="integer")
with open(os.path.join(out, "meta.json"), "w") as f:
    json.dump({"n_genes": n_genes, "n_cells": n_cells, "density": 0.06, "seed": 1234, "nnz": int(m.nnz)}, f)
print("wrote", out, m.shape, m.nnz)
