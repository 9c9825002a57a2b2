"""PCA-solver micro-benchmark: top-k eigenpairs of a synthetic H x H Gram matrix with a planted spectrum
(decaying top-k above a Marchenko-Pastur-like bulk), checked against numpy.linalg.eigh.
    python tools/pca_bench.py [H] [k] [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csb200 as cs

H = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 50
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n_cells = 500000
rng = np.random.default_rng(0)
top = np.geomspace(6.9, 1.24, k)
bulk = np.sort(rng.uniform(0.0, 1.0, H - k) ** 0.5 * 1.12)[::-1]
lam = np.r_[top, bulk] * (n_cells - 1)
U, _ = np.linalg.qr(rng.normal(size=(H, H)))
G = (U * lam) @ U.T
G = (G + G.T) / 2
ctx = cs.Context(0)
gram = ctx.vec_from(G.ravel())
colsum = ctx.vec(H, np.float64)
w = np.linalg.eigvalsh(G)[::-1][:k] / (n_cells - 1)
for r in range(reps):
    ctx.stage_times(reset=True)
    t0 = time.perf_counter()
    model = ctx.pca_solve(colsum, gram, H, n_cells, k)
    ctx.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms, launches = ctx.stage_times()
    info = model.info()
    err = np.abs(info["principalvars"] - w).max() / w[0]
    print(f"H={H} k={k}: solve {ms['pca_solve']:.2f} ms (wall {wall:.2f}), matmuls {model.iters}, resid {model.resid:.2e}, "
          f"eigenvalue err {err:.2e}, launches {launches}", flush=True)
