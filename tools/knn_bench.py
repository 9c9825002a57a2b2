"""kNN-only micro-benchmark (used for ncu captures of the scan kernel).
    python tools/knn_bench.py [n_cells] [dims] [k] [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csb200 as cs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 151552
d = int(sys.argv[2]) if len(sys.argv) > 2 else 50
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = cs.Context(0)
rng = np.random.default_rng(0)
cent = rng.normal(size=(40, d)).astype(np.float32) * 2
x = cent[rng.integers(0, 40, n)] + rng.normal(size=(n, d)).astype(np.float32)
emb = ctx.dense_upload(x)
for r in range(reps):
    ctx.stage_times(reset=True)
    t0 = time.perf_counter()
    idx, dist = ctx.knn(emb, emb, k)
    ctx.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ms, _ = ctx.stage_times()
    flops = 2.0 * n * n * d
    print(f"n={n} d={d} k={k}: knn {ms['knn']:.2f} ms, scan {ms['knn_scan']:.2f} ms, wall {wall:.2f} ms, "
          f"{flops / ms['knn_scan'] / 1e9:.1f} TFLOP/s, {n * n / ms['knn_scan'] / 1e6:.1f} Gcand/s, "
          f"fallback rows {ctx.knn_last_fallback_rows()}", flush=True)
