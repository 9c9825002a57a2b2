"""Host phases of csc_hvg_finish at whole-transcriptome size (33k genes): CSC_LOG_LEVEL=2 prints the checkpoints of
csrc/hvg.cu.   CSC_LOG_LEVEL=2 python tools/hvg_trace.py [n_genes]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csb200 as cs

G = int(sys.argv[1]) if len(sys.argv) > 1 else 33000
N = 1_000_000
rng = np.random.default_rng(1)
m = rng.lognormal(-3.0, 1.5, G)
s1 = np.floor(m * N) + 5
s2 = np.floor(s1 * (1.0 + m * rng.uniform(0.5, 1.0, G)) + s1 * s1 / N)
ctx = cs.Context(0)
stats = ctx.vec_from(np.concatenate([s1, s2]))
for it in range(4):
    ctx.synchronize()
    t0 = time.perf_counter()
    out = ctx.hvg_finish(stats, G, N, 2000, 0.3)
    print(f"hvg_finish {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr, flush=True)
