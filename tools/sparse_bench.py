"""Sparse-stage micro-benchmark: normalize (+ fused gene moments), scale statistics and the dense fill on a
synthetic count matrix generated on the device.
    python tools/sparse_bench.py [n_genes] [n_cells] [density] [n_features] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csb200 as cs

G = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
d = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
H = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ctx = cs.Context(0)
p = cs.synth.make_params(G, N, density=d, seed=1234)
raw = ctx.generate_counts(p, 0, N)
nnz = raw.nnz
print(f"G={G} N={N} nnz={nnz} ({nnz / (G * N):.4f})", flush=True)
peak = 6531.0
for r in range(reps):
    ctx.flush_l2()
    ctx.stage_times(reset=True)
    stats = ctx.vec(2 * G, np.float64)
    norm = ctx.normalize(raw, 1e4, cs.pipeline.norm_code("logarithm"), 1.0, False, stats)
    hv = ctx.hvg_finish(stats, G, N, H, 0.3)
    feats = hv["features"]
    rows = np.unique(feats)
    part = ctx.vec(2 * rows.size, np.float64)
    ctx.flush_l2()
    src, src_feats = norm, feats
    if G >= 4 * rows.size:                      # whole-transcriptome route of pipeline.run_path: subset_count first
        src, src_feats = ctx.select_rows(norm, feats), None
    ctx.scale_stats_partial(src, src_feats, part)
    ctx.flush_l2()
    scaled = ctx.scale_fill(src, src_feats, part, N)
    ctx.synchronize()
    ms, _ = ctx.stage_times()
    hn = rows.size
    b_norm = 12.0 * nnz + 8.0 * N
    b_stats = 8.0 * nnz + 8.0 * N
    b_fill = 8.0 * nnz + 8.0 * N + 4.0 * hn * N
    print(f"normalize+moments {ms['normalize']:.3f} ms {b_norm / ms['normalize'] / 1e6:.0f} GB/s ({b_norm / ms['normalize'] / 1e6 / peak:.2f}) | "
          f"scale_stats {ms['scale_stats']:.3f} ms {b_stats / ms['scale_stats'] / 1e6:.0f} GB/s ({b_stats / ms['scale_stats'] / 1e6 / peak:.2f}) | "
          f"scale_fill {ms['scale_fill']:.3f} ms {b_fill / ms['scale_fill'] / 1e6:.0f} GB/s ({b_fill / ms['scale_fill'] / 1e6 / peak:.2f}) | hvg {ms['hvg']:.2f} ms",
          flush=True)
    del norm, scaled, src
