import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csb200 as cs
ctx = cs.Context(0)
np.set_printoptions(linewidth=200, precision=4, suppress=False)
for (n, H, mode) in [(64, 256, "rand"), (48, 512, "rand"), (5000, 300, "rand"), (20000, 1000, "rand")]:
    rng = np.random.default_rng(1)
    z = rng.standard_normal((n, H)).astype(np.float32)
    dense = ctx.dense_upload(z)
    colsum, gram = ctx.vec(H), ctx.vec(H * H)
    ctx.pca_partial(dense, colsum, gram)
    g = gram.download().reshape(H, H)
    ref = z.astype(np.float64).T @ z.astype(np.float64)
    print(n, H, mode, "max abs err", np.abs(g - ref).max(), "ref max", np.abs(ref).max(), flush=True)
