"""Where the host-facing path spends its time: cProfile of one normalize_object..run_clustering pass with host
(pinned) CSC arrays.   python tools/e2e_profile.py [n_genes] [n_cells] [density]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csb200 as cs
import torch

G = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
d = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
ctx = cs.Context(0)
p = cs.synth.make_params(G, N, density=d, seed=1234)
raw = ctx.generate_counts(p, 0, N)
colptr, rowval, nzval = raw.download(0, N, 0)
del raw


def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
    v = t.numpy()
    v[...] = a
    return t, v


keep = [pinned(colptr), pinned(rowval), pinned(nzval)]
colptr, rowval, nzval = (k[1] for k in keep)


def one():
    m = cs.csc_from_arrays(nzval, rowval, colptr, (G, N))
    obj = cs.scRNAObject.__new__(cs.scRNAObject)
    obj.rawCount = cs.RawCountObject(m, np.arange(N), np.arange(G))
    obj.normCount = obj.scaleCount = obj.varGene = obj.dimReduction = obj.clustData = None
    obj.metaData, obj.imaging, obj._session = None, False, None
    t = [time.perf_counter()]
    obj = cs.normalize_object(obj, ctx=ctx); ctx.synchronize(); t.append(time.perf_counter())
    obj = cs.find_variable_genes(obj, nFeatures=2000, ctx=ctx); ctx.synchronize(); t.append(time.perf_counter())
    obj = cs.scale_object(obj, features=obj.varGene.var_gene, ctx=ctx); ctx.synchronize(); t.append(time.perf_counter())
    obj = cs.run_pca(obj, maxoutdim=50, ctx=ctx); ctx.synchronize(); t.append(time.perf_counter())
    obj = cs.run_clustering(obj, n_neighbors=20, graph="sum", ctx=ctx); ctx.synchronize(); t.append(time.perf_counter())
    return np.diff(t) * 1e3


one()
print("ms per call [normalize, hvg, scale, pca, clustering]:", np.round(one(), 1), flush=True)
pr = cProfile.Profile()
pr.enable()
one()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
