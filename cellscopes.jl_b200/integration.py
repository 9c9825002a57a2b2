"""run_harmony (src/integration/int_processing.jl:1-12) over libcellscopes_cuda.so: the host-side mirror of the
reference's entry point and of the HarmonyObject keywords (src/integration/int_objects.jl:153-166)."""
from __future__ import annotations

import numpy as np

from .processing import _session, default_context

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def seed_cells(n, k, random_state):
    """The K cells whose rows seed the k-means centres: the K smallest values of a counter-based hash of
    (random_state, cell).  Stands in for Clustering.jl's k-means++ under Random.seed!(random_state), which depends on
    Julia's RNG stream and cannot be reproduced outside Julia."""
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(random_state) ^ np.uint64(0xA0761D6478BD642F))
        h = _splitmix64(h ^ ((np.arange(n, dtype=np.uint64) * np.uint64(0x2545F4914F6CDD1D)) & _M64))
    return np.sort(np.argsort(h, kind="stable")[:k])


class HarmonyResult:
    """The fields of HarmonyObject a caller reads back (int_objects.jl:115-151)."""

    def __init__(self, Z_corr, R, objective_harmony, kmeans_rounds, K, theta, lamb, sigma):
        self.Z_corr, self.R = Z_corr, R
        self.objective_harmony, self.kmeans_rounds = objective_harmony, kmeans_rounds
        self.K, self.theta, self.lamb, self.sigma = K, theta, lamb, sigma


def harmony_matrix(data_mat, batch_values, theta=None, lamb=None, sigma=0.1, nclust=None, tau=0.0, block_size=0.05,
                   max_iter_harmony=10, max_iter_kmeans=20, epsilon_cluster=1e-5, epsilon_harmony=1e-4, verbose=True,
                   random_state=123, ctx=None, want_r=False):
    """HarmonyObject(data_mat, meta_data, vars_use; ...) on a cells x d matrix and the batch column."""
    c = ctx or default_context()
    z = np.ascontiguousarray(data_mat, dtype=np.float64)
    n = len(batch_values)
    if z.shape[0] != n:
        z = np.ascontiguousarray(z.T)                                        # int_objects.jl:169-171
    if z.shape[0] != n:
        raise ValueError("The number of cells in the metadata must match the number of cells in the count matrix.")
    levels, batch = np.unique(np.asarray(batch_values), return_inverse=True)  # categorical levels, sorted (get_dummies)
    nb = levels.size
    if nclust is None:
        nclust = min(int(round(n / 30.0)), 100)                               # :173-175
    k = int(nclust)
    sig = np.full(k, float(sigma)) if np.ndim(sigma) == 0 else np.asarray(sigma, dtype=np.float64)
    th = np.full(nb, 2.0 if theta is None else float(theta)) if np.ndim(theta) == 0 else np.asarray(theta, dtype=np.float64)
    lm = np.full(nb, 1.0 if lamb is None else float(lamb)) if np.ndim(lamb) == 0 else np.asarray(lamb, dtype=np.float64)
    n_b = np.bincount(batch, minlength=nb).astype(np.float64)
    if tau > 0:
        th = th * (1 - np.exp(-(n_b / (k * tau)) ** 2))                       # :195-197
    cen, _, _, _ = c.kmeans(z, k, seed_cells(n, k, random_state), 100, 1e-4, unit_rows=True)     # cluster_kmeans
    out = c.harmony(z, batch.astype(np.int32), nb, cen, th, lm, sig, block_size, max_iter_harmony, max_iter_kmeans,
                    epsilon_cluster, epsilon_harmony, random_state, want_r)
    return HarmonyResult(out["Z_corr"], out["R"], out["objective_harmony"], out["kmeans_rounds"], k, th, lm, sig)


def run_harmony(sc_obj, batch, ctx=None, **kwargs):
    """int_processing.jl:1-12: Harmony on the PCA embedding with the batch column ``batch`` of metaData; the corrected
    matrix replaces dimReduction.pca.cell_embedding (N x d Float64)."""
    if sc_obj.dimReduction is None or sc_obj.dimReduction.pca is None:
        raise ValueError("run_harmony: run_pca first")
    pca = sc_obj.dimReduction.pca
    meta = sc_obj.metaData
    ho = harmony_matrix(np.asarray(pca.cell_embedding, dtype=np.float64), meta[str(batch)].to_numpy(), ctx=ctx, **kwargs)
    pca.cell_embedding = ho.Z_corr
    pca.device = None                 # the device copy is the uncorrected embedding: run_clustering uploads the new one
    s = _session(sc_obj, ctx)
    s.emb = None
    sc_obj.harmony = ho
    return sc_obj
