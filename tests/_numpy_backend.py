"""TEST DOUBLE for the C-ABI backend: the same method surface as ``_capi.Context`` as far as
``pipeline.run_path`` uses it, computed with the CPU oracle.  It exists so that the multi-rank host
logic (sharding, partial/finish split, collectives) can be exercised with gloo on a box with no GPU.
It is never imported by the package."""
import numpy as np
import scipy.sparse as sp

import cellscopes_oracle as oracle


class NpVec:
    def __init__(self, a):
        self.a = a

    def __array__(self, dtype=None, copy=None):
        return self.a

    def download(self, n=None):
        return self.a.copy()

    def free(self):
        pass


class NpDense:
    def __init__(self, a):
        self.a = np.ascontiguousarray(a)

    def __array__(self, dtype=None, copy=None):
        return self.a

    @property
    def shape(self):
        return self.a.shape

    def download(self, dtype=np.float64, row_lo=0, row_hi=None):
        return self.a[row_lo:row_hi].astype(dtype)

    def free(self):
        pass


class NpSparse:
    def __init__(self, m):
        self.m = sp.csc_matrix(m)

    @property
    def shape(self):
        return self.m.shape

    @property
    def nnz(self):
        return self.m.nnz

    def free(self):
        pass


class NpPca:
    def __init__(self, d):
        self.d = d
        self.iters, self.resid = 0, 0.0

    def info(self):
        return self.d


class NpGraph:
    def __init__(self, m):
        self.m = m

    def to_scipy(self):
        return self.m


class NumpyBackend:
    def vec(self, n, dtype=np.float64):
        return NpVec(np.zeros(int(n), dtype=dtype))

    def normalize(self, raw, scale_factor, norm_method, pseudocount, nan_to_zero, gene_stats):
        m = raw.m
        g = m.shape[0]
        if gene_stats is not None:
            s1, s2, _, _ = oracle.gene_moments(m) if m.shape[1] else (np.zeros(g), np.zeros(g), None, None)
            gene_stats.a[:g] += s1
            gene_stats.a[g:2 * g] += s2
        out = oracle.normalize_object(m, scale_factor, "logarithm" if norm_method == 0 else "none", pseudocount,
                                      nan_to_zero)
        return NpSparse(out)

    def hvg_finish(self, gene_stats, n_genes, n_cells_total, n_features, span):
        s1, s2 = gene_stats.a[:n_genes], gene_stats.a[n_genes:2 * n_genes]
        n = n_cells_total
        mean = s1 / n
        var = (s2 - s1 * s1 / n) / (n - 1)
        lx, ly = np.log10(mean), np.log10(var)
        model = oracle.loess_fit(lx, ly, span)
        ve = 10.0 ** oracle.loess_predict(model, lx)
        vs = var / ve
        order = np.lexsort((np.arange(n_genes), -vs))
        return {"mean": mean, "variance": var, "variance_expected": ve, "variance_standardized": vs,
                "order": order, "features": order[:min(n_features, n_genes)].copy()}

    def scale_stats_partial(self, norm, features, partial):
        rows = np.unique(features)
        sub = norm.m.tocsr()[rows]
        h = rows.size
        partial.a[:h] += np.asarray(sub.sum(axis=1)).ravel()
        partial.a[h:2 * h] += np.asarray(sub.multiply(sub).sum(axis=1)).ravel()

    def scale_fill(self, norm, features, partial, n_cells_total, scale_max, do_scale, do_center):
        rows = np.unique(features)
        h = rows.size
        n = n_cells_total
        s1, s2 = partial.a[:h], partial.a[h:2 * h]
        mean = s1 / n
        sd = np.sqrt((s2 - s1 * s1 / n) / (n - 1))
        x = np.asarray(norm.m.tocsr()[rows].todense(), dtype=np.float64)
        if do_center:
            x = x - mean[:, None]
        if do_scale:
            x = x / sd[:, None]
        return NpDense(np.minimum(x, scale_max).T)

    def pca_partial(self, scaled, colsum, gram):
        z = scaled.a
        colsum.a += z.sum(axis=0)
        gram.a += (z.T @ z).ravel()

    def pca_solve(self, colsum, gram, n_feat, n_cells_total, maxoutdim, tol, max_iter):
        h, n = n_feat, n_cells_total
        g = gram.a.reshape(h, h) - np.outer(colsum.a, colsum.a) / n
        lam, u = np.linalg.eigh(g)
        lam, u = lam[::-1], u[:, ::-1]
        k = min(maxoutdim, h)
        return NpPca({"principalvars": lam[:k] / (n - 1), "tvar": np.trace(g) / (n - 1), "loadings": u[:, :k],
                      "mean": colsum.a / n, "k": k, "n_feat": h})

    def pca_transform(self, scaled, pca):
        d = pca.d
        return NpDense((scaled.a - d["mean"][None, :]) @ d["loadings"])

    def dense_wrap(self, ptr, n_rows, n_cols, keepalive=None):
        return NpDense(keepalive.numpy().reshape(n_rows, n_cols))

    def vec_wrap_i32(self, tensor):
        return NpVec(tensor.numpy().ravel())

    def knn(self, queries, keys, k, query_offset=0, dim_lo=0, dim_hi=None, metric=0):
        x = np.asarray(keys.a, dtype=np.float64)
        rows = np.arange(query_offset, query_offset + queries.shape[0])
        idx, dist = oracle.knn_exact(x, k, "cosine" if metric == 0 else "euclidean", queries=rows)
        return NpVec(idx.astype(np.int32).ravel()), NpVec(dist.astype(np.float32).ravel())

    def graph_build(self, knn_idx, n_cells_total, k, cell_lo, cell_hi, mode, snn_prune):
        idx = knn_idx.a.reshape(n_cells_total, k)
        full = (oracle.adjacency_small(idx), oracle.adjacency_atlas(idx), oracle.snn_jaccard(idx, snn_prune))[mode]
        return NpGraph(full[:, cell_lo:cell_hi])
