"""Host-side mirror of the reference's operator interface for the hot path
(src/scrna/processing.jl): same function names, keyword names, defaults, mutation-and-return
behaviour and error cases, with the arithmetic done by libcellscopes_cuda.so through the C ABI.

    obj = scRNAObject(RawCountObject(counts, cells, genes))
    obj = normalize_object(obj)                       # processing.jl:40
    obj = find_variable_genes(obj)                    # processing.jl:157
    obj = scale_object(obj, features=obj.varGene.var_gene)   # processing.jl:88
    obj = run_pca(obj, maxoutdim=50)                  # processing.jl:164
    obj = run_clustering(obj, n_neighbors=20)         # processing.jl:309 (graph part; Leiden is out of scope)

A session (one ``Context`` = one GPU + the uploaded raw counts and the intermediate device
handles) is attached to the object on first use, so consecutive stages reuse device-resident data.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.sparse as sp

from . import _capi
from .objects import (ClusteringObject, LazyDeviceMatrix, NormCountObject, PCAObject, RawCountObject,
                      ReductionObject, ScaleCountObject, UMAPObject, VariableGeneObject, scRNAObject)
from .pipeline import metric_code, norm_code

_default_ctx = None


def default_context(device: int = 0) -> _capi.Context:
    global _default_ctx
    if _default_ctx is None or _default_ctx.h is None or _default_ctx.device != device:
        _default_ctx = _capi.Context(device)
    return _default_ctx


class _Session:
    """Device-resident state of one object."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.raw = None          # Sparse
        self.raw_key = None
        self.gene_stats = None   # Vec fp64[2G] raw moments
        self.norm = None         # Sparse
        self.scaled = None       # _ScaledSource
        self.pca = None
        self.emb = None
        self.knn = None


def _session(sc_obj, ctx=None) -> _Session:
    s = getattr(sc_obj, "_session", None)
    if s is None or (ctx is not None and s.ctx is not ctx):
        s = _Session(ctx or default_context())
        sc_obj._session = s
    return s


def _ensure_raw(sc_obj, s: _Session):
    m = sc_obj.rawCount._mtx
    key = id(m)
    if s.raw is None or s.raw_key != key:
        if isinstance(m, LazyDeviceMatrix):
            s.raw = m.handle
        else:
            if not sp.issparse(m):
                m = sp.csc_matrix(np.asarray(m))      # dense input is legal (atac_processing.jl:87-89)
            s.raw = s.ctx.sparse_from_scipy(m)
        s.raw_key = key
        s.gene_stats = None
    return s.raw


# ---------------------------------------------------------------------------------------------
# normalize_object  (processing.jl:15-48)
# ---------------------------------------------------------------------------------------------
def normalize_object(x, scale_factor=10000, norm_method="logarithm", pseudocount=1, ctx=None):
    code = norm_code(norm_method)
    if isinstance(x, scRNAObject):
        s = _session(x, ctx)
        raw = _ensure_raw(x, s)
        g = raw.shape[0]
        s.gene_stats = s.ctx.vec(2 * g, np.float64)
        # processing.jl:44-46: imaging objects replace NaN by 0
        s.norm = s.ctx.normalize(raw, scale_factor, code, pseudocount, bool(getattr(x, "imaging", False)), s.gene_stats)
        pattern = x.rawCount._mtx if sp.issparse(x.rawCount._mtx) else None
        lazy = LazyDeviceMatrix(s.norm, "sparse", raw.shape, pattern)
        x.normCount = NormCountObject(lazy, x.rawCount.cell_name, x.rawCount.gene_name, scale_factor, norm_method,
                                      pseudocount)
        return x
    if isinstance(x, RawCountObject):
        mtx = normalize_object(x.count_mtx, scale_factor, norm_method, pseudocount, ctx)
        return NormCountObject(mtx, x.cell_name, x.gene_name, scale_factor, norm_method, pseudocount)
    # matrix method (processing.jl:15-31): returns a matrix with the same stored pattern
    c = ctx or default_context()
    m = sp.csc_matrix(x) if not sp.issparse(x) else x.tocsc()
    m.sort_indices()
    raw = c.sparse_from_scipy(m)
    norm = c.normalize(raw, scale_factor, code, pseudocount, False, None)
    val = norm.download(values_only=True)
    return sp.csc_matrix((val, m.indices.copy(), m.indptr.copy()), shape=m.shape)


# ---------------------------------------------------------------------------------------------
# find_variable_genes  (processing.jl:133-162)
# ---------------------------------------------------------------------------------------------
def _vst_frame(res, gene_name):
    order = res["order"]
    # one fancy-indexing pass per column (a Python loop over 33k gene names cost more than the device part of the call)
    genes = gene_name if isinstance(gene_name, np.ndarray) else np.asarray(gene_name, dtype=object)
    df = pd.DataFrame({
        "mean": res["mean"][order], "variance": res["variance"][order],
        "variance_expected": res["variance_expected"][order],
        "variance_standardized": res["variance_standardized"][order],
        "gene": genes[order],
    })
    return df


def find_variable_genes(x, nFeatures: int = 2000, span: float = 0.3, ctx=None):
    if isinstance(x, scRNAObject):
        s = _session(x, ctx)
        raw = _ensure_raw(x, s)
        g, n = raw.shape
        if s.gene_stats is None:
            s.gene_stats = s.ctx.vec(2 * g, np.float64)
            s.ctx.gene_stats_partial(raw, s.gene_stats)
        res = s.ctx.hvg_finish(s.gene_stats, g, n, nFeatures, span)
        genes = x.rawCount.gene_name
        vst = _vst_frame(res, genes)
        feats = [genes[i] for i in res["features"]]
        x.varGene = VariableGeneObject(feats, vst)
        x.varGene.feature_idx = res["features"]
        return x
    if isinstance(x, RawCountObject):
        c = ctx or default_context()
        m = x.count_mtx
        raw = c.sparse_from_scipy(m if sp.issparse(m) else sp.csc_matrix(np.asarray(m)))
        g, n = raw.shape
        st = c.vec(2 * g, np.float64)
        c.gene_stats_partial(raw, st)
        res = c.hvg_finish(st, g, n, nFeatures, span)
        return _vst_frame(res, x.gene_name), [x.gene_name[i] for i in res["features"]]
    raise TypeError("find_variable_genes expects an scRNAObject or a RawCountObject")


# ---------------------------------------------------------------------------------------------
# scale_object  (processing.jl:65-93)
# ---------------------------------------------------------------------------------------------
class _ScaledSource:
    """What defines a scaled matrix on the device: the normalised rows it is taken from (``src``: the normalised
    matrix itself with ``feats`` naming the rows, or the row subset made by csc_sparse_select_rows with
    ``feats = None``), the all-reduced moments ``part`` and the scale arguments.  The dense matrix (4 H bytes per cell)
    is only built when somebody reads it (``dense()``): run_pca takes the Gram straight from ``src``
    (csc_scale_gram_partial) and the embedding from the sparse rows (csc_pca_transform_sparse)."""

    def __init__(self, c, src, feats, part, rows, n_cells_total, scale_max, do_scale, do_center):
        self.c, self.src, self.feats, self.part, self.rows = c, src, feats, part, rows
        self.n_cells_total, self.scale_max, self.do_scale, self.do_center = n_cells_total, scale_max, do_scale, do_center
        self._dense = None

    @property
    def shape(self):
        return self.src.shape[1], int(self.rows.size)      # cells x selected genes (device layout)

    def dense(self):
        if self._dense is None:
            self._dense = self.c.scale_fill(self.src, self.feats, self.part, self.n_cells_total, self.scale_max,
                                            self.do_scale, self.do_center)
        return self._dense

    def download(self, dtype=np.float64, row_lo=0, row_hi=None):      # LazyDeviceMatrix("dense_t") reads through this
        return self.dense().download(dtype, row_lo, row_hi)


def _scale_device(c, norm_dev, feature_idx, n_cells_total, scale_max, do_scale, do_center):
    g = norm_dev.shape[0]
    rows = np.arange(g, dtype=np.int64) if feature_idx is None else np.unique(np.asarray(feature_idx, dtype=np.int64))
    part = c.vec(2 * rows.size, np.float64)
    src, feats = norm_dev, feature_idx
    if feature_idx is not None and g >= 4 * rows.size:
        # subset_count on the device (utils.jl:59-87): a whole-transcriptome matrix keeps ~6 % of its rows, every later
        # pass reads the subset
        src, feats = c.select_rows(norm_dev, feature_idx), None
    c.scale_stats_partial(src, feats, part)
    return _ScaledSource(c, src, feats, part, rows, n_cells_total, scale_max, do_scale, do_center)


def scale_object(x, features=None, scale_max=10.0, do_scale: bool = True, do_center: bool = True, ctx=None, **kwargs):
    if isinstance(x, scRNAObject):
        if x.normCount is None:
            raise ValueError("scale_object: normCount is missing; run normalize_object first")
        s = _session(x, ctx)
        if s.norm is None:
            s.norm = s.ctx.sparse_from_scipy(x.normCount.count_mtx)
        genes = x.normCount.gene_name
        fidx = None
        if features is not None:
            pos = {g: i for i, g in enumerate(genes)}
            fidx = np.array([pos[f] for f in features if f in pos], dtype=np.int64)
        # processing.jl:78-83 forwards only kwargs... to the matrix method, so the effective values are
        # always the defaults (10.0, true, true) while the user's values are recorded in the object.
        eff_max = kwargs.get("scale_max", 10.0)
        eff_scale = kwargs.get("do_scale", True)
        eff_center = kwargs.get("do_center", True)
        sc = _scale_device(s.ctx, s.norm, fidx, s.norm.shape[1], eff_max, eff_scale, eff_center)
        s.scaled = sc
        # scaleCount stays on the device, and is not even built there until it is read (count_mtx) or run_pca needs it
        lazy = LazyDeviceMatrix(sc, "dense_t", (sc.rows.size, s.norm.shape[1]))
        x.scaleCount = ScaleCountObject(lazy, x.normCount.cell_name, [genes[i] for i in sc.rows], do_scale, do_center,
                                        scale_max)
        return x
    # matrix method (processing.jl:65-76): genes x cells in, same shape out (dense)
    c = ctx or default_context()
    m = x.count_mtx if isinstance(x, NormCountObject) else x
    m = sp.csc_matrix(m) if not sp.issparse(m) else m.tocsc()
    dev = c.sparse_from_scipy(m)
    out = _scale_device(c, dev, None, m.shape[1], scale_max, do_scale, do_center).download(np.float64).T
    if isinstance(x, NormCountObject):
        return ScaleCountObject(out, x.cell_name, x.gene_name, do_scale, do_center, scale_max)
    return out


# ---------------------------------------------------------------------------------------------
# run_pca  (processing.jl:164-193)
# ---------------------------------------------------------------------------------------------
def _pca_device(c, scaled, n_cells_total, maxoutdim, tol=1e-9, max_iter=500):
    """``scaled``: a ``_ScaledSource`` (Gram straight from the sparse normalised rows, embedding from them too: the dense
    matrix is never built) or a dense device matrix (uploaded scaleCount)."""
    if isinstance(scaled, _ScaledSource):
        h = int(scaled.rows.size)
        colsum = c.vec(h, np.float64)
        gram = c.vec(h * h, np.float64)
        split = None
        if scaled._dense is not None:
            c.pca_partial(scaled._dense, colsum, gram)
        else:
            split = c.scale_gram_partial(scaled.src, scaled.feats, scaled.part, scaled.n_cells_total, colsum, gram,
                                         scaled.scale_max, scaled.do_scale, scaled.do_center, keep_operands=maxoutdim <= 64)
        model = c.pca_solve(colsum, gram, h, n_cells_total, maxoutdim, tol, max_iter)
        if split is not None:
            emb = c.pca_transform_split(split, model)
            split.free()
        else:
            emb = c.pca_transform_sparse(scaled.src, scaled.feats, scaled.part, scaled.n_cells_total, model, scaled.scale_max,
                                         scaled.do_scale, scaled.do_center)
        return model, emb
    h = scaled.shape[1]
    colsum = c.vec(h, np.float64)
    gram = c.vec(h * h, np.float64)
    c.pca_partial(scaled, colsum, gram)
    model = c.pca_solve(colsum, gram, h, n_cells_total, maxoutdim, tol, max_iter)
    return model, c.pca_transform(scaled, model)


def run_pca(sc_obj, method="svd", pratio=1, maxoutdim=10, package="MLJ", ctx=None):
    if method not in ("svd", "auto", "cov"):
        raise ValueError(f"unsupported PCA method {method!r}")
    if sc_obj.scaleCount is None or sc_obj.varGene is None:
        raise ValueError("run_pca: scaleCount / varGene missing; run scale_object and find_variable_genes first")
    s = _session(sc_obj, ctx)
    features = sc_obj.varGene.var_gene
    sgenes = sc_obj.scaleCount.gene_name
    if len(sgenes) == len(sc_obj.rawCount.gene_name) and len(features) != len(sgenes):
        # processing.jl:166-167: scaleCount still holds all genes -> subset to the variable genes.
        # Re-scaling the subset from normCount gives the same rows (scaling is per gene).
        if s.norm is None:
            s.norm = s.ctx.sparse_from_scipy(sc_obj.normCount.count_mtx)
        pos = {g: i for i, g in enumerate(sc_obj.normCount.gene_name)}
        fidx = np.array([pos[f] for f in features], dtype=np.int64)
        scaled = _scale_device(s.ctx, s.norm, fidx, s.norm.shape[1], 10.0, True, True)
    elif s.scaled is not None:
        scaled = s.scaled
    else:
        scaled = s.ctx.dense_upload(np.ascontiguousarray(np.asarray(sc_obj.scaleCount.count_mtx).T))
    n = scaled.shape[0]
    model, emb = _pca_device(s.ctx, scaled, n, maxoutdim)
    s.pca, s.emb = model, emb
    info = model.info()
    proj = emb.download(np.float64)                     # N x k Float64 (utils.jl:108-117, int_objects.jl:153)
    # default "MLJ" branch stores no model and no explained variance (processing.jl:180-181); the other branch
    # stores percent_var (:186).  Its loadings-as-embedding quirk (:185) is NOT replicated.
    percent_var = None if package == "MLJ" else info["principalvars"] / info["tvar"] * 100
    pca_obj = PCAObject(proj, None, percent_var, "PC", method, pratio, maxoutdim)
    pca_obj.device = emb
    pca_obj.principalvars = info["principalvars"]
    pca_obj.tvar = info["tvar"]
    pca_obj.loadings = info["loadings"]
    pca_obj.solver = {"iters": model.iters, "resid": model.resid}
    sc_obj.dimReduction = ReductionObject(pca_obj, None, None)
    return sc_obj


# ---------------------------------------------------------------------------------------------
# run_clustering  (processing.jl:195-318): kNN + adjacency.  Leiden (:233-234, :279-280) is out of scope.
# ---------------------------------------------------------------------------------------------
def run_clustering(sc_obj, n_neighbors=30, metric="cosine", res=0.06, seed_use=1234, graph=None, snn_prune=0.0,
                   partition_fn=None, leiden=True, ctx=None):
    """Builds the graph the reference hands to Leiden.

    n < 10000 cells -> the "small" binary union adjacency (:268-278); otherwise the "atlas" COO-summed
    adjacency where mutual pairs carry 2 (:218-232).  ``graph="snn"`` gives the Jaccard-weighted SNN graph.
    The kNN is exact (the reference's NN-descent is approximate and unseeded).  The communities come from the device
    Leiden (csc_leiden: CPM quality with resolution ``res``, like Leiden.jl; deterministic, ``seed_use`` is not needed);
    ``leiden=False`` stops after the graph, ``partition_fn(adj, res, seed)`` substitutes another community detection.
    """
    if sc_obj.dimReduction is None or sc_obj.dimReduction.pca is None:
        raise ValueError("run_clustering: run_pca first")
    s = _session(sc_obj, ctx)
    pca = sc_obj.dimReduction.pca
    emb = pca.device if pca.device is not None else s.ctx.dense_upload(np.asarray(pca.cell_embedding))
    n = emb.shape[0]
    if n != sc_obj.rawCount.shape[1]:
        raise ValueError("embedding and count matrix disagree on the number of cells")
    idx, dist = s.ctx.knn(emb, emb, n_neighbors, 0, 0, emb.shape[1], metric_code(metric))
    if graph is None:
        graph = "binary" if n < 10000 else "sum"       # processing.jl:310-317
    mode = {"binary": _capi.GRAPH_BINARY, "sum": _capi.GRAPH_SUM, "snn": _capi.GRAPH_SNN_JACCARD}[graph]
    gdev = s.ctx.graph_build(idx, n, n_neighbors, 0, n, mode, snn_prune)
    adj = gdev.to_scipy(np.float64 if graph == "snn" else np.int64)      # SparseMatrixCSC{Int64,Int64} in the reference
    s.knn = (idx, dist)
    clustering = None
    result = None
    if partition_fn is None and leiden:
        # Leiden.leiden(adj_mat, resolution = res) (processing.jl:233-234, 279-280) on the device graph
        ld = s.ctx.leiden(gdev, res)
        order = np.argsort(-np.bincount(ld["labels"], minlength=ld["n_communities"]), kind="stable")   # largest first
        members = [np.flatnonzero(ld["labels"] == c) for c in order]
        result = {"partition": members, "quality": ld["quality"], "levels": ld["levels"]}
        labels = np.empty(n, dtype=object)
        for i, mem in enumerate(members):
            labels[mem] = str(i + 1)
        clustering = pd.DataFrame({"cluster": labels, "cell_id": sc_obj.rawCount.cell_name})   # ordered like colnames(obj), :241
        if sc_obj.metaData is not None:
            sc_obj.metaData["cluster"] = labels
    elif partition_fn is not None:
        result = partition_fn(adj, res, seed_use)
        labels = np.empty(n, dtype=object)
        for i, members in enumerate(result["partition"]):
            labels[np.asarray(members)] = str(i + 1)
        clustering = pd.DataFrame({"cluster": labels, "cell_id": sc_obj.rawCount.cell_name})
        sc_obj.metaData["cluster"] = labels
    cl = ClusteringObject(clustering, str(metric), adj, result, res)
    # k x n like the reference's knn_matrices, 1-based; a Julia k x n matrix is column-major, i.e. the device's
    # [n][k] layout: convert on the contiguous buffers and hand out transposed VIEWS (arithmetic on the transposed
    # view would copy with a stride of k elements)
    # the 1-based rebasing and the Float64 widening happen on the device before the copy (csc_vec_download_as): a
    # host pass over the 10 M-entry tables of C2 costs ~5 ms each, the wider copy 0.7 ms
    cl.knn_indices = idx.download_as(np.int32, add=1).reshape(n, n_neighbors).T
    cl.knn_distances = dist.download_as(np.float64).reshape(n, n_neighbors).T
    cl.device_graph = gdev
    sc_obj.clustData = cl
    # hand-off to run_umap / run_clustering_atlas: the reference's atlas branch takes its neighbour table from
    # `dimReduction.umap.knn_data` when a UMAP object exists (processing.jl:196-197; filled by run_umap at :334-346 with
    # the k x n index matrix of UMAP's own NN-descent).  The exact table is put there, so a later run_umap-less
    # run_clustering of the reference (or anything else that reads knn_data) consumes the exact kNN.
    um = sc_obj.dimReduction.umap
    if um is None:
        sc_obj.dimReduction.umap = UMAPObject(None, "UMAP", None, emb.shape[1], int(n_neighbors), str(metric), None,
                                              cl.knn_indices)
    else:
        um.knn_data, um.n_neighbors, um.metric = cl.knn_indices, int(n_neighbors), str(metric)
    return sc_obj


# ---------------------------------------------------------------------------------------------
# find_markers / find_all_markers  (processing.jl:350-414)
# ---------------------------------------------------------------------------------------------
def _normal_two_sided(z):
    """2 * ccdf(Normal(), z) = erfc(z / sqrt 2), vectorised, without leaving numpy."""
    from scipy.special import erfc
    return erfc(np.asarray(z, dtype=np.float64) / np.sqrt(2.0))


def _exact_small_pvalues(c, norm_dev, cells_1, cells_2, genes):
    """The exact branch of HypothesisTests.MannWhitneyUTest (n1 + n2 <= 10, or <= 50 without ties): at most 50 cells,
    so the rows come to the host and the permutation distribution is counted there."""
    import itertools
    import scipy.stats
    cols = np.concatenate([cells_1, cells_2])
    sub = sp.hstack([norm_dev.to_scipy(int(j), int(j) + 1) for j in cols]).tocsr()[genes].toarray()
    nx, ny = len(cells_1), len(cells_2)
    out = np.full(len(genes), np.nan)
    for r, row in enumerate(sub):
        ranks = scipy.stats.rankdata(row)
        if nx + ny > 10 and np.unique(row).size < row.size:
            continue                                     # ties and more than 10 cells: the approximate test applies
        u = ranks[:nx].sum() - nx * (nx + 1) / 2
        if nx + ny <= 10:
            us = np.array([ranks[list(cb)].sum() - nx * (nx + 1) / 2 for cb in itertools.combinations(range(nx + ny), nx)])
        else:
            us = None
        mid = nx * ny / 2
        if us is not None:
            p = np.mean(us <= u) if u < mid else np.mean(us >= u)
        else:
            p = scipy.stats.mannwhitneyu(row[:nx], row[nx:], alternative="less" if u < mid else "greater", method="exact").pvalue
        out[r] = min(1.0, 2 * p)
    return out


def _marker_frame(genes, tabs, g1, n1, n2, rest, expr_cutoff, min_pct, p_cutoff, only_pos, exact_fn=None):
    """Result columns of find_markers from the per-(gene, group) tables.  ``g1``: column of cluster_1; cluster_2 is
    column ``rest`` (an int) or everything else (``rest is None``: the totals minus cluster_1)."""
    R, S, P, Z, X = (tabs[k] for k in ("rank_sum", "sum_expm1", "n_expr", "nnz", "sum_x"))
    n = float(n1 + n2)

    def col(t, j):
        return t[:, j] if j is not None else t.sum(axis=1) - t[:, g1]
    nnz_all = Z.sum(axis=1)
    n_zero_all = tabs["n_included"] - nnz_all + tabs["n_zero_stored"]              # the zero tie group of every gene
    zero_rank = tabs["n_neg"] + 0.5 * (n_zero_all + 1.0)
    r1 = R[:, g1] + (n1 - Z[:, g1]) * zero_rank                                    # + the implicit zeros of cluster_1
    tieadj = tabs["tie_adj"] + n_zero_all ** 3 - n_zero_all
    u = r1 - n1 * (n1 + 1) / 2.0
    mu = u - n1 * n2 / 2.0
    with np.errstate(invalid="ignore", divide="ignore"):
        sigma = np.sqrt(n1 * n2 * (n + 1.0 - tieadj / (n * (n - 1.0))) / 12.0)
        p = _normal_two_sided(np.abs(mu - 0.5 * np.sign(mu)) / sigma)
    p = np.where((mu == 0) & (sigma == 0), 1.0, p)
    sx1, sx2 = X[:, g1], col(X, rest)
    tested = np.flatnonzero((sx1 > 0.0) & (sx2 > 0.0))                             # :364-367
    if exact_fn is not None and tested.size:
        pe = exact_fn(tested)
        p = p.copy()
        p[tested] = np.where(np.isnan(pe), p[tested], pe)
    mean1 = np.log(S[:, g1] / n1 + 1.0)
    mean2 = np.log(col(S, rest) / n2 + 1.0)
    zero_counts = expr_cutoff < 0                                                   # an implicit zero exceeds a negative cut-off
    pct1 = (P[:, g1] + ((n1 - Z[:, g1]) if zero_counts else 0.0)) / n1
    pct2 = (col(P, rest) + ((n2 - col(Z, rest)) if zero_counts else 0.0)) / n2
    logfc = mean1 - mean2
    p_adj = np.minimum(p * tested.size, 1.0)                                        # Bonferroni over the tested genes (:384)
    sel = np.zeros(len(genes), dtype=bool)
    sel[tested] = True
    sel &= (pct1 > min_pct) & (p < p_cutoff)
    if only_pos:
        sel &= logfc > 0.0
    idx = np.flatnonzero(sel)
    idx = idx[np.argsort(-logfc[idx], kind="stable")]                               # sort!(:avg_logFC, rev=true) (:389)
    return pd.DataFrame({"gene": [genes[i] for i in idx], "p_val": p[idx], "avg_logFC": logfc[idx], "pct_1": pct1[idx],
                         "pct_2": pct2[idx], "p_val_adj": p_adj[idx]})


def _marker_inputs(sc_obj, anno, ctx):
    if sc_obj.clustData is None or sc_obj.clustData.clustering is None:
        raise ValueError("Clustering has not been done. Please complete the \"RunClustering\" first!")
    if sc_obj.normCount is None:
        raise ValueError("find_markers: normCount is missing; run normalize_object first")
    s = _session(sc_obj, ctx)
    if s.norm is None:
        s.norm = s.ctx.sparse_from_scipy(sc_obj.normCount.count_mtx)
    df = sc_obj.clustData.clustering
    pos = {c: i for i, c in enumerate(sc_obj.normCount.cell_name)}
    lab = np.empty(len(pos), dtype=object)
    lab[:] = None
    lab[[pos[c] for c in df["cell_id"]]] = df[str(anno)].to_numpy()               # extract_cluster_count (utils.jl:89-106)
    return s, lab


def find_markers(sc_obj, cluster_1=None, cluster_2=None, anno="cluster", expr_cutoff=0.0, min_pct=0.1, p_cutoff=0.05,
                 only_pos=True, ctx=None):
    if cluster_1 is None:
        raise ValueError("Please provide the name of the cell cluster for which you wish to obtain the differential genes. "
                         "The \"cluster_1\" parameter cannot be left blank.")
    if cluster_2 is None:
        raise ValueError("Please enter the name of the cell cluster you wish to compare against. The \"cluster_2\" parameter "
                         "cannot be left blank.")
    s, lab = _marker_inputs(sc_obj, anno, ctx)
    labels = np.where(lab == cluster_1, 0, np.where(lab == cluster_2, 1, -1)).astype(np.int32)
    n1, n2 = int((labels == 0).sum()), int((labels == 1).sum())
    if n1 == 0 or n2 == 0:
        raise ValueError(f"find_markers: cluster {cluster_1!r} has {n1} cells, cluster {cluster_2!r} has {n2}")
    tabs = s.ctx.marker_stats(s.norm, labels, 2, expr_cutoff)
    exact_fn = None
    if n1 + n2 <= 50:
        c1, c2 = np.flatnonzero(labels == 0), np.flatnonzero(labels == 1)
        exact_fn = lambda genes: _exact_small_pvalues(s.ctx, s.norm, c1, c2, genes)   # noqa: E731
    return _marker_frame(sc_obj.normCount.gene_name, tabs, 0, n1, n2, 1, expr_cutoff, min_pct, p_cutoff, only_pos, exact_fn)


def find_all_markers(sc_obj, anno="cluster", expr_cutoff=0.0, min_pct=0.1, p_cutoff=0.05, only_pos=True, ctx=None):
    """Every cluster (in order of first appearance) against all other cells.  Like the reference, the thresholds given
    here are NOT forwarded: find_markers runs with its defaults (processing.jl:408).  One device pass serves all
    clusters: with every cell taking part the ranks do not depend on the cluster, only the rank sums do."""
    s, lab = _marker_inputs(sc_obj, anno, ctx)
    if any(v is None for v in lab):
        raise ValueError("find_all_markers: a cell has no cluster label")
    clusters = list(pd.unique(sc_obj.clustData.clustering[str(anno)]))
    code = {c: i for i, c in enumerate(clusters)}
    labels = np.fromiter((code[v] for v in lab), dtype=np.int32, count=lab.size)
    sizes = np.bincount(labels, minlength=len(clusters))
    tabs = s.ctx.marker_stats(s.norm, labels, len(clusters), 0.0)
    frames = []
    n = int(labels.size)
    for c in clusters:
        j = code[c]
        if n <= 50:
            relab = np.where(labels == j, 0, 1).astype(np.int32)
            t2 = s.ctx.marker_stats(s.norm, relab, 2, 0.0)
            c1, c2 = np.flatnonzero(relab == 0), np.flatnonzero(relab == 1)
            f = _marker_frame(sc_obj.normCount.gene_name, t2, 0, len(c1), len(c2), 1, 0.0, 0.1, 0.05, True,
                              lambda genes: _exact_small_pvalues(s.ctx, s.norm, c1, c2, genes))
        else:
            f = _marker_frame(sc_obj.normCount.gene_name, tabs, j, int(sizes[j]), n - int(sizes[j]), None, 0.0, 0.1, 0.05, True)
        f["cluster"] = c
        frames.append(f)
    return pd.concat(frames, ignore_index=True) if frames else pd.DataFrame()
