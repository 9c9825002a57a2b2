"""CPU oracle for the counts -> normalise -> HVG -> scale -> PCA -> kNN/SNN path.

TEST INFRASTRUCTURE ONLY.  This module is a float64 NumPy/SciPy restatement of
the reference's Julia algorithm for the hot path.  It is the checker for the
CUDA library; it is never the thing shipped or measured.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.

PARITY UNPINNED.  The reference ships no tests, fixtures or golden vectors
 and Julia is not installed in the build image, so this
restatement could not be run against the reference itself.  Where the arithmetic
lives in a third-party Julia package that is not under /root/reference the
published algorithm of the pinned version is restated and marked [3P]:

  SparseArrays / Statistics stdlib 1.10.0   (sum, mean, var, sparse broadcast)
  Loess.jl 0.6.3                            (kd-tree vertices + cubic Hermite blending)
  MultivariateStats.jl 0.10.3               (PCA: centre, SVD, principalvars, tvar)
  Distances.jl 0.10.11                      (CosineDist)
  NearestNeighborDescent.jl 0.3.6           (replaced by EXACT kNN, see knn_exact)

Anchors outside this repository (tests/test_oracle.py): the Loess local regression and the Hermite blending
reproduce R's published loess values on `cars` wherever R's kd-tree and Loess.jl's share a vertex; PCA against
numpy.linalg.svd / scikit-learn; the Mann-Whitney p-values against scipy.stats; the kNN against an fp64 brute force.

All citations ``processing.jl:N`` are /root/reference/src/scrna/processing.jl,
``utils.jl:N`` are /root/reference/src/scrna/utils.jl.

Conventions: count matrices are genes x cells (``objects.jl:15-20``), held as
``scipy.sparse.csc_matrix`` so that a column is a cell, like Julia's
SparseMatrixCSC.  Indices returned by this module are 0-based.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg
import scipy.sparse as sp

__all__ = [
    "subset_matrix", "subset_count_rows", "normalize_object", "gene_moments",
    "loess_fit", "loess_predict", "find_variable_genes", "scale_object",
    "run_pca", "cosine_dist", "knn_exact", "adjacency_small", "adjacency_atlas",
    "snn_jaccard", "run_pipeline",
]


# ----------------------------------------------------------------------------
# helpers the path relies on (utils.jl)
# ----------------------------------------------------------------------------
def subset_matrix(count_mat, min_gene=0, min_cell=0):
    """utils.jl:313-320: keep genes with row sum > min_cell and cells with column
    sum > min_gene (STRICT >).  Returns (matrix, gene_kept, cell_kept)."""
    count_mat = sp.csc_matrix(count_mat)
    row_sum = np.asarray(count_mat.sum(axis=1)).ravel()
    col_sum = np.asarray(count_mat.sum(axis=0)).ravel()
    gene_kept = np.flatnonzero(row_sum > min_cell)
    cell_kept = np.flatnonzero(col_sum > min_gene)
    return count_mat[gene_kept][:, cell_kept].tocsc(), gene_kept, cell_kept


def subset_count_rows(n_genes, feature_idx):
    """utils.jl:59-76: subsetting by a boolean mask keeps the ORIGINAL gene order,
    whatever order ``genes`` was given in.  Returns the sorted unique row ids."""
    mask = np.zeros(n_genes, dtype=bool)
    mask[np.asarray(feature_idx, dtype=np.int64)] = True
    return np.flatnonzero(mask)


# ----------------------------------------------------------------------------
# a1  normalize_object (processing.jl:15-31)
# ----------------------------------------------------------------------------
def normalize_object(mtx, scale_factor=10000, norm_method="logarithm", pseudocount=1,
                     nan_to_zero=False):
    """processing.jl:15-31.

    ``inv = 1.0 ./ sum(mtx, dims=1)`` (:17-18), then
    ``log((mtx .* inv') * scale_factor + pseudocount)`` (:24) or, for "none",
    ``(mtx .* inv') * scale_factor`` (:26); anything else throws (:28).
    The result keeps the stored pattern.  A zero-sum column gives inv = Inf and
    0*Inf = NaN; the sparse broadcast then stores NaN for the WHOLE column [3P].
    ``nan_to_zero`` is the ``replace!(NaN=>0)`` of :44-46 (imaging objects).
    The NaN-column rule is only representable here on the stored pattern; a
    fully empty column has nothing stored so it stays empty (documented deviation:
    Julia would densify that column to NaN).
    """
    if norm_method not in ("logarithm", "none"):
        raise ValueError(f"Unknown normalization method: {norm_method}")
    m = sp.csc_matrix(mtx).astype(np.float64)        # :16 convert to Float64
    m.sort_indices()
    sum_val = np.asarray(m.sum(axis=0)).ravel()      # :17
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / sum_val                           # :18
        per_nz_inv = np.repeat(inv, np.diff(m.indptr))
        v = (m.data * per_nz_inv) * float(scale_factor)   # association order of :24
        if norm_method == "logarithm":
            v = np.log(v + float(pseudocount))
    if nan_to_zero:
        v = np.where(np.isnan(v), 0.0, v)
    return sp.csc_matrix((v, m.indices.copy(), m.indptr.copy()), shape=m.shape)


# ----------------------------------------------------------------------------
# a2  find_variable_genes (processing.jl:133-155) + Loess.jl 0.6.3 [3P]
# ----------------------------------------------------------------------------
def gene_moments(mtx):
    """Exact integer moments per gene, SURVEY A.3-7: S1 = sum x, S2 = sum x^2
    over all N cells (zeros included), then
        mean = S1/N ;  var = (S2 - S1*S1/N)/(N-1)          (corrected, n-1)
    in float64 with exactly this operation order (processing.jl:138-139 computes
    the same quantities with Statistics.mean/var)."""
    m = sp.csc_matrix(mtx)
    n = m.shape[1]
    d = m.data.astype(np.float64)
    s1 = np.bincount(m.indices, weights=d, minlength=m.shape[0])
    s2 = np.bincount(m.indices, weights=d * d, minlength=m.shape[0])
    mean = s1 / n
    var = (s2 - s1 * s1 / n) / (n - 1)
    return s1, s2, mean, var


# The two points of Loess.jl 0.6.3 that SURVEY A.3 could only recall, not verify (the package is not vendored in
# /root/reference and Julia is absent).  The oracle's DEFAULT is ("middle", True); the other settings exist so that
# tests can show whether the HVG set depends on them (tests/test_oracle.py, tests/test_gpu_configs.py):
#   odd_median  "middle": an odd-length node splits at its true median, the (k+1)-th of 2k+1 order statistics;
#               "lower" : at the k-th (`mid = length(perm) ÷ 2` read with 1-based indexing), left child = k points.
#   centred     True : the local quadratic is fitted in (x - v), so the intercept is the value at the vertex;
#               False: in raw x like `us[i, 1+l] = w * x^l`, value = b0 + b1 v + b2 v^2, slope = b1 + 2 b2 v
#                      (the same polynomial, different rounding).
LOESS_VARIANTS = [("middle", True), ("lower", True), ("middle", False), ("lower", False)]


def _kd_vertices(x, leaf_size_cutoff, odd_median="middle"):
    """Loess.jl 0.6.3 src/kd.jl [3P]: 1-D kd-tree.  Vertices = {min, max} plus every
    split median.  A node with <= leaf_size_cutoff points is a leaf.  Split at the
    median: even length -> mean of the two middle order statistics, left/right get
    n/2 each; odd length -> see ``odd_median`` above (default: the middle order
    statistic, left gets floor(n/2) points and right the rest, median included)."""
    xs = np.sort(np.asarray(x, dtype=np.float64), kind="stable")
    verts = {float(xs[0]), float(xs[-1])}
    stack = [(0, xs.size)]
    while stack:
        lo, hi = stack.pop()
        n = hi - lo
        if n <= leaf_size_cutoff or xs[hi - 1] - xs[lo] <= 0.0:
            continue
        mid = n // 2
        if n % 2 == 1:
            med = float(xs[lo + mid]) if odd_median == "middle" else float(xs[lo + mid - 1])
        else:
            med = float((xs[lo + mid - 1] + xs[lo + mid]) / 2)
        verts.add(med)
        stack.append((lo, lo + mid))
        stack.append((lo + mid, hi))
    return np.array(sorted(verts), dtype=np.float64)


def _pivoted_qr_solve3(us, vs):
    """Least squares ``us \\ vs`` through Householder QR with column pivoting
    (Julia: ``qr(us, ColumnNorm()) \\ vs``), written out so the C++ host code in
    the library can follow the identical operation order."""
    a = np.array(us, dtype=np.float64, copy=True)
    b = np.array(vs, dtype=np.float64, copy=True)
    q, ncol = a.shape
    perm = list(range(ncol))
    for k in range(ncol):
        norms = [float(np.dot(a[k:, j], a[k:, j])) for j in range(k, ncol)]
        p = k + int(np.argmax(norms))
        if p != k:
            a[:, [k, p]] = a[:, [p, k]]
            perm[k], perm[p] = perm[p], perm[k]
        x = a[k:, k]
        alpha = math.sqrt(float(np.dot(x, x)))
        if alpha == 0.0:
            continue
        if x[0] > 0:
            alpha = -alpha
        v = x.copy()
        v[0] -= alpha
        vnorm2 = float(np.dot(v, v))
        if vnorm2 == 0.0:
            continue
        for j in range(k, ncol):
            s = 2.0 * float(np.dot(v, a[k:, j])) / vnorm2
            a[k:, j] -= s * v
        s = 2.0 * float(np.dot(v, b[k:])) / vnorm2
        b[k:] -= s * v
    coef_p = np.zeros(ncol)
    for k in range(ncol - 1, -1, -1):
        s = b[k]
        for j in range(k + 1, ncol):
            s -= a[k, j] * coef_p[j]
        coef_p[k] = s / a[k, k] if a[k, k] != 0.0 else 0.0
    coef = np.zeros(ncol)
    for k in range(ncol):
        coef[perm[k]] = coef_p[k]
    return coef


def loess_fit(x, y, span=0.75, degree=2, cell=0.2, odd_median="middle", centred=True):
    """Loess.jl 0.6.3 ``loess(xs, ys; span, degree=2, cell=0.2)`` for one predictor [3P].

    q = max(1, floor(span*n)); kd-tree leaf cutoff = ceil(cell*span*n).  At every
    vertex v: the q nearest points by |x-v| (ties: lower index first), dmax = largest
    of those distances (1 if 0), weights tricube (1-(d/dmax)^3)^3, weighted least
    squares of y on [1, (x-v), (x-v)^2] with rows scaled by sqrt(w), solved by
    column-pivoted QR; the intercept is the fitted value at v, the linear
    coefficient its derivative.  Returns (verts, value_at_vert, slope_at_vert)."""
    if degree != 2:
        raise ValueError("oracle restates degree=2 only (the reference never changes it)")
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.size
    if n == 0:
        raise ValueError("input arrays are empty")
    if not (0 < span <= 1):
        raise ValueError("span must be in the range (0, 1].")
    q = max(1, int(math.floor(span * n)))
    cutoff = int(math.ceil(cell * span * n))
    verts = _kd_vertices(x, cutoff, odd_median)
    val = np.empty(verts.size)
    slope = np.empty(verts.size)
    for vi, v in enumerate(verts):
        val[vi], slope[vi] = loess_local_fit(x, y, v, q, centred)
    return verts, val, slope


def loess_local_fit(x, y, v, q, centred=True):
    """The local regression of loess_fit at ONE point v: (fitted value, derivative) from the q nearest points (ties:
    lower index first), tricube weights, degree 2.  Separate so that it can be checked on its own against published
    values of R's loess (same local fit, different kd-tree: tests/test_oracle.py)."""
    d = np.abs(x - v)
    order = np.argsort(d, kind="stable")[:q]
    dq = d[order]
    dmax = float(dq.max())
    if dmax == 0.0:
        dmax = 1.0
    u = dq / dmax
    w = np.sqrt((1.0 - u ** 3) ** 3)
    xc = x[order] - v if centred else x[order]
    us = np.stack([w, w * xc, w * xc * xc], axis=1)
    vs = y[order] * w
    coef = _pivoted_qr_solve3(us, vs)
    if centred:
        return coef[0], coef[1]
    return coef[0] + coef[1] * v + coef[2] * v * v, coef[1] + 2.0 * coef[2] * v


def loess_predict(model, z):
    """Loess.jl ``predict`` [3P]: at a vertex return its fitted value; otherwise the
    cubic Hermite interpolant through the two enclosing vertices."""
    verts, val, slope = model
    z = np.asarray(z, dtype=np.float64)
    out = np.empty(z.shape, dtype=np.float64)
    hi = np.searchsorted(verts, z, side="left")      # verts[hi-1] < z <= verts[hi]
    if np.any(z < verts[0]) or np.any(z > verts[-1]):
        raise ValueError("Loess cannot extrapolate outside the fitted range")
    hi = np.clip(hi, 1, verts.size - 1)
    lo = hi - 1
    v1, v2 = verts[lo], verts[hi]
    y1, y2 = val[lo], val[hi]
    d1, d2 = slope[lo], slope[hi]
    h = v2 - v1
    t = (z - v1) / h
    h00 = (1 + 2 * t) * (1 - t) ** 2
    h10 = t * (1 - t) ** 2
    h01 = t * t * (3 - 2 * t)
    h11 = t * t * (t - 1)
    out = h00 * y1 + h10 * h * d1 + h01 * y2 + h11 * h * d2
    exact_lo = z == v1
    exact_hi = z == v2
    out = np.where(exact_lo, y1, out)
    out = np.where(exact_hi, y2, out)
    return out


def hvg_from_moments(s1, s2, n_cells, nFeatures=2000, span=0.3, odd_median="middle", centred=True):
    """processing.jl:140-155 from the exact integer moments (S1 = sum x, S2 = sum x^2 per gene): everything of
    find_variable_genes after :138-139.  Separate so that the Loess variants can be run on the moment tables of the
    full-size benchmark matrices (33k points) without the matrices themselves."""
    s1 = np.asarray(s1, dtype=np.float64)
    s2 = np.asarray(s2, dtype=np.float64)
    n_genes = s1.size
    n = float(n_cells)
    nFeatures = min(int(nFeatures), n_genes)               # :134-137
    mean = s1 / n
    var = (s2 - s1 * s1 / n) / (n - 1)
    if np.any(~(var > 0.0)):
        raise ValueError("find_variable_genes: a gene has variance <= 0 "
                         "(reference precondition, processing.jl:143,151)")
    lx = np.log10(mean)
    ly = np.log10(var)
    model = loess_fit(lx, ly, span=span, odd_median=odd_median, centred=centred)
    var_exp = 10.0 ** loess_predict(model, lx)
    var_std = var / var_exp
    # sort(..., rev=true) is stable: ties keep ascending gene order (:152)
    order = np.lexsort((np.arange(n_genes), -var_std))
    return {
        "mean": mean, "variance": var, "variance_expected": var_exp,
        "variance_standardized": var_std, "order": order,
        "features": order[:nFeatures].copy(), "S1": s1, "S2": s2,
        "loess": model,
    }


def find_variable_genes(mtx, nFeatures=2000, span=0.3, odd_median="middle", centred=True):
    """processing.jl:133-155.

    mean/var of RAW counts per gene over all cells (:138-139, exact-moment rule of
    SURVEY A.3-7), loess of log10(var) on log10(mean) with span (:144),
    variance_expected = 10^predict (:145), variance_standardized =
    var((x - mean)/sqrt(variance_expected)) (:146-150) = var/variance_expected,
    stable descending sort (:152), first nFeatures (:153).
    The reference requires every gene to have variance > 0 (:143,:151)."""
    m = sp.csc_matrix(mtx)
    s1, s2, _, _ = gene_moments(m)
    return hvg_from_moments(s1, s2, m.shape[1], nFeatures, span, odd_median, centred)


# ----------------------------------------------------------------------------
# a3  scale_object (processing.jl:65-76)
# ----------------------------------------------------------------------------
def scale_object(count_mtx, scale_max=10.0, do_scale=True, do_center=True):
    """processing.jl:65-76 on a genes x cells matrix of NORMALISED values.

    rmean = mean(dims=2) (:66); rsd = sqrt.(var(dims=2)) corrected (:67);
    ``.-= rmean`` (:69); ``./ rsd`` (:72); ``min.(x, scale_max)`` (:74), an UPPER
    clip only.  Returns a dense float64 genes x cells array."""
    if sp.issparse(count_mtx):
        x = np.asarray(count_mtx.todense(), dtype=np.float64)
    else:
        x = np.array(count_mtx, dtype=np.float64, copy=True)
    rmean = x.mean(axis=1, keepdims=True)
    rsd = np.sqrt(x.var(axis=1, ddof=1, keepdims=True))
    if do_center:
        x = x - rmean
    with np.errstate(divide="ignore", invalid="ignore"):
        if do_scale:
            x = x / rsd
        x = np.where(np.isnan(x), x, np.minimum(x, scale_max))
    return x


# ----------------------------------------------------------------------------
# a4  run_pca (processing.jl:164-193; MultivariateStats 0.10.3 [3P])
# ----------------------------------------------------------------------------
def run_pca(scaled, maxoutdim=10, method="svd"):
    """processing.jl:164-193 with the MultivariateStats ``fit(PCA, X; method=:svd,
    pratio=1, maxoutdim)`` semantics [3P].  ``scaled`` is genes(H) x cells(N).

    m = mean over cells; Zc = Z - m; thin SVD Zc = U S V'; principalvars = s^2/(N-1);
    tvar = sum over ALL components; keep the first maxoutdim; embedding = Zc' * U_k
    (N x k).  ``method="gram"`` diagonalises Zc*Zc' instead (same result, used on
    the wide matrices where a full SVD would take minutes)."""
    z = np.asarray(scaled, dtype=np.float64)
    h, n = z.shape
    mean = z.mean(axis=1, keepdims=True)
    zc = z - mean
    k = min(int(maxoutdim), min(h, n))
    if method == "svd":
        u, s, _ = np.linalg.svd(zc, full_matrices=False)
        lam = s * s
    else:
        g = zc @ zc.T
        lam, u = scipy.linalg.eigh(g)
        lam = np.maximum(lam[::-1], 0.0)
        u = u[:, ::-1]
    principalvars = lam / (n - 1)
    tvar = float(principalvars.sum())
    uk = u[:, :k]
    emb = zc.T @ uk
    return {
        "embedding": emb, "principalvars": principalvars[:k].copy(), "tvar": tvar,
        "loadings": uk, "mean": mean.ravel(), "all_vars": principalvars,
    }


# ----------------------------------------------------------------------------
# a5/a6  kNN and graphs (processing.jl:195-232, 263-278)
# ----------------------------------------------------------------------------
def cosine_dist(a, b):
    """Distances.jl 0.10.11 CosineDist [3P]: max(1 - <a,b>/(|a||b|), 0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return max(1.0 - float(a @ b) / (math.sqrt(float(a @ a)) * math.sqrt(float(b @ b))), 0.0)


def knn_exact(emb, k, metric="cosine", block=2048, queries=None):
    """EXACT brute-force kNN in float64 (the reference uses the approximate, unseeded
    NN-descent at processing.jl:200,265; exact kNN is what the build computes, SURVEY
    finding 5).  Self excluded by index; ascending distance; ties by lower index.
    Returns (idx [n,k] int64 0-based, dist [n,k] float64).  ``queries`` restricts the
    query rows (for spot checks on large inputs)."""
    x = np.asarray(emb, dtype=np.float64)
    n = x.shape[0]
    k = int(k)
    if k > n - 1:
        raise ValueError("k must be < number of cells")
    if metric == "cosine":
        nrm = np.sqrt((x * x).sum(axis=1))
        xn = x / nrm[:, None]
    elif metric == "euclidean":
        xn = x
        sq = (x * x).sum(axis=1)
    else:
        raise ValueError(f"unknown metric {metric}")
    qrows = np.arange(n) if queries is None else np.asarray(queries, dtype=np.int64)
    idx = np.empty((qrows.size, k), dtype=np.int64)
    dist = np.empty((qrows.size, k), dtype=np.float64)
    for s in range(0, qrows.size, block):
        rows = qrows[s:s + block]
        if metric == "cosine":
            d = np.maximum(1.0 - xn[rows] @ xn.T, 0.0)
        else:
            d = np.sqrt(np.maximum(sq[rows, None] + sq[None, :] - 2.0 * (xn[rows] @ xn.T), 0.0))
        d[np.arange(rows.size), rows] = np.inf
        # exact ordering incl. ties by index: take every candidate <= the k-th smallest
        kth = np.partition(d, k - 1, axis=1)[:, k - 1]
        for r in range(rows.size):
            cand = np.flatnonzero(d[r] <= kth[r])
            o = np.lexsort((cand, d[r, cand]))[:k]
            idx[s + r] = cand[o]
            dist[s + r] = d[r, cand[o]]
    return idx, dist


def adjacency_small(idx, n=None):
    """processing.jl:268-278: ``adj[idx[j,i], i] = 1; adj[i, idx[j,i]] = 1``: the
    binary union-symmetrised kNN adjacency (Int64).  Returned sparse (CSC); the
    reference holds it dense (n < 10000 only)."""
    idx = np.asarray(idx, dtype=np.int64)
    n = idx.shape[0] if n is None else n
    k = idx.shape[1]
    cols = np.repeat(np.arange(idx.shape[0]), k)
    rows = idx.ravel()
    a = sp.coo_matrix((np.ones(rows.size * 2, dtype=np.int64),
                       (np.concatenate([rows, cols]), np.concatenate([cols, rows]))),
                      shape=(n, n)).tocsc()
    a.data[:] = 1
    a.sort_indices()
    return a


def adjacency_atlas(idx, n=None):
    """processing.jl:218-232: COO push of (idx,i,1) and (i,idx,1); ``sparse`` SUMS
    duplicates, so a mutual pair carries 2 and a one-directional pair 1."""
    idx = np.asarray(idx, dtype=np.int64)
    n = idx.shape[0] if n is None else n
    k = idx.shape[1]
    cols = np.repeat(np.arange(idx.shape[0]), k)
    rows = idx.ravel()
    a = sp.coo_matrix((np.ones(rows.size * 2, dtype=np.int64),
                       (np.concatenate([rows, cols]), np.concatenate([cols, rows]))),
                      shape=(n, n)).tocsc()          # duplicates are summed
    a.sort_indices()
    return a


def snn_jaccard(idx, prune=0.0):
    """SNN extension (not in the reference).  N+(i) = {i} U kNN(i).  For
    every edge (i,j) of the union-symmetrised kNN graph
        w_ij = |N+(i) & N+(j)| / |N+(i) U N+(j)| ;
    edges with w < prune are dropped.  Symmetric CSC with float64 weights."""
    idx = np.asarray(idx, dtype=np.int64)
    n, k = idx.shape
    pattern = adjacency_small(idx)
    plus = np.concatenate([np.arange(n)[:, None], idx], axis=1)
    coo = pattern.tocoo()
    inter = _intersections(plus, coo.row, coo.col)
    union = 2 * (k + 1) - inter          # kNN lists hold k distinct cells, none of them i
    w = inter.astype(np.float64) / union.astype(np.float64)
    keep = w >= prune
    out = sp.coo_matrix((w[keep], (coo.row[keep], coo.col[keep])), shape=(n, n)).tocsc()
    out.sort_indices()
    return out


def _intersections(plus, rows, cols, chunk=20000):
    out = np.empty(rows.size, dtype=np.int64)
    for s in range(0, rows.size, chunk):
        a = plus[rows[s:s + chunk]]
        b = plus[cols[s:s + chunk]]
        out[s:s + chunk] = (a[:, :, None] == b[:, None, :]).sum(axis=(1, 2))
    return out


# ----------------------------------------------------------------------------
# the whole path, in the atlas order of docs/scRNA_tutorial/README.md:260-280
# ----------------------------------------------------------------------------
def run_pipeline(counts, nFeatures=2000, span=0.3, scale_max=10.0, n_pcs=50, k=20,
                 scale_factor=10000, pca_method="svd", metric="cosine", graph="atlas"):
    """normalize_object -> find_variable_genes -> scale_object(features=HVG) ->
    run_pca(maxoutdim=n_pcs) -> exact kNN -> adjacency (+ SNN)."""
    counts = sp.csc_matrix(counts)
    norm = normalize_object(counts, scale_factor=scale_factor)
    hvg = find_variable_genes(counts, nFeatures=nFeatures, span=span)
    rows = subset_count_rows(counts.shape[0], hvg["features"])
    scaled = scale_object(norm.tocsr()[rows].tocsc(), scale_max=scale_max)
    pca = run_pca(scaled, maxoutdim=n_pcs, method=pca_method)
    idx, dist = knn_exact(pca["embedding"], k, metric=metric)
    adj = adjacency_atlas(idx) if graph == "atlas" else adjacency_small(idx)
    return {"norm": norm, "hvg": hvg, "rows": rows, "scaled": scaled, "pca": pca,
            "knn_idx": idx, "knn_dist": dist, "adj": adj}


# ----------------------------------------------------------------------------
# f3  find_markers / find_all_markers (processing.jl:350-414)
# ----------------------------------------------------------------------------
def _mwu_exact_pvalue(x, y):
    """HypothesisTests.ExactMannWhitneyUTest [3P]: the permutation distribution of U (two-sided: twice the smaller
    tail, capped at 1).  Only reached for nx + ny <= 10, or <= 50 without ties (HypothesisTests' MannWhitneyUTest
    dispatch); enumerates the C(n, nx) assignments for n <= 10 and uses the tie-free counting recursion otherwise."""
    import itertools
    import scipy.stats
    nx, ny = len(x), len(y)
    allv = np.concatenate([x, y])
    ranks = scipy.stats.rankdata(allv)
    u = ranks[:nx].sum() - nx * (nx + 1) / 2
    if nx + ny <= 10:
        us = np.array([ranks[list(c)].sum() - nx * (nx + 1) / 2 for c in itertools.combinations(range(nx + ny), nx)])
        mid = nx * ny / 2
        p = np.mean(us <= u) if u < mid else np.mean(us >= u)
        return float(min(1.0, 2 * p))
    # no ties: number of arrangements with statistic u (Mann & Whitney's recursion)
    cnt = np.zeros((nx + 1, ny + 1, nx * ny + 1))
    cnt[0, :, 0] = 1
    cnt[:, 0, 0] = 1
    for i in range(1, nx + 1):
        for j in range(1, ny + 1):
            for k in range(i * j + 1):
                cnt[i, j, k] = (cnt[i - 1, j, k - j] if k >= j else 0) + cnt[i, j - 1, k]
    pmf = cnt[nx, ny] / cnt[nx, ny].sum()
    ui = int(round(u))
    p = pmf[:ui + 1].sum() if u < nx * ny / 2 else pmf[ui:].sum()
    return float(min(1.0, 2 * p))


def mann_whitney_pvalue(x, y):
    """HypothesisTests.MannWhitneyUTest(x, y) -> pvalue (two-sided) [3P, HypothesisTests 0.11]:
    exact for nx + ny <= 10 or (nx + ny <= 50 and no ties), else the normal approximation with tie and continuity
    correction:  U = sum ranks(x) - nx (nx + 1) / 2;  mu = U - nx ny / 2;
    sigma = sqrt(nx ny (n + 1 - tieadj / (n (n - 1))) / 12),  tieadj = sum (t^3 - t);
    p = 2 ccdf(Normal, |mu - 0.5 sign(mu)| / sigma), and 1 when mu == sigma == 0."""
    import scipy.stats
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    nx, ny = x.size, y.size
    n = nx + ny
    ranks = scipy.stats.rankdata(np.concatenate([x, y]))
    _, t = np.unique(np.concatenate([x, y]), return_counts=True)
    tieadj = float(((t.astype(np.float64) ** 3) - t).sum())
    if n <= 10 or (n <= 50 and tieadj == 0):
        return _mwu_exact_pvalue(x, y)
    u = ranks[:nx].sum() - nx * (nx + 1) / 2
    mu = u - nx * ny / 2
    sigma = math.sqrt(nx * ny * (n + 1 - tieadj / (n * (n - 1))) / 12)
    if mu == 0 and sigma == 0:
        return 1.0
    z = abs(mu - 0.5 * np.sign(mu)) / sigma
    return float(2 * scipy.stats.norm.sf(z))


def find_markers(norm, cluster_of_cell, cluster_1, cluster_2, gene_name=None, expr_cutoff=0.0, min_pct=0.1,
                 p_cutoff=0.05, only_pos=True):
    """processing.jl:350-391 on a genes x cells NORMALISED matrix and a per-cell cluster label.

    genes with rowSum > 0 in BOTH clusters (:364-367); per gene MannWhitneyUTest(x, y) on the dense rows (:376-377);
    mean_i = log(mean(expm1(x)) + 1) (:378-379); avg_logFC = mean1 - mean2; pct_i = #(x > expr_cutoff)/n_i (:381-382);
    Bonferroni over the tested genes (:384); keep pct_1 > min_pct and p_val < p_cutoff (:385), avg_logFC > 0 if
    only_pos (:386-388); sorted by avg_logFC, descending (:389).  Returns a dict of columns."""
    m = sp.csr_matrix(norm)
    lab = np.asarray(cluster_of_cell)
    c1 = np.flatnonzero(lab == cluster_1)
    c2 = np.flatnonzero(lab == cluster_2)
    a = m[:, c1].toarray()
    b = m[:, c2].toarray()
    keep = np.flatnonzero((a.sum(axis=1) > 0.0) & (b.sum(axis=1) > 0.0))
    a, b = a[keep], b[keep]
    p = np.array([mann_whitney_pvalue(x, y) for x, y in zip(a, b)])
    mean1 = np.log(np.expm1(a).mean(axis=1) + 1)
    mean2 = np.log(np.expm1(b).mean(axis=1) + 1)
    logfc = mean1 - mean2
    pct1 = (a > expr_cutoff).mean(axis=1)
    pct2 = (b > expr_cutoff).mean(axis=1)
    p_adj = np.minimum(p * keep.size, 1.0)
    sel = (pct1 > min_pct) & (p < p_cutoff)
    if only_pos:
        sel &= logfc > 0.0
    idx = np.flatnonzero(sel)
    idx = idx[np.argsort(-logfc[idx], kind="stable")]
    genes = keep if gene_name is None else np.asarray(gene_name, dtype=object)[keep]
    return {"gene": genes[idx], "p_val": p[idx], "avg_logFC": logfc[idx], "pct_1": pct1[idx], "pct_2": pct2[idx],
            "p_val_adj": p_adj[idx], "tested": keep, "p_all": p, "logfc_all": logfc}


def find_all_markers(norm, cluster_of_cell, gene_name=None):
    """processing.jl:393-414: every cluster, in order of first appearance, against all other cells ("nonself");
    find_markers is called with its DEFAULT thresholds (the arguments of find_all_markers are not forwarded, :408)."""
    lab = np.asarray(cluster_of_cell)
    _, first = np.unique(lab, return_index=True)
    out = []
    for c in lab[np.sort(first)]:
        relab = np.where(lab == c, c, "nonself")
        r = find_markers(norm, relab, c, "nonself", gene_name)
        r["cluster"] = np.full(r["gene"].size, c, dtype=object)
        out.append(r)
    return out


# ----------------------------------------------------------------------------
# f4  ingest: merge_matrices (int_utils.jl:1-23), run_tf_idf (atac_processing.jl:1-13), check_duplicates (utils.jl:31-46)
# ----------------------------------------------------------------------------
def merge_matrices(matrices, row_names):
    """int_utils.jl:1-23: the matrices side by side on the SORTED union of their row names; every non-zero value
    (`if val != 0`, :11) is pushed as Float64; rows a matrix does not have stay empty in its columns."""
    all_rows = sorted(set(n for names in row_names for n in names))
    pos = {n: i for i, n in enumerate(all_rows)}
    blocks = []
    for m, names in zip(matrices, row_names):
        m = sp.coo_matrix(m)
        keep = m.data != 0
        rows = np.array([pos[names[r]] for r in m.row[keep]], dtype=np.int64)
        blocks.append(sp.coo_matrix((m.data[keep].astype(np.float64), (rows, m.col[keep])), shape=(len(all_rows), m.shape[1])))
    out = sp.hstack(blocks).tocsc()
    out.sort_indices()
    return out, all_rows


def run_tf_idf(mtx, scale_factor=10000):
    """atac_processing.jl:1-13: tf = x / colsum (:3-4); idf = n_cells / rowsum (:7-8); log1p(idf * tf * scale_factor)
    (:11-12), on the stored pattern."""
    m = sp.csc_matrix(mtx).astype(np.float64)
    m.sort_indices()
    npeaks = np.asarray(m.sum(axis=0)).ravel()
    rsums = np.asarray(m.sum(axis=1)).ravel()
    with np.errstate(divide="ignore", invalid="ignore"):
        tf = m.data * np.repeat(1.0 / npeaks, np.diff(m.indptr))
        idf = m.shape[1] / rsums
        v = np.log1p((idf[m.indices] * tf) * float(scale_factor))
    return sp.csc_matrix((v, m.indices.copy(), m.indptr.copy()), shape=m.shape)


def check_duplicates(arr):
    """utils.jl:31-46: (names in order of first appearance, {name: position of its SECOND occurrence})."""
    seen, uniq, removed = set(), [], {}
    for i, v in enumerate(arr):
        if v not in seen:
            seen.add(v)
            uniq.append(v)
        elif v not in removed:
            removed[v] = i
    return uniq, removed


# ----------------------------------------------------------------------------
# f2  run_harmony (int_processing.jl:1-12; HarmonyObject int_objects.jl:115-236; int_utils.jl:49-199)
# ----------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def hash_order(seed, rnd, n):
    """The update order that replaces Julia's randperm (not reproducible): stable argsort of a counter-based hash of
    (seed, round, cell) -- the same three splitmix64 steps as the library (harmony.cu: k_hm_perm_keys)."""
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(seed) ^ np.uint64(0xD1B54A32D192ED03))
        h = _splitmix64(h ^ np.uint64(rnd))
        h = _splitmix64(h ^ ((np.arange(n, dtype=np.uint64) * np.uint64(0x2545F4914F6CDD1D)) & _M64))
    return np.argsort(h, kind="stable")


def harmony_zcos(z):
    """int_objects.jl:201-203: every cell divided by its LARGEST coordinate, then scaled to unit L2 length."""
    z = np.asarray(z, dtype=np.float64)
    v = z / z.max(axis=1, keepdims=True)
    return v / np.sqrt((v * v).sum(axis=1, keepdims=True))


def kmeans_lloyd(x, k, init_idx, max_iter=100, tol=1e-4):
    """cluster_kmeans (int_utils.jl:49-56) restated as Lloyd's iteration from the rows ``init_idx`` (Clustering.jl's
    k-means++ seeding depends on Julia's RNG): stop when the objective changes by less than ``tol`` or after
    ``max_iter`` assignments; an empty cluster keeps its centre; ties go to the lower cluster id."""
    x = np.asarray(x, dtype=np.float64)
    cen = x[np.asarray(init_idx)].copy()
    prev = np.inf
    lab = None
    for it in range(1, max(1, max_iter) + 1):
        d2 = ((x[:, None, :] - cen[None, :, :]) ** 2).sum(axis=2) if x.shape[0] * k <= 2_000_000 else \
            (x * x).sum(1)[:, None] - 2 * x @ cen.T + (cen * cen).sum(1)[None, :]
        lab = d2.argmin(axis=1)
        obj = float(d2[np.arange(x.shape[0]), lab].sum())
        if abs(obj - prev) < tol:
            break
        prev = obj
        for c in range(k):
            m = lab == c
            if m.any():
                cen[c] = x[m].mean(axis=0)
    return cen, lab, obj, it


def run_harmony(z_orig, batch, y0, theta=None, lamb=None, sigma=0.1, block_size=0.05, max_iter_harmony=10,
                max_iter_kmeans=20, epsilon_cluster=1e-5, epsilon_harmony=1e-4, seed=123):
    """HarmonyObject(...) + harmonize! on a cells x d matrix (the reference's data_mat is its transpose) and an integer
    batch label per cell (level index in sorted order).  ``y0`` [K x d]: the k-means centres of harmony_zcos(z_orig).
    Follows the reference line by line, INCLUDING its departures from harmonypy: optimized_norm is the L1 norm
    (int_utils.jl:32-41), the convergence window of cluster! lags one iteration (:160-166), Z_cos after a correction is
    Z_corr with dimension j divided by the L1 norm of the old Z_cos of cell j (:197), the tail of the update order beyond
    n_blocks * floor(N / n_blocks) is skipped (:141).  Returns dict(Z_corr, R, objective_harmony, kmeans_rounds)."""
    zo = np.asarray(z_orig, dtype=np.float64).T.copy()           # d x N
    d, n = zo.shape
    batch = np.asarray(batch)
    nb = int(batch.max()) + 1
    phi = np.zeros((nb, n))
    phi[batch, np.arange(n)] = 1.0
    k = y0.shape[0]
    theta = np.full(nb, 2.0 if theta is None else theta, dtype=np.float64) if np.ndim(theta) == 0 else np.asarray(theta, float)
    lamb = np.full(nb, 1.0 if lamb is None else lamb, dtype=np.float64) if np.ndim(lamb) == 0 else np.asarray(lamb, float)
    sigma = np.full(k, sigma, dtype=np.float64) if np.ndim(sigma) == 0 else np.asarray(sigma, float)
    pr_b = phi.sum(axis=1) / n
    lamb_mat = np.diag(np.concatenate([[0.0], lamb]))
    phi_moe = np.vstack([np.ones((1, n)), phi])
    z_cos = harmony_zcos(zo.T).T
    z_corr = zo.copy()

    def l1(m, cols):
        return np.abs(m[:, cols]).sum(axis=0)

    # init_cluster! (:58-71)
    y = np.asarray(y0, dtype=np.float64).T.copy()                # d x K
    y = y / l1(y, np.arange(k))[None, :]
    dist = 2 * (1 - y.T @ z_cos)
    r = -dist / sigma[:, None]
    r -= r.max(axis=0, keepdims=True)
    r = np.exp(r)
    r /= r.sum(axis=0, keepdims=True)
    e = np.outer(r.sum(axis=1), pr_b)
    o = r @ phi.T
    obj_k, obj_h, rounds = [], [], []

    def objective():
        with np.errstate(divide="ignore", invalid="ignore"):
            ent = r * np.log(r)
        ent = np.where(np.isfinite(ent), ent, 0.0)
        w = (theta[None, :] * np.log((o + 1) / (e + 1))) @ phi
        obj_k.append(float((r * dist).sum() + (ent * sigma[:, None]).sum() + ((r * sigma[:, None]) * w).sum()))
    objective()
    obj_h.append(obj_k[-1])
    rnd = 0
    n_blocks = int(math.ceil(1 / block_size))
    per = n // n_blocks
    window = 3
    for _ in range(max_iter_harmony):
        j = 0
        for i in range(1, max_iter_kmeans + 1):
            y = z_cos @ r.T
            y = y / l1(y, np.arange(k))[None, :]
            dist = 2 * (1 - y.T @ z_cos)
            sd = -dist / sigma[:, None]
            sd -= sd.max(axis=0, keepdims=True)
            sd = np.exp(sd)
            order = hash_order(seed, rnd, n)
            rnd += 1
            for b in range(n_blocks):
                cells = order[per * b:min(per * (b + 1), n)]
                if cells.size == 0:
                    continue
                e -= np.outer(r[:, cells].sum(axis=1), pr_b)
                o -= r[:, cells] @ phi[:, cells].T
                rb = sd[:, cells] * (((e + 1) / (o + 1)) ** theta[None, :] @ phi[:, cells])
                rb = rb / np.abs(rb).sum(axis=0, keepdims=True)
                r[:, cells] = rb
                e += np.outer(rb.sum(axis=1), pr_b)
                o += rb @ phi[:, cells].T
            objective()
            if i > window:
                okl = len(obj_k)
                o_old = sum(obj_k[okl - 3 - w] for w in range(1, window))
                o_new = sum(obj_k[okl - 2 - w] for w in range(1, window))
                if abs(o_old - o_new) / abs(o_old) < epsilon_cluster:
                    break
            j = i
        rounds.append(j)
        obj_h.append(obj_k[-1])
        # moe_correct_ridge (:188-199)
        z_corr = zo.copy()
        for c in range(k):
            phi_rk = phi_moe * r[c][None, :]
            x = phi_rk @ phi_moe.T + lamb_mat
            w = np.linalg.inv(x) @ phi_rk @ zo.T
            w[0, :] = 0.0
            z_corr -= w.T @ phi_rk
        z_cos = z_corr / l1(z_cos, np.arange(d))[:, None]
        if (obj_h[-2] - obj_h[-1]) / abs(obj_h[-2]) < epsilon_harmony:
            break
    return {"Z_corr": z_corr.T.copy(), "R": r.T.copy(), "objective_harmony": np.array(obj_h), "kmeans_rounds": np.array(rounds)}


# ----------------------------------------------------------------------------
# f1  Leiden (processing.jl:233-234, 279-280: Leiden.leiden(adj_mat, resolution = res)) -- Traag, Waltman & van Eck 2019,
#     Constant Potts Model quality (what Leiden.jl optimises):  H = sum_c [ e_c - gamma * n_c (n_c - 1) / 2 ]
#     with e_c the total weight of the edges inside c and n_c its number of cells.
# ----------------------------------------------------------------------------
def cpm_quality(adj, labels, gamma):
    """CPM quality of a partition of a symmetric weighted graph (every undirected edge stored twice)."""
    a = sp.coo_matrix(adj)
    labels = np.asarray(labels)
    inside = labels[a.row] == labels[a.col]
    off = a.row != a.col
    e_in = a.data[inside & off].sum() / 2.0 + a.data[inside & ~off].sum()
    _, n_c = np.unique(labels, return_counts=True)
    return float(e_in - gamma * (n_c * (n_c - 1) / 2.0).sum())


def communities_connected(adj, labels):
    """True when every community induces a connected subgraph (the guarantee Leiden adds to Louvain)."""
    import scipy.sparse.csgraph as csg
    a = sp.csr_matrix(adj)
    labels = np.asarray(labels)
    coo = a.tocoo()
    keep = labels[coo.row] == labels[coo.col]
    sub = sp.csr_matrix((np.ones(keep.sum()), (coo.row[keep], coo.col[keep])), shape=a.shape)
    n_comp, comp = csg.connected_components(sub, directed=False)
    return n_comp == np.unique(labels).size


def leiden_cpm(adj, gamma, seed=1234, theta=0.01, max_levels=50):
    """Sequential Leiden with the CPM quality function: fast local moving (queue), refinement (random merges of
    well-connected singletons, temperature ``theta``), aggregation on the refined partition with the non-refined one as
    the initial assignment; repeated until nothing moves.  Returns (labels per node, quality)."""
    rng = np.random.default_rng(seed)
    a = sp.csr_matrix(adj).astype(np.float64)
    a.setdiag(0)
    a.eliminate_zeros()
    n0 = a.shape[0]
    node_of = np.arange(n0)                      # original node -> current aggregate node
    size = np.ones(n0)
    part = np.arange(n0)                          # aggregate node -> community
    g = a
    for _ in range(max_levels):
        n = g.shape[0]
        indptr, indices, data = g.indptr, g.indices, g.data
        csize = np.bincount(part, weights=size, minlength=n).astype(np.float64)
        # ---- fast local moving ----
        queue = list(rng.permutation(n))
        inq = np.ones(n, dtype=bool)
        moved_any = False
        while queue:
            v = queue.pop(0)
            inq[v] = False
            nb, w = indices[indptr[v]:indptr[v + 1]], data[indptr[v]:indptr[v + 1]]
            own = part[v]
            kvc = {}
            for u, ww in zip(nb, w):
                if u != v:
                    kvc[part[u]] = kvc.get(part[u], 0.0) + ww
            csize[own] -= size[v]
            best, best_gain = own, kvc.get(own, 0.0) - gamma * size[v] * csize[own]
            for c, k in kvc.items():
                gain = k - gamma * size[v] * csize[c]
                if gain > best_gain + 1e-12:
                    best, best_gain = c, gain
            if 0.0 > best_gain + 1e-12:              # an empty community
                empty = np.flatnonzero(csize == 0)
                if empty.size:
                    best, best_gain = int(empty[0]), 0.0
            csize[best] += size[v]
            if best != own:
                part[v] = best
                moved_any = True
                for u in nb:
                    if part[u] != best and not inq[u]:
                        queue.append(u)
                        inq[u] = True
        # ---- refinement ----
        ref = np.arange(n)
        rsize = size.copy()
        for c in np.unique(part):
            members = np.flatnonzero(part == c)
            if members.size == 1:
                continue
            nc = size[members].sum()
            in_c = np.zeros(n, dtype=bool)
            in_c[members] = True
            # weight from v to the rest of its community
            kin = np.array([data[indptr[v]:indptr[v + 1]][in_c[indices[indptr[v]:indptr[v + 1]]] & (indices[indptr[v]:indptr[v + 1]] != v)].sum() for v in members])
            well = kin >= gamma * size[members] * (nc - size[members])
            ext = {}                                  # refined community -> weight to the rest of c
            for v, k in zip(members, kin):
                ext[v] = k
            singleton = {v: True for v in members}
            for v in rng.permutation(members[well]):
                if not singleton[v]:
                    continue
                nb, w = indices[indptr[v]:indptr[v + 1]], data[indptr[v]:indptr[v + 1]]
                kvr = {}
                for u, ww in zip(nb, w):
                    if u != v and in_c[u]:
                        kvr[ref[u]] = kvr.get(ref[u], 0.0) + ww
                cands, gains = [], []
                for r, k in kvr.items():
                    if r == ref[v]:
                        continue
                    if ext[r] >= gamma * rsize[r] * (nc - rsize[r]):        # the target is well connected within c
                        gain = k - gamma * size[v] * rsize[r]
                        if gain >= 0:
                            cands.append(r)
                            gains.append(gain)
                if not cands:
                    continue
                p = np.exp((np.array(gains) - max(gains)) / theta)
                r = cands[int(rng.choice(len(cands), p=p / p.sum()))]
                # merge v into r
                k_vr = kvr[r]
                ext[r] = ext[r] + ext[v] - 2 * k_vr
                rsize[r] += size[v]
                singleton[r] = False if r in singleton else False
                for u in members:
                    pass
                ref[v] = r
                singleton[v] = False
                # the members of r are no longer singletons
                for u in np.flatnonzero(ref == r):
                    singleton[u] = False
        # ---- aggregate on the refined partition ----
        uniq, new_id = np.unique(ref, return_inverse=True)
        if uniq.size == n and not moved_any:
            break
        m = uniq.size
        coo = g.tocoo()
        ng = sp.coo_matrix((coo.data, (new_id[coo.row], new_id[coo.col])), shape=(m, m)).tocsr()
        ng.sum_duplicates()
        new_size = np.bincount(new_id, weights=size, minlength=m)
        new_part = np.zeros(m, dtype=np.int64)
        new_part[new_id] = part
        _, new_part = np.unique(new_part, return_inverse=True)
        node_of = new_id[node_of]
        g, size, part = ng, new_size, new_part
        if uniq.size == n:
            break
    labels = part[node_of]
    _, labels = np.unique(labels, return_inverse=True)
    return labels, cpm_quality(adj, labels, gamma)
