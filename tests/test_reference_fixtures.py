"""Parity against the REAL reference, for the day somebody has Julia: consumes the fixtures that
julia/make_fixtures.jl (the unmodified CellScopes.jl functions) wrote into tests/golden/ref_out/ for the matrix of
tests/golden/make_ref_inputs.py.  Without the fixtures the tests are skipped with that instruction -- the oracle stays
"parity unpinned" until then.  CPU part: the oracle against the reference.  GPU part: the CUDA
path against the reference, with the tolerances north_star states."""
import json
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
REF_IN = os.path.join(HERE, "golden", "ref_in")
REF_OUT = os.path.join(HERE, "golden", "ref_out")
needs_fixtures = pytest.mark.skipif(
    not os.path.exists(os.path.join(REF_OUT, "manifest.json")),
    reason="no reference fixtures: run `python tests/golden/make_ref_inputs.py` and `julia julia/make_fixtures.jl` "
           "(needs Julia 1.10 + CellScopes.jl; neither is in the build image)")


def load(name):
    man = json.load(open(os.path.join(REF_OUT, "manifest.json")))
    shape = tuple(reversed(man[name]))            # Julia column-major dims -> C order
    for ext, dt in ((".f64", np.float64), (".i64", np.int64)):
        p = os.path.join(REF_OUT, name + ext)
        if os.path.exists(p):
            return np.fromfile(p, dtype=dt).reshape(shape)
    raise FileNotFoundError(name)


def ref_counts():
    m = sp.csc_matrix(scipy.io.mmread(os.path.join(REF_IN, "matrix.mtx"))).astype(np.int64)
    return m[load("kept_genes")][:, load("kept_cells")].tocsc()


def rel_close(a, b, rtol, atol=0.0):
    return np.all(np.abs(np.asarray(a, float) - np.asarray(b, float)) <= atol + rtol * np.abs(np.asarray(b, float)))


def check_stages(got, k_pcs=50):
    """got: dict with norm (csc), hvg_order, features, rows, scaled (H x N), emb (N x k), principalvars, adj (csc)."""
    ref_norm = sp.csc_matrix((load("norm_nzval"), load("norm_rowval"), load("norm_colptr")), shape=got["norm"].shape)
    assert np.array_equal(got["norm"].indptr, ref_norm.indptr) and np.array_equal(got["norm"].indices, ref_norm.indices)
    assert rel_close(got["norm"].data, ref_norm.data, 1e-5)                         # normalised within 1e-5 relative
    kept = load("kept_genes")
    pos = {g: i for i, g in enumerate(kept)}
    ref_feat = np.array([pos[g] for g in load("var_gene")])
    assert set(ref_feat) == set(got["features"]), "HVG set bit-exact"
    ref_rows = np.array([pos[g] for g in load("scale_gene")])
    assert np.array_equal(ref_rows, got["rows"]), "scaled rows in ORIGINAL gene order (subset_count)"
    ref_scaled = load("scaled").T                                                     # [N][H] -> H x N
    assert rel_close(got["scaled"], ref_scaled, 1e-5, atol=2e-6)
    ref_emb = load("pca_embedding")
    sign = np.sign(np.sum(got["emb"] * ref_emb, axis=0))
    pv = load("principalvars")
    gaps = np.minimum(np.abs(np.diff(pv[:k_pcs + 1])), np.r_[np.inf, np.abs(np.diff(pv[:k_pcs]))]) / pv[:k_pcs]
    err = np.abs(got["emb"] * sign - ref_emb).max(axis=0) / np.sqrt((ref_emb ** 2).mean(axis=0))
    assert np.all(err[gaps >= 1e-2] < 1e-4)
    assert np.allclose(got["principalvars"], pv[:k_pcs], rtol=1e-4)
    ref_adj = sp.csc_matrix((load("adj_nzval"), load("adj_rowval"), load("adj_colptr")), shape=got["adj"].shape)
    assert (got["adj"] != ref_adj).nnz == 0, "adjacency identical"


@needs_fixtures
def test_oracle_against_reference_fixtures(oracle):
    counts = ref_counts()
    out = oracle.run_pipeline(counts, nFeatures=2000, n_pcs=50, k=20, pca_method="svd")
    idx = load("knn_exact_indices")
    check_stages({"norm": out["norm"], "features": out["hvg"]["features"], "rows": out["rows"], "scaled": out["scaled"],
                  "emb": out["pca"]["embedding"], "principalvars": out["pca"]["principalvars"],
                  "adj": oracle.adjacency_atlas(idx)})
    # the Loess / variance columns themselves (pins the recalled Loess.jl details: SURVEY A.3)
    order = np.array([{g: i for i, g in enumerate(load("kept_genes"))}[g] for g in load("vst_gene")])
    assert np.allclose(out["hvg"]["variance_expected"][order], load("vst_variance_expected"), rtol=1e-8)
    # exact kNN of the reference's embedding == the oracle's exact kNN of it
    widx, _ = oracle.knn_exact(load("pca_embedding"), 20)
    assert (widx == idx).mean() > 0.9999


@needs_fixtures
@pytest.mark.gpu
def test_cuda_path_against_reference_fixtures(cs, ctx):
    counts = ref_counts()
    n = counts.shape[1]
    raw = ctx.sparse_from_scipy(counts)
    params = cs.PathParams(n_features=2000, n_pcs=50, k=20, graph_mode="sum", keep_norm=True, keep_scaled=True)
    out = cs.run_path(ctx, cs.LocalComm(), raw, n, 0, params)
    idx = load("knn_exact_indices").astype(np.int32)
    adj = ctx.graph_build(ctx.vec_from(idx.ravel()), n, 20, 0, n, cs._capi.GRAPH_SUM).to_scipy(np.int64)
    check_stages({"norm": out["norm"].to_scipy(), "features": out["hvg"]["features"], "rows": out["rows"],
                  "scaled": out["scaled"].download(np.float64).T, "emb": out["embedding"].download(np.float64),
                  "principalvars": out["pca"].info()["principalvars"], "adj": adj})
    e = ctx.dense_upload(np.ascontiguousarray(load("pca_embedding")))
    got_idx, _ = ctx.knn(e, e, 20)
    assert (got_idx.download().reshape(n, 20) == idx).mean() > 0.9995       # ties within 1e-6 may swap
