"""f3: find_markers / find_all_markers (src/scrna/processing.jl:350-414) on the CUDA path against the oracle (dense
rows + scipy ranks, itself cross-checked against scipy.stats.mannwhitneyu in tests/test_oracle.py)."""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp
import scipy.stats

pytestmark = pytest.mark.gpu


def make_object(cs, oracle, n_genes=1500, n_cells=2400, n_clusters=5, seed=3):
    p = cs.synth.make_params(n_genes, n_cells, density=0.08, seed=seed)
    counts, _, _ = oracle.subset_matrix(cs.synth.sample_counts(p))
    _, _, typ = cs.synth.cell_scores(p, np.arange(counts.shape[1]))
    genes = [f"g{i}" for i in range(counts.shape[0])]
    cells = [f"c{i}" for i in range(counts.shape[1])]
    obj = cs.scRNAObject(cs.RawCountObject(counts, cells, genes))
    lab = np.array([f"k{t % n_clusters}" for t in typ], dtype=object)
    return obj, lab


def p_close(got, want):
    """p-values agree, except where rounding decides a tie.  Normalised values that are mathematically equal (3 counts
    among 3k, 1 count among k) tie or not depending on the last bit of (x * (1/colsum)) * sf: in Float64 (the reference,
    the oracle) they often come out distinct, in the device's fp32 storage they always tie; each such pair moves U by 1/2
    and a p-value of 1e-11 by up to ~1e-3 relative.  Everything else must match to 1e-6; a systematic error (a wrong
    tie correction, a rank offset) moves the p-values of most genes by far more than the allowance below."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    with np.errstate(divide="ignore"):
        lg, lw = np.log(np.maximum(got, 1e-300)), np.log(np.maximum(want, 1e-300))
    exact = np.isclose(got, want, rtol=1e-6, atol=1e-300)
    return bool(np.all(np.abs(lg - lw) <= 5e-3 * np.maximum(1.0, np.abs(lw)))) and exact.mean() > 0.8


def attach_clusters(cs, obj, lab):
    obj.clustData = cs.ClusteringObject(pd.DataFrame({"cluster": lab, "cell_id": obj.rawCount.cell_name}), "cosine", None, None, 0.06)
    return obj


def test_marker_stats_tables_against_dense_ranks(cs, ctx):
    """The device tables on a matrix with negative values, stored zeros, heavy ties and cells that take no part."""
    rng = np.random.default_rng(4)
    g, n = 37, 500
    dense = np.round(rng.normal(0.3, 1.0, size=(g, n)), 1) * (rng.random((g, n)) < 0.35)
    m = sp.csc_matrix(dense)
    # a few STORED zeros (legal in a SparseMatrixCSC) and an all-tied gene
    m.data[rng.integers(0, m.nnz, 40)] = 0.0
    dense = m.toarray()
    labels = rng.integers(-1, 3, size=n).astype(np.int32)            # -1: not in the test
    dev = ctx.sparse_from_scipy(m)
    t = ctx.marker_stats(dev, labels, 3, expr_cutoff=0.25)
    inc = labels >= 0
    assert t["n_included"] == inc.sum()
    for gi in range(g):
        row = dense[gi, inc].astype(np.float32).astype(np.float64)
        lab = labels[inc]
        ranks = scipy.stats.rankdata(row)
        # stored pattern of this gene
        csr = m.tocsr()
        st_cols = csr.indices[csr.indptr[gi]:csr.indptr[gi + 1]]
        is_stored = np.zeros(n, dtype=bool)
        is_stored[st_cols] = True
        is_stored = is_stored[inc]
        for c in range(3):
            sel = (lab == c) & is_stored
            assert np.isclose(t["rank_sum"][gi, c], ranks[sel].sum(), rtol=1e-12, atol=1e-9), (gi, c)
            assert np.isclose(t["sum_expm1"][gi, c], np.expm1(row[sel]).sum(), rtol=1e-6, atol=1e-9)
            assert t["n_expr"][gi, c] == (row[sel] > np.float32(0.25)).sum()
            assert t["nnz"][gi, c] == sel.sum()
            assert np.isclose(t["sum_x"][gi, c], row[sel].sum(), rtol=1e-6, atol=1e-9)
        _, cnt = np.unique(row[row != 0], return_counts=True)
        assert t["tie_adj"][gi] == float((cnt.astype(float) ** 3 - cnt).sum())
        assert t["n_neg"][gi] == (row < 0).sum()
        assert t["n_zero_stored"][gi] == ((row == 0) & is_stored).sum()


@pytest.mark.parametrize("kw", [{}, {"only_pos": False, "min_pct": 0.0, "p_cutoff": 1.1}, {"expr_cutoff": 0.5}])
def test_find_markers_matches_oracle(cs, oracle, ctx, kw):
    obj, lab = make_object(cs, oracle)
    obj = cs.normalize_object(obj, ctx=ctx)
    attach_clusters(cs, obj, lab)
    got = cs.find_markers(obj, cluster_1="k1", cluster_2="k3", ctx=ctx, **kw)
    norm = oracle.normalize_object(obj.rawCount.count_mtx)
    want = oracle.find_markers(norm, lab, "k1", "k3", obj.rawCount.gene_name, **kw)
    assert list(got.columns) == ["gene", "p_val", "avg_logFC", "pct_1", "pct_2", "p_val_adj"]
    assert len(got) > 20
    # the normalised values are fp32 on the device: a gene can move across a threshold or swap with a neighbour in the
    # logFC order only if it sits within rounding of it
    wg = list(want["gene"])
    common = [x for x in got["gene"] if x in set(wg)]
    assert len(common) >= 0.995 * max(len(wg), len(got))
    gi = got.set_index("gene").loc[common]
    wi = pd.DataFrame({k: want[k] for k in ("p_val", "avg_logFC", "pct_1", "pct_2", "p_val_adj")}, index=wg).loc[common]
    assert p_close(gi["p_val"], wi["p_val"])
    assert p_close(gi["p_val_adj"], wi["p_val_adj"])
    assert np.allclose(gi["avg_logFC"], wi["avg_logFC"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(gi["pct_1"], wi["pct_1"]) and np.array_equal(gi["pct_2"], wi["pct_2"])
    assert np.all(np.diff(got["avg_logFC"].to_numpy()) <= 1e-12), "sorted by avg_logFC, descending"


def test_find_all_markers_matches_oracle(cs, oracle, ctx):
    obj, lab = make_object(cs, oracle, n_genes=900, n_cells=1800, n_clusters=4, seed=8)
    obj = cs.normalize_object(obj, ctx=ctx)
    attach_clusters(cs, obj, lab)
    got = cs.find_all_markers(obj, ctx=ctx)
    want = oracle.find_all_markers(oracle.normalize_object(obj.rawCount.count_mtx), lab, obj.rawCount.gene_name)
    assert list(got["cluster"].unique()) == [w["cluster"][0] for w in want if w["gene"].size]
    for w in want:
        if not w["gene"].size:
            continue
        c = w["cluster"][0]
        g = got[got["cluster"] == c].set_index("gene")
        common = [x for x in w["gene"] if x in g.index]
        assert len(common) >= 0.995 * max(len(g), w["gene"].size)
        wi = pd.DataFrame({k: w[k] for k in ("p_val", "avg_logFC")}, index=list(w["gene"])).loc[common]
        assert p_close(g.loc[common, "p_val"], wi["p_val"])
        assert np.allclose(g.loc[common, "avg_logFC"], wi["avg_logFC"], rtol=1e-5, atol=1e-6)


def test_find_markers_small_groups_use_the_exact_test(cs, oracle, ctx):
    """n1 + n2 <= 10: HypothesisTests enumerates the permutation distribution."""
    rng = np.random.default_rng(1)
    counts = sp.csc_matrix(rng.poisson(1.2, size=(30, 9)).astype(np.int64))
    obj = cs.scRNAObject(cs.RawCountObject(counts, [f"c{i}" for i in range(9)], [f"g{i}" for i in range(30)]))
    obj = cs.normalize_object(obj, ctx=ctx)
    lab = np.array(["a"] * 4 + ["b"] * 5, dtype=object)
    attach_clusters(cs, obj, lab)
    got = cs.find_markers(obj, cluster_1="a", cluster_2="b", only_pos=False, min_pct=0.0, p_cutoff=1.1, ctx=ctx)
    want = oracle.find_markers(oracle.normalize_object(obj.rawCount.count_mtx), lab, "a", "b", obj.rawCount.gene_name,
                               only_pos=False, min_pct=0.0, p_cutoff=1.1)
    assert list(got["gene"]) == list(want["gene"])
    assert np.allclose(got["p_val"], want["p_val"], rtol=1e-9)


def test_find_markers_errors(cs, oracle, ctx):
    obj, lab = make_object(cs, oracle, n_genes=300, n_cells=400)
    obj = cs.normalize_object(obj, ctx=ctx)
    with pytest.raises(ValueError):
        cs.find_markers(obj, cluster_1="k1", cluster_2="k2", ctx=ctx)            # no clustering yet (processing.jl:358-360)
    attach_clusters(cs, obj, lab)
    with pytest.raises(ValueError):
        cs.find_markers(obj, cluster_2="k2", ctx=ctx)                             # :352-354
    with pytest.raises(ValueError):
        cs.find_markers(obj, cluster_1="k1", ctx=ctx)                             # :355-357
