"""Stage-isolated parity of the CUDA path (through the C ABI) against the CPU oracle, each stage on
the SAME input.  Tolerances are the ones BASELINE.json's north_star states."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def rel_close(a, b, rtol, atol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= atol + rtol * np.abs(b))


# ---- a1 normalize ---------------------------------------------------------------------------
@pytest.mark.parametrize("method,pc,sf", [("logarithm", 1, 10000), ("logarithm", 0.5, 1e4), ("none", 1, 1e6)])
def test_normalize_matches_oracle(cs, oracle, ctx, small_counts, method, pc, sf):
    raw = ctx.sparse_from_scipy(small_counts)
    norm = ctx.normalize(raw, sf, cs.pipeline.norm_code(method), pc)
    got = norm.to_scipy()
    want = oracle.normalize_object(small_counts, sf, method, pc)
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    # north_star: normalised matrix within 1e-5 relative (fp32 storage gives ~1e-7)
    assert rel_close(got.data, want.data, 1e-5)
    assert np.max(np.abs(got.data - want.data) / np.abs(want.data)) < 2e-6


def test_normalize_round_trip_property(cs, ctx, medium_counts):
    """expm1(norm)/sf * colsum recovers the counts (size-independent property)."""
    raw = ctx.sparse_from_scipy(medium_counts)
    norm = ctx.normalize(raw, 1e4, 0, 1.0).to_scipy()
    colsum = np.asarray(medium_counts.sum(axis=0)).ravel()
    rec = np.expm1(norm.data) / 1e4 * np.repeat(colsum, np.diff(norm.indptr))
    assert np.allclose(rec, medium_counts.data, rtol=2e-5)


def test_normalize_unknown_method_errors(cs, ctx, small_counts):
    raw = ctx.sparse_from_scipy(small_counts)
    with pytest.raises(cs.CellScopesCudaError) as e:
        ctx.normalize(raw, 1e4, 7, 1.0)
    assert "Unknown normalization method" in str(e.value)
    with pytest.raises(ValueError):
        cs.normalize_object(small_counts, norm_method="bogus")


def test_normalize_zero_sum_column_is_nan_then_zero(cs, oracle, ctx):
    """A column whose stored values sum to zero becomes NaN; imaging objects replace by 0."""
    m = sp.csc_matrix(np.array([[1, 0, 0], [2, 0, 3], [0, 0, 1]], dtype=np.int64))
    m = sp.csc_matrix((np.array([1, 2, 0, 0, 3, 1], dtype=np.int64), np.array([0, 1, 0, 2, 1, 2]),
                       np.array([0, 2, 4, 6])), shape=(3, 3))       # column 1 stores explicit zeros
    raw = ctx.sparse_from_scipy(m)
    got = ctx.normalize(raw, 1e4, 0, 1.0).to_scipy()
    want = oracle.normalize_object(m)
    assert np.isnan(got.data[2:4]).all() and np.isnan(want.data[2:4]).all()
    assert rel_close(got.data[[0, 1, 4, 5]], want.data[[0, 1, 4, 5]], 1e-6)
    got0 = ctx.normalize(raw, 1e4, 0, 1.0, nan_to_zero=True).to_scipy()
    assert np.all(got0.data[2:4] == 0.0)


def test_upload_int64_one_based_and_float_values(cs, ctx, small_counts):
    m = small_counts
    a = ctx.sparse_upload(m.shape[0], m.shape[1], m.indptr.astype(np.int64) + 1, m.indices.astype(np.int64) + 1,
                          m.data.astype(np.int64), index_base=1)
    b = ctx.sparse_upload(m.shape[0], m.shape[1], m.indptr.astype(np.int32), m.indices.astype(np.int32),
                          m.data.astype(np.float64), index_base=0)
    for h in (a, b):
        back = h.to_scipy()
        assert (back != m).nnz == 0
    cp, rv, v = a.download(10, 50, index_base=1)
    assert cp[0] == 1 and rv.min() >= 1
    sub = m[:, 10:50]
    assert np.array_equal(cp - 1, sub.indptr) and np.array_equal(rv - 1, sub.indices)


@pytest.mark.parametrize("piece", ["777", "100000"])
def test_upload_two_lanes_equals_plain_upload(cs, ctx, medium_counts, monkeypatch, piece):
    """csc_sparse_upload's two-lane route for large Int64 / Float64 inputs (host cores narrow pieces from the front --
    to uint16 where lossless -- while the bus carries wide pieces from the back): forced at test sizes with small pieces,
    it must leave exactly what the plain route leaves."""
    m = medium_counts
    monkeypatch.setenv("CSC_UPLOAD_HOST_NARROW", "2")
    monkeypatch.setenv("CSC_UPLOAD_PIECE", piece)
    cp, rv = m.indptr.astype(np.int64) + 1, m.indices.astype(np.int64) + 1
    frac = m.data.astype(np.float64) * 0.37          # not integers: the fp32 branch
    mixed = m.data.astype(np.float64)
    mixed[::1500] = 70000.5                          # a few pieces fail the uint16 test, the others pass
    for vals in (m.data.astype(np.int64), m.data.astype(np.float64), frac, mixed):
        h = ctx.sparse_upload(m.shape[0], m.shape[1], cp, rv, vals, index_base=1)
        gcp, grv, gv = h.download(0, m.shape[1], index_base=1)
        assert np.array_equal(gcp, cp) and np.array_equal(grv, rv)
        assert np.array_equal(gv.astype(np.float32), vals.astype(np.float32))
    # more than 65536 genes: indices travel as int32
    rng = np.random.default_rng(3)
    big = sp.random(70000, 300, density=0.01, format="csc", random_state=rng, data_rvs=lambda n: rng.integers(1, 9, n))
    h = ctx.sparse_upload(70000, 300, big.indptr.astype(np.int64), big.indices.astype(np.int64), big.data.astype(np.float64))
    assert (h.to_scipy() != big).nnz == 0
    # a row index out of range is caught in whichever lane it lands
    for pos in (5, rv.size // 2, rv.size - 3):
        bad = rv.copy()
        bad[pos] = m.shape[0] + 1
        with pytest.raises(cs.CellScopesCudaError):
            ctx.sparse_upload(m.shape[0], m.shape[1], cp, bad, mixed, index_base=1)


def test_upload_rejects_bad_rows(cs, ctx):
    with pytest.raises(cs.CellScopesCudaError):
        ctx.sparse_upload(3, 2, np.array([0, 1, 2]), np.array([0, 5]), np.array([1.0, 1.0]))


def test_upload_rejects_bad_colptr(cs, ctx):
    """interior column pointers must be monotone and within [0, nnz]"""
    rows, vals = np.array([0, 1, 2, 0]), np.ones(4)
    for cp in ([0, 3, 2, 4], [0, 5, 4, 4], [0, -1, 2, 4]):
        with pytest.raises(cs.CellScopesCudaError) as ei:
            ctx.sparse_upload(3, 3, np.array(cp), rows, vals)
        assert ei.value.code == -1


def test_empty_and_ragged_columns(cs, oracle, ctx):
    """Empty columns, a single-entry column and a full column."""
    rng = np.random.default_rng(0)
    d = np.zeros((40, 7), dtype=np.int64)
    d[:, 1] = rng.integers(1, 5, 40)      # full
    d[3, 2] = 9                           # single
    d[:17, 4] = rng.integers(0, 3, 17)
    d[5:, 6] = rng.integers(0, 2, 35)
    m = sp.csc_matrix(d)
    raw = ctx.sparse_from_scipy(m)
    got = ctx.normalize(raw, 1e4, 0, 1.0).to_scipy()
    want = oracle.normalize_object(m)
    assert np.array_equal(got.indptr, want.indptr)
    assert rel_close(got.data, want.data, 1e-5)


# ---- a2 HVG ---------------------------------------------------------------------------------
def test_gene_moments_exact(cs, oracle, ctx, medium_counts):
    raw = ctx.sparse_from_scipy(medium_counts)
    g = medium_counts.shape[0]
    st = ctx.vec(2 * g)
    ctx.gene_stats_partial(raw, st)
    got = st.download()
    s1, s2, _, _ = oracle.gene_moments(medium_counts)
    assert np.array_equal(got[:g], s1) and np.array_equal(got[g:], s2)      # integers: bit-exact
    # fused into normalize: same table
    st2 = ctx.vec(2 * g)
    ctx.normalize(raw, 1e4, 0, 1.0, False, st2)
    assert np.array_equal(st2.download(), got)


def test_gene_moments_large_counts(cs, oracle, ctx):
    """Counts above the shared-memory fast path (>65535) and non-integer values take the direct path."""
    rng = np.random.default_rng(1)
    d = rng.integers(0, 4, size=(50, 300)).astype(np.float64)
    d[7, 5] = 100000.0
    d[9, 6] = 70000.0
    d[11, :] *= 1.5
    m = sp.csc_matrix(d)
    raw = ctx.sparse_from_scipy(m)
    st = ctx.vec(100)
    ctx.gene_stats_partial(raw, st)
    s1, s2, _, _ = oracle.gene_moments(m)
    got = st.download()
    assert np.allclose(got[:50], s1, rtol=1e-12) and np.allclose(got[50:], s2, rtol=1e-12)


@pytest.mark.parametrize("n_genes", [1500, 30000, 60000])
def test_gene_moments_mixed_values_in_every_table_mode(cs, ctx, n_genes):
    """The per-entry moment update (raw_moment_update in csrc/sparse.cu) in its three table modes (both tables / the sum
    table + E slots for part of the genes / no table in shared memory) on values of every class it distinguishes: counts
    of 1, small counts, counts above 255 (x^2 - x leaves the E table), above 65535, non-integers, stored zeros, -0.0 and
    negative values.  All sums here are exactly representable, so the tables must equal the dense sums bit for bit, and
    the pass fused into normalize must equal the standalone one."""
    rng = np.random.default_rng(n_genes)
    n_cells = 64
    nnz_per = min(n_genes, 900)
    rows = np.concatenate([np.sort(rng.choice(n_genes, size=nnz_per, replace=False)) for _ in range(n_cells)])
    vals = rng.integers(1, 6, rows.size).astype(np.float64)
    pick = rng.random(rows.size)
    vals[pick < 0.02] = 300.0
    vals[(pick >= 0.02) & (pick < 0.03)] = 2.5
    vals[(pick >= 0.03) & (pick < 0.035)] = 70000.0
    vals[(pick >= 0.035) & (pick < 0.04)] = 0.0
    vals[(pick >= 0.04) & (pick < 0.045)] = -3.0
    vals[(pick >= 0.045) & (pick < 0.05)] = -0.0
    m = sp.csc_matrix((vals, rows, np.arange(n_cells + 1) * nnz_per), shape=(n_genes, n_cells))
    raw = ctx.sparse_upload(n_genes, n_cells, m.indptr, m.indices, m.data)
    st = ctx.vec(2 * n_genes)
    ctx.gene_stats_partial(raw, st)
    d = m.toarray()
    want = np.r_[d.sum(axis=1), (d * d).sum(axis=1)]
    assert np.array_equal(st.download(), want)
    st2 = ctx.vec(2 * n_genes)
    ctx.normalize(raw, 1e4, 1, 1.0, False, st2)
    assert np.array_equal(st2.download(), want)


@pytest.mark.parametrize("n_genes", [30000, 60000])
def test_gene_moments_exact_when_tables_do_not_fit(cs, oracle, ctx, n_genes):
    """Whole-transcriptome gene counts: with 30k genes only the sum table fits in shared memory (the squares go to
    the fp64 table in global memory), with 60k genes nothing does.  Moments stay bit-exact, normalised values within
    tolerance, and the fused pass equals the standalone one."""
    rng = np.random.default_rng(n_genes)
    n_cells = 400
    m = sp.random(n_genes, n_cells, density=0.01, format="csc", random_state=rng,
                  data_rvs=lambda k: rng.integers(1, 40, size=k).astype(np.float64))
    m.data[::97] = 300.0                                   # beyond the one-byte fast path of the squares
    raw = ctx.sparse_from_scipy(m)
    st = ctx.vec(2 * n_genes)
    ctx.gene_stats_partial(raw, st)
    s1, s2, _, _ = oracle.gene_moments(m)
    got = st.download()
    assert np.array_equal(got[:n_genes], s1) and np.array_equal(got[n_genes:], s2)
    st2 = ctx.vec(2 * n_genes)
    norm = ctx.normalize(raw, 1e4, 0, 1.0, False, st2)
    assert np.array_equal(st2.download(), got)
    want = oracle.normalize_object(m)
    assert rel_close(norm.download(values_only=True), want.data, 1e-5)


@pytest.mark.parametrize("nfeat", [200, 2000])
def test_hvg_set_bit_exact(cs, oracle, ctx, medium_counts, nfeat):
    raw = ctx.sparse_from_scipy(medium_counts)
    g, n = medium_counts.shape
    st = ctx.vec(2 * g)
    ctx.gene_stats_partial(raw, st)
    got = ctx.hvg_finish(st, g, n, nfeat, 0.3)
    want = oracle.find_variable_genes(medium_counts, nfeat, 0.3)
    assert np.array_equal(got["mean"], want["mean"])
    assert np.array_equal(got["variance"], want["variance"])
    assert np.allclose(got["variance_expected"], want["variance_expected"], rtol=1e-10)
    assert np.allclose(got["variance_standardized"], want["variance_standardized"], rtol=1e-10)
    # north_star: HVG set bit-exact (rank order included)
    assert np.array_equal(got["features"], want["features"])
    margin = want["variance_standardized"][want["order"][nfeat - 1]] - want["variance_standardized"][want["order"][nfeat]]
    print(f"HVG boundary margin at rank {nfeat}: {margin:.3e}")


def _tied_moment_table(n_genes, n_cells, seed):
    """Exact moment tables (S1, S2) with heavily repeated means: low-count genes share their S1, so the q-th nearest
    distance of a Loess vertex is attained by many points and the tie rule (lower gene index first) decides."""
    rng = np.random.default_rng(seed)
    s1 = rng.integers(3, 60, n_genes).astype(np.float64)
    s1[rng.random(n_genes) < 0.2] = float(rng.integers(100, 5000))      # a fifth of the genes share ONE mean
    big = rng.random(n_genes) < 0.3
    s1[big] = rng.integers(60, 20000, int(big.sum())).astype(np.float64)
    s2 = s1 + 2.0 * rng.integers(0, 40, n_genes) * (1.0 + np.floor(s1 / 50.0))
    assert np.all((s2 - s1 * s1 / n_cells) > 0)
    return s1, s2


@pytest.mark.parametrize("n_genes,span", [(1024, 0.3), (6000, 0.3), (20011, 0.3), (6000, 0.05), (6000, 1.0)])
def test_hvg_device_finish_matches_oracle_with_tied_means(cs, oracle, ctx, n_genes, span):
    """hvg_dev.cu (>= 1024 genes): window selection by binary searches, tied rim by merge path, block-reduced QR."""
    n_cells = 25000
    s1, s2 = _tied_moment_table(n_genes, n_cells, 5 + n_genes)
    st = ctx.vec_from(np.concatenate([s1, s2]))
    got = ctx.hvg_finish(st, n_genes, n_cells, 2000, span)
    want = oracle.hvg_from_moments(s1, s2, n_cells, 2000, span)
    assert np.array_equal(got["mean"], want["mean"])
    assert np.array_equal(got["variance"], want["variance"])
    assert np.allclose(got["variance_expected"], want["variance_expected"], rtol=1e-9)
    assert np.allclose(got["variance_standardized"], want["variance_standardized"], rtol=1e-9)
    # equal (S1, S2) pairs give equal variance_standardized: the stable rank (ties by gene index) must agree wherever
    # the oracle's neighbouring values are not within rounding of each other
    vs = want["variance_standardized"]
    o_w, o_g = want["order"], got["order"]
    differ = np.nonzero(o_w != o_g)[0]
    for r in differ:
        assert abs(vs[o_w[r]] - vs[o_g[r]]) <= 1e-9 * abs(vs[o_w[r]]), (r, o_w[r], o_g[r])
    print(f"{n_genes} genes span {span}: {differ.size} rank positions differ inside rounding ties")


def test_hvg_device_and_host_finish_agree(cs, ctx, monkeypatch):
    n_genes, n_cells = 12000, 40000
    s1, s2 = _tied_moment_table(n_genes, n_cells, 77)
    st = ctx.vec_from(np.concatenate([s1, s2]))
    dev = ctx.hvg_finish(st, n_genes, n_cells, 2000, 0.3)
    monkeypatch.setenv("CSC_HVG_HOST", "1")
    host = ctx.hvg_finish(st, n_genes, n_cells, 2000, 0.3)
    monkeypatch.delenv("CSC_HVG_HOST")
    assert np.array_equal(dev["mean"], host["mean"]) and np.array_equal(dev["variance"], host["variance"])
    assert np.allclose(dev["variance_expected"], host["variance_expected"], rtol=1e-11)
    vs = host["variance_standardized"]
    differ = np.nonzero(dev["order"] != host["order"])[0]
    for r in differ:
        assert abs(vs[dev["order"][r]] - vs[host["order"][r]]) <= 1e-11 * abs(vs[host["order"][r]])


def test_hvg_device_finish_rejects_zero_variance_gene(cs, ctx):
    n_genes, n_cells = 3000, 25000
    s1, s2 = _tied_moment_table(n_genes, n_cells, 3)
    s1[1234] = 7.0 * n_cells          # the same count in every cell: variance exactly 0
    s2[1234] = 49.0 * n_cells
    st = ctx.vec_from(np.concatenate([s1, s2]))
    with pytest.raises(cs.CellScopesCudaError) as ei:
        ctx.hvg_finish(st, n_genes, n_cells, 2000, 0.3)
    assert ei.value.code == -1 and "gene 1234" in str(ei.value)


def test_hvg_nfeatures_clamped(cs, ctx, small_counts):
    raw = ctx.sparse_from_scipy(small_counts)
    g, n = small_counts.shape
    st = ctx.vec(2 * g)
    ctx.gene_stats_partial(raw, st)
    got = ctx.hvg_finish(st, g, n, 10 * g, 0.3)       # processing.jl:134-137
    assert got["features"].size == g
    assert np.array_equal(np.sort(got["features"]), np.arange(g))


def test_hvg_rejects_zero_variance_gene(cs, oracle, ctx):
    """processing.jl:143,151: the reference throws when a gene does not vary; so do the oracle and the library."""
    d = np.array([[1, 0, 2, 0, 3], [2, 2, 2, 2, 2], [0, 1, 0, 4, 0]], dtype=np.int64)
    m = sp.csc_matrix(d)
    with pytest.raises(ValueError):
        oracle.find_variable_genes(m, 2)
    raw = ctx.sparse_from_scipy(m)
    st = ctx.vec(6)
    ctx.gene_stats_partial(raw, st)
    with pytest.raises(cs.CellScopesCudaError) as ei:
        ctx.hvg_finish(st, 3, 5, 2, 0.3)
    assert ei.value.code == -1 and "variance" in str(ei.value)


# ---- a3 scale -------------------------------------------------------------------------------
def _scaled_pair(cs, oracle, ctx, counts, nfeat, **kw):
    norm_o = oracle.normalize_object(counts)
    hv = oracle.find_variable_genes(counts, nfeat)
    rows = oracle.subset_count_rows(counts.shape[0], hv["features"])
    want = oracle.scale_object(norm_o.tocsr()[rows].tocsc(), **kw)
    raw = ctx.sparse_from_scipy(counts)
    norm = ctx.normalize(raw, 1e4, 0, 1.0)
    part = ctx.vec(2 * rows.size)
    ctx.scale_stats_partial(norm, hv["features"], part)
    dense = ctx.scale_fill(norm, hv["features"], part, counts.shape[1], kw.get("scale_max", 10.0),
                           kw.get("do_scale", True), kw.get("do_center", True))
    return dense, want, rows


@pytest.mark.parametrize("kw", [{}, {"scale_max": 3.0}, {"do_center": False}, {"do_scale": False}])
def test_scale_matches_oracle(cs, oracle, ctx, medium_counts, kw):
    dense, want, rows = _scaled_pair(cs, oracle, ctx, medium_counts, 800, **kw)
    got = dense.download(np.float64).T          # genes x cells
    assert got.shape == want.shape
    # north_star: scaled matrix within 1e-5 relative; z-scores near 0 need an absolute component
    assert rel_close(got, want, 1e-5, atol=2e-6)
    assert got.max() <= kw.get("scale_max", 10.0) + 1e-6
    if not kw:
        assert (want >= 10.0).any(), "the clip must actually be exercised"


@pytest.mark.parametrize("method,sf", [("logarithm", 1e4), ("none", 1e4), ("none", 1.0), ("logarithm", 1e6)])
def test_scale_moments_match_oracle(cs, oracle, ctx, medium_counts, method, sf, monkeypatch):
    """(sum y, sum y^2) of the selected rows -- 64-bit fixed-point accumulators for the logarithm method, fp64 sums when
    the value range leaves too few fractional bits ("none" with a large scale factor) -- against float64 sums of the
    oracle's normalised matrix, and the scaled matrix built from them (the 32-bit tables of round 1 resolved
    y^2 to ~40 for "none" and to 3e-5 for the logarithm)."""
    g, n = medium_counts.shape
    raw = ctx.sparse_from_scipy(medium_counts)
    norm = ctx.normalize(raw, sf, cs.pipeline.norm_code(method), 1.0)
    want = oracle.normalize_object(medium_counts, sf, method, 1.0).tocsr()
    feats = np.arange(0, g, 3)
    s1 = np.asarray(want[feats].sum(axis=1)).ravel()
    s2 = np.asarray(want[feats].multiply(want[feats]).sum(axis=1)).ravel()
    got = {}
    for impl in ("", "fp64"):
        monkeypatch.setenv("CSC_SCALE_STATS_IMPL", impl)
        part = ctx.vec(2 * feats.size)
        ctx.scale_stats_partial(norm, feats, part)
        got[impl] = part.download()
        # the device values are fp32 (relative 6e-8 each): sums agree to ~1e-6 relative
        assert np.allclose(got[impl][:feats.size], s1, rtol=2e-6, atol=1e-9 * sf)
        assert np.allclose(got[impl][feats.size:], s2, rtol=4e-6, atol=1e-9 * sf * sf)
    # the two device routes see the SAME fp32 values: they differ by the fixed-point resolution only
    assert np.allclose(got[""], got["fp64"], rtol=1e-9, atol=1e-6)
    monkeypatch.setenv("CSC_SCALE_STATS_IMPL", "")
    part = ctx.vec(2 * feats.size)
    ctx.scale_stats_partial(norm, feats, part)
    z = ctx.scale_fill(norm, feats, part, n).download(np.float64).T
    wz = oracle.scale_object(want[feats].tocsc())
    assert rel_close(z, wz, 1e-5, atol=2e-6)
    # sharding invariance: two cell ranges add up to the whole
    a, b = ctx.sparse_from_scipy(medium_counts[:, :n // 3]), ctx.sparse_from_scipy(medium_counts[:, n // 3:])
    p2 = ctx.vec(2 * feats.size)
    for shard in (a, b):
        ctx.scale_stats_partial(ctx.normalize(shard, sf, cs.pipeline.norm_code(method), 1.0), feats, p2)
    assert np.allclose(p2.download(), got[""], rtol=1e-12, atol=1e-7)


def test_scale_moments_rare_gene_in_deep_cells(cs, oracle, ctx):
    """A gene seen in three very deep cells only (y ~ 0.1): its sd must still come out to 1e-6 (the case ADVICE r1
    computed a 1e-4..1e-3 relative sd error for)."""
    rng = np.random.default_rng(5)
    n, g = 3000, 40
    d = rng.poisson(0.3, size=(g, n)).astype(np.int64)
    d[:, :3] += rng.poisson(3000, size=(g, 3))          # three deep cells (~1.2e5 counts each)
    d[7] = 0
    d[7, :3] = 1                                          # y = log1p(1e4 / 1.2e5) ~ 0.08
    m = sp.csc_matrix(d)
    raw = ctx.sparse_from_scipy(m)
    norm = ctx.normalize(raw, 1e4, 0, 1.0)
    part = ctx.vec(2 * g)
    ctx.scale_stats_partial(norm, None, part)
    got = part.download()
    y = oracle.normalize_object(m).toarray()
    assert np.allclose(got[:g], y.sum(axis=1), rtol=1e-6)
    assert np.allclose(got[g:], (y * y).sum(axis=1), rtol=1e-6)
    sd_got = np.sqrt((got[g:] - got[:g] ** 2 / n) / (n - 1))
    assert np.allclose(sd_got, y.std(axis=1, ddof=1), rtol=2e-6)
    z = ctx.scale_fill(norm, None, part, n).download(np.float64).T
    assert rel_close(z, oracle.scale_object(sp.csc_matrix(y)), 1e-5, atol=2e-6)


def test_scale_rows_in_original_gene_order(cs, oracle, ctx, small_counts):
    """features given in rank order; rows come out in original gene order (utils.jl:72-76)."""
    dense, want, rows = _scaled_pair(cs, oracle, ctx, small_counts, 100)
    assert np.all(np.diff(rows) > 0)
    assert dense.shape == (small_counts.shape[1], 100)


def test_scale_zero_variance_gene_gives_nan(cs, oracle, ctx):
    d = np.zeros((4, 6))
    d[0] = [1, 2, 3, 4, 5, 6]
    d[1] = [2, 2, 2, 2, 2, 2]     # constant -> sd 0 -> NaN (0/0)
    d[2] = [0, 1, 0, 1, 0, 7]
    d[3] = [5, 0, 0, 0, 0, 1]
    m = sp.csc_matrix(d)
    dev = ctx.sparse_from_scipy(m)
    part = ctx.vec(8)
    ctx.scale_stats_partial(dev, None, part)
    got = ctx.scale_fill(dev, None, part, 6).download(np.float64).T
    want = oracle.scale_object(m)
    assert np.isnan(got[1]).all() and np.isnan(want[1]).all()
    assert rel_close(got[[0, 2, 3]], want[[0, 2, 3]], 1e-5, atol=1e-6)


# ---- a4 PCA ---------------------------------------------------------------------------------
def _align(got, want):
    sign = np.sign(np.sum(got * want, axis=0))
    sign[sign == 0] = 1
    return got * sign


def _check_pca(got_emb, info, want, k):
    lam_w = want["principalvars"]
    assert np.allclose(info["principalvars"], lam_w, rtol=1e-4), "explained variance"
    assert abs(info["tvar"] - want["tvar"]) <= 1e-5 * want["tvar"]
    allv = want["all_vars"]
    gaps = np.minimum(np.abs(np.diff(allv[:k + 1])), np.r_[np.inf, np.abs(np.diff(allv[:k]))]) / allv[:k]
    emb_w = want["embedding"]
    got = _align(got_emb, emb_w)
    scale = np.sqrt((emb_w ** 2).mean(axis=0))
    err = np.abs(got - emb_w).max(axis=0) / scale
    # an eigenvector is defined to ~eps/gap (measured on this path: error ~ 1.8e-6 / gap, i.e. 1.1e-4 at a gap of
    # 1.6e-2): components are compared one by one where the gap is >= 2e-2, as a subspace below
    well = gaps >= 2e-2
    assert well.sum() >= 1
    # north_star: PCA embeddings within a FLAT 1e-4 after per-component sign alignment (no gap-dependent slack)
    worst = int(np.argmax(np.where(well, err, 0.0)))
    assert np.all(err[well] < 1e-4), f"component {worst}: err {err[worst]:.3e} at gap {gaps[worst]:.3e}; all: " + \
        " ".join(f"{e:.1e}@{g:.1e}" for e, g in zip(err, gaps))
    # ill-separated components: compare the subspace instead
    qg, _ = np.linalg.qr(got_emb)
    qw, _ = np.linalg.qr(emb_w)
    s = np.linalg.svd(qg.T @ qw, compute_uv=False)
    if gaps.min() >= 1e-3:
        assert s.min() > 1 - 1e-6
    return err, gaps


def test_pca_direct_small_h(cs, oracle, ctx):
    """H <= 160: full Jacobi eigendecomposition."""
    rng = np.random.default_rng(3)
    n, h, k = 3000, 96, 20
    lat = rng.normal(size=(n, 25)) * np.linspace(6, 1.5, 25)
    z = lat @ rng.normal(size=(25, h)) + rng.normal(size=(n, h))
    z = np.minimum((z - z.mean(0)) / z.std(0, ddof=1), 10.0)
    dense = ctx.dense_upload(z.astype(np.float32))
    z32 = dense.download(np.float64)
    want = oracle.run_pca(z32.T, k)
    model, emb = cs.processing._pca_device(ctx, dense, n, k)
    _check_pca(emb.download(np.float64), model.info(), want, k)


def test_pca_block_solver(cs, oracle, ctx):
    """H > 160: Gram + block subspace iteration with Rayleigh-Ritz."""
    rng = np.random.default_rng(4)
    n, h, k = 6000, 600, 30
    lat = rng.normal(size=(n, 40)) * (8.0 * 0.9 ** np.arange(40))
    z = lat @ rng.normal(size=(40, h)) / np.sqrt(h) * 6 + rng.normal(size=(n, h))
    z = np.minimum((z - z.mean(0)) / z.std(0, ddof=1), 10.0)
    dense = ctx.dense_upload(z.astype(np.float32))
    z32 = dense.download(np.float64)
    want = oracle.run_pca(z32.T, k)
    model, emb = cs.processing._pca_device(ctx, dense, n, k)
    print("eigensolver iters", model.iters, "resid", model.resid)
    assert model.resid < 1e-7
    _check_pca(emb.download(np.float64), model.info(), want, k)


def test_pca_solve_reports_non_finite_gram_and_non_convergence(cs, ctx):
    """CSC_ERR_NUMERIC (-5) instead of NaN loadings returned as OK."""
    rng = np.random.default_rng(2)
    h, n = 300, 500
    z = rng.standard_normal((n, h))
    gram = z.T @ z
    colsum = z.sum(axis=0)
    bad = gram.copy()
    bad[17, :] = np.nan
    bad[:, 17] = np.nan
    with pytest.raises(cs.CellScopesCudaError) as ei:
        ctx.pca_solve(ctx.vec_from(colsum), ctx.vec_from(bad.ravel()), h, n, 10)
    assert ei.value.code == -5
    # a flat noise spectrum cannot be resolved to 1e-13 in two products
    with pytest.raises(cs.CellScopesCudaError) as ei:
        ctx.pca_solve(ctx.vec_from(colsum), ctx.vec_from(gram.ravel()), h, n, 10, tol=1e-13, max_iter=2)
    assert ei.value.code == -5
    ok = ctx.pca_solve(ctx.vec_from(colsum), ctx.vec_from(gram.ravel()), h, n, 10)
    assert ok.resid <= 1e-9


def test_pca_on_pipeline_data(cs, oracle, ctx, medium_counts):
    dense, want_scaled, rows = _scaled_pair(cs, oracle, ctx, medium_counts, 1000)
    n = medium_counts.shape[1]
    z32 = dense.download(np.float64)
    want = oracle.run_pca(z32.T, 30, method="gram")
    model, emb = cs.processing._pca_device(ctx, dense, n, 30)
    err, gaps = _check_pca(emb.download(np.float64), model.info(), want, 30)
    print("pipeline-data PCA: max err", err.max(), "min gap", gaps.min(), "iters", model.iters)


def test_sparse_transform_equals_dense_transform(cs, oracle, ctx, medium_counts):
    """(Z - 1m')U taken from the sparse normalised matrix (Z = S + 1 z0') equals the dense product, also with a
    component count that is not a multiple of 32 and with clipping active."""
    raw = ctx.sparse_from_scipy(medium_counts)
    g, n = medium_counts.shape
    norm = ctx.normalize(raw, 1e4, 0, 1.0)
    feats = np.arange(1, g, 2)
    part = ctx.vec(2 * feats.size)
    ctx.scale_stats_partial(norm, feats, part)
    for k, smax in ((20, 10.0), (50, 10.0), (33, 1.5)):
        dense = ctx.scale_fill(norm, feats, part, n, smax)
        colsum, gram = ctx.vec(feats.size), ctx.vec(feats.size ** 2)
        ctx.pca_partial(dense, colsum, gram)
        model = ctx.pca_solve(colsum, gram, feats.size, n, k)
        a = ctx.pca_transform(dense, model).download(np.float64)
        b = ctx.pca_transform_sparse(norm, feats, part, n, model, smax).download(np.float64)
        scale = np.abs(a).max(axis=0)
        assert np.all(np.abs(a - b).max(axis=0) <= 2e-5 * scale), (k, smax, np.abs(a - b).max(axis=0) / scale)
        # and against the fp64 product of the downloaded matrix
        z = dense.download(np.float64)
        info = model.info()
        want = (z - info["mean"]) @ info["loadings"]
        assert np.all(np.abs(b - want).max(axis=0) <= 2e-5 * scale)


@pytest.mark.gpu
def test_select_rows_is_subset_count(cs, oracle, ctx, medium_counts):
    """csc_sparse_select_rows == rows of the matrix in ORIGINAL gene order (src/scrna/utils.jl:59-87), bit for bit; an
    empty selection of a cell stays an empty column; unsorted and repeated feature ids are fine."""
    raw = ctx.sparse_from_scipy(medium_counts)
    G, N = medium_counts.shape
    rng = np.random.default_rng(3)
    feats = rng.choice(G, size=137, replace=False)
    feats = np.r_[feats, feats[:5]]                       # repeats
    sub = ctx.select_rows(raw, feats).to_scipy()
    rows = np.unique(feats)
    want = medium_counts.tocsr()[rows].tocsc()
    want.sort_indices()
    assert sub.shape == (rows.size, N)
    assert np.array_equal(sub.indptr, want.indptr) and np.array_equal(sub.indices, want.indices)
    assert np.array_equal(sub.data, want.data.astype(np.float32).astype(np.float64))


@pytest.mark.gpu
@pytest.mark.parametrize("nfeat,kw", [(2000, {}), (2000, {"scale_max": 3.0}), (333, {}), (1200, {"do_center": False})])
def test_fused_scale_gram_equals_fill_then_gram(cs, oracle, ctx, medium_counts, nfeat, kw):
    """csc_scale_gram_partial (no dense scaled matrix) == csc_scale_fill + csc_pca_partial, and both == the fp64 Gram of
    the oracle's scaled matrix.  333 features exercise the fallback (H not a multiple of 4), 1200 the zero-padded atom."""
    raw = ctx.sparse_from_scipy(medium_counts)
    G, N = medium_counts.shape
    st = ctx.vec(2 * G)
    norm = ctx.normalize(raw, 1e4, cs.pipeline.norm_code("logarithm"), 1.0, False, st)
    feats = ctx.hvg_finish(st, G, N, nfeat, 0.3)["features"]
    H = np.unique(feats).size
    part = ctx.vec(2 * H)
    ctx.scale_stats_partial(norm, feats, part)
    args = dict(scale_max=10.0, do_scale=True, do_center=True)
    args.update(kw)
    c1, g1 = ctx.vec(H), ctx.vec(H * H)
    scaled = ctx.scale_fill(norm, feats, part, N, args["scale_max"], args["do_scale"], args["do_center"])
    ctx.pca_partial(scaled, c1, g1)
    c2, g2 = ctx.vec(H), ctx.vec(H * H)
    ctx.scale_gram_partial(norm, feats, part, N, c2, g2, args["scale_max"], args["do_scale"], args["do_center"])
    z = scaled.download(np.float64)
    ref_g = z.T @ z
    ref_c = z.sum(0)
    scale = np.sqrt(np.outer(np.diag(ref_g), np.diag(ref_g))) + 1e-30
    for c, g in ((c1, g1), (c2, g2)):
        gg = g.download().reshape(H, H)
        assert np.abs(gg - ref_g).max() / scale.max() < 2e-6
        # the fused path takes the column sums out of the Gram kernel (a column of ones): split-precision accuracy
        assert np.allclose(c.download(), ref_c, rtol=0, atol=1e-6 * np.sqrt(N * np.diag(ref_g)).max())
    assert np.abs(g1.download() - g2.download()).max() / scale.max() < 1e-6



def test_gram_partials_add_up(cs, ctx):
    """Sharding invariance: partials of two cell ranges sum to the partial of the whole (multi-GPU allreduce)."""
    rng = np.random.default_rng(5)
    z = rng.normal(size=(5000, 300)).astype(np.float32)
    full, a, b = ctx.dense_upload(z), ctx.dense_upload(z[:1777]), ctx.dense_upload(z[1777:])
    outs = []
    for parts in ([full], [a, b]):
        cs_, g = ctx.vec(300), ctx.vec(300 * 300)
        for d in parts:
            ctx.pca_partial(d, cs_, g)
        outs.append((cs_.download(), g.download()))
    assert np.allclose(outs[0][0], outs[1][0], rtol=1e-12, atol=1e-9)
    assert np.allclose(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-3)
    ref = z.astype(np.float64).T @ z.astype(np.float64)
    assert np.allclose(outs[0][1].reshape(300, 300), ref, rtol=1e-5, atol=2e-3)


# ---- a5/a6 kNN ------------------------------------------------------------------------------
def _knn_check(idx, dist, widx, wdist, tie=1e-6):
    """Bit-exact neighbour sets except at distance ties within `tie` (north_star)."""
    n, k = idx.shape
    assert np.allclose(dist, wdist, atol=2e-6)
    same = idx == widx
    bad_rows = np.flatnonzero(~same.all(axis=1))
    for r in bad_rows:
        for j in np.flatnonzero(~same[r]):
            # a mismatch is only allowed inside a tie band: the two candidates' distances differ by <= tie
            assert abs(wdist[r, j] - dist[r, j]) <= tie, (r, j)
            lo = wdist[r, j] - tie
            hi = wdist[r, j] + tie
            assert lo <= dist[r, j] <= hi
        # as sets, everything strictly inside the k-th distance must agree
        kth = wdist[r, -1]
        strict_w = set(widx[r][wdist[r] < kth - tie])
        assert strict_w <= set(idx[r])
    return bad_rows.size


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
@pytest.mark.parametrize("n,d,k", [(3000, 50, 20), (1111, 10, 30), (700, 33, 5), (2500, 50, 24), (2500, 50, 25),
                                   (2500, 50, 30), (2500, 50, 32), (900, 20, 48)])
def test_knn_exact_vs_oracle(cs, oracle, ctx, metric, n, d, k):
    rng = np.random.default_rng(n + d)
    centers = rng.normal(size=(12, d)) * 3
    x = (centers[rng.integers(0, 12, n)] + rng.normal(size=(n, d))).astype(np.float32)
    dense = ctx.dense_upload(x)
    idx, dist = ctx.knn(dense, dense, k, metric=cs.pipeline.metric_code(metric))
    idx = idx.download().reshape(n, k)
    dist = dist.download().reshape(n, k).astype(np.float64)
    widx, wdist = oracle.knn_exact(x.astype(np.float64), k, metric)
    nbad = _knn_check(idx, dist, widx, wdist, tie=1e-6 if metric == "cosine" else 1e-5)
    assert nbad <= max(2, n // 200)
    assert np.all(idx != np.arange(n)[:, None]), "self excluded"
    assert np.all(np.diff(dist, axis=1) >= -1e-7), "ascending distance"


def test_knn_ties_broken_by_lower_index(cs, ctx):
    """Duplicated points: equal distances must come out in ascending id order."""
    rng = np.random.default_rng(9)
    base = rng.normal(size=(40, 8)).astype(np.float32)
    x = np.concatenate([base, base, base], axis=0)      # every point has two exact duplicates
    dense = ctx.dense_upload(x)
    idx, dist = ctx.knn(dense, dense, 4)
    idx = idx.download().reshape(-1, 4)
    dist = dist.download().reshape(-1, 4)
    for i in range(x.shape[0]):
        dup = sorted(j for j in (i % 40, i % 40 + 40, i % 40 + 80) if j != i)
        assert list(idx[i, :2]) == dup
        assert np.all(dist[i, :2] <= 1e-6)


def test_knn_dims_subset_and_query_shard(cs, oracle, ctx):
    """dims_use (run_umap dims 1:10) and a query shard against all keys (multi-GPU layout)."""
    rng = np.random.default_rng(11)
    x = rng.normal(size=(2000, 30)).astype(np.float32)
    keys = ctx.dense_upload(x)
    widx, wdist = oracle.knn_exact(x[:, :10].astype(np.float64), 15, "cosine")
    q = ctx.dense_upload(x[500:1300])
    idx, dist = ctx.knn(q, keys, 15, query_offset=500, dim_lo=0, dim_hi=10)
    idx = idx.download().reshape(800, 15)
    dist = dist.download().reshape(800, 15).astype(np.float64)
    _knn_check(idx, dist, widx[500:1300], wdist[500:1300])


def test_knn_paths_agree_at_scale(cs, ctx, monkeypatch):
    """Size-independent property at a size the oracle cannot reach: the fp16 filter/prove/refine scan with the
    locality order, the same without the order, with and without the key-range split of the last wave, and the exact
    3xTF32 scan give IDENTICAL neighbour tables (ids bit-exact, distances to rounding), for the whole set and for a
    query shard.  Includes exact duplicates."""
    rng = np.random.default_rng(21)
    n, d, k = 40000, 32, 15
    centers = rng.normal(size=(25, d)) * 2.5
    x = (centers[rng.integers(0, 25, n)] + rng.normal(size=(n, d))).astype(np.float32)
    x[1000:1200] = x[3000:3200]                          # exact duplicates: ties must resolve by lower id
    keys = ctx.dense_upload(x)

    def run(env):
        for kk in ("CSC_KNN_ORDER", "CSC_KNN_IMPL", "CSC_KNN_SPLIT", "CSC_KNN_FB_DIRECT", "CSC_KNN_L"):
            monkeypatch.delenv(kk, raising=False)
        for kk, v in env.items():
            monkeypatch.setenv(kk, v)
        i, dd = ctx.knn(keys, keys, k)
        fb = ctx.knn_last_fallback_rows()
        q = ctx.dense_upload(x[12345:23456])
        iq, dq = ctx.knn(q, keys, k, query_offset=12345)
        return (i.download().reshape(n, k), dd.download().reshape(n, k), iq.download().reshape(-1, k),
                dq.download().reshape(-1, k), fb)

    a = run({})
    b = run({"CSC_KNN_ORDER": "0"})
    c = run({"CSC_KNN_IMPL": "tf32"})
    # the blocks of the scan's thin last wave cut into key-range parts (default: on, 157 blocks -> 9 blocks x 8 parts):
    # off, and forced to 2, 3 and 8 parts with and without the locality order
    splits = [run({"CSC_KNN_SPLIT": "0"}), run({"CSC_KNN_SPLIT": "2"}), run({"CSC_KNN_SPLIT": "3", "CSC_KNN_ORDER": "0"}),
              run({"CSC_KNN_SPLIT": "8"})]
    assert a[4] < n // 20, "the proof should cover almost every row"
    # the unproven rows: direct exact scan (default when there are few) against the tensor-core re-scan; k = 15 with 32
    # candidates leaves a handful, the same k with the proof starved (k close to L is emulated by k = 24) leaves more
    splits.append(run({"CSC_KNN_FB_DIRECT": "0"}))
    print("fallback rows:", a[4])
    for other in [b, c] + splits:
        assert np.array_equal(a[0], other[0])
        assert np.array_equal(a[2], other[2])
        assert np.allclose(a[1], other[1], atol=1e-6) and np.allclose(a[3], other[3], atol=1e-6)
    assert np.array_equal(a[2], a[0][12345:23456])
    assert np.all(a[0] != np.arange(n)[:, None])
    # duplicates: row 1000+j and 3000+j are each other's nearest neighbour at distance 0
    assert np.array_equal(a[0][1000:1200, 0], np.arange(3000, 3200)) and np.all(a[1][1000:1200, 0] <= 1e-6)


@pytest.mark.parametrize("k", [24, 25, 30, 32])
def test_knn_default_k_stays_on_the_fp16_path(cs, ctx, monkeypatch, k):
    """run_clustering's default is n_neighbors = 30 (processing.jl:309), the 4M-cell config uses 30: k = 25..32 run the
    fp16 filter with 48 candidates per row.  Same table as the exact 3xTF32 scan, and the completeness proof must cover
    all but a sliver of the rows (otherwise the 'fast path' is the fallback in disguise)."""
    rng = np.random.default_rng(77)
    n, d = 60000, 50
    centers = rng.normal(size=(30, d)) * 2.0
    x = (centers[rng.integers(0, 30, n)] + rng.normal(size=(n, d)) * np.linspace(1.5, 0.3, d)).astype(np.float32)
    keys = ctx.dense_upload(x)
    monkeypatch.delenv("CSC_KNN_IMPL", raising=False)
    i1, d1 = ctx.knn(keys, keys, k)
    fb = ctx.knn_last_fallback_rows()
    assert fb < n // 100, f"k={k}: {fb} of {n} rows fell back to the exact re-scan"
    monkeypatch.setenv("CSC_KNN_IMPL", "tf32")
    i2, d2 = ctx.knn(keys, keys, k)
    i1, i2 = i1.download().reshape(n, k), i2.download().reshape(n, k)
    d1, d2 = d1.download().reshape(n, k), d2.download().reshape(n, k)
    # independent referee: fp64 brute force (torch, no library code) on every row where the two paths disagree plus a
    # spread of 256 others; a difference is only legal between distances within fp32 noise of each other
    import torch
    xt = torch.as_tensor(x, device="cuda:0").double()
    unit = xt / xt.norm(dim=1, keepdim=True)
    differ = np.flatnonzero((i1 != i2).any(axis=1))
    rows = np.unique(np.concatenate([differ[:2048], np.linspace(0, n - 1, 256).astype(np.int64)]))
    rt = torch.as_tensor(rows, device="cuda:0")
    sim = unit[rt] @ unit.T
    sim[torch.arange(len(rows)), rt] = -2.0
    top = torch.topk(sim, k + 1, dim=1)
    ref_idx, ref_sim = top.indices.cpu().numpy(), top.values.cpu().numpy()
    bad1 = bad2 = 0
    for j, r in enumerate(rows):
        want = 1.0 - ref_sim[j, :k]
        assert np.allclose(d1[r], want, atol=2e-6), ("fp16 path distances", r)
        assert np.allclose(d2[r], want, atol=2e-6), ("tf32 path distances", r)
        if ref_sim[j, k - 1] - ref_sim[j, k] < 1e-6:
            continue
        bad1 += set(i1[r].tolist()) != set(ref_idx[j, :k].tolist())
        bad2 += set(i2[r].tolist()) != set(ref_idx[j, :k].tolist())
    assert bad1 == 0 and bad2 == 0, f"k={k}: rows with a wrong neighbour SET: fp16 path {bad1}, tf32 path {bad2} (of {rows.size} checked)"
    # the ORDER inside a row may differ only between (near-)equal distances
    for r in differ:
        assert np.allclose(np.sort(d1[r]), np.sort(d2[r]), atol=2e-6)
    assert differ.size <= n // 1000, f"{differ.size} rows differ between the two paths"


def test_knn_few_unproven_rows_take_the_direct_scan(cs, ctx, monkeypatch):
    """Rows whose candidate list cannot be proven complete are re-scanned exactly: directly (k_knn_fb_scan) when they are
    few, through the 3xTF32 tensor-core scan otherwise.  Find a k that leaves between 1 and 512 such rows (32 candidates
    per row forced), and require the two routes to give the same table, equal to an fp64 brute force on those rows."""
    import torch
    rng = np.random.default_rng(5)
    n, d = 50000, 50
    centers = rng.normal(size=(40, d)) * 1.5
    x = (centers[rng.integers(0, 40, n)] + rng.normal(size=(n, d)) * np.linspace(1.2, 0.2, d)).astype(np.float32)
    keys = ctx.dense_upload(x)
    monkeypatch.setenv("CSC_KNN_L", "32")
    found = None
    for k in (24, 26, 27, 28, 29, 30):
        monkeypatch.setenv("CSC_KNN_FB_DIRECT", "1")
        i1, d1 = ctx.knn(keys, keys, k)
        fb = ctx.knn_last_fallback_rows()
        if 1 <= fb <= 512:
            found = (k, fb, i1.download().reshape(n, k), d1.download().reshape(n, k))
            break
    assert found is not None, "no k left between 1 and 512 unproven rows: the direct scan was not exercised"
    k, fb, i1, d1 = found
    monkeypatch.setenv("CSC_KNN_FB_DIRECT", "0")
    i2, d2 = ctx.knn(keys, keys, k)
    i2, d2 = i2.download().reshape(n, k), d2.download().reshape(n, k)
    assert ctx.knn_last_fallback_rows() == fb
    differ = np.flatnonzero((i1 != i2).any(axis=1))
    for r in differ:                                       # only the order of (near-)equal distances may differ
        assert np.allclose(np.sort(d1[r]), np.sort(d2[r]), atol=2e-6)
    assert np.allclose(d1, d2, atol=2e-6)
    xt = torch.as_tensor(x, device="cuda:0").double()
    unit = xt / xt.norm(dim=1, keepdim=True)
    sim = unit @ unit.T
    sim.fill_diagonal_(-2.0)
    top = torch.topk(sim, k + 1, dim=1)
    ref_idx, ref_sim = top.indices.cpu().numpy(), top.values.cpu().numpy()
    sep = ref_sim[:, k - 1] - ref_sim[:, k] >= 1e-6
    wrong = [r for r in np.flatnonzero(sep) if set(i1[r].tolist()) != set(ref_idx[r, :k].tolist())]
    assert not wrong, f"k={k}, {fb} unproven rows: {len(wrong)} rows with a wrong neighbour set"
    print(f"direct scan exercised: k={k}, {fb} unproven rows, {differ.size} rows ordered differently by the two routes")


def test_knn_bad_k(cs, ctx):
    x = ctx.dense_upload(np.random.default_rng(0).normal(size=(10, 4)).astype(np.float32))
    with pytest.raises(cs.CellScopesCudaError):
        ctx.knn(x, x, 10)


# ---- adjacency / SNN ------------------------------------------------------------------------
@pytest.mark.parametrize("n,k", [(500, 10), (3000, 20)])
def test_graphs_identical_to_oracle(cs, oracle, ctx, n, k):
    rng = np.random.default_rng(n)
    x = rng.normal(size=(n, 16))
    widx, _ = oracle.knn_exact(x, k, "cosine")
    tab = ctx.vec_from(widx.astype(np.int32).ravel())
    for mode, want in ((cs._capi.GRAPH_BINARY, oracle.adjacency_small(widx)),
                       (cs._capi.GRAPH_SUM, oracle.adjacency_atlas(widx)),
                       (cs._capi.GRAPH_SNN_JACCARD, oracle.snn_jaccard(widx))):
        got = ctx.graph_build(tab, n, k, mode=mode).to_scipy()
        got.sort_indices()
        assert np.array_equal(got.indptr, want.indptr)
        assert np.array_equal(got.indices, want.indices)          # SNN graph edges identical
        assert np.array_equal(got.data, want.data.astype(np.float64))
        assert (got != got.T).nnz == 0                            # symmetric
    pruned = ctx.graph_build(tab, n, k, mode=cs._capi.GRAPH_SNN_JACCARD, snn_prune=1 / 15).to_scipy()
    wantp = oracle.snn_jaccard(widx, prune=1 / 15)
    assert np.array_equal(pruned.indices, wantp.indices) and np.array_equal(pruned.data, wantp.data)


def test_graph_rejects_ids_out_of_range(cs, ctx):
    idx = np.array([[1, 2], [0, 2], [0, 7]], dtype=np.int32)       # 7 >= n
    for mode in (cs._capi.GRAPH_SUM, cs._capi.GRAPH_SNN_JACCARD):
        with pytest.raises(cs.CellScopesCudaError) as ei:
            ctx.graph_build(ctx.vec_from(idx.ravel()), 3, 2, 0, 3, mode)
        assert ei.value.code == -1


def test_pinned_results_outlive_their_context(cs):
    """arrays returned in page-locked blocks stay valid after Context.close()"""
    c = cs.Context(0)
    v = c.vec_from(np.arange(1 << 18, dtype=np.float64))      # 2 MiB -> pinned download
    out = v.download()
    v.free()
    c.close()
    assert out[12345] == 12345.0 and out.sum() == (1 << 18) * ((1 << 18) - 1) / 2
    del out


def test_graph_column_shards_concatenate(cs, oracle, ctx):
    n, k = 1200, 12
    widx, _ = oracle.knn_exact(np.random.default_rng(2).normal(size=(n, 8)), k)
    tab = ctx.vec_from(widx.astype(np.int32).ravel())
    full = ctx.graph_build(tab, n, k, mode=cs._capi.GRAPH_SUM).to_scipy()
    parts = [ctx.graph_build(tab, n, k, lo, hi, cs._capi.GRAPH_SUM).to_scipy() for lo, hi in ((0, 400), (400, 1200))]
    cat = sp.hstack(parts).tocsc()
    assert (cat != full).nnz == 0


# ---- device generator -----------------------------------------------------------------------
def test_device_generator_density_and_determinism(cs, ctx):
    p = cs.synth.make_params(3000, 50000, density=0.05, seed=1234)
    a = ctx.generate_counts(p, 1000, 3000)
    b = ctx.generate_counts(p, 1000, 3000)
    ma, mb = a.to_scipy(), b.to_scipy()
    assert (ma != mb).nnz == 0
    dens = ma.nnz / (ma.shape[0] * ma.shape[1])
    assert 0.03 < dens < 0.08, dens
    assert np.all(ma.data >= 1) and np.all(ma.data == np.round(ma.data))
    assert np.all(np.diff(ma.indptr) > 0)
    # a shard is a slice of the whole: cells 1000..3000 of [0,4000) equal the direct generation
    whole = ctx.generate_counts(p, 0, 4000).to_scipy()
    assert (whole[:, 1000:3000] != ma).nnz == 0
    # statistically the same model as the NumPy twin
    twin = cs.synth.sample_counts(p, 1000, 3000)
    assert abs(twin.nnz - ma.nnz) / ma.nnz < 0.02


def test_vec_download_as_converts_on_the_device(cs, ctx):
    """csc_vec_download_as: the 1-based / Int64 neighbour table and the Float64 distance table of knn_matrices
    (processing.jl:196-201) leave the device already converted; bit-exact against the host conversion."""
    rng = np.random.default_rng(5)
    for n in (0, 1, 1000, 300_001):          # below and above the pinned-buffer threshold
        i32 = rng.integers(0, 2 ** 31 - 2, size=n, dtype=np.int32)
        f32 = rng.normal(size=n).astype(np.float32)
        vi, vf = ctx.vec_from(i32), ctx.vec_from(f32)
        assert np.array_equal(vi.download_as(np.int32, add=1), i32 + 1)
        a = vi.download_as(np.int64, add=1)
        assert a.dtype == np.int64 and np.array_equal(a, i32.astype(np.int64) + 1)
        b = vf.download_as(np.float64)
        assert b.dtype == np.float64 and np.array_equal(b, f32.astype(np.float64))
        assert np.array_equal(vf.download_as(np.float32), f32)
    with pytest.raises(Exception):
        ctx.vec_from(np.zeros(4, np.float32)).download_as(np.int64)


# ---- multi-GPU plumbing on one GPU ------------------------------------------------------------
def test_nccl_collectives_inside_the_library_single_rank(cs, ctx):
    """csc_comm_*: NCCL is dlopen'ed by the library and the collectives run on the context's stream.  One rank is all
    a 1-GPU box can offer (the N > 1 equality is checked by bench.py's `one_vs_n` diag and tools/multigpu_check.py):
    a world of one must leave an all-reduce unchanged and an all-gather equal to its input, through the real NCCL calls."""
    c = cs.Context(0)
    comm = cs.NcclComm(c, 0, 1, lambda uid: uid)
    info = c.comm_info()
    assert info["world"] == 1 and info["rank"] == 0 and info["nccl_version"] >= 21800
    v = c.vec_from(np.arange(1000, dtype=np.float64))
    comm.allreduce_sum(v)
    assert np.array_equal(v.download(), np.arange(1000.0))
    d = c.dense_upload(np.arange(600, dtype=np.float32).reshape(200, 3))
    g = comm.allgather_rows(c, d, np.array([0, 200]))
    assert g.shape == (200, 3) and np.array_equal(g.download(np.float32), np.arange(600, dtype=np.float32).reshape(200, 3))
    t = c.vec_from(np.arange(80, dtype=np.int32))
    gi = comm.allgather_i32(c, t, np.array([0, 20]), 4)
    assert np.array_equal(gi.download(), np.arange(80, dtype=np.int32))
    with pytest.raises(cs.CellScopesCudaError):
        comm.allgather_rows(c, d, np.array([0, 199]))          # offsets disagree with the local row count
    with pytest.raises(cs.CellScopesCudaError):
        c.comm_init(c.comm_unique_id(), 0, 1)                   # a context has one communicator
    c.close()


@pytest.mark.parametrize("nfeat,k,smax", [(2000, 50, 10.0), (1200, 20, 3.0), (2000, 64, 10.0), (400, 7, 10.0)])
def test_split_transform_equals_sparse_and_dense_transform(cs, oracle, ctx, medium_counts, nfeat, k, smax):
    """csc_pca_transform_split (tensor-core GEMM over the fp16 hi/lo operands kept by csc_scale_gram_partial) against the
    sparse transform, the dense CUDA-core transform and the fp64 product of the downloaded scaled matrix."""
    raw = ctx.sparse_from_scipy(medium_counts)
    G, N = medium_counts.shape
    st = ctx.vec(2 * G)
    norm = ctx.normalize(raw, 1e4, 0, 1.0, False, st)
    feats = ctx.hvg_finish(st, G, N, nfeat, 0.3)["features"]
    H = np.unique(feats).size
    part = ctx.vec(2 * H)
    ctx.scale_stats_partial(norm, feats, part)
    colsum, gram = ctx.vec(H), ctx.vec(H * H)
    split = ctx.scale_gram_partial(norm, feats, part, N, colsum, gram, smax, keep_operands=True)
    assert split is not None
    model = ctx.pca_solve(colsum, gram, H, N, k)
    a = ctx.pca_transform_split(split, model).download(np.float64)
    b = ctx.pca_transform_sparse(norm, feats, part, N, model, smax).download(np.float64)
    dense = ctx.scale_fill(norm, feats, part, N, smax)
    c = ctx.pca_transform(dense, model).download(np.float64)
    info = model.info()
    z = dense.download(np.float64)
    want = (z - info["mean"][None, :]) @ info["loadings"]
    scale = np.sqrt((want ** 2).mean(axis=0))
    # the tensor core's adds into its fp32 accumulator truncate: one accumulator over 2000 genes cost 6.8e-5 of the
    # component's rms (measured); the gene axis is therefore cut into four accumulators that are added round-to-nearest
    for name, got, tol in (("split", a, 3e-5), ("sparse", b, 2e-5), ("dense", c, 2e-5)):
        err = np.abs(got - want).max(axis=0) / scale
        assert err.max() < tol, (name, err.max())
    assert a.shape == (N, k)
