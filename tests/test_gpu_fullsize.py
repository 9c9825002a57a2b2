"""Parity at BASELINE.json's full single-GPU size (configs[1]: 500k cells x 5k genes, 5 % dense) through properties
that do not need the oracle to finish a 500k-cell run: additivity of the moment tables over cell shards, the
normalisation round trip, orthogonality / variance of the embedding, a brute-force fp64 check of sampled kNN rows,
sortedness, and the exact checksum of the summed adjacency.  The CUDA path only; one run shared by all tests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G, N, H, K_PCS, K_NN = 5000, 500_000, 2000, 50, 20


@pytest.fixture(scope="module")
def full(cs, ctx):
    import torch
    p = cs.synth.make_params(G, N, density=0.05, seed=1234)
    raw = ctx.generate_counts(p, 0, N)
    params = cs.pipeline.PathParams(n_features=H, n_pcs=K_PCS, k=K_NN, graph_mode="sum", keep_norm=True)
    out = cs.pipeline.run_path(ctx, cs.pipeline.LocalComm(), raw, N, 0, params)
    ctx.synchronize()
    yield {"p": p, "raw": raw, "out": out, "torch": torch}
    del out, raw


def test_moment_tables_add_up_over_cell_shards(cs, ctx, full):
    """sum_x and sum_x^2 per gene are integers: the tables of three ragged cell shards add up to the table of the whole
    matrix EXACTLY (the property the multi-GPU all-reduce relies on), and so does the count of stored entries."""
    whole = full["out"]["gene_stats"].download()
    acc = np.zeros_like(whole)
    nnz = 0
    for lo, hi in [(0, 123_457), (123_457, 400_001), (400_001, N)]:
        shard = ctx.generate_counts(full["p"], lo, hi)
        st = ctx.vec(2 * G, np.float64)
        ctx.normalize(shard, 1e4, cs.pipeline.norm_code("logarithm"), 1.0, False, st)
        acc += st.download()
        nnz += shard.nnz
    assert nnz == full["raw"].nnz
    assert np.array_equal(acc, whole)


def test_two_lane_upload_round_trip_at_full_size(cs, ctx, full):
    """The 125 M stored entries of this matrix as the reference holds them (Int64 indices 1-based, Float64 values: 2 GB) go
    through csc_sparse_upload's two-lane route (taken from 32 M entries on: host threads narrow from the front, the copy
    engine carries wide pieces from the back) and must come back exactly: encode -> decode round trip over the whole
    matrix, wherever the two lanes met; then the same with Int64 values and with a value no uint16 can hold in every
    few-millionth place (pieces falling back to fp32)."""
    raw = full["raw"]
    cp, rv, val = raw.download(0, N, 1)                       # Int64 colptr / rowval (1-based), Float64 values
    assert cp.dtype == np.int64 and rv.dtype == np.int64 and val.dtype == np.float64 and raw.nnz >= (32 << 20)
    for values in (val, val.astype(np.int64)):
        up = ctx.sparse_upload(G, N, cp, rv, values, index_base=1)
        cp2, rv2, val2 = up.download(0, N, 1)
        assert np.array_equal(cp2, cp) and np.array_equal(rv2, rv) and np.array_equal(val2, val)
        del up, cp2, rv2, val2
    odd = val.copy()
    odd[::3_000_017] = 70000.5
    up = ctx.sparse_upload(G, N, cp, rv, odd, index_base=1)
    assert np.array_equal(up.download(0, N, 1, values_only=True), odd.astype(np.float32).astype(np.float64))
    del up
    rv[rv.size // 3] = G + 1                                  # one row index out of range, somewhere in the middle
    with pytest.raises(cs.CellScopesCudaError):
        ctx.sparse_upload(G, N, cp, rv, odd, index_base=1)


def test_normalisation_round_trip(cs, ctx, full):
    """expm1(norm) / sf * colsum recovers the integer counts on a slice of cells in the middle of the matrix."""
    lo, hi = 250_000, 252_000
    cp, rv, val = full["raw"].download(lo, hi, 0)
    ncp, nrv, nval = full["out"]["norm"].download(lo, hi, 0)
    assert np.array_equal(cp, ncp) and np.array_equal(rv, nrv)
    colsum = np.add.reduceat(val, cp[:-1])
    rec = np.expm1(nval) / 1e4 * np.repeat(colsum, np.diff(cp))
    assert np.allclose(rec, val, rtol=2e-5)        # fp32 log1p/expm1 round trip


def test_embedding_is_a_pca(cs, ctx, full):
    """Columns of the embedding are centred, mutually orthogonal and carry the principal variances (fp64 on the GPU)."""
    torch = full["torch"]
    emb = torch.as_tensor(full["out"]["embedding"], device="cuda:0").double()
    info = full["out"]["pca"].info()
    pv = torch.as_tensor(info["principalvars"], device="cuda:0")
    mean = emb.mean(0)
    cov = (emb - mean).T @ (emb - mean) / (N - 1)
    lam1 = float(pv[0])
    assert float(mean.abs().max()) < 2e-4 * lam1 ** 0.5
    assert float((cov.diagonal() - pv).abs().max()) < 2e-4 * lam1
    off = cov - torch.diag(cov.diagonal())
    assert float(off.abs().max()) < 2e-4 * lam1
    assert bool((pv[:-1] >= pv[1:]).all())


def test_sampled_knn_rows_against_fp64_brute_force(cs, ctx, full):
    """512 query rows spread over the matrix: exact cosine top-20 against ALL 500k keys in fp64 (torch on the GPU, no
    library code), ids must agree wherever the 20th and 21st neighbours are separated by more than fp32 noise."""
    torch = full["torch"]
    emb = torch.as_tensor(full["out"]["embedding"], device="cuda:0").double()
    unit = emb / emb.norm(dim=1, keepdim=True)
    idx = full["out"]["knn_idx"].download().reshape(N, K_NN)
    dist = full["out"]["knn_dist"].download().reshape(N, K_NN)
    rows = np.linspace(0, N - 1, 512).astype(np.int64)
    sim = unit[torch.as_tensor(rows, device="cuda:0")] @ unit.T
    sim[torch.arange(len(rows)), torch.as_tensor(rows, device="cuda:0")] = -2.0       # self excluded
    top = torch.topk(sim, K_NN + 1, dim=1)
    ref_idx = top.indices.cpu().numpy()
    ref_sim = top.values.cpu().numpy()
    checked = 0
    for i, r in enumerate(rows):
        if ref_sim[i, K_NN - 1] - ref_sim[i, K_NN] < 1e-6:
            continue                                                                 # a tie at the boundary
        assert set(idx[r].tolist()) == set(ref_idx[i, :K_NN].tolist()), f"row {r}"
        assert np.allclose(dist[r], 1.0 - ref_sim[i, :K_NN], atol=2e-6)
        checked += 1
    assert checked > 480
    assert ctx.knn_last_fallback_rows() < N // 100


def test_knn_table_is_sorted_and_in_range(cs, ctx, full):
    idx = full["out"]["knn_idx"].download().reshape(N, K_NN)
    dist = full["out"]["knn_dist"].download().reshape(N, K_NN)
    assert idx.min() >= 0 and idx.max() < N
    assert not (idx == np.arange(N)[:, None]).any()                     # no self neighbours
    assert (np.diff(dist, axis=1) >= 0).all()                           # ascending distance
    assert (np.sort(idx, axis=1)[:, 1:] != np.sort(idx, axis=1)[:, :-1]).all()      # no duplicates in a row


def test_summed_adjacency_checksum_and_symmetry(cs, ctx, full):
    """A = K + K' for the kNN indicator K: entries are 1 or 2, the grand total is exactly 2 N k, the diagonal is empty and
    A is symmetric (its transpose has the same sorted triplets)."""
    import scipy.sparse as sp
    colptr, rowval, val = full["out"]["graph"].download(0)
    a = sp.csc_matrix((val, rowval, colptr), shape=(N, N))
    assert set(np.unique(a.data).tolist()) <= {1, 2}
    assert int(a.data.sum()) == 2 * N * K_NN
    assert a.diagonal().sum() == 0
    at = a.T.tocsc()
    at.sort_indices()
    a.sort_indices()
    assert np.array_equal(a.indptr, at.indptr) and np.array_equal(a.indices, at.indices) and np.array_equal(a.data, at.data)
    # mutual pairs carry 2: their number follows from the totals
    n_mutual_entries = int((a.data == 2).sum())
    assert a.nnz == 2 * N * K_NN - n_mutual_entries
