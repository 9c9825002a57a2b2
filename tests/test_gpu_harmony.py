"""f2: run_harmony (src/integration/int_processing.jl:1-12, HarmonyObject int_objects.jl:115-236, int_utils.jl:49-199)
on the CUDA path against the oracle's line-by-line restatement.  Julia's RNG cannot be reproduced, so both sides take
the same k-means centres and walk the update order of the same counter-based hash; everything else is deterministic
fp64 arithmetic and must agree to rounding."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def batchy_embedding(n=6000, d=20, n_types=7, n_batch=3, seed=0):
    rng = np.random.default_rng(seed)
    cent = rng.normal(size=(n_types, d)) * 2.5
    lab = rng.integers(0, n_types, n)
    batch = rng.integers(0, n_batch, n)
    shift = rng.normal(size=(n_batch, d)) * 0.8
    z = cent[lab] + rng.normal(size=(n, d)) + shift[batch]
    return z, lab, batch


@pytest.mark.parametrize("n,d,k,n_batch", [(6000, 20, 25, 3), (3001, 50, 100, 2), (900, 7, 5, 4)])
def test_harmony_matches_oracle(cs, oracle, ctx, n, d, k, n_batch):
    z, _, batch = batchy_embedding(n, d, 7, n_batch, seed=n)
    zc = oracle.harmony_zcos(z)
    idx = cs.integration.seed_cells(n, k, 123)
    y0, _, _, _ = oracle.kmeans_lloyd(zc, k, idx)
    want = oracle.run_harmony(z, batch, y0, seed=123)
    got = ctx.harmony(z, batch, n_batch, y0, np.full(n_batch, 2.0), np.full(n_batch, 1.0), np.full(k, 0.1), seed=123, want_r=True)
    assert np.array_equal(got["kmeans_rounds"], want["kmeans_rounds"])
    assert np.allclose(got["objective_harmony"], want["objective_harmony"], rtol=1e-9)
    assert np.allclose(got["R"], want["R"], rtol=1e-7, atol=1e-12)
    scale = np.abs(want["Z_corr"]).max()
    assert np.abs(got["Z_corr"] - want["Z_corr"]).max() <= 1e-9 * scale
    assert np.abs(want["Z_corr"] - z).max() > 1e-2 * scale, "the correction must actually move the cells"


def test_harmony_options_match_oracle(cs, oracle, ctx):
    """theta / lamb / sigma / block_size / iteration limits; a cell count that leaves a tail outside the blocks."""
    n, d, k = 2503, 12, 12
    z, _, batch = batchy_embedding(n, d, 5, 3, seed=5)
    y0, _, _, _ = oracle.kmeans_lloyd(oracle.harmony_zcos(z), k, cs.integration.seed_cells(n, k, 7))
    kw = dict(theta=1.0, lamb=0.5, sigma=0.2, block_size=0.3, max_iter_harmony=4, max_iter_kmeans=6, seed=99)
    want = oracle.run_harmony(z, batch, y0, **kw)
    got = ctx.harmony(z, batch, 3, y0, np.full(3, 1.0), np.full(3, 0.5), np.full(k, 0.2), block_size=0.3, max_iter_harmony=4,
                      max_iter_kmeans=6, seed=99)
    assert np.array_equal(got["kmeans_rounds"], want["kmeans_rounds"])
    assert np.allclose(got["objective_harmony"], want["objective_harmony"], rtol=1e-9)
    assert np.abs(got["Z_corr"] - want["Z_corr"]).max() <= 1e-9 * np.abs(want["Z_corr"]).max()


def test_kmeans_matches_oracle(cs, oracle, ctx):
    z, _, _ = batchy_embedding(5000, 16, 9, 2, seed=3)
    zc = oracle.harmony_zcos(z)
    idx = cs.integration.seed_cells(5000, 9, 1)
    cen, lab, obj, it = ctx.kmeans(z, 9, idx, unit_rows=True)
    wcen, wlab, wobj, wit = oracle.kmeans_lloyd(zc, 9, idx)
    assert it == wit and np.isclose(obj, wobj, rtol=1e-10)
    assert np.array_equal(lab, wlab) and np.allclose(cen, wcen, rtol=1e-10, atol=1e-13)
    cen2, _, _, _ = ctx.kmeans(zc, 9, idx, unit_rows=False)
    assert np.allclose(cen2, wcen, rtol=1e-10, atol=1e-13)


def test_run_harmony_object_api(cs, oracle, ctx):
    """run_harmony(sc_obj, :batch): the corrected matrix replaces pca.cell_embedding (N x d Float64) and the batches mix."""
    z, lab, batch = batchy_embedding(4000, 15, 6, 2, seed=11)
    obj = cs.scRNAObject.__new__(cs.scRNAObject)
    obj.metaData = pd.DataFrame({"Cell_id": [f"c{i}" for i in range(4000)], "batch": [f"s{b}" for b in batch]})
    pca = cs.PCAObject(z.copy(), None, None, "PC", "svd", 1, 15)
    obj.dimReduction = cs.ReductionObject(pca, None, None)
    obj._session = None
    obj = cs.run_harmony(obj, "batch", ctx=ctx, verbose=False)
    out = obj.dimReduction.pca.cell_embedding
    assert out.shape == (4000, 15) and out.dtype == np.float64

    def batch_gap(m):
        return np.mean([np.linalg.norm(m[(lab == t) & (batch == 0)].mean(0) - m[(lab == t) & (batch == 1)].mean(0)) for t in range(6)])
    assert batch_gap(out) < 0.5 * batch_gap(z), (batch_gap(out), batch_gap(z))
    with pytest.raises(ValueError):
        cs.harmony_matrix(z[:100], batch, ctx=ctx)
