"""f4: the data formats on the way into the path -- read_10x (MatrixMarket, src/fileio.jl:1-37), merge_matrices
(src/integration/int_utils.jl:1-23), run_tf_idf (src/scatac/atac_processing.jl:1-13) -- on the CUDA path against the
oracle / scipy."""
import gzip
import os
import shutil

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def write_10x(tmp, m, genes, cells, version):
    os.makedirs(tmp, exist_ok=True)
    scipy.io.mmwrite(os.path.join(tmp, "matrix.mtx"), m, field="integer")
    with open(os.path.join(tmp, "barcodes.tsv"), "w") as f:
        f.write("".join(f"{c}\n" for c in cells))
    gname = "genes.tsv" if version == "v2" else "features.tsv"
    with open(os.path.join(tmp, gname), "w") as f:
        f.write("".join(f"ENSG{i:08d}\t{g}\tGene Expression\n" for i, g in enumerate(genes)))
    if version == "v3":
        for name in ("matrix.mtx", "barcodes.tsv", "features.tsv"):
            with open(os.path.join(tmp, name), "rb") as src, gzip.open(os.path.join(tmp, name + ".gz"), "wb") as dst:
                shutil.copyfileobj(src, dst)
            os.remove(os.path.join(tmp, name))


@pytest.mark.parametrize("version", ["v2", "v3"])
def test_read_10x_matches_reference_semantics(cs, oracle, ctx, tmp_path, version):
    rng = np.random.default_rng(5)
    g, n = 700, 450
    m = sp.random(g, n, density=0.07, random_state=3, data_rvs=lambda k: rng.integers(1, 40, k)).tocsc().astype(np.int64)
    m[13, :] = 0
    m[:, 7] = 0                    # an all-zero gene and an all-zero cell: dropped by the strict > 0 filters (:22-28)
    m.eliminate_zeros()
    genes = [f"G{i}" for i in range(g)]
    genes[100] = genes[40]          # a duplicated name: its second occurrence is dropped (:29-34)
    cells = [f"AAAC{i:05d}-1" for i in range(n)]
    d = str(tmp_path / version)
    write_10x(d, m, genes, cells, version)
    obj = cs.read_10x(d, version=version, ctx=ctx)
    got = obj.count_mtx                                   # materialises the device matrix
    keep_g = [i for i in range(g) if i not in (13, 100)]
    keep_c = [j for j in range(n) if j != 7]
    want = m[keep_g][:, keep_c]
    assert got.shape == want.shape and (got != want).nnz == 0
    assert obj.gene_name == [genes[i] for i in keep_g] and obj.cell_name == [cells[j] for j in keep_c]
    # thresholds are strict (> min_cell / > min_gene)
    obj2 = cs.read_10x(d, version=version, min_gene=int(np.median(np.asarray(m.sum(axis=0)))), min_cell=30, ctx=ctx)
    rs, csum = np.asarray(m.sum(axis=1)).ravel(), np.asarray(m.sum(axis=0)).ravel()
    assert obj2.shape[1] == int((csum > int(np.median(csum))).sum())
    assert obj2.shape[0] <= int((rs > 30).sum())


def test_read_mtx_general_real_pattern_and_duplicates(cs, ctx, tmp_path):
    """sparse(I, J, V) semantics of MatrixMarket.mmread: duplicates are summed, explicit zeros stay stored; real and
    pattern fields; comments and blank lines; entries in any order."""
    p = tmp_path / "m.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% a comment\n%another\n4 3 7\n"
                 "4 3 2.5\n1 1 1\n\n2 2 0\n1 1 3\n3 1 -1.5\n  2 3 1e2\n4 3 0.5\n")
    got = ctx.read_mtx(str(p)).to_scipy().toarray()
    want = np.zeros((4, 3))
    want[3, 2], want[0, 0], want[2, 0], want[1, 2] = 3.0, 4.0, -1.5, 100.0
    assert np.array_equal(got, want)
    assert ctx.read_mtx(str(p)).nnz == 5                   # the explicit zero at (2, 2) stays stored
    q = tmp_path / "p.mtx"
    q.write_text("%%MatrixMarket matrix coordinate pattern general\n3 2 3\n1 1\n3 2\n2 1\n")
    assert np.array_equal(ctx.read_mtx(str(q)).to_scipy().toarray(), np.array([[1, 0], [1, 0], [0, 1]], dtype=float))
    for bad in ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
                "%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n",
                "%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n",
                "no banner\n"):
        b = tmp_path / "bad.mtx"
        b.write_text(bad)
        with pytest.raises(cs.CellScopesCudaError):
            ctx.read_mtx(str(b))


def test_read_mtx_large_file_parsed_on_many_threads(cs, oracle, ctx, tmp_path):
    p = cs.synth.make_params(3000, 4000, density=0.05, seed=2)
    m = cs.synth.sample_counts(p)
    path = str(tmp_path / "big.mtx")
    scipy.io.mmwrite(path, m, field="integer")
    got = ctx.read_mtx(path).to_scipy()
    assert got.shape == m.shape and (got != m).nnz == 0
    assert np.array_equal(got.indptr, m.indptr) and np.array_equal(got.indices, m.indices), "row indices sorted within columns"


def test_merge_matrices_matches_oracle(cs, oracle, ctx):
    rng = np.random.default_rng(9)
    names = [f"gene{i}" for i in range(60)]
    mats, rows = [], []
    for k, (g, n) in enumerate([(40, 25), (33, 50), (50, 7)]):
        sel = list(rng.choice(60, size=g, replace=False))
        a = sp.random(g, n, density=0.3, random_state=k, data_rvs=lambda q: rng.integers(0, 5, q)).tocsc().astype(np.int64)
        mats.append(a)                                     # holds stored zeros (value 0 drawn): dropped by `val != 0`
        rows.append([names[i] for i in sel])
    got, got_rows = cs.merge_matrices(mats, rows, ctx=ctx)
    want, want_rows = oracle.merge_matrices(mats, rows)
    assert got_rows == want_rows == sorted(set(sum(rows, [])))
    assert got.shape == want.shape and got.dtype == np.float64
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.array_equal(got.data, want.data)
    with pytest.raises(ValueError):
        cs.merge_matrices(mats, rows[:2], ctx=ctx)


def test_tf_idf_matches_oracle(cs, oracle, ctx):
    p = cs.synth.make_params(2500, 1800, density=0.04, seed=6)
    m, _, _ = oracle.subset_matrix(cs.synth.sample_counts(p))
    got = cs.run_tf_idf(m, ctx=ctx)
    want = oracle.run_tf_idf(m)
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.allclose(got.data, want.data, rtol=1e-5)     # normalised values within 1e-5 relative
    got2 = cs.run_tf_idf(m, scale_factor=100, ctx=ctx)
    assert np.allclose(got2.data, oracle.run_tf_idf(m, 100).data, rtol=1e-5)
    raw = cs.RawCountObject(m, [f"c{i}" for i in range(m.shape[1])], [f"p{i}" for i in range(m.shape[0])])
    nc = cs.run_tf_idf(raw, ctx=ctx)
    assert nc.norm_method == "TF-IDF" and nc.scale_factor == 10000 and nc.shape == m.shape


def test_column_sums_and_select_cols(cs, ctx, small_counts):
    dev = ctx.sparse_from_scipy(small_counts)
    assert np.array_equal(dev.column_sums(), np.asarray(small_counts.sum(axis=0)).ravel())
    cols = np.array([5, 0, 999, 5, 17])
    sub = ctx.select_cols(dev, cols).to_scipy()
    assert (sub != small_counts[:, cols]).nnz == 0
    with pytest.raises(cs.CellScopesCudaError):
        ctx.select_cols(dev, np.array([small_counts.shape[1]]))
