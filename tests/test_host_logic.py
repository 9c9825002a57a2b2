"""Host-side logic that needs no GPU: partitioning, container semantics, synthetic generator, argument errors."""
import numpy as np
import pytest
import scipy.sparse as sp


def test_partition_cells_by_count(cs):
    off = cs.partition_cells(10, 3)
    assert list(off) == [0, 4, 7, 10]
    assert list(cs.partition_cells(5, 1)) == [0, 5]
    off = cs.partition_cells(2, 4)
    assert off[0] == 0 and off[-1] == 2 and np.all(np.diff(off) >= 0)
    with pytest.raises(ValueError):
        cs.partition_cells(5, 0)


def test_partition_cells_by_nnz(cs):
    colptr = np.array([0, 100, 100, 110, 120, 200, 200, 300])      # 7 cells, ragged incl. empty columns
    off = cs.partition_cells(7, 2, colptr)
    assert off[0] == 0 and off[-1] == 7
    left = colptr[off[1]] - colptr[0]
    assert abs(left - 150) <= 80
    off8 = cs.partition_cells(7, 8, colptr)
    assert off8.size == 9 and np.all(np.diff(off8) >= 0) and off8[-1] == 7


def test_objects_dimension_checks(cs):
    m = sp.csc_matrix(np.ones((3, 2)))
    with pytest.raises(ValueError, match="cell ID do not match"):
        cs.RawCountObject(m, ["a"], ["g1", "g2", "g3"])
    with pytest.raises(ValueError, match="gene name do not match"):
        cs.RawCountObject(m, ["a", "b"], ["g1"])
    r = cs.RawCountObject(m, ["a", "b"], ["g1", "g2", "g3"])
    assert r.shape == (3, 2) and r.device is None


def test_screate_filters_like_subset_matrix(cs):
    d = np.array([[0, 0, 0, 0], [1, 0, 2, 0], [0, 0, 3, 1], [2, 0, 0, 0]])
    raw = cs.RawCountObject(sp.csc_matrix(d), ["c1", "c2", "c3", "c4"], ["A", "B", "C", "B"])
    obj = cs.scRNAObject(raw)
    # zero gene A and zero cell c2 dropped (strict >), duplicated gene name B keeps its first occurrence
    assert obj.rawCount.gene_name == ["B", "C"] and obj.rawCount.cell_name == ["c1", "c3", "c4"]
    assert list(obj.metaData["nFeatures"]) == [1, 5, 1]           # total counts per cell (objects.jl:148)
    assert list(obj.metaData["nGenes"]) == [1, 2, 1]
    obj2 = cs.scRNAObject(raw, min_gene=1, prefix="s1")
    assert obj2.rawCount.cell_name == ["s1_c1", "s1_c3"]


def test_subset_count_keeps_original_order(cs):
    m = sp.csc_matrix(np.arange(12).reshape(4, 3))
    n = cs.NormCountObject(m, ["a", "b", "c"], ["g1", "g2", "g3", "g4"], 1e4, "logarithm", 1)
    s = cs.subset_count(n, genes=["g4", "g1"], cells=["c", "a"])
    assert s.gene_name == ["g1", "g4"] and s.cell_name == ["a", "c"]
    assert np.array_equal(s.count_mtx.toarray(), np.array([[0, 2], [9, 11]]))
    assert isinstance(s, cs.NormCountObject) and s.scale_factor == 1e4


def test_enum_mapping_and_errors(cs):
    assert cs.pipeline.norm_code("logarithm") == 0 and cs.pipeline.norm_code("none") == 1
    with pytest.raises(ValueError, match="Unknown normalization method"):
        cs.pipeline.norm_code("clr")
    assert cs.pipeline.metric_code("CosineDist()") == 0 and cs.pipeline.metric_code("euclidean") == 1
    with pytest.raises(ValueError):
        cs.pipeline.metric_code("manhattan")


def test_synth_params_and_sampler(cs):
    p = cs.synth.make_params(800, 5000, density=0.05, seed=1)
    q = cs.synth.make_params(800, 5000, density=0.05, seed=1)
    assert np.array_equal(p.a, q.a) and np.array_equal(p.module, q.module)
    m = cs.synth.sample_counts(p, 100, 400)
    assert m.shape == (800, 300)
    dens = m.nnz / (800 * 300)
    assert 0.03 < dens < 0.08
    assert np.all(np.diff(m.indptr) > 0), "no empty cells"
    # any cell range is reproducible anywhere
    sub = cs.synth.sample_counts(p, 150, 200)
    assert (m[:, 50:100] != sub).nnz == 0


def test_lazy_matrix_is_not_touched_by_shape(cs):
    class Boom:
        def download(self, *a, **k):
            raise AssertionError("must not download")
    lz = cs.objects.LazyDeviceMatrix(Boom(), "dense_t", (5, 7))
    n = cs.ScaleCountObject(lz, list("abcdefg"), list("vwxyz"), True, True, 10.0)
    assert n.shape == (5, 7) and n.device is lz.handle


@pytest.mark.parametrize("variant", [0, 1])
def test_upload_host_lane_narrowing(cs, variant):
    """Host lane of csc_sparse_upload (csrc/host_narrow.cpp): Int64 / Float64 pieces narrowed to uint16 when that is
    lossless, else to int32 / fp32; range violations of row indices flagged.  Variant 0 = generic loops, 1 = AVX-512 when the
    CPU has it.  Expected values are NumPy casts (the device-side narrowing kernels do the same casts)."""
    import ctypes
    lib = ctypes.CDLL(cs._capi.LIB_PATH)
    fn = lib.csc_test_host_narrow
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int,
                   ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int)]
    rng = np.random.default_rng(5)

    def run(kind, src, idx16=0, base=0, n_rows=0):
        stage = np.zeros(src.size * 4 + 64, np.uint8)
        bad = ctypes.c_int(0)
        elem = fn(variant, kind, src.ctypes.data, src.size, stage.ctypes.data, idx16, base, n_rows, ctypes.byref(bad))
        assert not stage[src.size * elem:].any()           # nothing written past the piece
        return elem, stage[:src.size * elem], bad.value

    for n in (0, 1, 15, 16, 17, 1000, 4099):
        # row indices, 1-based, 33k genes -> uint16
        idx = rng.integers(1, 33001, n).astype(np.int64)
        elem, out, bad = run(0, idx, idx16=1, base=1, n_rows=33000)
        assert (elem, bad) == (2, 0) and np.array_equal(out.view(np.uint16), (idx - 1).astype(np.uint16))
        # 100k genes -> int32
        idx = rng.integers(0, 100000, n).astype(np.int64)
        elem, out, bad = run(0, idx, idx16=0, base=0, n_rows=100000)
        assert (elem, bad) == (4, 0) and np.array_equal(out.view(np.int32), idx.astype(np.int32))
        if n:
            for wrong in (-1, 100000, 1 << 40):
                idx2 = idx.copy()
                idx2[n // 2] = wrong
                assert run(0, idx2, idx16=0, base=0, n_rows=100000)[2] == 1
                assert run(0, idx2 + 1, idx16=1, base=1, n_rows=100000)[2] == 1
        # counts as Float64 -> uint16; one non-integer / large / negative / -0.0 / NaN value sends the piece to fp32
        val = rng.integers(0, 65536, n).astype(np.float64)
        elem, out, _ = run(1, val)
        assert elem == 2 and np.array_equal(out.view(np.uint16), val.astype(np.uint16))
        for odd in (0.5, 65536.0, -3.0, -0.0, np.nan, np.inf, 1e300):
            if not n:
                break
            v2 = val.copy()
            v2[n - 1 - n // 3] = odd
            elem, out, _ = run(1, v2)
            with np.errstate(over="ignore"):
                want = v2.astype(np.float32)
            assert elem == 4 and np.array_equal(out.view(np.uint32), want.view(np.uint32)), odd
        # counts as Int64
        vi = rng.integers(0, 65536, n).astype(np.int64)
        elem, out, _ = run(2, vi)
        assert elem == 2 and np.array_equal(out.view(np.uint16), vi.astype(np.uint16))
        for odd in (-1, 65536, (1 << 53) + 1):
            if not n:
                break
            v2 = vi.copy()
            v2[n // 2] = odd
            elem, out, _ = run(2, v2)
            assert elem == 4 and np.array_equal(out.view(np.float32), v2.astype(np.float32)), odd


def test_read_10x_h5_layout_and_missing_h5py(cs, monkeypatch):
    """fileio.jl:353-364 through a stand-in for h5py (the image has no HDF5 library): 0-based CSC arrays, names from
    matrix/features/name and matrix/barcodes, stored zeros dropped like sparse(Matrix(...)) does; without h5py the
    function says what is missing."""
    import sys
    import types

    import scipy.sparse as sp
    dense = np.array([[0, 2, 0, 1], [3, 0, 0, 0], [0, 0, 5, 4]], dtype=np.int32)
    m = sp.csc_matrix(dense)
    data = m.data.copy()
    data[1] = 0                                            # an explicitly stored zero (entry (0, 1))
    store = {"matrix/data": data, "matrix/indices": m.indices.astype(np.int64), "matrix/indptr": m.indptr.astype(np.int64),
             "matrix/shape": np.array([3, 4], dtype=np.int32),
             "matrix/features/name": np.array([b"G1", b"G2", b"G3"]), "matrix/barcodes": np.array([b"AAA-1", b"CCC-1", b"GGG-1", b"TTT-1"])}

    class FakeFile:
        def __init__(self, path, mode):
            assert mode == "r"

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def __getitem__(self, key):
            return store[key]

    fake = types.ModuleType("h5py")
    fake.File = FakeFile
    monkeypatch.setitem(sys.modules, "h5py", fake)
    obj = cs.read_10x_h5("whatever.h5", keep_on_device=False)
    assert obj.gene_name == ["G1", "G2", "G3"] and obj.cell_name == ["AAA-1", "CCC-1", "GGG-1", "TTT-1"]
    want = dense.copy()
    want[0, 1] = 0
    assert np.array_equal(obj.count_mtx.toarray(), want) and obj.count_mtx.nnz == 4
    store["matrix/barcodes"] = store["matrix/barcodes"][:3]
    with pytest.raises(ValueError):
        cs.read_10x_h5("whatever.h5", keep_on_device=False)
    monkeypatch.setitem(sys.modules, "h5py", None)          # import h5py -> ImportError
    with pytest.raises(ImportError) as ei:
        cs.read_10x_h5("whatever.h5", keep_on_device=False)
    assert "h5py" in str(ei.value)
