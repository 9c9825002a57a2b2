"""Data formats on the way into the path, mirroring the reference's readers and helpers:

    read_10x(tenx_dir; version, min_gene, min_cell)        src/fileio.jl:1-37
    merge_matrices(matrices, row_indices)                  src/integration/int_utils.jl:1-23
    run_tf_idf(mtx | RawCountObject; scale_factor)         src/scatac/atac_processing.jl:1-13 (+ wrapper :15-20)

The matrix work (MatrixMarket parsing on the host threads of the library, COO -> CSC assembly, row/column filters, the
merge and the TF-IDF transform) runs through libcellscopes_cuda.so; names and bookkeeping stay here.

    read_10x_h5(h5file_path)                               src/fileio.jl:353-364

is host I/O in the reference too (HDF5.jl's h5read of six datasets); here it needs ``h5py``, which this image does not
have: the function raises ImportError without it, and the arrays it reads enter the device through csc_sparse_upload like
any other matrix.  (In the Julia drop-in the reference's own read_10x_h5 stands unchanged.)"""
from __future__ import annotations

import csv
import gzip
import os

import numpy as np
import scipy.sparse as sp

from . import _capi
from .objects import LazyDeviceMatrix, NormCountObject, RawCountObject
from .processing import default_context


def check_duplicates(arr):
    """utils.jl:31-46: names in order of first appearance and, per duplicated name, the position of its SECOND occurrence."""
    seen, uniq, removed = set(), [], {}
    for i, v in enumerate(arr):
        if v not in seen:
            seen.add(v)
            uniq.append(v)
        elif v not in removed:
            removed[v] = i
    return uniq, removed


def _read_tsv_column(path, col):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt", newline="") as f:
        return [row[col] for row in csv.reader(f, delimiter="\t") if row]


def read_10x(tenx_dir, version="v2", min_gene=0, min_cell=0, ctx=None, keep_on_device=True):
    """fileio.jl:1-37.  v2: matrix.mtx / barcodes.tsv / genes.tsv; v3: the .gz files with features.tsv.gz.  Genes are
    named by the SECOND column, cells by the first (:20-21); genes with row sum > min_cell and cells with column sum >
    min_gene are kept (:22-28); the second occurrence of a duplicated gene name is dropped (:29-34)."""
    if version == "v2":
        mtx, cells_f, genes_f = "matrix.mtx", "barcodes.tsv", "genes.tsv"
    elif version == "v3":
        mtx, cells_f, genes_f = "matrix.mtx.gz", "barcodes.tsv.gz", "features.tsv.gz"
    else:
        raise ValueError("version can only be v2 or v3!")
    c = ctx or default_context()
    dev = c.read_mtx(os.path.join(tenx_dir, mtx))
    cells = _read_tsv_column(os.path.join(tenx_dir, cells_f), 0)
    genes = _read_tsv_column(os.path.join(tenx_dir, genes_f), 1)
    g, n = dev.shape
    if len(genes) != g or len(cells) != n:
        raise ValueError(f"matrix is {g} x {n} but there are {len(genes)} gene names and {len(cells)} barcodes")
    # row / column sums on the device (exact for integer counts)
    st = c.vec(2 * g, np.float64)
    c.gene_stats_partial(dev, st)
    row_sum = st.download()[:g]
    col_sum = dev.column_sums()
    gene_kept = np.flatnonzero(row_sum > min_cell)
    cell_kept = np.flatnonzero(col_sum > min_gene)
    genes = [genes[i] for i in gene_kept]
    cells = [cells[i] for i in cell_kept]
    uniq, removed = check_duplicates(genes)
    drop = set(removed.values())
    keep_pos = [i for i in range(len(genes)) if i not in drop]
    rows = gene_kept[keep_pos]
    if rows.size != g or cell_kept.size != n:
        dev = c.select_rows(dev, rows) if rows.size != g else dev
        if cell_kept.size != n:
            dev = c.select_cols(dev, cell_kept)
    m = LazyDeviceMatrix(dev, "sparse", dev.shape) if keep_on_device else dev.to_scipy()
    return RawCountObject(m, cells, uniq)


def read_10x_h5(h5file_path, ctx=None, keep_on_device=True):
    """fileio.jl:353-364: the CSC arrays ``matrix/{data,indices,indptr,shape}`` (0-based in the file), gene names from
    ``matrix/features/name``, cell names from ``matrix/barcodes``; no filtering.  The reference rebuilds the matrix through
    ``sparse(Matrix(...))`` (:361), which drops stored zeros: so does this."""
    try:
        import h5py
    except ImportError as e:                               # no HDF5 library: say so instead of guessing at the format
        raise ImportError("read_10x_h5 needs the h5py package (HDF5); read_10x reads the MatrixMarket form of the same "
                          "data without it") from e

    def names(ds):
        return [x.decode() if isinstance(x, (bytes, np.bytes_)) else str(x) for x in ds[...]]

    with h5py.File(h5file_path, "r") as f:
        data = np.asarray(f["matrix/data"][...])
        indices = np.asarray(f["matrix/indices"][...])
        indptr = np.asarray(f["matrix/indptr"][...])
        shape = [int(v) for v in f["matrix/shape"][...]]
        genes = names(f["matrix/features/name"])
        cells = names(f["matrix/barcodes"])
    if len(shape) != 2 or indptr.size != shape[1] + 1 or len(genes) != shape[0] or len(cells) != shape[1]:
        raise ValueError(f"{h5file_path}: matrix/shape {shape} does not match indptr ({indptr.size}), the {len(genes)} gene "
                         f"names or the {len(cells)} barcodes")
    m = sp.csc_matrix((data, indices.astype(np.int64), indptr.astype(np.int64)), shape=(shape[0], shape[1]))
    m.eliminate_zeros()
    if not keep_on_device:
        return RawCountObject(m, cells, genes)
    dev = (ctx or default_context()).sparse_from_scipy(m)
    return RawCountObject(LazyDeviceMatrix(dev, "sparse", dev.shape), cells, genes)


def merge_matrices(matrices, row_indices, ctx=None, keep_on_device=False):
    """int_utils.jl:1-23: ``matrices`` (genes_i x cells_i sparse) side by side on the sorted union of ``row_indices``
    (one list of row names per matrix).  Returns (merged matrix with Float64 values, all_row_indices)."""
    if len(matrices) != len(row_indices):
        raise ValueError("one list of row names per matrix")
    c = ctx or default_context()
    all_rows = sorted(set(n for names in row_indices for n in names))
    pos = {n: i for i, n in enumerate(all_rows)}
    devs, maps = [], []
    for m, names in zip(matrices, row_indices):
        if m.shape[0] != len(names):
            raise ValueError("row names do not match the matrix")
        devs.append(m if isinstance(m, _capi.Sparse) else c.sparse_from_scipy(sp.csc_matrix(m)))
        maps.append(np.array([pos[x] for x in names], dtype=np.int64))
    out = c.sparse_merge(devs, maps, len(all_rows))
    return (out if keep_on_device else out.to_scipy()), all_rows


def run_tf_idf(x, scale_factor=10000, ctx=None):
    """atac_processing.jl:1-13 on a matrix; :15-20 on a RawCountObject (returns a NormCountObject with norm_method
    "TF-IDF")."""
    c = ctx or default_context()
    if isinstance(x, RawCountObject):
        m = x.count_mtx
        out = run_tf_idf(m, scale_factor, ctx)
        return NormCountObject(out, x.cell_name, x.gene_name, scale_factor, "TF-IDF", 1)
    m = sp.csc_matrix(x) if not sp.issparse(x) else x.tocsc()
    m.sort_indices()
    dev = c.sparse_from_scipy(m)
    st = c.vec(2 * m.shape[0], np.float64)
    c.gene_stats_partial(dev, st)
    val = c.tf_idf(dev, st, m.shape[1], scale_factor).download(values_only=True)
    return sp.csc_matrix((val, m.indices.copy(), m.indptr.copy()), shape=m.shape)
