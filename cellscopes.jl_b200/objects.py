"""Python mirror of the reference's container types (src/scrna/objects.jl:10-165).

Field names, constructor checks and error messages follow the Julia structs so that the host
code above the C ABI reads like the reference.  ``count_mtx`` is genes x cells; sparse matrices
are ``scipy.sparse.csc_matrix`` (the analogue of SparseMatrixCSC: a column is a cell).

Results that live on the GPU are held as *lazy* matrices: the object keeps the device handle and
only copies to the host when ``count_mtx`` is actually read.  That is what lets
normalize -> HVG -> scale -> PCA -> clustering run back to back without moving the 4 B/nnz
normalised values or the dense scaled matrix over PCIe, while plots and other consumers that read
the fields still see ordinary arrays.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.sparse as sp


def _check_dims(count_mtx, cell_name, gene_name):
    # objects.jl:15-20
    shape = count_mtx.shape
    if len(cell_name) != shape[1]:
        raise ValueError("The total number of cells in the count matrix and the cell ID do not match!")
    if len(gene_name) != shape[0]:
        raise ValueError("The total number of genes in the count matrix and the gene name do not match!")


class LazyDeviceMatrix:
    """A matrix whose values still live on the device.  ``.shape`` is free; ``materialize()``
    downloads once and caches.  ``kind`` is "sparse" (pattern of ``pattern`` + device values) or
    "dense_t" (device dense cell-major, exposed as genes x cells)."""

    def __init__(self, handle, kind, shape, pattern=None):
        self.handle = handle
        self.kind = kind
        self.shape = tuple(shape)
        self.pattern = pattern
        self._host = None

    def materialize(self):
        if self._host is None:
            if self.kind == "sparse":
                if self.pattern is not None:
                    val = self.handle.download(values_only=True)
                    self._host = sp.csc_matrix((val, self.pattern.indices, self.pattern.indptr), shape=self.shape)
                else:
                    self._host = self.handle.to_scipy()
            else:
                # device [cells x genes] row-major == Julia genes x cells column-major
                self._host = self.handle.download(np.float64).T
        return self._host


class _CountBase:
    def __init__(self, count_mtx, cell_name, gene_name):
        _check_dims(count_mtx, cell_name, gene_name)
        self._mtx = count_mtx
        self.cell_name = list(cell_name) if not isinstance(cell_name, np.ndarray) else cell_name
        self.gene_name = list(gene_name) if not isinstance(gene_name, np.ndarray) else gene_name

    @property
    def count_mtx(self):
        if isinstance(self._mtx, LazyDeviceMatrix):
            return self._mtx.materialize()
        return self._mtx

    @count_mtx.setter
    def count_mtx(self, m):
        self._mtx = m

    @property
    def device(self):
        """The device handle behind a lazy matrix (None if the matrix is host-only)."""
        return self._mtx.handle if isinstance(self._mtx, LazyDeviceMatrix) else None

    @property
    def shape(self):
        return self._mtx.shape


class RawCountObject(_CountBase):
    """objects.jl:10-24"""


class NormCountObject(_CountBase):
    """objects.jl:26-43"""

    def __init__(self, count_mtx, cell_name, gene_name, scale_factor, norm_method, pseudocount):
        super().__init__(count_mtx, cell_name, gene_name)
        self.scale_factor = scale_factor
        self.norm_method = norm_method
        self.pseudocount = pseudocount


class ScaleCountObject(_CountBase):
    """objects.jl:45-62"""

    def __init__(self, count_mtx, cell_name, gene_name, do_scale, do_center, scale_max):
        super().__init__(count_mtx, cell_name, gene_name)
        self.do_scale = do_scale
        self.do_center = do_center
        self.scale_max = scale_max


class PCAObject:
    """objects.jl:64-73"""

    def __init__(self, cell_embedding, pca_model, percent_var, key, method, pratio, maxoutdim):
        self.cell_embedding = cell_embedding
        self.pca_model = pca_model
        self.percent_var = percent_var
        self.key = key
        self.method = method
        self.pratio = pratio
        self.maxoutdim = maxoutdim
        self.device = None          # device-resident fp32 embedding, if any


class UMAPObject:
    """objects.jl:86-95 (only knn_data matters to this path, processing.jl:196-197)"""

    def __init__(self, cell_embedding, key, n_components, n_dimensions, n_neighbors, metric, min_dist, knn_data):
        self.cell_embedding = cell_embedding
        self.key = key
        self.n_components = n_components
        self.n_dimensions = n_dimensions
        self.n_neighbors = n_neighbors
        self.metric = metric
        self.min_dist = min_dist
        self.knn_data = knn_data


class ReductionObject:
    """objects.jl:97-102"""

    def __init__(self, pca, tsne, umap):
        self.pca = pca
        self.tsne = tsne
        self.umap = umap


class ClusteringObject:
    """objects.jl:104-111"""

    def __init__(self, clustering, metric, adj_mat, ledein_res, resolution):
        self.clustering = clustering
        self.metric = metric
        self.adj_mat = adj_mat
        self.ledein_res = ledein_res
        self.resolution = resolution
        self.knn_indices = None      # k x n, 1-based like knn_matrices (processing.jl:201,266)
        self.knn_distances = None


class VariableGeneObject:
    """objects.jl:113-117"""

    def __init__(self, var_gene, vst_data):
        self.var_gene = var_gene
        self.vst_data = vst_data


def subset_matrix(count_mat, gene_name, cell_name, min_gene, min_cell):
    """utils.jl:313-327: genes with row sum > min_cell and cells with column sum > min_gene (strict),
    then duplicated gene names are dropped (first occurrence kept)."""
    count_mat = sp.csc_matrix(count_mat)
    row_sum = np.asarray(count_mat.sum(axis=1)).ravel()
    col_sum = np.asarray(count_mat.sum(axis=0)).ravel()
    gene_kept = np.flatnonzero(row_sum > min_cell)
    cell_kept = np.flatnonzero(col_sum > min_gene)
    gene_name = [gene_name[i] for i in gene_kept]
    cell_name = [cell_name[i] for i in cell_kept]
    if gene_kept.size != count_mat.shape[0] or cell_kept.size != count_mat.shape[1]:
        count_mat = count_mat[gene_kept][:, cell_kept].tocsc()
    if len(set(gene_name)) != len(gene_name):
        seen, keep = set(), []
        for i, g in enumerate(gene_name):
            if g not in seen:
                seen.add(g)
                keep.append(i)
        gene_name = [gene_name[i] for i in keep]
        count_mat = count_mat[keep, :].tocsc()
    count_mat.sort_indices()
    return count_mat, gene_name, cell_name


def subset_count(ct_obj, genes=None, cells=None):
    """utils.jl:59-87: boolean-mask subsetting, ORIGINAL order kept."""
    all_genes, all_cells = ct_obj.gene_name, ct_obj.cell_name
    gset = set(all_genes if genes is None else genes)
    cset = set(all_cells if cells is None else cells)
    check_gene = np.fromiter((g in gset for g in all_genes), dtype=bool, count=len(all_genes))
    check_cell = np.fromiter((c in cset for c in all_cells), dtype=bool, count=len(all_cells))
    gene_name = [g for g, k in zip(all_genes, check_gene) if k]
    cell_name = [c for c, k in zip(all_cells, check_cell) if k]
    m = ct_obj.count_mtx
    new_count = m[np.flatnonzero(check_gene)][:, np.flatnonzero(check_cell)]
    if isinstance(ct_obj, NormCountObject):
        return NormCountObject(new_count, cell_name, gene_name, ct_obj.scale_factor, ct_obj.norm_method, ct_obj.pseudocount)
    if isinstance(ct_obj, ScaleCountObject):
        return ScaleCountObject(new_count, cell_name, gene_name, ct_obj.do_scale, ct_obj.do_center, ct_obj.scale_max)
    return RawCountObject(new_count, cell_name, gene_name)


class scRNAObject:
    """objects.jl:128-165.  ``imaging=True`` marks the Merfish/Cartana/Xenium behaviour of
    normalize_object (NaN -> 0, processing.jl:44-46); the other spatial containers carry the same
    fields and go through the same functions."""

    def __init__(self, raw_count: RawCountObject, meta_data=None, min_gene: int = 0, min_cell: int = 0,
                 prefix=None, postfix=None, imaging: bool = False):
        count_mat, genes, cells = subset_matrix(raw_count.count_mtx, raw_count.gene_name, raw_count.cell_name,
                                                min_gene, min_cell)
        if meta_data is None:
            n_features = np.asarray(count_mat.sum(axis=0)).ravel()
            n_genes = np.diff(count_mat.indptr) if not (count_mat.data <= 0).any() else \
                np.asarray((count_mat > 0).sum(axis=0)).ravel()
            meta_data = pd.DataFrame({"Cell_id": cells, "nFeatures": n_features, "nGenes": n_genes})
        if prefix is not None:
            cells = [f"{prefix}_{c}" for c in cells]
        if postfix is not None:
            cells = [f"{c}_{postfix}" for c in cells]
        self.rawCount = RawCountObject(count_mat, cells, genes)
        self.normCount = None
        self.scaleCount = None
        self.metaData = meta_data
        self.varGene = None
        self.dimReduction = None
        self.clustData = None
        self.undefinedData = None
        self.imaging = imaging
        self._session = None
