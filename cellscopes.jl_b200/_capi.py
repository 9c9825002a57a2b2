"""ctypes binding of libcellscopes_cuda.so (include/cellscopes_cuda.h).

This is the only way Python reaches the CUDA path, and it is the same C ABI the Julia shim
(julia/CellScopesCUDA.jl) binds with ``ccall``.  There is no CPU fallback: if the shared library
is missing or no B200 is visible, calls raise ``CellScopesCudaError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcellscopes_cuda.so")

# enums of the header
DT_F32, DT_F64, DT_I32, DT_I64 = 0, 1, 2, 3
NORM_LOGARITHM, NORM_NONE = 0, 1
METRIC_COSINE, METRIC_EUCLIDEAN = 0, 1
GRAPH_BINARY, GRAPH_SUM, GRAPH_SNN_JACCARD = 0, 1, 2
STAGES = ["upload", "normalize", "gene_stats", "hvg", "scale_stats", "scale_fill", "pca_gram", "pca_solve",
          "pca_transform", "knn", "graph", "generate", "knn_scan", "gram_mma", "alloc_host", "markers", "harmony", "leiden"]
T_COUNT = len(STAGES)

_NP_DT = {np.dtype(np.float32): DT_F32, np.dtype(np.float64): DT_F64,
          np.dtype(np.int32): DT_I32, np.dtype(np.int64): DT_I64}
_DT_NP = {v: k for k, v in _NP_DT.items()}


class CellScopesCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libcellscopes_cuda error {code}: {msg}")
        self.code = code


# every exported symbol: name -> (restype, argtypes).  tests/test_abi.py checks this table against the header.
_P = C.c_void_p
_I32, _I64, _F64, _U64, _F32 = C.c_int32, C.c_int64, C.c_double, C.c_uint64, C.c_float
_PP = C.POINTER(C.c_void_p)
SIGNATURES = {
    "csc_abi_version": (_I32, []),
    "csc_create": (_I32, [_I32, _PP]),
    "csc_destroy": (_I32, [_P]),
    "csc_last_error": (C.c_char_p, [_P]),
    "csc_synchronize": (_I32, [_P]),
    "csc_device_info": (_I32, [_P, C.POINTER(_I32), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I32), C.POINTER(_I32)]),
    "csc_timer_start": (_I32, [_P]),
    "csc_timer_stop": (_I32, [_P, C.POINTER(_F64)]),
    "csc_stage_times": (_I32, [_P, C.POINTER(_F64), C.POINTER(_I64), _I32]),
    "csc_flush_l2": (_I32, [_P]),
    "csc_host_alloc": (_I32, [_P, _I64, _PP]),
    "csc_host_free": (_I32, [_P, _P]),
    "csc_vec_alloc": (_I32, [_P, _I64, _I32, _PP]),
    "csc_vec_wrap": (_I32, [_P, _P, _I64, _I32, _PP]),
    "csc_vec_free": (_I32, [_P]),
    "csc_vec_info": (_I32, [_P, C.POINTER(_I64), C.POINTER(_I32), _PP]),
    "csc_vec_upload": (_I32, [_P, _P, _I64]),
    "csc_vec_download": (_I32, [_P, _P, _I64]),
    "csc_vec_download_as": (_I32, [_P, _P, _I64, _I32, _I64]),
    "csc_sparse_upload": (_I32, [_P, _I64, _I64, _P, _P, _I32, _P, _I32, _I32, _PP]),
    "csc_sparse_info": (_I32, [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "csc_sparse_range_nnz": (_I32, [_P, _I64, _I64, C.POINTER(_I64)]),
    "csc_sparse_download": (_I32, [_P, _I64, _I64, _P, _P, _P, _I32]),
    "csc_sparse_free": (_I32, [_P]),
    "csc_generate_counts": (_I32, [_P, _I64, _I64, _I64, _U64, _P, _P, _P, _P, _I32, _I32, _F32, _F32, _F32, _PP]),
    "csc_normalize": (_I32, [_P, _P, _F64, _I32, _F64, _I32, _P, _PP]),
    "csc_gene_stats_partial": (_I32, [_P, _P, _P]),
    "csc_hvg_finish": (_I32, [_P, _P, _I64, _I64, _I64, _F64, _P, _P, _P, _P, _P]),
    "csc_hvg_finish_collective": (_I32, [_P, _P, _I64, _I64, _I64, _F64, _P, _P, _P, _P, _P]),
    "csc_sparse_select_rows": (_I32, [_P, _P, _P, _I64, _PP]),
    "csc_scale_stats_partial": (_I32, [_P, _P, _P, _I64, _P]),
    "csc_scale_fill": (_I32, [_P, _P, _P, _I64, _P, _I64, _F64, _I32, _I32, _PP]),
    "csc_scale_gram_partial": (_I32, [_P, _P, _P, _I64, _P, _I64, _F64, _I32, _I32, _P, _P, _PP]),
    "csc_split_free": (_I32, [_P]),
    "csc_pca_transform_split": (_I32, [_P, _P, _P, _PP]),
    "csc_dense_alloc": (_I32, [_P, _I64, _I64, _PP]),
    "csc_dense_upload": (_I32, [_P, _P, _I32, _I64, _I64, _PP]),
    "csc_dense_download": (_I32, [_P, _P, _I32, _I64, _I64]),
    "csc_dense_info": (_I32, [_P, C.POINTER(_I64), C.POINTER(_I64), _PP]),
    "csc_dense_wrap": (_I32, [_P, _P, _I64, _I64, _PP]),
    "csc_dense_free": (_I32, [_P]),
    "csc_pca_partial": (_I32, [_P, _P, _P, _P]),
    "csc_pca_solve": (_I32, [_P, _P, _P, _I64, _I64, _I32, _F64, _I32, _PP, C.POINTER(_I32), C.POINTER(_F64)]),
    "csc_pca_info": (_I32, [_P, C.POINTER(_I64), C.POINTER(_I32), _P, C.POINTER(_F64), _P, _P]),
    "csc_pca_transform": (_I32, [_P, _P, _P, _PP]),
    "csc_pca_transform_sparse": (_I32, [_P, _P, _P, _I64, _P, _I64, _F64, _I32, _I32, _P, _PP]),
    "csc_pca_free": (_I32, [_P]),
    "csc_knn": (_I32, [_P, _P, _P, _I64, _I32, _I32, _I32, _I32, _PP, _PP]),
    "csc_knn_last_fallback_rows": (_I32, [_P, C.POINTER(_I64)]),
    "csc_graph_build": (_I32, [_P, _P, _I64, _I32, _I64, _I64, _I32, _F64, _PP]),
    "csc_graph_info": (_I32, [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "csc_graph_download": (_I32, [_P, _P, _P, _P, _I32, _I32]),
    "csc_graph_free": (_I32, [_P]),
    "csc_leiden": (_I32, [_P, _P, _F64, _I32, _I32, _P, C.POINTER(_I32), C.POINTER(_F64), C.POINTER(_I32)]),
    "csc_harmony": (_I32, [_P, _P, _I64, _I32, _P, _I32, _P, _I32, _P, _P, _P, _F64, _I32, _I32, _F64, _F64, _U64, _P, _P,
                           C.POINTER(_I32), _P, C.POINTER(_I32), _P]),
    "csc_kmeans": (_I32, [_P, _P, _I64, _I32, _I32, _P, _I32, _F64, _I32, _P, _P, C.POINTER(_F64), C.POINTER(_I32)]),
    "csc_read_mtx": (_I32, [_P, C.c_char_p, _PP]),
    "csc_sparse_col_sums": (_I32, [_P, _P]),
    "csc_sparse_select_cols": (_I32, [_P, _P, _P, _I64, _PP]),
    "csc_sparse_merge": (_I32, [_P, _P, _P, _I32, _I64, _PP]),
    "csc_tf_idf": (_I32, [_P, _P, _P, _I64, _F64, _PP]),
    "csc_marker_stats": (_I32, [_P, _P, _P, _I32, _F64, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_I64)]),
    "csc_comm_unique_id": (_I32, [_P, _I64]),
    "csc_comm_init": (_I32, [_P, _P, _I64, _I32, _I32]),
    "csc_comm_info": (_I32, [_P, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32)]),
    "csc_comm_destroy": (_I32, [_P]),
    "csc_comm_allreduce_sum": (_I32, [_P, _P]),
    "csc_comm_allgather_rows": (_I32, [_P, _P, _P, _PP]),
    "csc_comm_allgather_i32": (_I32, [_P, _P, _P, _I32, _PP]),
}

_lib = None


def csc_from_arrays(data, indices, indptr, shape):
    """scipy CSC matrix around existing (already valid, sorted) arrays without the validation passes and index
    down-casts of the constructor, which cost more than the device work for a 10^7-entry graph."""
    import scipy.sparse as sp
    m = sp.csc_matrix(shape, dtype=data.dtype)
    m.data, m.indices, m.indptr = data, indices, indptr
    m.has_sorted_indices = True
    return m


def _host_free(ctx, ptr):
    # a block that outlives its context (Context.close() leaves live blocks alone) goes back with a NULL context
    ctx.lib.csc_host_free(ctx.h if ctx.h is not None else None, C.c_void_p(ptr))


def load_library(path: str | None = None):
    """dlopen the shared library and declare every signature.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("CELLSCOPES_CUDA_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise CellScopesCudaError(
            -100, f"{p} not found: build it with `python cellscopes.jl_b200/build.py` "
                  "(the CUDA extension is required; there is no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.csc_abi_version() != 2:
        raise CellScopesCudaError(-101, "ABI version mismatch")
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU + stream (csc_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.csc_create(device, C.byref(h))
        if rc != 0:
            raise CellScopesCudaError(rc, (self.lib.csc_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise CellScopesCudaError(rc, (self.lib.csc_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.csc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- misc ------------------------------------------------------------------------------
    def pinned_empty(self, shape, dtype=np.float64):
        """Uninitialised numpy array in page-locked memory of this context's host cache (results land in it at
        PCIe rate).  The block returns to the cache when the array (and every view of it) is garbage-collected."""
        import weakref
        dtype = np.dtype(dtype)
        shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if nbytes == 0:
            return np.empty(shape, dtype=dtype)
        p = C.c_void_p()
        self.check(self.lib.csc_host_alloc(self.h, nbytes, C.byref(p)))
        buf = (C.c_char * nbytes).from_address(p.value)
        weakref.finalize(buf, _host_free, self, p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def synchronize(self):
        self.check(self.lib.csc_synchronize(self.h))

    def device_info(self):
        sm, cmaj, cmin = _I32(), _I32(), _I32()
        fr, tot = _I64(), _I64()
        self.check(self.lib.csc_device_info(self.h, C.byref(sm), C.byref(fr), C.byref(tot), C.byref(cmaj), C.byref(cmin)))
        return {"sm_count": sm.value, "free_bytes": fr.value, "total_bytes": tot.value, "cc": (cmaj.value, cmin.value)}

    def timer_start(self):
        self.check(self.lib.csc_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = _F64()
        self.check(self.lib.csc_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def stage_times(self, reset=False):
        ms = (_F64 * T_COUNT)()
        n = _I64()
        self.check(self.lib.csc_stage_times(self.h, ms, C.byref(n), 1 if reset else 0))
        return {STAGES[i]: ms[i] for i in range(T_COUNT)}, n.value

    def flush_l2(self):
        self.check(self.lib.csc_flush_l2(self.h))

    def leiden(self, graph, resolution=0.06, max_levels=0, max_sweeps=0):
        """csc_leiden: dict(labels int32[n], n_communities, quality, levels)"""
        n = graph.info()[0]
        lab = np.empty(n, dtype=np.int32)
        nc, lv, q = _I32(), _I32(), _F64()
        self.check(self.lib.csc_leiden(self.h, graph.h, float(resolution), int(max_levels), int(max_sweeps), _ptr(lab), C.byref(nc),
                                       C.byref(q), C.byref(lv)))
        return {"labels": lab, "n_communities": nc.value, "quality": q.value, "levels": lv.value}

    def kmeans(self, x, k, init_idx, max_iter=100, tol=1e-4, unit_rows=False):
        """csc_kmeans: (centres [k x d], labels [n], objective, iterations)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        idx = np.ascontiguousarray(init_idx, dtype=np.int64)
        cen = np.empty((int(k), x.shape[1]), dtype=np.float64)
        lab = np.empty(x.shape[0], dtype=np.int32)
        obj, it = _F64(), _I32()
        self.check(self.lib.csc_kmeans(self.h, _ptr(x), x.shape[0], x.shape[1], int(k), _ptr(idx), int(max_iter), float(tol),
                                       1 if unit_rows else 0, _ptr(cen), _ptr(lab), C.byref(obj), C.byref(it)))
        return cen, lab, obj.value, it.value

    def harmony(self, z, batch, n_batch, y0, theta, lamb, sigma, block_size=0.05, max_iter_harmony=10, max_iter_kmeans=20,
                eps_kmeans=1e-5, eps_harmony=1e-4, seed=123, want_r=False):
        """csc_harmony: dict(Z_corr [n x d], objective_harmony, kmeans_rounds, R [n x K] if asked)"""
        z = np.ascontiguousarray(z, dtype=np.float64)
        n, d = z.shape
        b = np.ascontiguousarray(batch, dtype=np.int32)
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        k = y0.shape[0]
        th = np.ascontiguousarray(theta, dtype=np.float64)
        lm = np.ascontiguousarray(lamb, dtype=np.float64)
        sg = np.ascontiguousarray(sigma, dtype=np.float64)
        out = np.empty_like(z)
        obj = np.zeros(int(max_iter_harmony) + 2, dtype=np.float64)
        rounds = np.zeros(int(max_iter_harmony) + 1, dtype=np.int32)
        n_obj, n_rounds = _I32(obj.size), _I32(rounds.size)
        r = np.empty((n, k), dtype=np.float64) if want_r else None
        self.check(self.lib.csc_harmony(self.h, _ptr(z), n, d, _ptr(b), int(n_batch), _ptr(y0), k, _ptr(th), _ptr(lm), _ptr(sg),
                                        float(block_size), int(max_iter_harmony), int(max_iter_kmeans), float(eps_kmeans),
                                        float(eps_harmony), int(seed), _ptr(out), _ptr(obj), C.byref(n_obj), _ptr(rounds),
                                        C.byref(n_rounds), _ptr(r)))
        return {"Z_corr": out, "objective_harmony": obj[:n_obj.value].copy(), "kmeans_rounds": rounds[:n_rounds.value].copy(),
                "R": r}

    def read_mtx(self, path):
        """MatrixMarket coordinate file (plain or .gz) -> device CSC (rows = genes, columns = cells)."""
        h = C.c_void_p()
        self.check(self.lib.csc_read_mtx(self.h, os.fsencode(path), C.byref(h)))
        return Sparse(self, h)

    def select_cols(self, m, cols):
        """the columns ``cols`` (0-based) of a device matrix as a new device matrix"""
        c = np.ascontiguousarray(cols, dtype=np.int64)
        h = C.c_void_p()
        self.check(self.lib.csc_sparse_select_cols(self.h, m.h, _ptr(c), c.size, C.byref(h)))
        return Sparse(self, h)

    def sparse_merge(self, mats, row_maps, n_rows_merged):
        """merge_matrices: ``mats`` side by side; row r of mats[i] becomes row row_maps[i][r] of the result."""
        n = len(mats)
        maps = [np.ascontiguousarray(m, dtype=np.int64) for m in row_maps]
        hp = (C.c_void_p * n)(*[m.h for m in mats])
        mp = (C.c_void_p * n)(*[m.ctypes.data for m in maps])
        h = C.c_void_p()
        self.check(self.lib.csc_sparse_merge(self.h, hp, mp, n, int(n_rows_merged), C.byref(h)))
        return Sparse(self, h)

    def tf_idf(self, raw, gene_stats, n_cells_total, scale_factor=10000.0):
        h = C.c_void_p()
        self.check(self.lib.csc_tf_idf(self.h, raw.h, gene_stats.h, int(n_cells_total), float(scale_factor), C.byref(h)))
        return Sparse(self, h)

    def marker_stats(self, norm, labels, n_groups, expr_cutoff=0.0):
        """csc_marker_stats: the per-(gene, group) tables of the Mann-Whitney tests and of the result columns."""
        g = norm.shape[0]
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        tabs = {k: np.zeros((g, n_groups), dtype=np.float64) for k in ("rank_sum", "sum_expm1", "n_expr", "nnz", "sum_x")}
        per_gene = {k: np.zeros(g, dtype=np.float64) for k in ("tie_adj", "n_neg", "n_zero_stored")}
        n_incl = _I64()
        self.check(self.lib.csc_marker_stats(self.h, norm.h, _ptr(lab), int(n_groups), float(expr_cutoff),
                                             _ptr(tabs["rank_sum"]), _ptr(tabs["sum_expm1"]), _ptr(tabs["n_expr"]),
                                             _ptr(tabs["nnz"]), _ptr(tabs["sum_x"]), _ptr(per_gene["tie_adj"]),
                                             _ptr(per_gene["n_neg"]), _ptr(per_gene["n_zero_stored"]), C.byref(n_incl)))
        tabs.update(per_gene)
        tabs["n_included"] = n_incl.value
        return tabs

    # ---- multi-GPU: NCCL inside the library ------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        """128-byte NCCL unique id (rank 0 makes it; the host moves it to the other ranks)."""
        buf = C.create_string_buffer(128)
        rc = self.lib.csc_comm_unique_id(buf, 128)
        if rc != 0:
            raise CellScopesCudaError(rc, (self.lib.csc_last_error(None) or b"").decode())
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self.check(self.lib.csc_comm_init(self.h, buf, 128, int(rank), int(world)))

    def comm_info(self):
        r, w, v = _I32(), _I32(), _I32()
        self.check(self.lib.csc_comm_info(self.h, C.byref(r), C.byref(w), C.byref(v)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value}

    def comm_allreduce_sum(self, vec):
        self.check(self.lib.csc_comm_allreduce_sum(self.h, vec.h))
        return vec

    def comm_allgather_rows(self, dense, offsets):
        o = np.ascontiguousarray(offsets, dtype=np.int64)
        h = C.c_void_p()
        self.check(self.lib.csc_comm_allgather_rows(self.h, dense.h, _ptr(o), C.byref(h)))
        return Dense(self, h)

    def comm_allgather_i32(self, vec, offsets, width):
        o = np.ascontiguousarray(offsets, dtype=np.int64)
        h = C.c_void_p()
        self.check(self.lib.csc_comm_allgather_i32(self.h, vec.h, _ptr(o), int(width), C.byref(h)))
        return Vec(self, h)

    # ---- vectors ---------------------------------------------------------------------------
    def vec(self, n, dtype=np.float64):
        h = C.c_void_p()
        self.check(self.lib.csc_vec_alloc(self.h, int(n), _NP_DT[np.dtype(dtype)], C.byref(h)))
        return Vec(self, h)

    def vec_from(self, arr):
        arr = np.ascontiguousarray(arr)
        v = self.vec(arr.size, arr.dtype)
        v.upload(arr)
        return v

    def vec_wrap(self, device_ptr, n, dtype, keepalive=None):
        h = C.c_void_p()
        self.check(self.lib.csc_vec_wrap(self.h, C.c_void_p(int(device_ptr)), int(n), _NP_DT[np.dtype(dtype)], C.byref(h)))
        v = Vec(self, h)
        v._keepalive = keepalive
        return v

    def vec_wrap_i32(self, tensor):
        """Non-owning int32 view of a contiguous torch CUDA tensor."""
        return self.vec_wrap(tensor.data_ptr(), tensor.numel(), np.int32, keepalive=tensor)

    # ---- sparse ----------------------------------------------------------------------------
    def sparse_upload(self, n_genes, n_cells, colptr, rowval, nzval, index_base=0):
        colptr = np.ascontiguousarray(colptr)
        rowval = np.ascontiguousarray(rowval)
        nzval = np.ascontiguousarray(nzval)
        if colptr.dtype not in (np.int64, np.int32) or rowval.dtype != colptr.dtype:
            colptr = colptr.astype(np.int64)
            rowval = rowval.astype(np.int64)
        if nzval.dtype not in _NP_DT:
            nzval = nzval.astype(np.float64)
        h = C.c_void_p()
        self.check(self.lib.csc_sparse_upload(self.h, int(n_genes), int(n_cells), _ptr(colptr), _ptr(rowval),
                                              _NP_DT[colptr.dtype], _ptr(nzval), _NP_DT[nzval.dtype],
                                              int(index_base), C.byref(h)))
        return Sparse(self, h)

    def sparse_from_scipy(self, m):
        import scipy.sparse as sp
        # the SparseMatrixCSC invariant (sorted row indices) is the caller's, as in the reference; nothing on the
        # path depends on it, so no host-side pass over the indices is made here
        m = m if sp.isspmatrix_csc(m) else sp.csc_matrix(m)
        return self.sparse_upload(m.shape[0], m.shape[1], m.indptr, m.indices, m.data, 0)

    def generate_counts(self, params, cell_lo, cell_hi):
        """Device generator; ``params`` is a synth.SynthParams."""
        h = C.c_void_p()
        a = np.ascontiguousarray(params.a, dtype=np.float32)
        mod = np.ascontiguousarray(params.module, dtype=np.int32)
        load = np.ascontiguousarray(params.loading, dtype=np.float32)
        tm = np.ascontiguousarray(params.type_mean, dtype=np.float32)
        self.check(self.lib.csc_generate_counts(self.h, params.n_genes, int(cell_lo), int(cell_hi), params.seed,
                                                _ptr(a), _ptr(mod), _ptr(load), _ptr(tm), params.n_types,
                                                params.n_modules, params.lib_sd, params.type_w, params.noise_w,
                                                C.byref(h)))
        return Sparse(self, h)

    # ---- stages ----------------------------------------------------------------------------
    def normalize(self, raw, scale_factor=10000.0, norm_method=NORM_LOGARITHM, pseudocount=1.0, nan_to_zero=False,
                  gene_stats=None):
        h = C.c_void_p()
        self.check(self.lib.csc_normalize(self.h, raw.h, float(scale_factor), int(norm_method), float(pseudocount),
                                          1 if nan_to_zero else 0, gene_stats.h if gene_stats is not None else None,
                                          C.byref(h)))
        return Sparse(self, h)

    def gene_stats_partial(self, raw, gene_stats):
        self.check(self.lib.csc_gene_stats_partial(self.h, raw.h, gene_stats.h))

    def hvg_finish(self, gene_stats, n_genes, n_cells_total, n_features=2000, span=0.3, collective=False):
        """``collective``: every rank of the context's communicator makes this call (csc_hvg_finish_collective)."""
        g = int(n_genes)
        out = {k: np.empty(g, dtype=np.float64) for k in
               ("mean", "variance", "variance_expected", "variance_standardized")}
        order = np.empty(g, dtype=np.int64)
        fn = self.lib.csc_hvg_finish_collective if collective else self.lib.csc_hvg_finish
        self.check(fn(self.h, gene_stats.h, g, int(n_cells_total), int(n_features), float(span),
                                           _ptr(out["mean"]), _ptr(out["variance"]), _ptr(out["variance_expected"]),
                                           _ptr(out["variance_standardized"]), _ptr(order)))
        out["order"] = order
        out["features"] = order[:min(int(n_features), g)].copy()
        return out

    def select_rows(self, m, features):
        """subset_count on the device: rows `features` in original gene order, renumbered 0..H-1"""
        f = np.ascontiguousarray(features, dtype=np.int64)
        h = C.c_void_p()
        self.check(self.lib.csc_sparse_select_rows(self.h, m.h, _ptr(f), f.size, C.byref(h)))
        return Sparse(self, h)

    def scale_stats_partial(self, norm, features, partial):
        f = None if features is None else np.ascontiguousarray(features, dtype=np.int64)
        self.check(self.lib.csc_scale_stats_partial(self.h, norm.h, _ptr(f), 0 if f is None else f.size, partial.h))

    def scale_fill(self, norm, features, partial, n_cells_total, scale_max=10.0, do_scale=True, do_center=True):
        f = None if features is None else np.ascontiguousarray(features, dtype=np.int64)
        h = C.c_void_p()
        self.check(self.lib.csc_scale_fill(self.h, norm.h, _ptr(f), 0 if f is None else f.size, partial.h,
                                           int(n_cells_total), float(scale_max), 1 if do_scale else 0,
                                           1 if do_center else 0, C.byref(h)))
        return Dense(self, h)

    def scale_gram_partial(self, norm, features, partial, n_cells_total, colsum, gram, scale_max=10.0, do_scale=True,
                           do_center=True, keep_operands=False):
        """scale_fill + pca_partial without the dense scaled matrix: colsum += column sums, gram += Z'Z of this shard.
        ``keep_operands``: returns the fp16 hi/lo split of the scaled matrix (a ``Split``, or None when the fallback ran)
        for pca_transform_split."""
        f = None if features is None else np.ascontiguousarray(features, dtype=np.int64)
        h = C.c_void_p()
        self.check(self.lib.csc_scale_gram_partial(self.h, norm.h, _ptr(f), 0 if f is None else f.size, partial.h,
                                                   int(n_cells_total), float(scale_max), 1 if do_scale else 0,
                                                   1 if do_center else 0, colsum.h, gram.h,
                                                   C.byref(h) if keep_operands else None))
        return Split(self, h) if (keep_operands and h.value) else None

    def pca_transform_split(self, split, pca):
        h = C.c_void_p()
        self.check(self.lib.csc_pca_transform_split(self.h, split.h, pca.h, C.byref(h)))
        return Dense(self, h)

    def dense_upload(self, arr):
        arr = np.ascontiguousarray(arr)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        h = C.c_void_p()
        self.check(self.lib.csc_dense_upload(self.h, _ptr(arr), _NP_DT[arr.dtype], arr.shape[0], arr.shape[1], C.byref(h)))
        return Dense(self, h)

    def dense_alloc(self, n_rows, n_cols):
        h = C.c_void_p()
        self.check(self.lib.csc_dense_alloc(self.h, int(n_rows), int(n_cols), C.byref(h)))
        return Dense(self, h)

    def dense_wrap(self, device_ptr, n_rows, n_cols, keepalive=None):
        h = C.c_void_p()
        self.check(self.lib.csc_dense_wrap(self.h, C.c_void_p(int(device_ptr)), int(n_rows), int(n_cols), C.byref(h)))
        d = Dense(self, h)
        d._keepalive = keepalive
        return d

    def pca_partial(self, scaled, colsum, gram):
        self.check(self.lib.csc_pca_partial(self.h, scaled.h, colsum.h, gram.h))

    def pca_solve(self, colsum, gram, n_feat, n_cells_total, maxoutdim, tol=1e-9, max_iter=500):
        h = C.c_void_p()
        it, res = _I32(), _F64()
        self.check(self.lib.csc_pca_solve(self.h, colsum.h, gram.h, int(n_feat), int(n_cells_total), int(maxoutdim),
                                          float(tol), int(max_iter), C.byref(h), C.byref(it), C.byref(res)))
        p = Pca(self, h)
        p.iters, p.resid = it.value, res.value
        return p

    def pca_transform(self, scaled, pca):
        h = C.c_void_p()
        self.check(self.lib.csc_pca_transform(self.h, scaled.h, pca.h, C.byref(h)))
        return Dense(self, h)

    def pca_transform_sparse(self, norm, features, partial, n_cells_total, pca, scale_max=10.0, do_scale=True,
                             do_center=True):
        f = None if features is None else np.ascontiguousarray(features, dtype=np.int64)
        h = C.c_void_p()
        self.check(self.lib.csc_pca_transform_sparse(self.h, norm.h, _ptr(f), 0 if f is None else f.size, partial.h,
                                                     int(n_cells_total), float(scale_max), 1 if do_scale else 0,
                                                     1 if do_center else 0, pca.h, C.byref(h)))
        return Dense(self, h)

    def knn(self, queries, keys, k, query_offset=0, dim_lo=0, dim_hi=None, metric=METRIC_COSINE):
        hi, hd = C.c_void_p(), C.c_void_p()
        if dim_hi is None:
            dim_hi = keys.shape[1]
        self.check(self.lib.csc_knn(self.h, queries.h, keys.h, int(query_offset), int(dim_lo), int(dim_hi), int(k),
                                    int(metric), C.byref(hi), C.byref(hd)))
        return Vec(self, hi), Vec(self, hd)

    def knn_last_fallback_rows(self) -> int:
        n = _I64()
        self.check(self.lib.csc_knn_last_fallback_rows(self.h, C.byref(n)))
        return n.value

    def graph_build(self, knn_idx, n_cells_total, k, cell_lo=0, cell_hi=None, mode=GRAPH_SUM, snn_prune=0.0):
        h = C.c_void_p()
        cell_hi = n_cells_total if cell_hi is None else cell_hi
        self.check(self.lib.csc_graph_build(self.h, knn_idx.h, int(n_cells_total), int(k), int(cell_lo), int(cell_hi),
                                            int(mode), float(snn_prune), C.byref(h)))
        return Graph(self, h)


class _Handle:
    _free = None

    def __init__(self, ctx: Context, h):
        self.ctx = ctx
        self.h = h

    def free(self):
        if getattr(self, "h", None):
            getattr(self.ctx.lib, self._free)(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Vec(_Handle):
    _free = "csc_vec_free"

    def info(self):
        n, dt, p = _I64(), _I32(), C.c_void_p()
        self.ctx.check(self.ctx.lib.csc_vec_info(self.h, C.byref(n), C.byref(dt), C.byref(p)))
        return n.value, _DT_NP[dt.value], p.value

    @property
    def n(self):
        return self.info()[0]

    @property
    def dtype(self):
        return self.info()[1]

    @property
    def device_ptr(self):
        return self.info()[2]

    def upload(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        self.ctx.check(self.ctx.lib.csc_vec_upload(self.h, _ptr(arr), arr.size))

    def download(self, n=None):
        total, dt, _ = self.info()
        n = total if n is None else n
        out = self.ctx.pinned_empty(n, dt) if n * np.dtype(dt).itemsize >= (1 << 20) else np.empty(n, dtype=dt)
        self.ctx.check(self.ctx.lib.csc_vec_download(self.h, _ptr(out), n))
        return out

    def download_as(self, dtype, add=0, n=None):
        """Download converted on the device first: int32 -> int32/int64 with ``add`` on every element, float32 -> float64."""
        total, _, _ = self.info()
        n = total if n is None else n
        dtype = np.dtype(dtype)
        out = self.ctx.pinned_empty(n, dtype) if n * dtype.itemsize >= (1 << 20) else np.empty(n, dtype=dtype)
        self.ctx.check(self.ctx.lib.csc_vec_download_as(self.h, _ptr(out), n, _NP_DT[dtype], int(add)))
        return out

    @property
    def __cuda_array_interface__(self):
        n, dt, p = self.info()
        return {"shape": (n,), "typestr": np.dtype(dt).str, "data": (p or 0, False), "version": 3, "strides": None}


class Sparse(_Handle):
    _free = "csc_sparse_free"

    @property
    def shape(self):
        g, n, _ = self.info()
        return g, n

    def info(self):
        g, n, z = _I64(), _I64(), _I64()
        self.ctx.check(self.ctx.lib.csc_sparse_info(self.h, C.byref(g), C.byref(n), C.byref(z)))
        return g.value, n.value, z.value

    @property
    def nnz(self):
        return self.info()[2]

    def column_sums(self):
        """colSum (utils.jl:1-2), fp64 on the host"""
        out = np.zeros(self.shape[1], dtype=np.float64)
        self.ctx.check(self.ctx.lib.csc_sparse_col_sums(self.h, _ptr(out)))
        return out

    def range_nnz(self, lo, hi):
        z = _I64()
        self.ctx.check(self.ctx.lib.csc_sparse_range_nnz(self.h, int(lo), int(hi), C.byref(z)))
        return z.value

    def download(self, cell_lo=0, cell_hi=None, index_base=0, values_only=False):
        g, n, _ = self.info()
        cell_hi = n if cell_hi is None else cell_hi
        nnz = self.range_nnz(cell_lo, cell_hi)
        val = self.ctx.pinned_empty(nnz, np.float64)
        if values_only:
            self.ctx.check(self.ctx.lib.csc_sparse_download(self.h, int(cell_lo), int(cell_hi), None, None, _ptr(val),
                                                            int(index_base)))
            return val
        colptr = self.ctx.pinned_empty(cell_hi - cell_lo + 1, np.int64)
        rowval = self.ctx.pinned_empty(nnz, np.int64)
        self.ctx.check(self.ctx.lib.csc_sparse_download(self.h, int(cell_lo), int(cell_hi), _ptr(colptr), _ptr(rowval),
                                                        _ptr(val), int(index_base)))
        return colptr, rowval, val

    def to_scipy(self, cell_lo=0, cell_hi=None):
        import scipy.sparse as sp
        g, n, _ = self.info()
        cell_hi = n if cell_hi is None else cell_hi
        colptr, rowval, val = self.download(cell_lo, cell_hi, 0)
        return sp.csc_matrix((val, rowval, colptr), shape=(g, cell_hi - cell_lo))


class Dense(_Handle):
    _free = "csc_dense_free"

    def info(self):
        r, c, p = _I64(), _I64(), C.c_void_p()
        self.ctx.check(self.ctx.lib.csc_dense_info(self.h, C.byref(r), C.byref(c), C.byref(p)))
        return r.value, c.value, p.value

    @property
    def shape(self):
        r, c, _ = self.info()
        return r, c

    @property
    def device_ptr(self):
        return self.info()[2]

    def download(self, dtype=np.float64, row_lo=0, row_hi=None):
        r, c, _ = self.info()
        row_hi = r if row_hi is None else row_hi
        out = self.ctx.pinned_empty((row_hi - row_lo, c), dtype)
        self.ctx.check(self.ctx.lib.csc_dense_download(self.h, _ptr(out), _NP_DT[np.dtype(dtype)], int(row_lo), int(row_hi)))
        return out

    @property
    def __cuda_array_interface__(self):
        r, c, p = self.info()
        return {"shape": (r, c), "typestr": "<f4", "data": (p or 0, False), "version": 3, "strides": None}


class Split(_Handle):
    _free = "csc_split_free"


class Pca(_Handle):
    _free = "csc_pca_free"
    iters = 0
    resid = 0.0

    def info(self):
        nf, k = _I64(), _I32()
        self.ctx.check(self.ctx.lib.csc_pca_info(self.h, C.byref(nf), C.byref(k), None, None, None, None))
        pv = np.empty(k.value, dtype=np.float64)
        load = np.empty((k.value, nf.value), dtype=np.float64)   # column-major H x k == row-major k x H
        mean = np.empty(nf.value, dtype=np.float64)
        tv = _F64()
        self.ctx.check(self.ctx.lib.csc_pca_info(self.h, None, None, _ptr(pv), C.byref(tv), _ptr(load), _ptr(mean)))
        return {"n_feat": nf.value, "k": k.value, "principalvars": pv, "tvar": tv.value, "loadings": load.T.copy(),
                "mean": mean}


class Graph(_Handle):
    _free = "csc_graph_free"

    def info(self):
        r, c, z = _I64(), _I64(), _I64()
        self.ctx.check(self.ctx.lib.csc_graph_info(self.h, C.byref(r), C.byref(c), C.byref(z)))
        return r.value, c.value, z.value

    def download(self, index_base=0, val_dtype=np.float64):
        r, c, z = self.info()
        colptr = self.ctx.pinned_empty(c + 1, np.int64)
        rowval = self.ctx.pinned_empty(z, np.int64)
        val = self.ctx.pinned_empty(z, val_dtype)
        self.ctx.check(self.ctx.lib.csc_graph_download(self.h, _ptr(colptr), _ptr(rowval), _ptr(val),
                                                       _NP_DT[np.dtype(val_dtype)], int(index_base)))
        return colptr, rowval, val

    def to_scipy(self, val_dtype=np.float64):
        r, c, _ = self.info()
        colptr, rowval, val = self.download(0, val_dtype)
        return csc_from_arrays(val, rowval, colptr, (r, c))
