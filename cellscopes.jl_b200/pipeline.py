"""Device-resident orchestration of the hot path, cell-sharded over ranks.

``run_path`` is the one place where the stage order and the exchange steps of the path are written
down.  It talks to a *backend* (the ctypes ``Context`` of ``_capi``: every call is one C-ABI entry
point) and to a *communicator* (``LocalComm`` for one GPU, ``TorchComm`` for one process per GPU over
``torch.distributed``).  The collectives are exactly the natural ones:

  all-reduce  fp64 [2G]      per-gene sum x, sum x^2            (find_variable_genes)
  all-reduce  fp64 [2H]      per-HVG sum y, sum y^2             (scale_object)
  all-reduce  fp64 [H], [H*H] column sums and Gram of Z         (run_pca)
  all-gather  fp32 [N x k_pc] embedding                          (kNN keys)
  all-gather  int32 [N x k]   neighbour table                    (adjacency / SNN)

Normalisation, the dense fill, the PCA transform and the kNN scan itself need no exchange.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import _capi


@dataclasses.dataclass
class PathParams:
    scale_factor: float = 10000.0
    norm_method: str = "logarithm"
    pseudocount: float = 1.0
    n_features: int = 2000
    span: float = 0.3
    scale_max: float = 10.0
    do_scale: bool = True
    do_center: bool = True
    n_pcs: int = 50
    k: int = 20
    metric: str = "cosine"
    graph_mode: str = "sum"        # "binary" (small), "sum" (atlas), "snn"
    snn_prune: float = 0.0
    pca_tol: float = 1e-9
    pca_max_iter: int = 500
    keep_norm: bool = False
    keep_scaled: bool = False      # materialise the dense scaled matrix (scaleCount) as well


_NORM = {"logarithm": _capi.NORM_LOGARITHM, "none": _capi.NORM_NONE}
_METRIC = {"cosine": _capi.METRIC_COSINE, "euclidean": _capi.METRIC_EUCLIDEAN}
_GRAPH = {"binary": _capi.GRAPH_BINARY, "sum": _capi.GRAPH_SUM, "snn": _capi.GRAPH_SNN_JACCARD}


def norm_code(name):
    if name not in _NORM:
        # processing.jl:28 throws ArgumentError("Unknown normalization method: ...")
        raise ValueError(f"Unknown normalization method: {name}")
    return _NORM[name]


def metric_code(name):
    name = str(name).lower().replace("dist", "").replace("()", "")
    if name not in _METRIC:
        raise ValueError(f"unsupported metric {name!r} (cosine, euclidean)")
    return _METRIC[name]


def partition_cells(n_cells: int, world: int, colptr=None):
    """Contiguous cell ranges per rank.  With ``colptr`` the cut points balance the number of
    non-zeros (the sparse stages stream nnz); otherwise they balance the cell count (kNN work is
    proportional to the number of local queries).  Returns ``world + 1`` offsets."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if colptr is None:
        base, rem = divmod(int(n_cells), world)
        sizes = [base + (1 if r < rem else 0) for r in range(world)]
        return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    colptr = np.asarray(colptr, dtype=np.int64)
    nnz = colptr[-1] - colptr[0]
    targets = colptr[0] + (nnz * np.arange(1, world)) // world
    cuts = np.searchsorted(colptr, targets, side="left")
    offs = np.concatenate([[0], cuts, [n_cells]]).astype(np.int64)
    return np.maximum.accumulate(offs)


class LocalComm:
    """world == 1: every exchange is the identity."""
    rank = 0
    world = 1

    def allreduce_sum(self, vec):
        return vec

    def allgather_rows(self, be, dense, offsets):
        return dense

    def allgather_i32(self, be, vec, offsets, k):
        return vec

    def barrier(self):
        pass


class NcclComm:
    """One process per GPU; every collective is a csc_comm_* call: NCCL inside libcellscopes_cuda.so, enqueued on the
    context's own stream, no host synchronisation around it.  ``bootstrap`` moves the 128-byte unique id from rank 0
    to the others (any callable ``bytes|None -> bytes``; ``from_torch`` uses a torch.distributed broadcast, a Julia
    host would use MPI.jl / Distributed.jl)."""

    def __init__(self, be, rank, world, bootstrap):
        self.be, self.rank, self.world = be, int(rank), int(world)
        uid = bootstrap(be.comm_unique_id() if self.rank == 0 else None)
        be.comm_init(uid, self.rank, self.world)

    @classmethod
    def from_torch(cls, be):
        import torch.distributed as dist

        def bootstrap(uid):
            box = [uid]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        return cls(be, dist.get_rank(), dist.get_world_size(), bootstrap)

    def allreduce_sum(self, vec):
        return self.be.comm_allreduce_sum(vec)

    def allgather_rows(self, be, dense, offsets):
        return be.comm_allgather_rows(dense, offsets)

    def allgather_i32(self, be, vec, offsets, k):
        return be.comm_allgather_i32(vec, offsets, k)

    def barrier(self):
        self.be.synchronize()


class TorchComm:
    """One process per GPU; collectives through torch.distributed on tensors that alias the
    library's device buffers (``__cuda_array_interface__``), so no staging copy is made for the
    all-reduces."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.device = device

    def _tensor(self, obj):
        torch = self.torch
        if isinstance(obj, torch.Tensor):
            return obj
        if hasattr(obj, "__cuda_array_interface__") and self.device is not None:
            return torch.as_tensor(obj, device=self.device)
        return torch.from_numpy(np.asarray(obj))

    def allreduce_sum(self, vec):
        t = self._tensor(vec)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        if t.is_cuda:
            self.torch.cuda.synchronize(t.device)
        return vec

    def _gather(self, local, offsets, width, dtype):
        torch = self.torch
        sizes = np.diff(np.asarray(offsets, dtype=np.int64))
        max_rows = int(sizes.max())
        pad = torch.zeros((max_rows, width), dtype=dtype, device=local.device)
        pad[: local.shape[0]] = local.reshape(-1, width)
        out = torch.empty((self.world, max_rows, width), dtype=dtype, device=local.device)
        self.dist.all_gather_into_tensor(out.view(-1), pad.view(-1))
        full = torch.cat([out[r, : int(sizes[r])] for r in range(self.world)], dim=0).contiguous()
        if full.is_cuda:
            torch.cuda.synchronize(full.device)
        return full

    def allgather_rows(self, be, dense, offsets):
        n, d = dense.shape
        full = self._gather(self._tensor(dense).reshape(n, d), offsets, d, self.torch.float32)
        return be.dense_wrap(full.data_ptr(), full.shape[0], d, keepalive=full)

    def allgather_i32(self, be, vec, offsets, k):
        t = self._tensor(vec).reshape(-1, k)
        full = self._gather(t, offsets, k, self.torch.int32)
        return be.vec_wrap_i32(full)

    def barrier(self):
        self.dist.barrier()


def run_path(be, comm, raw, n_cells_total, cell_offset, p: PathParams, offsets=None):
    """counts (device CSC shard ``raw``) -> normalised -> HVG -> scaled dense -> PCA -> kNN -> graph.

    ``offsets``: the ``world + 1`` global cell offsets of the shards (needed when world > 1).
    Returns a dict of device handles plus the small host-side results."""
    G, n_loc = raw.shape
    if offsets is None:
        offsets = np.array([0, n_cells_total], dtype=np.int64)
    # a1 + the raw moments of a2, one pass over the counts
    stats = be.vec(2 * G, np.float64)
    norm = be.normalize(raw, p.scale_factor, norm_code(p.norm_method), p.pseudocount, False, stats)
    comm.allreduce_sum(stats)
    # a2
    if isinstance(comm, NcclComm) and comm.world > 1:
        hvg = be.hvg_finish(stats, G, n_cells_total, p.n_features, p.span, collective=True)   # Loess fits dealt over the ranks
    else:
        hvg = be.hvg_finish(stats, G, n_cells_total, p.n_features, p.span)
    feats = hvg["features"]
    rows = np.unique(feats)                       # subset_count keeps original gene order
    H = int(rows.size)
    # a3
    part = be.vec(2 * H, np.float64)
    src, src_feats = norm, feats                  # what the scale / Gram / transform passes read
    if hasattr(be, "select_rows") and G >= 4 * H:
        # whole-transcriptome matrix: subset_count first (few of the rows are kept), everything after reads the subset
        src, src_feats = be.select_rows(norm, feats), None
    # the normalised moments of the selected rows (64-bit fixed-point accumulators on the device)
    be.scale_stats_partial(src, src_feats, part)
    comm.allreduce_sum(part)
    colsum = be.vec(H, np.float64)
    gram = be.vec(H * H, np.float64)
    split = None
    if p.keep_scaled or not hasattr(be, "scale_gram_partial"):
        scaled = be.scale_fill(src, src_feats, part, n_cells_total, p.scale_max, p.do_scale, p.do_center)
        # a4
        be.pca_partial(scaled, colsum, gram)
    else:
        # a3 + a4 fused: nothing after the Gram reads the dense scaled matrix (the embedding comes from the sparse one)
        scaled = None
        split = be.scale_gram_partial(src, src_feats, part, n_cells_total, colsum, gram, p.scale_max, p.do_scale, p.do_center,
                                      keep_operands=p.n_pcs <= 64 and hasattr(be, "pca_transform_split"))
    comm.allreduce_sum(colsum)
    comm.allreduce_sum(gram)
    pca = be.pca_solve(colsum, gram, H, n_cells_total, p.n_pcs, p.pca_tol, p.pca_max_iter)
    # the embedding comes straight from the sparse normalised matrix (Z = S + 1 z0' is sparse-plus-rank-one)
    if split is not None:
        # the fp16 hi/lo operands the Gram kernel consumed are still in HBM: the embedding is one tensor-core GEMM over them
        emb = be.pca_transform_split(split, pca)
        split.free()
    elif hasattr(be, "pca_transform_sparse"):
        emb = be.pca_transform_sparse(src, src_feats, part, n_cells_total, pca, p.scale_max, p.do_scale, p.do_center)
    else:
        emb = be.pca_transform(scaled, pca)
    nnz_selected = src.nnz if hasattr(src, "nnz") else None
    if src is not norm:
        src.free()
    if not p.keep_norm:
        norm.free()
        norm = None
    # a5/a6
    keys = comm.allgather_rows(be, emb, offsets)
    idx, dist = be.knn(emb, keys, p.k, query_offset=cell_offset, metric=metric_code(p.metric))
    idx_all = comm.allgather_i32(be, idx, offsets, p.k)
    graph = be.graph_build(idx_all, n_cells_total, p.k, cell_offset, cell_offset + n_loc, _GRAPH[p.graph_mode],
                           p.snn_prune)
    return {"norm": norm, "hvg": hvg, "rows": rows, "scaled": scaled, "pca": pca, "embedding": emb,
            "knn_idx": idx, "knn_dist": dist, "graph": graph, "gene_stats": stats,
            # the all-gathered tables (the local ones again when world == 1) and the size of what the scale / Gram /
            # transform passes read (the row subset of a whole-transcriptome matrix)
            "keys": keys, "knn_idx_all": idx_all, "nnz_selected": nnz_selected}
