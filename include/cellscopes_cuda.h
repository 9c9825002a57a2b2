/*
 * libcellscopes_cuda.so -- C ABI of the B200 (sm_100a) implementation of the
 * CellScopes.jl preprocessing-to-graph path:
 *
 *   normalize_object -> find_variable_genes -> scale_object -> run_pca -> run_clustering(kNN/adjacency)
 *
 * The reference (HaojiaWu/CellScopes.jl) is pure Julia and has no FFI of its own
 * (SURVEY.md section 0.1): the boundary it offers is Julia multiple dispatch on the
 * function names in src/scrna/processing.jl.  Every entry point below therefore cites
 * the reference FUNCTION it replaces (file:line under /root/reference); the Julia
 * `ccall` shim that binds them is julia/CellScopesCUDA.jl and is described in
 * INTEGRATION.md.  The same symbols are bound from Python by ctypes
 * (cellscopes.jl_b200/_capi.py), which is how every test and benchmark reaches them.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  No exceptions, abort() or signals cross
 *     the boundary.  Every function returns csc_status (0 = OK, < 0 = error class);
 *     csc_last_error() gives the message.
 *   - Host buffers are caller-allocated and caller-owned; the library never keeps a
 *     host pointer after the call returns.  Device objects are opaque handles owned by
 *     the library and released with the matching *_free or with csc_destroy.
 *   - A context owns one GPU, one stream.  It is not thread-safe; distinct contexts are
 *     independent.  Multi-GPU = one context per process per GPU (cells sharded in
 *     contiguous column ranges); the few exchange steps are device-resident "partials"
 *     that are all-reduced / all-gathered between the *_partial and the *_finish call
 *     with the csc_comm_* entry points (NCCL inside the library, on the context's
 *     stream).  With one GPU the host simply skips the collective.
 *   - Matrices are genes x cells, compressed sparse COLUMN: a column is a cell
 *     (src/scrna/objects.jl:10-24).  Indices at the ABI are int64 with an explicit
 *     index_base (1 for Julia, 0 for C/Python).  On the device they are narrowed to
 *     int64 colptr / int32 rowval / fp32 values.
 */
#ifndef CELLSCOPES_CUDA_H
#define CELLSCOPES_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSC_ABI_VERSION 2

typedef int32_t csc_status;
enum {
  CSC_OK = 0,
  CSC_ERR_ARG = -1,         /* bad argument (the reference throws ArgumentError / DimensionMismatch) */
  CSC_ERR_CUDA = -2,        /* a CUDA runtime call failed */
  CSC_ERR_OOM = -3,         /* device or host allocation failed */
  CSC_ERR_UNSUPPORTED = -4, /* valid in the reference, not built here */
  CSC_ERR_NUMERIC = -5,     /* e.g. eigensolver did not converge */
  CSC_ERR_INTERNAL = -6
};

typedef struct csc_ctx csc_ctx;       /* one GPU + stream */
typedef struct csc_sparse csc_sparse; /* device CSC shard: genes x local cells */
typedef struct csc_dense csc_dense;   /* device dense fp32, cell-major: rows = cells */
typedef struct csc_vec csc_vec;       /* device buffer of fp64/fp32/int32 (partials, tables) */
typedef struct csc_pca csc_pca;       /* PCA model: loadings, variances */
typedef struct csc_graph csc_graph;   /* device CSC graph: cells x local cells */
typedef struct csc_split csc_split;   /* fp16 hi/lo operands of the scaled matrix (Gram / transform GEMMs) */

enum { CSC_DT_F32 = 0, CSC_DT_F64 = 1, CSC_DT_I32 = 2, CSC_DT_I64 = 3 };
enum { CSC_NORM_LOGARITHM = 0, CSC_NORM_NONE = 1 };       /* processing.jl:23-29 */
enum { CSC_METRIC_COSINE = 0, CSC_METRIC_EUCLIDEAN = 1 }; /* Distances.jl CosineDist / Euclidean */
enum {
  CSC_GRAPH_BINARY = 0,     /* run_clustering_small, processing.jl:268-278: union, value 1 */
  CSC_GRAPH_SUM = 1,        /* run_clustering_atlas, processing.jl:218-232: mutual pairs carry 2 */
  CSC_GRAPH_SNN_JACCARD = 2 /* extension: Jaccard weight of the shared neighbourhoods */
};
/* stage slots of csc_stage_times */
enum {
  CSC_T_UPLOAD = 0, CSC_T_NORMALIZE, CSC_T_GENE_STATS, CSC_T_HVG, CSC_T_SCALE_STATS, CSC_T_SCALE_FILL,
  CSC_T_PCA_GRAM, CSC_T_PCA_SOLVE, CSC_T_PCA_TRANSFORM, CSC_T_KNN, CSC_T_GRAPH, CSC_T_GENERATE,
  CSC_T_KNN_SCAN, /* the all-pairs scan kernel alone (inside CSC_T_KNN): the dominant kernel of the path */
  CSC_T_GRAM_MMA, /* the tcgen05 Gram kernel alone (inside CSC_T_PCA_GRAM) */
  CSC_T_ALLOC_HOST, /* HOST milliseconds spent inside the device allocator (pool growth shows up here) */
  CSC_T_MARKERS,    /* csc_marker_stats */
  CSC_T_HARMONY,    /* csc_harmony, csc_kmeans */
  CSC_T_LEIDEN,     /* csc_leiden */
  CSC_T_COUNT
};

/* ---- lifecycle ------------------------------------------------------------------------- */
int32_t csc_abi_version(void);
csc_status csc_create(int32_t device, csc_ctx** ctx);
csc_status csc_destroy(csc_ctx* ctx);
/* message of the last failure on ctx (ctx == NULL: last failure of csc_create on this thread) */
const char* csc_last_error(const csc_ctx* ctx);
csc_status csc_synchronize(csc_ctx* ctx);
csc_status csc_device_info(csc_ctx* ctx, int32_t* sm_count, int64_t* free_bytes, int64_t* total_bytes,
                           int32_t* cc_major, int32_t* cc_minor);
/* CUDA-event stopwatch on the context's stream (the stream every kernel is launched on) */
csc_status csc_timer_start(csc_ctx* ctx);
csc_status csc_timer_stop(csc_ctx* ctx, double* elapsed_ms);
/* device milliseconds spent in each stage since the last reset, and kernels launched */
csc_status csc_stage_times(csc_ctx* ctx, double* ms /*[CSC_T_COUNT]*/, int64_t* kernel_launches, int32_t reset);
/* scratch-flush: overwrite a buffer larger than L2 (bench hygiene between timed iterations) */
csc_status csc_flush_l2(csc_ctx* ctx);

/* ---- pinned host memory (optional) -------------------------------------------------------
 * Results copied into pageable memory move at ~5 GB/s (staging + first-touch page faults); into page-locked
 * memory they move at PCIe rate.  A caller that wants that allocates its result arrays here; blocks are cached
 * per context and handed back by csc_host_free.  Every download entry point accepts either kind of memory.
 * Lifetime: csc_destroy does NOT free blocks that are still live (they back result arrays the caller may still
 * read); release those with csc_host_free(NULL, p) once the context is gone. */
csc_status csc_host_alloc(csc_ctx* ctx, int64_t bytes, void** out);
csc_status csc_host_free(csc_ctx* ctx, void* p);

/* ---- device buffers (partials that the host reduces/gathers between ranks) ------------- */
csc_status csc_vec_alloc(csc_ctx* ctx, int64_t n, int32_t dtype, csc_vec** out); /* zero-filled */
/* non-owning view of caller-managed device memory (e.g. an all-gathered table held by torch) */
csc_status csc_vec_wrap(csc_ctx* ctx, void* device_ptr, int64_t n, int32_t dtype, csc_vec** out);
csc_status csc_vec_free(csc_vec* v);
csc_status csc_vec_info(const csc_vec* v, int64_t* n, int32_t* dtype, void** device_ptr);
csc_status csc_vec_upload(csc_vec* v, const void* host, int64_t n);
csc_status csc_vec_download(const csc_vec* v, void* host, int64_t n);
/* download with the conversion the host side of the reference needs done on the device first: out_dtype I32|I64 for an
 * I32 vector (every element + add: the 1-based k x N neighbour table of knn_matrices, processing.jl:196-201), F64 for
 * an F32 vector (the distance table); out_dtype == the vector's dtype with add == 0 is csc_vec_download. */
csc_status csc_vec_download_as(const csc_vec* v, void* host, int64_t n, int32_t out_dtype, int64_t add);

/* ---- sparse count matrix: rawCount.count_mtx (src/scrna/objects.jl:10-24) --------------- */
/* colptr[n_cells+1], rowval[nnz], nzval[nnz] of a SparseMatrixCSC; idx_dtype I64|I32,
 * val_dtype I64|F64|F32|I32 (readers yield Int64, merge_matrices Float64: SURVEY A.1). */
csc_status csc_sparse_upload(csc_ctx* ctx, int64_t n_genes, int64_t n_cells, const void* colptr,
                             const void* rowval, int32_t idx_dtype, const void* nzval, int32_t val_dtype,
                             int32_t index_base, csc_sparse** out);
csc_status csc_sparse_info(const csc_sparse* m, int64_t* n_genes, int64_t* n_cells, int64_t* nnz);
/* columns [cell_lo, cell_hi): colptr_out[cell_hi-cell_lo+1] rebased to index_base; rowval/nzval sized by
 * csc_sparse_range_nnz.  Any output pointer may be NULL. */
csc_status csc_sparse_range_nnz(const csc_sparse* m, int64_t cell_lo, int64_t cell_hi, int64_t* nnz);
csc_status csc_sparse_download(const csc_sparse* m, int64_t cell_lo, int64_t cell_hi, int64_t* colptr_out,
                               int64_t* rowval_out, double* nzval_out, int32_t index_base);
csc_status csc_sparse_free(csc_sparse* m);
/* synthetic module-structured Poisson counts for global cells [cell_lo, cell_hi) (the
 * per-gene tables come from cellscopes.jl_b200/synth.py::make_params) */
csc_status csc_generate_counts(csc_ctx* ctx, int64_t n_genes, int64_t cell_lo, int64_t cell_hi, uint64_t seed,
                               const float* gene_base, const int32_t* gene_module, const float* gene_loading,
                               const float* type_mean, int32_t n_types, int32_t n_modules, float lib_sd,
                               float type_w, float noise_w, csc_sparse** out);

/* ---- a1: normalize_object(mtx) src/scrna/processing.jl:15-31 ---------------------------- */
/* out shares the sparsity pattern of raw.  nan_to_zero = the replace!(NaN=>0) of :44-46.
 * gene_stats (optional, fp64[2*n_genes]) receives += per-gene (sum x, sum x^2) of the RAW counts of this
 * shard, fused into the same pass (find_variable_genes :138-139 needs them; exact for integer counts).
 * (ABI 1 also took a norm_stats table here, filled in 32-bit fixed point; it was too coarse -- see
 * csc_scale_stats_partial, which now takes the normalised moments of the selected rows with 64-bit accumulators.) */
csc_status csc_normalize(csc_ctx* ctx, const csc_sparse* raw, double scale_factor, int32_t norm_method,
                         double pseudocount, int32_t nan_to_zero, csc_vec* gene_stats, csc_sparse** out);

/* ---- a2: find_variable_genes src/scrna/processing.jl:133-155 ---------------------------- */
/* per-gene (sum x, sum x^2) of this shard, += into gene_stats fp64[2*n_genes] (standalone K2) */
csc_status csc_gene_stats_partial(csc_ctx* ctx, const csc_sparse* raw, csc_vec* gene_stats);
/* gene_stats already summed over ranks.  Outputs (host, any may be NULL): mean/variance/
 * variance_expected/variance_standardized [n_genes] = the vst_data columns of :141; order[n_genes] =
 * gene ids (0-based) sorted by variance_standardized descending, stable (:152); the first n_features of
 * it are Features (:153).  n_features is clamped to n_genes (:134-137).  From 1024 genes on (and span >= 0.02) the whole
 * finish -- sorts, Loess vertex fits, prediction, rank -- runs on the device (hvg_dev.cu, ~0.5 ms at 33k genes); smaller
 * tables and CSC_HVG_HOST=1 take the host implementation of the same algorithm (hvg.cu). */
csc_status csc_hvg_finish(csc_ctx* ctx, const csc_vec* gene_stats, int64_t n_genes, int64_t n_cells_total,
                          int64_t n_features, double span, double* mean, double* variance,
                          double* variance_expected, double* variance_standardized, int64_t* order);

/* The same for one process per GPU: called by EVERY rank of the context's communicator (csc_comm_init) with the
 * all-reduced table.  On the device path every rank runs the (deterministic) finish on its own copy and nothing is
 * exchanged; on the host path the Loess vertex fits are dealt out over the ranks and their (value, slope) pairs
 * all-reduced, so every rank returns bit-identical outputs without N copies of the same host job competing for the
 * host cores. */
csc_status csc_hvg_finish_collective(csc_ctx* ctx, const csc_vec* gene_stats, int64_t n_genes, int64_t n_cells_total,
                                     int64_t n_features, double span, double* mean, double* variance,
                                     double* variance_expected, double* variance_standardized, int64_t* order);

/* subset_count (src/scrna/utils.jl:59-87): the rows features[n_feat] (0-based, any order) of m in ORIGINAL gene order,
 * renumbered 0..H-1, as a new matrix with the same cells.  scale_object works on this subset in the reference; here it
 * pays off when few of the genes are kept (whole-transcriptome matrices): the passes after it take features == NULL. */
csc_status csc_sparse_select_rows(csc_ctx* ctx, const csc_sparse* m, const int64_t* features, int64_t n_feat,
                                  csc_sparse** out);

/* ---- a3: scale_object src/scrna/processing.jl:65-93 ------------------------------------- */
/* features[n_feat]: 0-based gene ids in any order; rows are taken in ORIGINAL gene order like
 * subset_count (src/scrna/utils.jl:72-76).  features == NULL -> all genes.
 * partial: fp64[2*H] += (sum y, sum y^2) of the normalised values of this shard: 64-bit fixed-point accumulators
 * (resolution 2^-28 of y and of y^2 for the logarithm method; deterministic) when the matrix came out of
 * csc_normalize / csc_sparse_select_rows with a usable bound on |y|, exact fp64 sums otherwise. */
csc_status csc_scale_stats_partial(csc_ctx* ctx, const csc_sparse* norm, const int64_t* features, int64_t n_feat,
                                   csc_vec* partial);
/* partial already summed over ranks; out = dense fp32 [n_cells_local x H], cell-major, which is the
 * memory layout of Julia's genes x cells column-major Matrix: min((y-mean)/sd, scale_max) (:69-74). */
csc_status csc_scale_fill(csc_ctx* ctx, const csc_sparse* norm, const int64_t* features, int64_t n_feat,
                          const csc_vec* partial, int64_t n_cells_total, double scale_max, int32_t do_scale,
                          int32_t do_center, csc_dense** out);
/* a3 + a4 fused for callers that only need the PCA of the scaled matrix (run_pca after scale_object when scaleCount
 * itself is not asked for; src/scrna/processing.jl:162-191 -> MultivariateStats.fit(PCA, ...)): adds the column sums
 * of this shard's scaled matrix into colsum[H] and its Gram matrix Z'Z into gram[H*H] (both fp64, cells of this shard
 * only, ready for the all-reduce) without materialising the dense matrix.  Same arguments as csc_scale_fill. */
csc_status csc_scale_gram_partial(csc_ctx* ctx, const csc_sparse* norm, const int64_t* features, int64_t n_feat,
                                  const csc_vec* partial, int64_t n_cells_total, double scale_max, int32_t do_scale,
                                  int32_t do_center, csc_vec* colsum, csc_vec* gram, csc_split** operands);
/* operands (optional): the fp16 hi/lo split of this shard's scaled matrix that the Gram kernel consumed, kept for
 * csc_pca_transform_split (4 H bytes per cell of device memory until csc_split_free). */
csc_status csc_split_free(csc_split* s);


/* ---- dense cell-major matrices: scaleCount.count_mtx', pca.cell_embedding ---------------- */
csc_status csc_dense_alloc(csc_ctx* ctx, int64_t n_rows, int64_t n_cols, csc_dense** out);
/* host is row-major [n_rows x n_cols] (== Julia column-major [n_cols x n_rows]) */
csc_status csc_dense_upload(csc_ctx* ctx, const void* host, int32_t dtype, int64_t n_rows, int64_t n_cols,
                            csc_dense** out);
csc_status csc_dense_download(const csc_dense* d, void* host, int32_t dtype, int64_t row_lo, int64_t row_hi);
csc_status csc_dense_info(const csc_dense* d, int64_t* n_rows, int64_t* n_cols, void** device_ptr);
/* non-owning view of caller-managed device memory (e.g. an all-gathered embedding held by torch) */
csc_status csc_dense_wrap(csc_ctx* ctx, void* device_ptr, int64_t n_rows, int64_t n_cols, csc_dense** out);
csc_status csc_dense_free(csc_dense* d);

/* ---- a4: run_pca src/scrna/processing.jl:164-193 (MultivariateStats PCA, method=:svd) ---- */
/* colsum fp64[H] += sum over local cells of z; gram fp64[H*H] += Z'Z of this shard (symmetric, full) */
csc_status csc_pca_partial(csc_ctx* ctx, const csc_dense* scaled, csc_vec* colsum, csc_vec* gram);
/* colsum/gram already summed over ranks.  Top-k eigenpairs of (Z-m)'(Z-m); principalvars = lambda/(N-1),
 * tvar = trace/(N-1).  iters_out/resid_out report the eigensolver (may be NULL). */
csc_status csc_pca_solve(csc_ctx* ctx, const csc_vec* colsum, const csc_vec* gram, int64_t n_feat,
                         int64_t n_cells_total, int32_t maxoutdim, double tol, int32_t max_iter, csc_pca** out,
                         int32_t* iters_out, double* resid_out);
csc_status csc_pca_info(const csc_pca* p, int64_t* n_feat, int32_t* k, double* principalvars /*[k]*/, double* tvar,
                        double* loadings /*[n_feat*k], column-major H x k*/, double* mean /*[n_feat]*/);
/* embedding of the local cells: (Z - 1 m') U_k, dense fp32 [n_cells_local x k] */
csc_status csc_pca_transform(csc_ctx* ctx, const csc_dense* scaled, const csc_pca* p, csc_dense** out);
/* the same embedding straight from the normalised sparse matrix: the scaled matrix is sparse-plus-rank-one
 * (z = s + 1 z0'), so (Z - 1 m')U = S U + 1 ((z0 - m)'U) costs ~nnz(selected rows) x k multiply-adds instead of
 * N x H x k.  Arguments as csc_scale_fill (features/partial/scale parameters define Z) plus the model. */
csc_status csc_pca_transform_sparse(csc_ctx* ctx, const csc_sparse* norm, const int64_t* features, int64_t n_feat,
                                    const csc_vec* partial, int64_t n_cells_total, double scale_max, int32_t do_scale,
                                    int32_t do_center, const csc_pca* p, csc_dense** out);
/* the same embedding as a tensor-core GEMM over the split operands kept by csc_scale_gram_partial: E = Z U - 1 (m'U),
 * three tcgen05 kind::f16 passes (hi Uhi', hi Ulo', lo Uhi'), HBM-bound (4 H bytes per cell read once); k <= 64 */
csc_status csc_pca_transform_split(csc_ctx* ctx, const csc_split* operands, const csc_pca* p, csc_dense** out);
csc_status csc_pca_free(csc_pca* p);

/* ---- a5/a6: kNN inside run_clustering src/scrna/processing.jl:195-201, 263-266 ----------- */
/* Exact brute force (the reference's nndescent is approximate and unseeded).  queries: local cells
 * [n_q x d]; keys: all cells [n_keys x d] (the same object with one GPU); query_offset = global id of
 * query row 0 (its own key, excluded).  Columns [dim_lo, dim_hi) are used (run_umap dims_use, :335).
 * idx int32 [n_q x k] 0-based global ids, ascending distance, ties by lower id; dist fp32 [n_q x k]. */
csc_status csc_knn(csc_ctx* ctx, const csc_dense* queries, const csc_dense* keys, int64_t query_offset,
                   int32_t dim_lo, int32_t dim_hi, int32_t k, int32_t metric, csc_vec** idx, csc_vec** dist);

/* rows of the last csc_knn call whose fp16 filter could not prove its candidate list complete and that were
 * re-scanned with the exact 3xTF32 kernel (diagnostic; the result is exact either way) */
csc_status csc_knn_last_fallback_rows(csc_ctx* ctx, int64_t* n_rows);

/* ---- adjacency fed to Leiden src/scrna/processing.jl:218-232, 268-278 ------------------- */
/* knn_idx: int32 [n_cells_total x k] of ALL cells (all-gathered); builds columns [cell_lo, cell_hi). */
csc_status csc_graph_build(csc_ctx* ctx, const csc_vec* knn_idx, int64_t n_cells_total, int32_t k,
                           int64_t cell_lo, int64_t cell_hi, int32_t mode, double snn_prune, csc_graph** out);
csc_status csc_graph_info(const csc_graph* g, int64_t* n_rows, int64_t* n_cols, int64_t* nnz);
/* nzval: fp64 or int64 [nnz] as val_dtype says (the reference's adjacency is SparseMatrixCSC{Int64,Int64}) */
csc_status csc_graph_download(const csc_graph* g, int64_t* colptr, int64_t* rowval, void* nzval, int32_t val_dtype,
                              int32_t index_base);
csc_status csc_graph_free(csc_graph* g);

/* ---- f1: community detection on the graph src/scrna/processing.jl:233-234, 279-280 (Leiden.leiden(adj_mat, resolution)) --
 * Leiden scheme (local moving, refinement inside the communities, aggregation on the refined partition) for the
 * Constant Potts Model quality H = sum_c [e_c - resolution * n_c (n_c - 1) / 2] on a graph built by csc_graph_build
 * with all columns on this device (cell_lo = 0, cell_hi = n).  Parallel and deterministic (the reference's is
 * sequential and seeded from Julia's RNG: partitions are comparable by quality, not label by label).
 * labels_out: host int32[n], communities numbered 0.. in order of first appearance; quality = H of the result. */
csc_status csc_leiden(csc_ctx* ctx, const csc_graph* g, double resolution, int32_t max_levels, int32_t max_sweeps,
                      int32_t* labels_out, int32_t* n_communities, double* quality, int32_t* levels_out);

/* ---- f2: run_harmony src/integration/int_processing.jl:1-12 -> HarmonyObject src/integration/int_objects.jl:115-236,
 *      init_cluster! / cluster! / update_R! / compute_objective! / moe_correct_ridge src/integration/int_utils.jl:49-199
 * z_orig: host fp64 [n x d], cell-major (the memory of the reference's d x N data_mat); batch[n] in [0, n_batch) = the
 * level of the batch column in sorted order (get_dummies, int_utils.jl:202-211); y0: host fp64 [K x d], the k-means
 * centres of the unit-length cells (cluster_kmeans :49-56: csc_kmeans below, or the caller's own); theta[n_batch],
 * lamb[n_batch] (the ridge diagonal without its leading 0), sigma[K]; the remaining arguments as the keywords of
 * HarmonyObject (:153-166).  seed drives the update order of update_R! (the stable argsort of a counter-based hash
 * of (seed, round, cell): Julia's randperm cannot be reproduced).  Outputs: z_corr_out host fp64 [n x d] (Z_corr, what
 * run_harmony stores as the new cell_embedding, :9-10); objective_harmony[*n_objective] (in: capacity); kmeans_rounds
 * [*n_rounds] (in: capacity); r_out (optional) host fp64 [n x K], the final soft assignment.  fp64 throughout. */
csc_status csc_harmony(csc_ctx* ctx, const double* z_orig, int64_t n, int32_t d, const int32_t* batch, int32_t n_batch,
                       const double* y0, int32_t K, const double* theta, const double* lamb, const double* sigma,
                       double block_size, int32_t max_iter_harmony, int32_t max_iter_kmeans, double eps_kmeans,
                       double eps_harmony, uint64_t seed, double* z_corr_out, double* objective_harmony,
                       int32_t* n_objective, int32_t* kmeans_rounds, int32_t* n_rounds, double* r_out);
/* Lloyd's k-means in fp64 from the rows init_idx[K] of x (host fp64 [n x d]); unit_rows != 0 applies csc_harmony's
 * Z_cos transform to the rows first.  Stops when the objective changes by less than tol (Clustering.jl's rule) or after
 * max_iter assignments.  centers_out host fp64 [K x d]; labels_out (optional) int32[n]. */
csc_status csc_kmeans(csc_ctx* ctx, const double* x, int64_t n, int32_t d, int32_t K, const int64_t* init_idx,
                      int32_t max_iter, double tol, int32_t unit_rows, double* centers_out, int32_t* labels_out,
                      double* objective_out, int32_t* iters_out);

/* ---- f4: data formats on the way into the path -----------------------------------------------------------------
 * csc_read_mtx: the MatrixMarket.mmread of read_10x (src/fileio.jl:7,16): a "coordinate integer|real|pattern general"
 * file, plain or gzip (matrix.mtx / matrix.mtx.gz), parsed on the host threads and assembled on the device like
 * sparse(I, J, V) (duplicates summed, explicit zeros kept).  rows = genes, columns = cells.
 * csc_sparse_merge: merge_matrices (src/integration/int_utils.jl:1-23): the matrices side by side, rows mapped onto
 * the sorted union of the row names by row_maps[i][r] (0-based row of the merged matrix for row r of matrix i; the
 * host owns the names); zero values are dropped (`if val != 0`, :11).
 * csc_tf_idf: run_tf_idf (src/scatac/atac_processing.jl:1-13): log1p(idf_g * (x / colsum_c) * scale_factor) with
 * idf_g = n_cells_total / rowsum_g on the stored pattern; gene_stats[0:n_genes] = the row sums, already summed over the
 * ranks (csc_gene_stats_partial fills them). */
csc_status csc_read_mtx(csc_ctx* ctx, const char* path, csc_sparse** out);
/* colSum (src/scrna/utils.jl:1-2): host fp64[n_cells]; and the column half of subset_matrix / subset_count
 * (utils.jl:59-87, 313-327): the columns cols[n_sel] (0-based, any order, repeats allowed) as a new matrix */
csc_status csc_sparse_col_sums(const csc_sparse* m, double* out);
csc_status csc_sparse_select_cols(csc_ctx* ctx, const csc_sparse* m, const int64_t* cols, int64_t n_sel, csc_sparse** out);
csc_status csc_sparse_merge(csc_ctx* ctx, const csc_sparse* const* mats, const int64_t* const* row_maps, int32_t n_mats,
                            int64_t n_rows_merged, csc_sparse** out);
csc_status csc_tf_idf(csc_ctx* ctx, const csc_sparse* raw, const csc_vec* gene_stats, int64_t n_cells_total,
                      double scale_factor, csc_sparse** out);

/* ---- f3: find_markers / find_all_markers src/scrna/processing.jl:350-414 --------------------------------------
 * Sufficient statistics of the per-gene Mann-Whitney U tests (HypothesisTests.MannWhitneyUTest, :376) and of the
 * columns of the result frame (:378-382) for EVERY group of cells at once, from one sort of the stored entries of the
 * normalised matrix.  labels: host int32[n_cells]; -1 = the cell takes no part (find_markers: cluster_1 -> 0,
 * cluster_2 -> 1; find_all_markers: label = cluster, "cluster c against all other cells" then needs only c's rank sum).
 * Outputs, host fp64, [n_genes x n_groups] row-major: rank_sum = sum of the (tie-averaged, 1-based) ranks of the
 * STORED entries of the group among all participating cells; sum_expm1 = sum expm1(x) (:378-379); n_expr = #(x >
 * expr_cutoff) (:381-382); nnz = #stored; sum_x = sum x (the rowSum > 0 filter of :364-366).  Per gene [n_genes]:
 * tie_adj = sum (t^3 - t) over the tie groups of the non-zero values; n_neg = #negative values; n_zero_stored = #stored
 * zeros.  The implicit zeros of a gene are one more tie group, ranked n_neg + (n0 + 1) / 2 with n0 = *n_included - (all
 * stored entries of the gene) + n_zero_stored: the caller adds (n_label - nnz) of them per group and n0^3 - n0 to the
 * tie term.  Limits: n_genes <= 65535, n_groups <= 65535 (16 bits each in the sort key). */
csc_status csc_marker_stats(csc_ctx* ctx, const csc_sparse* norm, const int32_t* labels, int32_t n_groups,
                            double expr_cutoff, double* rank_sum, double* sum_expm1, double* n_expr, double* nnz,
                            double* sum_x, double* tie_adj, double* n_neg, double* n_zero_stored, int64_t* n_included);

/* ---- multi-GPU: the exchange steps of the path, NCCL over NVLink inside the library ----------------
 * One process per GPU, one context each; cells sharded in contiguous column ranges.  The reference is single-process
 * Julia (no counterpart to cite); these are the collectives BASELINE.json's north_star names: all-reduce of the per-gene
 * statistics (processing.jl:138-139, :66-67 taken per shard) and of the Gram / column sums (run_pca :164-193), all-gather
 * of the embedding before the kNN query pass and of the k x N neighbour table before the adjacency (:218-232).
 * NCCL is dlopen'ed on first use (libnccl.so.2; CSC_NCCL_LIB overrides); without it these return CSC_ERR_UNSUPPORTED.
 * Rendezvous: rank 0 calls csc_comm_unique_id, the HOST moves the 128 bytes to every rank (MPI.jl, Distributed.jl,
 * a shared file, torch.distributed ...), every rank calls csc_comm_init with the same id.
 * All collectives are enqueued on the context's stream and return without waiting: the kernels of the next stage are
 * ordered behind them on the same stream. */
csc_status csc_comm_unique_id(void* id_out /*[128]*/, int64_t id_bytes);
csc_status csc_comm_init(csc_ctx* ctx, const void* unique_id, int64_t id_bytes, int32_t rank, int32_t world);
csc_status csc_comm_info(const csc_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version);
csc_status csc_comm_destroy(csc_ctx* ctx); /* also done by csc_destroy */
/* v <- sum over ranks of v, in place (fp64 / fp32 / int32 / int64) */
csc_status csc_comm_allreduce_sum(csc_ctx* ctx, csc_vec* v);
/* rows of every rank, concatenated in rank order; row_offsets[world + 1] = global row ranges of the ranks (ragged
 * shards need no padding).  out: dense [row_offsets[world] x n_cols] / int32 [row_offsets[world] * width]. */
csc_status csc_comm_allgather_rows(csc_ctx* ctx, const csc_dense* local, const int64_t* row_offsets, csc_dense** out);
csc_status csc_comm_allgather_i32(csc_ctx* ctx, const csc_vec* local, const int64_t* row_offsets, int32_t width,
                                  csc_vec** out);

#ifdef __cplusplus
}
#endif
#endif /* CELLSCOPES_CUDA_H */
