// Adjacency fed to Leiden (src/scrna/processing.jl:218-232 atlas, 268-278 small) and the SNN/Jaccard
// extension, built on the device from the exact kNN table.
//
// Column i of the result = kNN(i) U { j : i in kNN(j) }.  Every directed kNN edge contributes the two
// COO entries (idx, i) and (i, idx) exactly like the reference's push! loop; they are encoded as 64-bit
// keys (column << row_bits | row, only as many bits as the sizes need: 39 for 500k cells, i.e. 5 radix passes
// instead of 8), radix-sorted, and run-length encoded: the run length (1 or 2) is the value
// the reference's sparse(I, J, V) would have summed.
#include <cub/cub.cuh>

#include "common.cuh"


// forward entries (col = local cell, row = neighbour) and reverse entries (col = neighbour if local, row = cell)
__global__ void k_graph_keys(const int32_t* __restrict__ idx, int64_t n_total, int32_t k, int64_t lo, int64_t hi, int row_bits,
                             unsigned long long sentinel, unsigned long long* __restrict__ keys,
                             unsigned long long* __restrict__ n_valid) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_fwd = (hi - lo) * k;
  const int64_t n_all = n_fwd + n_total * k;
  unsigned long long key = sentinel;             // sorts after every valid key
  if (e < n_fwd) {
    const int64_t cell = lo + e / k;
    const int32_t nb = idx[cell * k + e % k];
    // an id outside [0, n_total) (a table gathered from elsewhere, a placeholder of a failed kNN) would spill into the
    // column bits of the key and index the SNN lists out of bounds: flag it, the host turns the flag into CSC_ERR_ARG
    if (nb < 0 || nb >= n_total) n_valid[1] = 1ull;
    else key = ((unsigned long long)(cell - lo) << row_bits) | (unsigned int)nb;
  } else if (e < n_all) {
    const int64_t r = e - n_fwd;
    const int64_t cell = r / k;
    const int32_t nb = idx[r];
    if (nb < 0 || nb >= n_total) n_valid[1] = 1ull;
    if (nb >= lo && nb < hi) key = ((unsigned long long)(nb - lo) << row_bits) | (unsigned int)cell;
  }
  if (e < n_all) keys[e] = key;
  // one atomic per CTA: one per warp was 625k same-address atomics at C2, 0.4 of the kernel's 0.5 ms
  const int valid = __syncthreads_count(key != sentinel);
  if (threadIdx.x == 0 && valid) atomicAdd(n_valid, (unsigned long long)valid);
}

__global__ void k_graph_heads(const unsigned long long* __restrict__ keys, int64_t n, int64_t* __restrict__ head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// unique entries: row, multiplicity
__global__ void k_graph_compact(const unsigned long long* __restrict__ keys, const int64_t* __restrict__ pos, int64_t n,
                                unsigned long long* __restrict__ ukeys, int32_t* __restrict__ mult) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool is_head = (i == 0 || keys[i] != keys[i - 1]);
  if (!is_head) return;
  int64_t j = i + 1;
  while (j < n && keys[j] == keys[i]) ++j;
  ukeys[pos[i]] = keys[i];
  mult[pos[i]] = (int32_t)(j - i);
}

// sorted {i} U kNN(i) per cell (ascending id), one thread per cell, k+1 <= 129
__global__ void k_sort_lists(const int32_t* __restrict__ idx, int64_t n_total, int32_t k, int32_t* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_total) return;
  int32_t* o = out + c * (k + 1);
  o[0] = (int32_t)c;
  for (int j = 0; j < k; ++j) {
    const int32_t v = idx[c * k + j];
    int p = j + 1;
    while (p > 0 && o[p - 1] > v) {
      o[p] = o[p - 1];
      --p;
    }
    o[p] = v;
  }
}

// value of every unique entry; keep flag for pruning
__global__ void k_graph_values(const unsigned long long* __restrict__ ukeys, const int32_t* __restrict__ mult, int64_t nnz,
                               int64_t lo, int32_t k, int mode, double prune, const int32_t* __restrict__ lists, int row_bits,
                               double* __restrict__ val, int64_t* __restrict__ keep) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  double v;
  if (mode == CSC_GRAPH_BINARY) {
    v = 1.0;
  } else if (mode == CSC_GRAPH_SUM) {
    v = (double)mult[e];
  } else {
    const int64_t col = (int64_t)(ukeys[e] >> row_bits) + lo;
    const int64_t row = (int64_t)(ukeys[e] & ((1ull << row_bits) - 1));
    const int32_t* a = lists + col * (k + 1);
    const int32_t* b = lists + row * (k + 1);
    int i = 0, j = 0, inter = 0;
    while (i <= k && j <= k) {
      const int32_t x = a[i], y = b[j];
      if (x == y) {
        ++inter;
        ++i;
        ++j;
      } else if (x < y) {
        ++i;
      } else {
        ++j;
      }
    }
    v = (double)inter / (double)(2 * (k + 1) - inter);
  }
  val[e] = v;
  keep[e] = (mode != CSC_GRAPH_SNN_JACCARD || v >= prune) ? 1 : 0;
}

__global__ void k_graph_emit(const unsigned long long* __restrict__ ukeys, const double* __restrict__ val,
                             const int64_t* __restrict__ keep, const int64_t* __restrict__ pos, int64_t nnz, int row_bits,
                             int32_t* __restrict__ rowval, double* __restrict__ out_val, int32_t* __restrict__ out_col) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz || !keep[e]) return;
  const int64_t p = pos[e];
  rowval[p] = (int32_t)(ukeys[e] & ((1ull << row_bits) - 1));
  out_val[p] = val[e];
  out_col[p] = (int32_t)(ukeys[e] >> row_bits);
}

// colptr[c] = first entry whose column >= c (entries are sorted by column)
__global__ void k_graph_colptr(const int32_t* __restrict__ cols, int64_t nnz, int64_t n_cols, int64_t* __restrict__ colptr) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_cols) return;
  int64_t a = 0, b = nnz;
  while (a < b) {
    const int64_t m = (a + b) >> 1;
    if (cols[m] < c) a = m + 1;
    else b = m;
  }
  colptr[c] = a;
}

static inline unsigned nblk(int64_t n) { return (unsigned)std::max<int64_t>(1, ceil_div64(n, 256)); }

static csc_status exclusive_scan(csc_ctx* ctx, const int64_t* in, int64_t* out, int64_t n) {
  size_t tmp_bytes = 0;
  CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, n, ctx->stream));
  DevPtr tmp;
  CSC_TRY(dev_alloc(ctx, tmp_bytes, &tmp));
  CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp->p, tmp_bytes, in, out, n, ctx->stream));
  ctx->launches += 2;
  return CSC_OK;
}

extern "C" csc_status csc_graph_build(csc_ctx* ctx, const csc_vec* knn_idx, int64_t n_cells_total, int32_t k,
                                      int64_t cell_lo, int64_t cell_hi, int32_t mode, double snn_prune, csc_graph** out) {
  if (!ctx || !knn_idx || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, knn_idx->dtype == CSC_DT_I32 && knn_idx->n >= n_cells_total * k, "csc_graph_build: knn_idx must be int32[n_cells_total*k]");
  CSC_REQUIRE(ctx, k >= 1 && k <= 128, "csc_graph_build: k=%d outside [1,128]", k);
  CSC_REQUIRE(ctx, 0 <= cell_lo && cell_lo <= cell_hi && cell_hi <= n_cells_total, "csc_graph_build: bad cell range");
  CSC_REQUIRE(ctx, mode >= CSC_GRAPH_BINARY && mode <= CSC_GRAPH_SNN_JACCARD, "csc_graph_build: unknown mode %d", mode);
  CSC_REQUIRE(ctx, n_cells_total < ((int64_t)1 << 31), "csc_graph_build: more than 2^31 cells");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_GRAPH);
  const int64_t n_loc = cell_hi - cell_lo;
  const int64_t n_all = n_loc * k + n_cells_total * k;
  const int32_t* idx = knn_idx->as<int32_t>();
  CSC_TRACE(ctx, "graph: begin");

  std::unique_ptr<csc_graph> g(new csc_graph());
  g->ctx = ctx;
  g->n_rows = n_cells_total;
  g->n_cols = n_loc;
  g->col_lo = cell_lo;
  g->nnz = 0;
  CSC_TRY(dev_alloc(ctx, (size_t)(n_loc + 1) * 8, &g->colptr, true));
  if (n_loc == 0 || n_all == 0) {
    CSC_TRY(dev_alloc(ctx, 0, &g->rowval));
    CSC_TRY(dev_alloc(ctx, 0, &g->val));
    *out = g.release();
    return CSC_OK;
  }
  DevPtr keys_a, keys_b, counter;
  CSC_TRY(dev_alloc(ctx, (size_t)n_all * 8, &keys_a));
  CSC_TRY(dev_alloc(ctx, (size_t)n_all * 8, &keys_b));
  CSC_TRY(dev_alloc(ctx, 16, &counter, true));      // [0] valid keys, [1] bad-id flag
  CSC_TRACE(ctx, "graph: alloc keys");
  int row_bits = 1, col_bits = 1;
  while (((int64_t)1 << row_bits) < n_cells_total) ++row_bits;
  while (((int64_t)1 << col_bits) < n_loc) ++col_bits;
  // "no entry" key: sorts after every valid key.  The all-ones pattern of the key's own bits is free unless the largest
  // valid key (column n_loc - 1, row n_cells_total - 1) is all ones itself, i.e. both sizes are powers of two; only then
  // does the sentinel need a bit of its own (at 1M cells that bit was a sixth radix pass over 40 M keys).
  const bool free_top = (((int64_t)1 << row_bits) != n_cells_total) || (((int64_t)1 << col_bits) != n_loc);
  const unsigned long long sentinel = free_top ? (1ull << (row_bits + col_bits)) - 1ull : 1ull << (row_bits + col_bits);
  k_graph_keys<<<nblk(n_all), 256, 0, ctx->stream>>>(idx, n_cells_total, k, cell_lo, cell_hi, row_bits, sentinel,
                                                      (unsigned long long*)keys_a->p, (unsigned long long*)counter->p);
  CSC_LAUNCHED(ctx);
  {
    const int key_bits = row_bits + col_bits + (free_top ? 0 : 1);
    size_t tmp_bytes = 0;
    CSC_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (const unsigned long long*)keys_a->p,
                                                 (unsigned long long*)keys_b->p, n_all, 0, key_bits, ctx->stream));
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, tmp_bytes, &tmp));
    // sentinels sort last
    CSC_CUDA(ctx, cub::DeviceRadixSort::SortKeys(tmp->p, tmp_bytes, (const unsigned long long*)keys_a->p,
                                                 (unsigned long long*)keys_b->p, n_all, 0, key_bits, ctx->stream));
    ctx->launches += 1 + (key_bits + 7) / 8;
  }
  CSC_TRACE(ctx, "graph: keys+sort");
  unsigned long long h_counter[2] = {0, 0};
  CSC_CUDA(ctx, cudaMemcpyAsync(h_counter, counter->p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CSC_REQUIRE(ctx, h_counter[1] == 0, "csc_graph_build: a neighbour id is outside [0, n_cells_total)");
  const int64_t nv = (int64_t)h_counter[0];
  const unsigned long long* skeys = (const unsigned long long*)keys_b->p;
  // run-length encode
  DevPtr head, pos;
  CSC_TRY(dev_alloc(ctx, (size_t)(nv + 1) * 8, &head, true));
  CSC_TRY(dev_alloc(ctx, (size_t)(nv + 1) * 8, &pos));
  k_graph_heads<<<nblk(nv), 256, 0, ctx->stream>>>(skeys, nv, (int64_t*)head->p);
  CSC_LAUNCHED(ctx);
  CSC_TRY(exclusive_scan(ctx, (const int64_t*)head->p, (int64_t*)pos->p, nv + 1));
  CSC_TRACE(ctx, "graph: heads+scan");
  int64_t n_unique = 0;
  CSC_CUDA(ctx, cudaMemcpyAsync(&n_unique, (int64_t*)pos->p + nv, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  DevPtr ukeys, mult, val, keep, kpos;
  CSC_TRY(dev_alloc(ctx, (size_t)n_unique * 8, &ukeys));
  CSC_TRY(dev_alloc(ctx, (size_t)n_unique * 4, &mult));
  CSC_TRY(dev_alloc(ctx, (size_t)n_unique * 8, &val));
  CSC_TRY(dev_alloc(ctx, (size_t)(n_unique + 1) * 8, &keep, true));
  CSC_TRY(dev_alloc(ctx, (size_t)(n_unique + 1) * 8, &kpos));
  k_graph_compact<<<nblk(nv), 256, 0, ctx->stream>>>(skeys, (const int64_t*)pos->p, nv, (unsigned long long*)ukeys->p,
                                                      (int32_t*)mult->p);
  CSC_LAUNCHED(ctx);
  DevPtr lists;
  if (mode == CSC_GRAPH_SNN_JACCARD) {
    CSC_TRY(dev_alloc(ctx, (size_t)n_cells_total * (k + 1) * 4, &lists));
    k_sort_lists<<<nblk(n_cells_total), 256, 0, ctx->stream>>>(idx, n_cells_total, k, (int32_t*)lists->p);
    CSC_LAUNCHED(ctx);
  }
  k_graph_values<<<nblk(n_unique), 256, 0, ctx->stream>>>((const unsigned long long*)ukeys->p, (const int32_t*)mult->p,
                                                           n_unique, cell_lo, k, mode, snn_prune,
                                                           lists ? (const int32_t*)lists->p : nullptr, row_bits,
                                                           (double*)val->p, (int64_t*)keep->p);
  CSC_LAUNCHED(ctx);
  CSC_TRY(exclusive_scan(ctx, (const int64_t*)keep->p, (int64_t*)kpos->p, n_unique + 1));
  CSC_TRACE(ctx, "graph: compact+values+scan");
  int64_t nnz = 0;
  CSC_CUDA(ctx, cudaMemcpyAsync(&nnz, (int64_t*)kpos->p + n_unique, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  g->nnz = nnz;
  DevPtr cols;
  CSC_TRY(dev_alloc(ctx, (size_t)nnz * 4, &g->rowval));
  CSC_TRY(dev_alloc(ctx, (size_t)nnz * 8, &g->val));
  CSC_TRY(dev_alloc(ctx, (size_t)nnz * 4, &cols));
  k_graph_emit<<<nblk(n_unique), 256, 0, ctx->stream>>>((const unsigned long long*)ukeys->p, (const double*)val->p,
                                                         (const int64_t*)keep->p, (const int64_t*)kpos->p, n_unique, row_bits,
                                                         (int32_t*)g->rowval->p, (double*)g->val->p, (int32_t*)cols->p);
  CSC_LAUNCHED(ctx);
  k_graph_colptr<<<nblk(n_loc + 1), 256, 0, ctx->stream>>>((const int32_t*)cols->p, nnz, n_loc, (int64_t*)g->colptr->p);
  CSC_LAUNCHED(ctx);
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CSC_TRACE(ctx, "graph: emit+colptr");
  *out = g.release();
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_graph_info(const csc_graph* g, int64_t* n_rows, int64_t* n_cols, int64_t* nnz) {
  if (!g) return CSC_ERR_ARG;
  if (n_rows) *n_rows = g->n_rows;
  if (n_cols) *n_cols = g->n_cols;
  if (nnz) *nnz = g->nnz;
  return CSC_OK;
}

__global__ void k_i32_to_i64_base(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n, int64_t base) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int64_t)in[i] + base;
}
__global__ void k_i64_add_base(const int64_t* __restrict__ in, int64_t* __restrict__ out, int64_t n, int64_t base) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] + base;
}

__global__ void k_f64_to_i64(const double* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int64_t)llrint(in[i]);
}

extern "C" csc_status csc_graph_download(const csc_graph* g, int64_t* colptr, int64_t* rowval, void* nzval, int32_t val_dtype,
                                         int32_t index_base) {
  if (!g) return CSC_ERR_ARG;
  csc_ctx* ctx = g->ctx;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, val_dtype == CSC_DT_F64 || val_dtype == CSC_DT_I64, "csc_graph_download: val_dtype must be F64 or I64");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (colptr) {
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, (size_t)(g->n_cols + 1) * 8, &tmp));
    k_i64_add_base<<<nblk(g->n_cols + 1), 256, 0, ctx->stream>>>((const int64_t*)g->colptr->p, (int64_t*)tmp->p, g->n_cols + 1,
                                                                  index_base);
    CSC_LAUNCHED(ctx);
    CSC_TRY(d2h_copy(ctx, colptr, tmp->p, (size_t)(g->n_cols + 1) * 8));
  }
  if (rowval && g->nnz > 0) {
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, (size_t)g->nnz * 8, &tmp));
    k_i32_to_i64_base<<<nblk(g->nnz), 256, 0, ctx->stream>>>((const int32_t*)g->rowval->p, (int64_t*)tmp->p, g->nnz, index_base);
    CSC_LAUNCHED(ctx);
    CSC_TRY(d2h_copy(ctx, rowval, tmp->p, (size_t)g->nnz * 8));
  }
  if (nzval && g->nnz > 0) {
    if (val_dtype == CSC_DT_F64) {
      CSC_TRY(d2h_copy(ctx, nzval, g->val->p, (size_t)g->nnz * 8));
    } else {
      DevPtr tmp;
      CSC_TRY(dev_alloc(ctx, (size_t)g->nnz * 8, &tmp));
      k_f64_to_i64<<<nblk(g->nnz), 256, 0, ctx->stream>>>((const double*)g->val->p, (int64_t*)tmp->p, g->nnz);
      CSC_LAUNCHED(ctx);
      CSC_TRY(d2h_copy(ctx, nzval, tmp->p, (size_t)g->nnz * 8));
    }
  }
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_graph_free(csc_graph* g) {
  delete g;
  return CSC_OK;
}
