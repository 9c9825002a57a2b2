// Data formats on the way INTO the path: the MatrixMarket reader behind read_10x (src/fileio.jl:1-37),
// merge_matrices (src/integration/int_utils.jl:1-23) and the TF-IDF normalisation of scATAC counts
// (src/scatac/atac_processing.jl:1-13).  All three end in the device CSC layout the path starts from.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <thread>

#include <cub/cub.cuh>

#include "common.cuh"

namespace {

__global__ void k_coo_keys(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int64_t n, int row_bits,
                           unsigned long long* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((unsigned long long)cols[i] << row_bits) | (unsigned int)rows[i];
}

// sorted (col, row) keys + values -> rowval / val; duplicates are SUMMED into their first occurrence like sparse(I, J, V)
// (head[i] = 1 where a new (col,row) starts; pos = exclusive scan of head)
__global__ void k_coo_heads(const unsigned long long* __restrict__ keys, const float* __restrict__ vals, int64_t n, int drop_zero,
                            int64_t* __restrict__ head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool h = i == 0 || keys[i] != keys[i - 1];
  bool keep = h;
  if (h && drop_zero) {
    // a stored entry whose duplicates sum to zero is still stored by sparse(); only literal zeros of the input are
    // dropped where the reference tests `val != 0` (merge_matrices)
    float s = 0.f;
    for (int64_t j = i; j < n && keys[j] == keys[i]; ++j) s += vals[j];
    keep = s != 0.f;
  }
  head[i] = keep ? 1 : 0;
}

__global__ void k_coo_emit(const unsigned long long* __restrict__ keys, const float* __restrict__ vals, const int64_t* __restrict__ head,
                           const int64_t* __restrict__ pos, int64_t n, int row_bits, int32_t* __restrict__ rowval,
                           float* __restrict__ val, int32_t* __restrict__ col_of) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !head[i]) return;
  float s = 0.f;
  for (int64_t j = i; j < n && keys[j] == keys[i]; ++j) s += vals[j];
  const int64_t p = pos[i];
  rowval[p] = (int32_t)(keys[i] & ((1ull << row_bits) - 1ull));
  val[p] = s;
  col_of[p] = (int32_t)(keys[i] >> row_bits);
}

// colptr[c] = first entry whose column is >= c (entries sorted by column)
__global__ void k_colptr_from_cols(const int32_t* __restrict__ col_of, int64_t nnz, int64_t n_cols, int64_t* __restrict__ colptr) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_cols) return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col_of[mid] < c) lo = mid + 1;
    else hi = mid;
  }
  colptr[c] = lo;
}

inline unsigned nblk(int64_t n) { return (unsigned)std::max<int64_t>(1, ceil_div64(n, 256)); }

// device COO (0-based int32 rows / cols, fp32 values; any order, duplicates allowed) -> csc_sparse
csc_status coo_to_csc(csc_ctx* ctx, const int32_t* d_rows, const int32_t* d_cols, const float* d_vals, int64_t n, int64_t n_rows,
                      int64_t n_cols, bool drop_zero, csc_sparse** out) {
  std::unique_ptr<csc_sparse> m(new csc_sparse());
  m->ctx = ctx;
  m->n_genes = n_rows;
  m->n_cells = n_cols;
  m->nnz = 0;
  CSC_TRY(dev_alloc(ctx, (size_t)(n_cols + 1) * 8, &m->colptr, true));
  if (n == 0) {
    CSC_TRY(dev_alloc(ctx, 0, &m->rowval));
    CSC_TRY(dev_alloc(ctx, 0, &m->val));
    *out = m.release();
    return CSC_OK;
  }
  int row_bits = 1, col_bits = 1;
  while (((int64_t)1 << row_bits) < n_rows) ++row_bits;
  while (((int64_t)1 << col_bits) < n_cols) ++col_bits;
  DevPtr ka, kb, va, vb, head, pos;
  CSC_TRY(dev_alloc(ctx, (size_t)n * 8, &ka));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 8, &kb));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &vb));
  k_coo_keys<<<nblk(n), 256, 0, ctx->stream>>>(d_rows, d_cols, n, row_bits, (unsigned long long*)ka->p);
  CSC_LAUNCHED(ctx);
  {
    size_t tb = 0;
    CSC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tb, (const unsigned long long*)ka->p, (unsigned long long*)kb->p, d_vals,
                                                  (float*)vb->p, n, 0, row_bits + col_bits, ctx->stream));
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, tb, &tmp));
    CSC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp->p, tb, (const unsigned long long*)ka->p, (unsigned long long*)kb->p, d_vals,
                                                  (float*)vb->p, n, 0, row_bits + col_bits, ctx->stream));
    ctx->launches += 1 + (row_bits + col_bits + 7) / 8;
  }
  CSC_TRY(dev_alloc(ctx, (size_t)(n + 1) * 8, &head, true));
  CSC_TRY(dev_alloc(ctx, (size_t)(n + 1) * 8, &pos));
  k_coo_heads<<<nblk(n), 256, 0, ctx->stream>>>((const unsigned long long*)kb->p, (const float*)vb->p, n, drop_zero ? 1 : 0,
                                                (int64_t*)head->p);
  CSC_LAUNCHED(ctx);
  {
    size_t tb = 0;
    CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int64_t*)head->p, (int64_t*)pos->p, n + 1, ctx->stream));
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, tb, &tmp));
    CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp->p, tb, (const int64_t*)head->p, (int64_t*)pos->p, n + 1, ctx->stream));
    ctx->launches += 2;
  }
  int64_t nnz = 0;
  CSC_CUDA(ctx, cudaMemcpyAsync(&nnz, (const int64_t*)pos->p + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  m->nnz = nnz;
  DevPtr col_of;
  CSC_TRY(dev_alloc(ctx, (size_t)nnz * 4, &m->rowval));
  CSC_TRY(dev_alloc(ctx, (size_t)nnz * 4, &m->val));
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(nnz, 1) * 4, &col_of));
  k_coo_emit<<<nblk(n), 256, 0, ctx->stream>>>((const unsigned long long*)kb->p, (const float*)vb->p, (const int64_t*)head->p,
                                               (const int64_t*)pos->p, n, row_bits, (int32_t*)m->rowval->p, m->v(), (int32_t*)col_of->p);
  CSC_LAUNCHED(ctx);
  k_colptr_from_cols<<<nblk(n_cols + 1), 256, 0, ctx->stream>>>((const int32_t*)col_of->p, nnz, n_cols, (int64_t*)m->colptr->p);
  CSC_LAUNCHED(ctx);
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = m.release();
  return CSC_OK;
}

// ------------------------------------------------------------------------------------------
// MatrixMarket "coordinate" text -> COO on host threads
// ------------------------------------------------------------------------------------------
struct MtxChunk {
  std::vector<int32_t> rows, cols;
  std::vector<float> vals;
  int64_t bad_line = -1;
};

// parses the lines that START inside [b, e) of the data section
void parse_chunk(const char* p, const char* end_all, const char* b, const char* e, bool pattern, int64_t n_rows, int64_t n_cols,
                 MtxChunk* out) {
  const char* q = b;
  while (q < e) {
    // skip blank space at the start of a line
    while (q < end_all && (*q == ' ' || *q == '\t' || *q == '\r' || *q == '\n')) ++q;
    if (q >= e || q >= end_all) break;
    const char* line = q;
    if (*q == '%') {
      while (q < end_all && *q != '\n') ++q;
      continue;
    }
    int64_t i = 0, j = 0;
    bool ok = false;
    while (q < end_all && *q >= '0' && *q <= '9') { i = i * 10 + (*q - '0'); ++q; ok = true; }
    while (q < end_all && (*q == ' ' || *q == '\t')) ++q;
    bool ok2 = false;
    while (q < end_all && *q >= '0' && *q <= '9') { j = j * 10 + (*q - '0'); ++q; ok2 = true; }
    double v = 1.0;
    bool ok3 = pattern;
    if (!pattern) {
      while (q < end_all && (*q == ' ' || *q == '\t')) ++q;
      char* endp = nullptr;
      v = strtod(q, &endp);
      ok3 = endp != q;
      q = endp;
    }
    if (!ok || !ok2 || !ok3 || i < 1 || i > n_rows || j < 1 || j > n_cols) {
      out->bad_line = (int64_t)(line - p);
      return;
    }
    out->rows.push_back((int32_t)(i - 1));
    out->cols.push_back((int32_t)(j - 1));
    out->vals.push_back((float)v);
    while (q < end_all && *q != '\n') ++q;
  }
}

// whole file (plain or gzip) into memory
csc_status slurp(csc_ctx* ctx, const char* path, std::vector<char>* buf) {
  gzFile f = gzopen(path, "rb");          // reads plain files transparently
  if (!f) return csc_fail(ctx, CSC_ERR_ARG, "cannot open %s", path);
  gzbuffer(f, 1 << 20);
  size_t used = 0;
  buf->resize((size_t)64 << 20);
  while (true) {
    if (used == buf->size()) buf->resize(buf->size() * 2);
    const int got = gzread(f, buf->data() + used, (unsigned)std::min<size_t>(buf->size() - used, (size_t)1 << 30));
    if (got < 0) {
      gzclose(f);
      return csc_fail(ctx, CSC_ERR_ARG, "read error in %s", path);
    }
    if (got == 0) break;
    used += (size_t)got;
  }
  gzclose(f);
  buf->resize(used);
  return CSC_OK;
}

}  // namespace

extern "C" csc_status csc_read_mtx(csc_ctx* ctx, const char* path, csc_sparse** out) {
  if (!ctx || !path || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_UPLOAD);
  std::vector<char> buf;
  CSC_TRY(slurp(ctx, path, &buf));
  const char* p = buf.data();
  const char* end_all = p + buf.size();
  // banner: %%MatrixMarket matrix coordinate <integer|real|pattern> general
  const char* q = p;
  const char* eol = (const char*)memchr(q, '\n', buf.size());
  CSC_REQUIRE(ctx, eol && buf.size() > 14 && strncmp(q, "%%MatrixMarket", 14) == 0, "%s: no MatrixMarket banner", path);
  std::string banner(q, eol);
  for (auto& ch : banner) ch = (char)tolower(ch);
  CSC_REQUIRE(ctx, banner.find("coordinate") != std::string::npos, "%s: only the coordinate format is supported", path);
  CSC_REQUIRE(ctx, banner.find("general") != std::string::npos, "%s: only general (unsymmetric) matrices are supported", path);
  const bool pattern = banner.find("pattern") != std::string::npos;
  q = eol + 1;
  while (q < end_all && *q == '%') {
    const char* e2 = (const char*)memchr(q, '\n', (size_t)(end_all - q));
    q = e2 ? e2 + 1 : end_all;
  }
  long long n_rows = 0, n_cols = 0, n_ent = 0;
  {
    const char* e2 = (const char*)memchr(q, '\n', (size_t)(end_all - q));
    std::string size_line(q, e2 ? e2 : end_all);
    CSC_REQUIRE(ctx, sscanf(size_line.c_str(), "%lld %lld %lld", &n_rows, &n_cols, &n_ent) == 3 && n_rows > 0 && n_cols >= 0 && n_ent >= 0,
                "%s: bad size line", path);
    q = e2 ? e2 + 1 : end_all;
  }
  CSC_REQUIRE(ctx, n_rows < ((int64_t)1 << 31) && n_cols < ((int64_t)1 << 31), "%s: dimensions must fit int32", path);
  // the data section, cut at line starts, parsed on the host threads
  const size_t n_thr = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), 32));
  const size_t data_bytes = (size_t)(end_all - q);
  std::vector<MtxChunk> chunks(n_thr);
  std::vector<const char*> cut(n_thr + 1, end_all);
  cut[0] = q;
  for (size_t t = 1; t < n_thr; ++t) {
    const char* c = q + data_bytes * t / n_thr;
    const char* nl = (const char*)memchr(c, '\n', (size_t)(end_all - c));
    cut[t] = nl ? nl + 1 : end_all;
  }
  {
    std::vector<std::thread> pool;
    for (size_t t = 0; t < n_thr; ++t) {
      chunks[t].rows.reserve((size_t)n_ent / n_thr + 16);
      chunks[t].cols.reserve((size_t)n_ent / n_thr + 16);
      chunks[t].vals.reserve((size_t)n_ent / n_thr + 16);
    }
    for (size_t t = 1; t < n_thr; ++t)
      pool.emplace_back(parse_chunk, p, end_all, cut[t], cut[t + 1], pattern, (int64_t)n_rows, (int64_t)n_cols, &chunks[t]);
    parse_chunk(p, end_all, cut[0], cut[1], pattern, (int64_t)n_rows, (int64_t)n_cols, &chunks[0]);
    for (auto& th : pool) th.join();
  }
  int64_t n = 0;
  for (auto& c : chunks) {
    CSC_REQUIRE(ctx, c.bad_line < 0, "%s: malformed or out-of-range entry at byte %lld", path, (long long)c.bad_line);
    n += (int64_t)c.rows.size();
  }
  CSC_REQUIRE(ctx, n == n_ent, "%s: the size line announces %lld entries, the file holds %lld", path, n_ent, (long long)n);
  DevPtr d_rows, d_cols, d_vals;
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4, &d_rows));
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4, &d_cols));
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4, &d_vals));
  int64_t o = 0;
  for (auto& c : chunks) {
    const size_t k = c.rows.size();
    if (!k) continue;
    CSC_CUDA(ctx, cudaMemcpyAsync((int32_t*)d_rows->p + o, c.rows.data(), k * 4, cudaMemcpyHostToDevice, ctx->stream));
    CSC_CUDA(ctx, cudaMemcpyAsync((int32_t*)d_cols->p + o, c.cols.data(), k * 4, cudaMemcpyHostToDevice, ctx->stream));
    CSC_CUDA(ctx, cudaMemcpyAsync((float*)d_vals->p + o, c.vals.data(), k * 4, cudaMemcpyHostToDevice, ctx->stream));
    o += (int64_t)k;
  }
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // MatrixMarket.mmread builds sparse(I, J, V): duplicates are summed, explicit zeros stay stored
  return coo_to_csc(ctx, (const int32_t*)d_rows->p, (const int32_t*)d_cols->p, (const float*)d_vals->p, n, n_rows, n_cols, false, out);
  CSC_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------
// merge_matrices (int_utils.jl:1-23): matrices side by side on the union of their rows
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
k_merge_fill(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const float* __restrict__ val,
             const int32_t* __restrict__ row_map, int64_t n_cells, int64_t col_offset, int64_t out_offset,
             int32_t* __restrict__ rows, int32_t* __restrict__ cols, float* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t base = colptr[0];
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    const int64_t b = colptr[c], e = colptr[c + 1];
    for (int64_t i = b + lane; i < e; i += 32) {
      const int64_t o = out_offset + (i - base);
      rows[o] = row_map[rowval[i]];
      cols[o] = (int32_t)(col_offset + c);
      vals[o] = val[i];
    }
  }
}
}  // namespace

extern "C" csc_status csc_sparse_merge(csc_ctx* ctx, const csc_sparse* const* mats, const int64_t* const* row_maps, int32_t n_mats,
                                       int64_t n_rows_merged, csc_sparse** out) {
  if (!ctx || !mats || !row_maps || !out || n_mats < 1) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_UPLOAD);
  CSC_REQUIRE(ctx, n_rows_merged > 0 && n_rows_merged < ((int64_t)1 << 31), "csc_sparse_merge: bad merged row count");
  int64_t n = 0, n_cols = 0;
  for (int i = 0; i < n_mats; ++i) {
    CSC_REQUIRE(ctx, mats[i] && row_maps[i], "csc_sparse_merge: matrix %d is NULL", i);
    for (int64_t r = 0; r < mats[i]->n_genes; ++r)
      CSC_REQUIRE(ctx, row_maps[i][r] >= 0 && row_maps[i][r] < n_rows_merged, "csc_sparse_merge: row map of matrix %d out of range", i);
    n += mats[i]->nnz;
    n_cols += mats[i]->n_cells;
  }
  CSC_REQUIRE(ctx, n_cols < ((int64_t)1 << 31), "csc_sparse_merge: more than 2^31 columns");
  DevPtr d_rows, d_cols, d_vals;
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4, &d_rows));
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4, &d_cols));
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4, &d_vals));
  int64_t o = 0, co = 0;
  for (int i = 0; i < n_mats; ++i) {
    const csc_sparse* m = mats[i];
    std::vector<int32_t> map32((size_t)m->n_genes);
    for (int64_t r = 0; r < m->n_genes; ++r) map32[r] = (int32_t)row_maps[i][r];
    DevPtr d_map;
    CSC_TRY(dev_alloc(ctx, (size_t)m->n_genes * 4, &d_map));
    CSC_CUDA(ctx, cudaMemcpyAsync(d_map->p, map32.data(), map32.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (m->nnz > 0) {
      k_merge_fill<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(m->cp(), m->rv(), m->v(), (const int32_t*)d_map->p, m->n_cells, co, o,
                                                               (int32_t*)d_rows->p, (int32_t*)d_cols->p, (float*)d_vals->p);
      CSC_LAUNCHED(ctx);
    }
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // map32 leaves scope
    o += m->nnz;
    co += m->n_cells;
  }
  // the reference pushes only `val != 0` (int_utils.jl:11): stored zeros of the inputs are dropped
  return coo_to_csc(ctx, (const int32_t*)d_rows->p, (const int32_t*)d_cols->p, (const float*)d_vals->p, n, n_rows_merged, n_cols, true, out);
  CSC_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------
// colSum (utils.jl:1-2) and the column half of subset_matrix / subset_count (utils.jl:59-87, 313-327)
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
k_col_sums(const int64_t* __restrict__ colptr, const float* __restrict__ val, int64_t n_cells, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    double s = 0.0;
    for (int64_t i = colptr[c] + lane; i < colptr[c + 1]; i += 32) s += (double)val[i];
    s = warp_sum(s);
    if (lane == 0) out[c] = s;
  }
}
__global__ void k_sel_counts(const int64_t* __restrict__ colptr, const int64_t* __restrict__ cols, int64_t n_sel,
                             int64_t* __restrict__ counts) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_sel) counts[j] = colptr[cols[j] + 1] - colptr[cols[j]];
}
__global__ void __launch_bounds__(256)
k_sel_copy(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const float* __restrict__ val,
           const int64_t* __restrict__ cols, const int64_t* __restrict__ out_colptr, int64_t n_sel, int32_t* __restrict__ out_rv,
           float* __restrict__ out_v) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = warp; j < n_sel; j += nwarp) {
    const int64_t b = colptr[cols[j]], e = colptr[cols[j] + 1], o = out_colptr[j];
    for (int64_t i = b + lane; i < e; i += 32) {
      out_rv[o + (i - b)] = rowval[i];
      out_v[o + (i - b)] = val[i];
    }
  }
}
}  // namespace

extern "C" csc_status csc_sparse_col_sums(const csc_sparse* m, double* out) {
  if (!m || !out) return CSC_ERR_ARG;
  csc_ctx* ctx = m->ctx;
  CSC_GUARD_BEGIN
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (m->n_cells == 0) return CSC_OK;
  DevPtr d;
  CSC_TRY(dev_alloc(ctx, (size_t)m->n_cells * 8, &d));
  k_col_sums<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(m->cp(), m->v(), m->n_cells, (double*)d->p);
  CSC_LAUNCHED(ctx);
  return d2h_copy(ctx, out, d->p, (size_t)m->n_cells * 8);
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_sparse_select_cols(csc_ctx* ctx, const csc_sparse* m, const int64_t* cols, int64_t n_sel, csc_sparse** out) {
  if (!ctx || !m || !out || (n_sel > 0 && !cols)) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int64_t j = 0; j < n_sel; ++j)
    CSC_REQUIRE(ctx, cols[j] >= 0 && cols[j] < m->n_cells, "csc_sparse_select_cols: column %lld outside [0,%lld)", (long long)cols[j],
                (long long)m->n_cells);
  std::unique_ptr<csc_sparse> r(new csc_sparse());
  r->ctx = ctx;
  r->n_genes = m->n_genes;
  r->n_cells = n_sel;
  r->nnz = 0;
  r->vmax_bound = m->vmax_bound;
  DevPtr d_cols, counts;
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(n_sel, 1) * 8, &d_cols));
  CSC_TRY(dev_alloc(ctx, (size_t)(n_sel + 1) * 8, &counts, true));
  CSC_TRY(dev_alloc(ctx, (size_t)(n_sel + 1) * 8, &r->colptr, true));
  if (n_sel > 0) {
    CSC_CUDA(ctx, cudaMemcpyAsync(d_cols->p, cols, (size_t)n_sel * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_sel_counts<<<nblk(n_sel), 256, 0, ctx->stream>>>(m->cp(), (const int64_t*)d_cols->p, n_sel, (int64_t*)counts->p);
    CSC_LAUNCHED(ctx);
    size_t tb = 0;
    CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int64_t*)counts->p, (int64_t*)r->colptr->p, n_sel + 1, ctx->stream));
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, tb, &tmp));
    CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp->p, tb, (const int64_t*)counts->p, (int64_t*)r->colptr->p, n_sel + 1, ctx->stream));
    ctx->launches += 2;
    CSC_CUDA(ctx, cudaMemcpyAsync(&r->nnz, (const int64_t*)r->colptr->p + n_sel, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  CSC_TRY(dev_alloc(ctx, (size_t)r->nnz * 4, &r->rowval));
  CSC_TRY(dev_alloc(ctx, (size_t)r->nnz * 4, &r->val));
  if (r->nnz > 0) {
    k_sel_copy<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(m->cp(), m->rv(), m->v(), (const int64_t*)d_cols->p, r->cp(), n_sel,
                                                          (int32_t*)r->rowval->p, r->v());
    CSC_LAUNCHED(ctx);
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  *out = r.release();
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------
// run_tf_idf (atac_processing.jl:1-13): log1p( idf_g * (x / colsum_c) * scale_factor ), idf_g = n_cells / rowsum_g
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
k_tfidf(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const float* __restrict__ raw,
        const double* __restrict__ row_sum, int64_t n_cells, double n_cells_total, double sf, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    const int64_t b = colptr[c], e = colptr[c + 1];
    double s = 0.0;
    for (int64_t i = b + lane; i < e; i += 32) s += (double)raw[i];
    s = warp_sum(s);
    const double inv = 1.0 / s;                                  // 1 ./ npeaks (:4)
    for (int64_t i = b + lane; i < e; i += 32) {
      const double tf = (double)raw[i] * inv;                    // mtx * Diagonal(1 ./ npeaks) (:4)
      const double idf = n_cells_total / row_sum[rowval[i]];     // size(mtx)[2] ./ rsums (:8)
      out[i] = (float)log1p((idf * tf) * sf);                    // log1p.(norm_data .* scale_factor) (:11-12)
    }
  }
}
}  // namespace

extern "C" csc_status csc_tf_idf(csc_ctx* ctx, const csc_sparse* raw, const csc_vec* gene_stats, int64_t n_cells_total,
                                 double scale_factor, csc_sparse** out) {
  if (!ctx || !raw || !gene_stats || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, gene_stats->dtype == CSC_DT_F64 && gene_stats->n >= raw->n_genes, "csc_tf_idf: gene_stats must be fp64[>= n_genes]");
  CSC_REQUIRE(ctx, n_cells_total >= 1, "csc_tf_idf: n_cells_total must be >= 1");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_NORMALIZE);
  std::unique_ptr<csc_sparse> m(new csc_sparse());
  m->ctx = ctx;
  m->n_genes = raw->n_genes;
  m->n_cells = raw->n_cells;
  m->nnz = raw->nnz;
  m->colptr = raw->colptr;
  m->rowval = raw->rowval;
  CSC_TRY(dev_alloc(ctx, (size_t)raw->nnz * 4, &m->val));
  if (raw->n_cells > 0) {
    k_tfidf<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(raw->cp(), raw->rv(), raw->v(), gene_stats->as<double>(), raw->n_cells,
                                                        (double)n_cells_total, scale_factor, m->v());
    CSC_LAUNCHED(ctx);
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  *out = m.release();
  return CSC_OK;
  CSC_GUARD_END(ctx)
}
