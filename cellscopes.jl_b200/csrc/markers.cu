// find_markers / find_all_markers (src/scrna/processing.jl:350-414): per-gene Mann-Whitney U test between cell groups
// on the NORMALISED matrix, plus the group means (log(mean(expm1 x) + 1)) and the expressing fractions.
//
// The reference densifies both clusters (Matrix(cl1_ct), :371-375) and runs HypothesisTests.MannWhitneyUTest row by
// row; find_all_markers repeats that for every cluster against all other cells (:393-414).  Here the sufficient
// statistics of EVERY group come out of one sort of the stored entries:
//   * the cells that take part carry a group label (find_markers: cluster_1 -> 0, cluster_2 -> 1, everything else is
//     left out; find_all_markers: label = cluster, every cell takes part, and "cluster c against the rest" needs only
//     the rank sum of c because the ranks are those of all cells together);
//   * every stored entry of such a cell becomes one 64-bit key [gene:16 | order-preserving image of the value:32 |
//     label:16]; a radix sort over the upper 48 bits groups the genes and orders the values, the label rides along;
//   * a pass over the sorted keys finds each entry's tie group (neighbour compare; binary search only inside real ties)
//     and adds the average rank of the entry to its (gene, label) cell of the table.  The implicit zeros of a gene form
//     one more tie group whose rank follows from counts alone; the host adds it (n_label - nnz[gene][label] zeros each).
// Output tables (fp64, [gene][label]): rank sum of the stored entries, sum expm1(x), #(x > expr_cutoff), #stored; per
// gene: sum (t^3 - t) over the tie groups of non-zero values, #negative values, #stored zeros.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t ord32(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord32(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
constexpr uint32_t ORD_NEG_ZERO = 0x7FFFFFFFu;   // ord32(-0.0f)
constexpr uint32_t ORD_POS_ZERO = 0x80000000u;   // ord32(+0.0f)

__global__ void k_mk_count(const int64_t* __restrict__ colptr, const int32_t* __restrict__ labels, int64_t n_cells,
                           int64_t* __restrict__ counts) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  counts[c] = labels[c] >= 0 ? colptr[c + 1] - colptr[c] : 0;
}

__global__ void __launch_bounds__(256)
k_mk_fill(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const float* __restrict__ val,
          const int32_t* __restrict__ labels, const int64_t* __restrict__ offs, int64_t n_cells,
          unsigned long long* __restrict__ keys) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    const int lab = labels[c];
    if (lab < 0) continue;
    const int64_t b = colptr[c], e = colptr[c + 1], o = offs[c];
    for (int64_t i = b + lane; i < e; i += 32) {
      const unsigned long long g = (unsigned long long)__ldcs(rowval + i);
      keys[o + (i - b)] = (g << 48) | ((unsigned long long)ord32(__ldcs(val + i)) << 16) | (unsigned long long)lab;
    }
  }
}

// first index in [lo, hi) whose (key >> 16) is >= target
__device__ __forceinline__ int64_t lower_bound48(const unsigned long long* __restrict__ keys, int64_t lo, int64_t hi,
                                                 unsigned long long target) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((keys[mid] >> 16) < target) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// per gene: segment start, #negative values, #stored zeros
__global__ void k_mk_bounds(const unsigned long long* __restrict__ keys, int64_t n, int32_t n_genes, int64_t* __restrict__ gene_ptr,
                            double* __restrict__ n_neg, double* __restrict__ n_zero_stored) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_genes) return;
  const int64_t s = lower_bound48(keys, 0, n, (unsigned long long)g << 32);
  gene_ptr[g] = s;
  if (g < n_genes) {
    const int64_t e = lower_bound48(keys, s, n, (unsigned long long)(g + 1) << 32);
    const int64_t a = lower_bound48(keys, s, e, ((unsigned long long)g << 32) | ORD_NEG_ZERO);
    const int64_t b = lower_bound48(keys, a, e, (((unsigned long long)g << 32) | ORD_POS_ZERO) + 1ull);
    n_neg[g] = (double)(a - s);
    n_zero_stored[g] = (double)(b - a);
  }
}

struct MkTables {
  double* rank_sum;   // [G][ng]
  double* sum_expm1;  // [G][ng]
  double* n_expr;     // [G][ng]
  double* nnz;        // [G][ng]
  double* sum_x;      // [G][ng]
  double* tie_adj;    // [G]
};

// adds v0..v3 of every lane into table cell `cell` (lanes with the same cell are combined first: the keys are sorted
// by gene, so a warp meets at most a handful of distinct cells and unaggregated atomics would all hit the same lines)
__device__ __forceinline__ void warp_table_add(unsigned active, bool valid, long long cell, const MkTables& t, double r, double x,
                                               double p, double z, double y) {
  const unsigned peers = __match_any_sync(active, valid ? cell : -1ll);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  double sr = 0.0, sx = 0.0, sp = 0.0, sz = 0.0, sy = 0.0;
  for (int src = 0; src < 32; ++src) {
    if (!((active >> src) & 1u)) continue;
    const double a = __shfl_sync(active, r, src), b = __shfl_sync(active, x, src), c = __shfl_sync(active, p, src),
                 d = __shfl_sync(active, z, src), f = __shfl_sync(active, y, src);
    if ((peers >> src) & 1u) {
      sr += a;
      sx += b;
      sp += c;
      sz += d;
      sy += f;
    }
  }
  if (valid && lane == leader) {
    atomicAdd(t.rank_sum + cell, sr);
    atomicAdd(t.sum_expm1 + cell, sx);
    if (sp != 0.0) atomicAdd(t.n_expr + cell, sp);
    atomicAdd(t.nnz + cell, sz);
    atomicAdd(t.sum_x + cell, sy);
  }
}

__global__ void __launch_bounds__(256)
k_mk_rank(const unsigned long long* __restrict__ keys, int64_t n, const int64_t* __restrict__ gene_ptr,
          const double* __restrict__ n_zero_stored, const double* __restrict__ n_neg, int32_t n_groups, double n_incl,
          float expr_cutoff, MkTables t) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = (n + 31) & ~(int64_t)31;        // whole warps stay together for the shuffles
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool valid = i < n;
    const unsigned active = __activemask();
    long long cell = 0;
    double rank = 0.0, ex = 0.0, pos = 0.0, xv = 0.0;
    if (valid) {
      const unsigned long long key = keys[i];
      const unsigned long long k48 = key >> 16;
      const int g = (int)(key >> 48);
      const int lab = (int)(key & 0xFFFFull);
      const uint32_t o = (uint32_t)(k48 & 0xFFFFFFFFull);
      const float v = unord32(o);
      const int64_t s = gene_ptr[g], e = gene_ptr[g + 1];
      const double nnz_g = (double)(e - s);
      const double n_zero_all = n_incl - nnz_g + n_zero_stored[g];      // the zero tie group: implicit + stored zeros
      if (o == ORD_NEG_ZERO || o == ORD_POS_ZERO) {
        rank = n_neg[g] + 0.5 * (n_zero_all + 1.0);
      } else {
        int64_t lo = i, hi = i + 1;
        if (i > s && (keys[i - 1] >> 16) == k48) lo = lower_bound48(keys, s, i, k48);
        if (hi < e && (keys[hi] >> 16) == k48) hi = lower_bound48(keys, hi, e, k48 + 1ull);
        const double tt = (double)(hi - lo);
        // entries below: (lo - s) stored ones, of which the stored zeros are already part of n_zero_all for a positive
        // value; a negative value has only stored negatives below it
        const double below = v > 0.f ? (double)(lo - s) - n_zero_stored[g] + n_zero_all : (double)(lo - s);
        rank = below + 0.5 * (tt + 1.0);
        if (lo == i && tt > 1.0) atomicAdd(t.tie_adj + g, tt * tt * tt - tt);
      }
      ex = expm1((double)v);
      xv = (double)v;
      pos = v > expr_cutoff ? 1.0 : 0.0;
      cell = (long long)g * n_groups + lab;
    }
    warp_table_add(active, valid, cell, t, rank, ex, pos, valid ? 1.0 : 0.0, xv);
  }
}

}  // namespace

extern "C" csc_status csc_marker_stats(csc_ctx* ctx, const csc_sparse* norm, const int32_t* labels, int32_t n_groups,
                                       double expr_cutoff, double* rank_sum, double* sum_expm1, double* n_expr, double* nnz_out,
                                       double* sum_x, double* tie_adj, double* n_neg, double* n_zero_stored, int64_t* n_included) {
  if (!ctx || !norm || !labels || !rank_sum || !sum_expm1 || !n_expr || !nnz_out || !sum_x || !tie_adj || !n_neg || !n_zero_stored)
    return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  const int64_t G = norm->n_genes, N = norm->n_cells;
  CSC_REQUIRE(ctx, n_groups >= 1 && n_groups <= 65535, "csc_marker_stats: n_groups=%d outside [1, 65535]", n_groups);
  CSC_REQUIRE(ctx, G <= 65535, "csc_marker_stats: more than 65535 genes is not supported (the sort key holds 16 gene bits)");
  int64_t n_incl = 0;
  for (int64_t c = 0; c < N; ++c) {
    CSC_REQUIRE(ctx, labels[c] < n_groups, "csc_marker_stats: label %d of cell %lld >= n_groups=%d", labels[c], (long long)c, n_groups);
    n_incl += labels[c] >= 0;
  }
  if (n_included) *n_included = n_incl;
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_MARKERS);
  const size_t tab = (size_t)G * n_groups;
  DevPtr d_lab, d_cnt, d_off, d_tab, d_gene, d_ptr;
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(N, 1) * 4, &d_lab));
  CSC_TRY(dev_alloc(ctx, (size_t)(N + 1) * 8, &d_cnt, true));
  CSC_TRY(dev_alloc(ctx, (size_t)(N + 1) * 8, &d_off));
  CSC_TRY(dev_alloc(ctx, (5 * tab + 3 * (size_t)G) * 8, &d_tab, true));
  CSC_TRY(dev_alloc(ctx, (size_t)(G + 1) * 8, &d_ptr));
  CSC_CUDA(ctx, cudaMemcpyAsync(d_lab->p, labels, (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
  double* T = (double*)d_tab->p;
  MkTables t{T, T + tab, T + 2 * tab, T + 3 * tab, T + 4 * tab, T + 5 * tab};
  double* d_neg = T + 5 * tab + G;
  double* d_zs = T + 5 * tab + 2 * G;
  int64_t n = 0;
  if (N > 0) {
    k_mk_count<<<(unsigned)ceil_div64(N, 256), 256, 0, ctx->stream>>>(norm->cp(), (const int32_t*)d_lab->p, N, (int64_t*)d_cnt->p);
    CSC_LAUNCHED(ctx);
    size_t tb = 0;
    CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int64_t*)d_cnt->p, (int64_t*)d_off->p, N + 1, ctx->stream));
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, tb, &tmp));
    CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp->p, tb, (const int64_t*)d_cnt->p, (int64_t*)d_off->p, N + 1, ctx->stream));
    ctx->launches += 2;
    CSC_CUDA(ctx, cudaMemcpyAsync(&n, (const int64_t*)d_off->p + N, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  DevPtr keys_a, keys_b;
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 8, &keys_a));
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 8, &keys_b));
  const unsigned long long* skeys = (const unsigned long long*)keys_b->p;
  if (n > 0) {
    k_mk_fill<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(norm->cp(), norm->rv(), norm->v(), (const int32_t*)d_lab->p,
                                                         (const int64_t*)d_off->p, N, (unsigned long long*)keys_a->p);
    CSC_LAUNCHED(ctx);
    int gene_bits = 1;
    while (((int64_t)1 << gene_bits) < G) ++gene_bits;
    size_t tb = 0;
    CSC_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tb, (const unsigned long long*)keys_a->p, (unsigned long long*)keys_b->p, n,
                                                 16, 48 + gene_bits, ctx->stream));
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, tb, &tmp));
    CSC_CUDA(ctx, cub::DeviceRadixSort::SortKeys(tmp->p, tb, (const unsigned long long*)keys_a->p, (unsigned long long*)keys_b->p, n,
                                                 16, 48 + gene_bits, ctx->stream));
    ctx->launches += 1 + (32 + gene_bits + 7) / 8;
  }
  k_mk_bounds<<<(unsigned)ceil_div64(G + 1, 256), 256, 0, ctx->stream>>>(skeys, n, (int32_t)G, (int64_t*)d_ptr->p, d_neg, d_zs);
  CSC_LAUNCHED(ctx);
  if (n > 0) {
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(n, 256), (int64_t)ctx->sm_count * 16);
    k_mk_rank<<<grid, 256, 0, ctx->stream>>>(skeys, n, (const int64_t*)d_ptr->p, d_zs, d_neg, n_groups, (double)n_incl,
                                             (float)expr_cutoff, t);
    CSC_LAUNCHED(ctx);
  }
  CSC_TRY(d2h_copy(ctx, rank_sum, t.rank_sum, tab * 8));
  CSC_TRY(d2h_copy(ctx, sum_expm1, t.sum_expm1, tab * 8));
  CSC_TRY(d2h_copy(ctx, n_expr, t.n_expr, tab * 8));
  CSC_TRY(d2h_copy(ctx, nnz_out, t.nnz, tab * 8));
  CSC_TRY(d2h_copy(ctx, sum_x, t.sum_x, tab * 8));
  CSC_TRY(d2h_copy(ctx, tie_adj, t.tie_adj, (size_t)G * 8));
  CSC_TRY(d2h_copy(ctx, n_neg, d_neg, (size_t)G * 8));
  CSC_TRY(d2h_copy(ctx, n_zero_stored, d_zs, (size_t)G * 8));
  return CSC_OK;
  CSC_GUARD_END(ctx)
}
