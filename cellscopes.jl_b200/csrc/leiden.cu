// Community detection on the device for the graph run_clustering builds (src/scrna/processing.jl:233-234, 279-280:
// `Leiden.leiden(adj_mat, resolution = res)`, 590 s on the 400k-cell tutorial): the Leiden scheme of Traag, Waltman &
// van Eck (2019) -- local moving, refinement inside the communities, aggregation on the REFINED partition with the
// unrefined one as the initial assignment -- for the Constant Potts Model quality that Leiden.jl optimises,
//     H = sum_c [ e_c - gamma n_c (n_c - 1) / 2 ],     e_c = weight inside c, n_c = cells in c, gamma = `res`.
// A parallel, deterministic variant (the reference's is sequential and seeded from Julia's RNG, so labels can never be
// compared one to one; tests compare the quality and the guarantees):
//   * local moving: every sweep scores ALL (node, neighbouring community) pairs at once -- one key per edge entry
//     [node | community of the neighbour], radix sort, reduce-by-key -> k(v, c) --, each node of the sweep's half
//     (nodes alternate by parity of a hash, so that two neighbours never swap places in the same sweep) takes the
//     community with the largest gain k(v,c) - gamma s_v n_c, or its own empty home slot when every gain is negative;
//   * refinement: inside every community the nodes start as singletons; a singleton that is well connected to its
//     community (k(v, c - v) >= gamma s_v (n_c - s_v)) merges into the neighbouring refined set of the same community
//     with the largest non-negative gain, provided that set is well connected too -- greedy where the paper draws at
//     random; merges only follow edges, so refined sets (and after aggregation the communities) are connected;
//   * aggregation: refined sets become nodes (sizes and edge weights summed by sort + reduce-by-key), the community of a
//     new node is the community its members had.
// Everything is sort / scan / reduce plus thin kernels; a sweep over the 2 x 30M edge entries of a 1M-cell graph costs
// a few tens of milliseconds.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

typedef unsigned long long u64;

struct Level {
  int64_t n = 0, m = 0;
  DevPtr rowptr, src, col, w, size, comm, csize;      // int64[n+1], int32[m], int32[m], double[m], double[n], int32[n], double[n]
};

inline unsigned lb(int64_t n) { return (unsigned)std::max<int64_t>(1, ceil_div64(n, 256)); }

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void k_ld_fill_src(const int64_t* __restrict__ rowptr, int64_t n, int32_t* __restrict__ src) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = warp; v < n; v += nwarp)
    for (int64_t e = rowptr[v] + lane; e < rowptr[v + 1]; e += 32) src[e] = (int32_t)v;
}
__global__ void k_ld_iota_ones(int32_t* __restrict__ a, double* __restrict__ ones, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    a[i] = (int32_t)i;
    if (ones) ones[i] = 1.0;
  }
}
// key = [node | label of the neighbour]; an entry that must not count (self loop, neighbour outside the node's
// community when `same_comm_only`) gets label 0xFFFFFFFF and weight 0
__global__ void k_ld_keys(const int32_t* __restrict__ src, const int32_t* __restrict__ col, const double* __restrict__ w,
                          const int32_t* __restrict__ label, const int32_t* __restrict__ comm, int same_comm_only, int64_t m,
                          u64* __restrict__ keys, double* __restrict__ vals) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int v = src[e], u = col[e];
  const bool ok = u != v && (!same_comm_only || comm[u] == comm[v]);
  keys[e] = ((u64)(uint32_t)v << 32) | (ok ? (uint32_t)label[u] : 0xFFFFFFFFu);
  vals[e] = ok ? w[e] : 0.0;
}
__device__ __forceinline__ int64_t lower_bound_u64(const u64* __restrict__ a, int64_t n, u64 t) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < t) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// local moving decision of one sweep: thread per node over its (node, community) segment of the reduced keys
__global__ void k_ld_decide(const u64* __restrict__ ukeys, const double* __restrict__ uvals, int64_t n_unique, int64_t n,
                            const int32_t* __restrict__ comm, const double* __restrict__ size, const double* __restrict__ csize,
                            double gamma, int sweep, int32_t* __restrict__ newc) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int own = comm[v];
  newc[v] = own;
  if (((mix32((uint32_t)v) + (uint32_t)sweep) & 1u) != 0u) return;       // the other half moves in the next sweep
  const double s = size[v];
  const int64_t b = lower_bound_u64(ukeys, n_unique, (u64)v << 32), e = lower_bound_u64(ukeys, n_unique, (u64)(v + 1) << 32);
  double own_gain = -gamma * s * (csize[own] - s);
  double best_gain = -INFINITY;
  int best = own;
  for (int64_t i = b; i < e; ++i) {
    const uint32_t c = (uint32_t)(ukeys[i] & 0xFFFFFFFFull);
    if (c == 0xFFFFFFFFu) continue;
    const double k = uvals[i];
    if ((int)c == own) {
      own_gain = k - gamma * s * (csize[own] - s);
    } else {
      const double g = k - gamma * s * csize[c];
      if (g > best_gain || (g == best_gain && (int)c < best)) {
        best_gain = g;
        best = (int)c;
      }
    }
  }
  // three options: stay (own_gain), join the best neighbouring community (best_gain), or be alone in the node's own
  // home slot when that is empty (gain 0)
  const bool alone = csize[own] <= s;
  const bool can_leave_alone = !alone && csize[v] == 0.0;
  int choice = own;
  double choice_gain = own_gain;
  if (best != own && best_gain > choice_gain + 1e-12) {
    choice = best;
    choice_gain = best_gain;
  }
  if (can_leave_alone && 0.0 > choice_gain + 1e-12) choice = (int32_t)v;
  if (choice == best && best != own && alone) {
    // two lone nodes of the same half that prefer each other would trade places for ever: a lone node only moves
    // "upwards" in the total order (community size, then lower community id), which admits no cycle
    if (!(csize[best] > s || (csize[best] == s && best < own))) choice = own;
  }
  newc[v] = choice;
}
__global__ void k_ld_apply(const int32_t* __restrict__ newc, int32_t* __restrict__ comm, const double* __restrict__ size,
                           double* __restrict__ csize, int64_t n, int* __restrict__ moved) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int a = comm[v], b = newc[v];
  if (a != b) {
    comm[v] = b;
    atomicAdd(csize + b, size[v]);
    atomicAdd(csize + a, -size[v]);
    atomicAdd(moved, 1);
  }
}

// refinement helpers ------------------------------------------------------------------------------------------
// ext[ref[v]] += w for every edge entry inside the community that leaves the refined set
__global__ void k_ld_ext(const int32_t* __restrict__ src, const int32_t* __restrict__ col, const double* __restrict__ w,
                         const int32_t* __restrict__ comm, const int32_t* __restrict__ ref, int64_t m, double* __restrict__ ext) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  const int v = src[e], u = col[e];
  if (u != v && comm[u] == comm[v] && ref[u] != ref[v]) atomicAdd(ext + ref[v], w[e]);
}
// proposal of one refinement round: thread per node over its (node, refined set) segment (same-community entries only)
__global__ void k_ld_propose(const u64* __restrict__ ukeys, const double* __restrict__ uvals, int64_t n_unique, int64_t n,
                             const int32_t* __restrict__ comm, const int32_t* __restrict__ ref, const double* __restrict__ size,
                             const double* __restrict__ rsize, const double* __restrict__ csize, const double* __restrict__ ext,
                             const int32_t* __restrict__ members, double gamma, int32_t* __restrict__ prop) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  prop[v] = -1;
  if (ref[v] != v || members[v] != 1) return;                              // only singletons merge
  const double s = size[v], nc = csize[comm[v]];
  if (ext[v] < gamma * s * (nc - s)) return;                               // v is not well connected to its community
  const int64_t b = lower_bound_u64(ukeys, n_unique, (u64)v << 32), e = lower_bound_u64(ukeys, n_unique, (u64)(v + 1) << 32);
  double best_gain = -INFINITY;
  int best = -1;
  for (int64_t i = b; i < e; ++i) {
    const uint32_t r = (uint32_t)(ukeys[i] & 0xFFFFFFFFull);
    if (r == 0xFFFFFFFFu || (int)r == (int)v) continue;
    if (ext[r] < gamma * rsize[r] * (nc - rsize[r])) continue;             // the target set is not well connected
    const double g = uvals[i] - gamma * s * rsize[r];
    if (g >= 0.0 && (g > best_gain || (g == best_gain && (int)r < best))) {
      best_gain = g;
      best = (int)r;
    }
  }
  prop[v] = best;
}
__global__ void k_ld_merge(const int32_t* __restrict__ prop, int32_t* __restrict__ ref, const double* __restrict__ size,
                           double* __restrict__ rsize, const int32_t* __restrict__ members_before, int32_t* __restrict__ members,
                           int64_t n, int* __restrict__ merged) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int r = prop[v];
  if (r < 0) return;
  // a singleton target must stay where it is in this round: it has no proposal of its own, or the two singletons chose
  // each other (the usual case), and then the one with the larger id moves
  // (decided on the member counts as they were when the round started: the result does not depend on the order in
  // which the threads of this kernel run)
  const bool target_single = members_before[r] == 1;
  if (target_single && (r > (int)v || (prop[r] >= 0 && prop[r] != (int)v))) return;
  ref[v] = r;
  atomicAdd(rsize + r, size[v]);
  atomicAdd(rsize + v, -size[v]);
  atomicAdd(members + r, 1);
  atomicSub(members + v, 1);
  atomicAdd(merged, 1);
}

// aggregation helpers -----------------------------------------------------------------------------------------
__global__ void k_ld_mark(const int32_t* __restrict__ ref, int64_t n, int64_t* __restrict__ used) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) used[ref[v]] = 1;
}
__global__ void k_ld_new_nodes(const int32_t* __restrict__ ref, const int64_t* __restrict__ newid, const int32_t* __restrict__ comm,
                               const double* __restrict__ size, int64_t n, double* __restrict__ nsize, int32_t* __restrict__ cmin) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int j = (int)newid[ref[v]];
  atomicAdd(nsize + j, size[v]);
  atomicMin(cmin + comm[v], j);                                           // new id of a community: its smallest new node
}
__global__ void k_ld_new_comm(const int32_t* __restrict__ ref, const int64_t* __restrict__ newid, const int32_t* __restrict__ comm,
                              const int32_t* __restrict__ cmin, int64_t n, int32_t* __restrict__ ncomm) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) ncomm[newid[ref[v]]] = cmin[comm[v]];
}
__global__ void k_ld_edge_keys(const int32_t* __restrict__ src, const int32_t* __restrict__ col, const int32_t* __restrict__ ref,
                               const int64_t* __restrict__ newid, int64_t m, u64* __restrict__ keys) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m) keys[e] = ((u64)newid[ref[src[e]]] << 32) | (u64)newid[ref[col[e]]];
}
__global__ void k_ld_split_keys(const u64* __restrict__ keys, int64_t m, int32_t* __restrict__ src, int32_t* __restrict__ col) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m) {
    src[e] = (int32_t)(keys[e] >> 32);
    col[e] = (int32_t)(keys[e] & 0xFFFFFFFFull);
  }
}
__global__ void k_ld_rowptr(const int32_t* __restrict__ src, int64_t m, int64_t n, int64_t* __restrict__ rowptr) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v > n) return;
  int64_t lo = 0, hi = m;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (src[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  rowptr[v] = lo;
}
__global__ void k_ld_csize(const int32_t* __restrict__ comm, const double* __restrict__ size, int64_t n, double* __restrict__ csize) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) atomicAdd(csize + comm[v], size[v]);
}
__global__ void k_ld_map_nodes(int32_t* __restrict__ node_of, const int32_t* __restrict__ ref, const int64_t* __restrict__ newid,
                               int64_t n0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n0) node_of[i] = (int32_t)newid[ref[node_of[i]]];
}
__global__ void k_ld_labels(const int32_t* __restrict__ node_of, const int32_t* __restrict__ comm, int64_t n0, int32_t* __restrict__ lab) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n0) lab[i] = comm[node_of[i]];
}
__global__ void k_ld_quality_edges(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ w,
                                   const int32_t* __restrict__ lab, int64_t n, double* __restrict__ e_in) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double s = 0.0;
  for (int64_t v = warp; v < n; v += nwarp)
    for (int64_t e = rowptr[v] + lane; e < rowptr[v + 1]; e += 32) {
      const int u = col[e];
      if (lab[u] == lab[v]) s += (u == (int)v) ? w[e] : 0.5 * w[e];
    }
  s = warp_sum(s);
  if (lane == 0 && s != 0.0) atomicAdd(e_in, s);
}

// (keys, vals) -> sorted unique keys with summed values; returns the number of unique keys
csc_status sort_reduce(csc_ctx* ctx, DevPtr& keys, DevPtr& vals, int64_t m, int key_bits, DevPtr& ukeys, DevPtr& uvals, int64_t* n_unique) {
  DevPtr k2, v2, cnt, tmp;
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(m, 1) * 8, &k2));
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(m, 1) * 8, &v2));
  CSC_TRY(dev_alloc(ctx, 8, &cnt, true));
  size_t tb = 0;
  CSC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tb, (const u64*)keys->p, (u64*)k2->p, (const double*)vals->p, (double*)v2->p, m, 0,
                                                key_bits, ctx->stream));
  CSC_TRY(dev_alloc(ctx, tb, &tmp));
  CSC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp->p, tb, (const u64*)keys->p, (u64*)k2->p, (const double*)vals->p, (double*)v2->p, m, 0,
                                                key_bits, ctx->stream));
  tb = 0;
  CSC_CUDA(ctx, cub::DeviceReduce::ReduceByKey(nullptr, tb, (const u64*)k2->p, (u64*)keys->p, (const double*)v2->p, (double*)vals->p,
                                               (int64_t*)cnt->p, cub::Sum(), m, ctx->stream));
  DevPtr tmp2;
  CSC_TRY(dev_alloc(ctx, tb, &tmp2));
  CSC_CUDA(ctx, cub::DeviceReduce::ReduceByKey(tmp2->p, tb, (const u64*)k2->p, (u64*)keys->p, (const double*)v2->p, (double*)vals->p,
                                               (int64_t*)cnt->p, cub::Sum(), m, ctx->stream));
  ctx->launches += 12;
  CSC_CUDA(ctx, cudaMemcpyAsync(n_unique, cnt->p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ukeys = keys;
  uvals = vals;
  return CSC_OK;
}

int bits_for(int64_t n) {
  int b = 1;
  while (((int64_t)1 << b) < n) ++b;
  return b;
}

}  // namespace

extern "C" csc_status csc_leiden(csc_ctx* ctx, const csc_graph* g, double resolution, int32_t max_levels, int32_t max_sweeps,
                                 int32_t* labels_out, int32_t* n_communities, double* quality, int32_t* levels_out) {
  if (!ctx || !g || !labels_out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, g->n_rows == g->n_cols && g->col_lo == 0, "csc_leiden: the whole graph must be on this device (n x n, all columns)");
  CSC_REQUIRE(ctx, g->n_rows >= 1 && g->n_rows < ((int64_t)1 << 31), "csc_leiden: bad node count");
  CSC_REQUIRE(ctx, resolution >= 0.0, "csc_leiden: resolution must be >= 0");
  if (max_levels <= 0) max_levels = 20;
  if (max_sweeps <= 0) max_sweeps = 40;
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_LEIDEN);
  cudaStream_t s = ctx->stream;
  const int64_t n0 = g->n_rows;
  const double gamma = resolution;
  // level 0: the adjacency as it is (symmetric: column c of the CSC is the neighbour list of node c)
  Level L;
  L.n = n0;
  L.m = g->nnz;
  L.rowptr = g->colptr;
  L.col = g->rowval;
  L.w = g->val;
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(L.m, 1) * 4, &L.src));
  CSC_TRY(dev_alloc(ctx, (size_t)n0 * 8, &L.size));
  CSC_TRY(dev_alloc(ctx, (size_t)n0 * 4, &L.comm));
  CSC_TRY(dev_alloc(ctx, (size_t)n0 * 8, &L.csize));
  if (L.m > 0) {
    k_ld_fill_src<<<ctx->sm_count * 8, 256, 0, s>>>((const int64_t*)L.rowptr->p, n0, (int32_t*)L.src->p);
    CSC_LAUNCHED(ctx);
  }
  k_ld_iota_ones<<<lb(n0), 256, 0, s>>>((int32_t*)L.comm->p, (double*)L.size->p, n0);
  CSC_LAUNCHED(ctx);
  CSC_CUDA(ctx, cudaMemcpyAsync(L.csize->p, L.size->p, (size_t)n0 * 8, cudaMemcpyDeviceToDevice, s));
  DevPtr node_of, flag;
  CSC_TRY(dev_alloc(ctx, (size_t)n0 * 4, &node_of));
  CSC_TRY(dev_alloc(ctx, 8, &flag));
  k_ld_iota_ones<<<lb(n0), 256, 0, s>>>((int32_t*)node_of->p, nullptr, n0);
  CSC_LAUNCHED(ctx);
  int level = 0;
  for (; level < max_levels; ++level) {
    const int64_t n = L.n, m = L.m;
    const int nbits = bits_for(n);
    const int32_t* src = (const int32_t*)L.src->p;
    const int32_t* col = (const int32_t*)L.col->p;
    const double* w = (const double*)L.w->p;
    int32_t* comm = (int32_t*)L.comm->p;
    double* size = (double*)L.size->p;
    double* csize = (double*)L.csize->p;
    // ---------------- local moving ----------------
    int total_moved = 0, quiet = 0;
    DevPtr newc;
    CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &newc));
    for (int sweep = 0; sweep < max_sweeps && m > 0; ++sweep) {
      DevPtr keys, vals, uk, uv;
      CSC_TRY(dev_alloc(ctx, (size_t)m * 8, &keys));
      CSC_TRY(dev_alloc(ctx, (size_t)m * 8, &vals));
      k_ld_keys<<<lb(m), 256, 0, s>>>(src, col, w, comm, comm, 0, m, (u64*)keys->p, (double*)vals->p);
      CSC_LAUNCHED(ctx);
      int64_t nu = 0;
      CSC_TRY(sort_reduce(ctx, keys, vals, m, 32 + nbits, uk, uv, &nu));
      CSC_CUDA(ctx, cudaMemsetAsync(flag->p, 0, 8, s));
      k_ld_decide<<<lb(n), 256, 0, s>>>((const u64*)uk->p, (const double*)uv->p, nu, n, comm, size, csize, gamma, sweep, (int32_t*)newc->p);
      CSC_LAUNCHED(ctx);
      k_ld_apply<<<lb(n), 256, 0, s>>>((const int32_t*)newc->p, comm, size, csize, n, (int*)flag->p);
      CSC_LAUNCHED(ctx);
      int moved = 0;
      CSC_CUDA(ctx, cudaMemcpyAsync(&moved, flag->p, 4, cudaMemcpyDeviceToHost, s));
      CSC_CUDA(ctx, cudaStreamSynchronize(s));
      total_moved += moved;
      quiet = moved == 0 ? quiet + 1 : 0;
      if (quiet >= 2) break;                       // both halves had their turn and nothing moved
    }
    // ---------------- refinement ----------------
    DevPtr ref, rsize, members, members_before, ext, prop;
    CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &ref));
    CSC_TRY(dev_alloc(ctx, (size_t)n * 8, &rsize));
    CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &members));
    CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &members_before));
    CSC_TRY(dev_alloc(ctx, (size_t)n * 8, &ext));
    CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &prop));
    k_ld_iota_ones<<<lb(n), 256, 0, s>>>((int32_t*)ref->p, nullptr, n);
    CSC_LAUNCHED(ctx);
    CSC_CUDA(ctx, cudaMemcpyAsync(rsize->p, size, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    {
      std::vector<int32_t> ones((size_t)n, 1);
      CSC_CUDA(ctx, cudaMemcpyAsync(members->p, ones.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
      CSC_CUDA(ctx, cudaStreamSynchronize(s));
    }
    int total_merged = 0;
    for (int round = 0; round < 24 && m > 0; ++round) {
      CSC_CUDA(ctx, cudaMemsetAsync(ext->p, 0, (size_t)n * 8, s));
      k_ld_ext<<<lb(m), 256, 0, s>>>(src, col, w, comm, (const int32_t*)ref->p, m, (double*)ext->p);
      CSC_LAUNCHED(ctx);
      DevPtr keys, vals, uk, uv;
      CSC_TRY(dev_alloc(ctx, (size_t)m * 8, &keys));
      CSC_TRY(dev_alloc(ctx, (size_t)m * 8, &vals));
      k_ld_keys<<<lb(m), 256, 0, s>>>(src, col, w, (const int32_t*)ref->p, comm, 1, m, (u64*)keys->p, (double*)vals->p);
      CSC_LAUNCHED(ctx);
      int64_t nu = 0;
      CSC_TRY(sort_reduce(ctx, keys, vals, m, 32 + nbits, uk, uv, &nu));
      k_ld_propose<<<lb(n), 256, 0, s>>>((const u64*)uk->p, (const double*)uv->p, nu, n, comm, (const int32_t*)ref->p, size,
                                         (const double*)rsize->p, csize, (const double*)ext->p, (const int32_t*)members->p, gamma,
                                         (int32_t*)prop->p);
      CSC_LAUNCHED(ctx);
      CSC_CUDA(ctx, cudaMemsetAsync(flag->p, 0, 8, s));
      CSC_CUDA(ctx, cudaMemcpyAsync(members_before->p, members->p, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
      k_ld_merge<<<lb(n), 256, 0, s>>>((const int32_t*)prop->p, (int32_t*)ref->p, size, (double*)rsize->p,
                                       (const int32_t*)members_before->p, (int32_t*)members->p, n, (int*)flag->p);
      CSC_LAUNCHED(ctx);
      int merged = 0;
      CSC_CUDA(ctx, cudaMemcpyAsync(&merged, flag->p, 4, cudaMemcpyDeviceToHost, s));
      CSC_CUDA(ctx, cudaStreamSynchronize(s));
      total_merged += merged;
      if (merged == 0) break;
    }
    if (total_merged == 0 && total_moved == 0) break;          // stable: the partition of this level is the answer
    if (total_merged == 0) {
      // nothing to aggregate: every refined set is a single node; the communities found by the moves stand
      ++level;
      break;
    }
    // ---------------- aggregation on the refined partition ----------------
    DevPtr used, newid;
    CSC_TRY(dev_alloc(ctx, (size_t)(n + 1) * 8, &used, true));
    CSC_TRY(dev_alloc(ctx, (size_t)(n + 1) * 8, &newid));
    k_ld_mark<<<lb(n), 256, 0, s>>>((const int32_t*)ref->p, n, (int64_t*)used->p);
    CSC_LAUNCHED(ctx);
    {
      size_t tb = 0;
      CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int64_t*)used->p, (int64_t*)newid->p, n + 1, s));
      DevPtr tmp;
      CSC_TRY(dev_alloc(ctx, tb, &tmp));
      CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp->p, tb, (const int64_t*)used->p, (int64_t*)newid->p, n + 1, s));
      ctx->launches += 2;
    }
    int64_t n_new = 0;
    CSC_CUDA(ctx, cudaMemcpyAsync(&n_new, (const int64_t*)newid->p + n, 8, cudaMemcpyDeviceToHost, s));
    CSC_CUDA(ctx, cudaStreamSynchronize(s));
    Level N2;
    N2.n = n_new;
    CSC_TRY(dev_alloc(ctx, (size_t)n_new * 8, &N2.size, true));
    CSC_TRY(dev_alloc(ctx, (size_t)n_new * 4, &N2.comm));
    CSC_TRY(dev_alloc(ctx, (size_t)n_new * 8, &N2.csize, true));
    DevPtr cmin;
    CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &cmin));
    CSC_CUDA(ctx, cudaMemsetAsync(cmin->p, 0x7f, (size_t)n * 4, s));
    k_ld_new_nodes<<<lb(n), 256, 0, s>>>((const int32_t*)ref->p, (const int64_t*)newid->p, comm, size, n, (double*)N2.size->p,
                                         (int32_t*)cmin->p);
    CSC_LAUNCHED(ctx);
    k_ld_new_comm<<<lb(n), 256, 0, s>>>((const int32_t*)ref->p, (const int64_t*)newid->p, comm, (const int32_t*)cmin->p, n,
                                        (int32_t*)N2.comm->p);
    CSC_LAUNCHED(ctx);
    k_ld_csize<<<lb(n_new), 256, 0, s>>>((const int32_t*)N2.comm->p, (const double*)N2.size->p, n_new, (double*)N2.csize->p);
    CSC_LAUNCHED(ctx);
    k_ld_map_nodes<<<lb(n0), 256, 0, s>>>((int32_t*)node_of->p, (const int32_t*)ref->p, (const int64_t*)newid->p, n0);
    CSC_LAUNCHED(ctx);
    {
      DevPtr keys, vals, uk, uv;
      CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(m, 1) * 8, &keys));
      CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(m, 1) * 8, &vals));
      k_ld_edge_keys<<<lb(m), 256, 0, s>>>(src, col, (const int32_t*)ref->p, (const int64_t*)newid->p, m, (u64*)keys->p);
      CSC_LAUNCHED(ctx);
      CSC_CUDA(ctx, cudaMemcpyAsync(vals->p, w, (size_t)m * 8, cudaMemcpyDeviceToDevice, s));
      int64_t nu = 0;
      CSC_TRY(sort_reduce(ctx, keys, vals, m, 32 + bits_for(n_new), uk, uv, &nu));
      N2.m = nu;
      CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(nu, 1) * 4, &N2.src));
      CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(nu, 1) * 4, &N2.col));
      CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(nu, 1) * 8, &N2.w));
      CSC_TRY(dev_alloc(ctx, (size_t)(n_new + 1) * 8, &N2.rowptr));
      k_ld_split_keys<<<lb(nu), 256, 0, s>>>((const u64*)uk->p, nu, (int32_t*)N2.src->p, (int32_t*)N2.col->p);
      CSC_LAUNCHED(ctx);
      CSC_CUDA(ctx, cudaMemcpyAsync(N2.w->p, uv->p, (size_t)nu * 8, cudaMemcpyDeviceToDevice, s));
      k_ld_rowptr<<<lb(n_new + 1), 256, 0, s>>>((const int32_t*)N2.src->p, nu, n_new, (int64_t*)N2.rowptr->p);
      CSC_LAUNCHED(ctx);
      CSC_CUDA(ctx, cudaStreamSynchronize(s));
    }
    L = N2;
  }
  // ---------------- labels of the original nodes, compacted to 0..C-1 in order of first appearance ----------------
  DevPtr lab;
  CSC_TRY(dev_alloc(ctx, (size_t)n0 * 4, &lab));
  k_ld_labels<<<lb(n0), 256, 0, s>>>((const int32_t*)node_of->p, (const int32_t*)L.comm->p, n0, (int32_t*)lab->p);
  CSC_LAUNCHED(ctx);
  std::vector<int32_t> h((size_t)n0);
  CSC_CUDA(ctx, cudaMemcpyAsync(h.data(), lab->p, (size_t)n0 * 4, cudaMemcpyDeviceToHost, s));
  CSC_CUDA(ctx, cudaStreamSynchronize(s));
  std::vector<int32_t> remap((size_t)L.n, -1);
  std::vector<double> count;
  int32_t nc = 0;
  for (int64_t i = 0; i < n0; ++i) {
    int32_t& r = remap[h[i]];
    if (r < 0) {
      r = nc++;
      count.push_back(0.0);
    }
    labels_out[i] = r;
    count[r] += 1.0;
  }
  if (n_communities) *n_communities = nc;
  if (levels_out) *levels_out = level;
  if (quality) {
    CSC_CUDA(ctx, cudaMemcpyAsync(lab->p, labels_out, (size_t)n0 * 4, cudaMemcpyHostToDevice, s));
    CSC_CUDA(ctx, cudaMemsetAsync(flag->p, 0, 8, s));
    if (g->nnz > 0) {
      k_ld_quality_edges<<<ctx->sm_count * 8, 256, 0, s>>>((const int64_t*)g->colptr->p, (const int32_t*)g->rowval->p,
                                                           (const double*)g->val->p, (const int32_t*)lab->p, n0, (double*)flag->p);
      CSC_LAUNCHED(ctx);
    }
    double e_in = 0.0;
    CSC_CUDA(ctx, cudaMemcpyAsync(&e_in, flag->p, 8, cudaMemcpyDeviceToHost, s));
    CSC_CUDA(ctx, cudaStreamSynchronize(s));
    double pen = 0.0;
    for (double c : count) pen += c * (c - 1.0) / 2.0;
    *quality = e_in - gamma * pen;
  }
  return CSC_OK;
  CSC_GUARD_END(ctx)
}
