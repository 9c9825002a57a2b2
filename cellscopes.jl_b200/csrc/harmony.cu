// run_harmony (src/integration/int_processing.jl:1-12 -> HarmonyObject, src/integration/int_objects.jl:115-236, with
// init_cluster! / cluster! / update_R! / compute_objective! / moe_correct_ridge of src/integration/int_utils.jl:49-199):
// soft k-means of the cells in PCA space with a batch-diversity penalty, then a ridge regression per cluster that
// removes the batch terms.  Everything is a reduction over cells or a small dense product (K <= 256 clusters, d <= 64
// dimensions, B <= 64 batches); it runs in fp64 like the reference (Array{Float64,2} throughout).
//
// Layout: every N x something matrix is cell-major ([N][d], [N][K]) -- the memory of the reference's column-major
// d x N / K x N arrays.  Two kernel shapes:
//   * warp per cell (k_hm_dist_scale, k_hm_update_block, k_hm_objective, k_hm_apply): lanes own clusters k = lane,
//     lane + 32, ...; the K x d centroid table / the K x B tables sit in shared memory; per-cell sums by shuffles;
//   * thread per cluster, a CTA walks its cells (k_hm_wsum, k_hm_rsum): the per-cluster accumulators live in registers
//     (d of them) or in a thread-owned shared-memory column, no atomics inside the loop, one flush per CTA.
// The reference's quirks are kept because a drop-in must return what the reference returns: its `optimized_norm` is
// the L1 norm (int_utils.jl:32-41: `norm(view, 1)`), so centroids are scaled to unit L1 length (:60-61, :117-118); the
// first step divides every cell by its largest coordinate before the L2 normalisation (int_objects.jl:201-203: a sign
// flip for a cell whose coordinates are all negative); after a correction Z_cos is Z_corr with dimension j divided by
// the L1 norm of the old Z_cos of CELL j (int_utils.jl:197: `optimized_norm(Z_cos, 1:size(Z_cos, 1))`), not a
// renormalisation; cells beyond n_blocks * floor(N / n_blocks) of the random order are not updated in a round (:141).
// What cannot be reproduced is Julia's RNG: the k-means initialisation (Clustering.kmeans after Random.seed!) is the
// caller's (initial centroids come in through the ABI), and the update order of a round is the stable argsort of a
// counter-based hash of (seed, round, cell) -- the oracle uses the same hash, so both sides walk the same blocks.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int HM_DMAX = 64;

__device__ __forceinline__ uint64_t hm_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void k_hm_perm_keys(uint64_t seed, uint64_t round, int64_t n, unsigned long long* __restrict__ keys, int32_t* __restrict__ ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t h = hm_splitmix64(seed ^ 0xD1B54A32D192ED03ull);
  h = hm_splitmix64(h ^ round);
  h = hm_splitmix64(h ^ ((uint64_t)i * 0x2545F4914F6CDD1Dull));
  keys[i] = h;
  ids[i] = (int32_t)i;
}

// Z_cos = Z / max_j Z, then unit L2 length (int_objects.jl:201-203)
__global__ void k_hm_zcos_init(const double* __restrict__ z, int64_t n, int d, double* __restrict__ zc) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double m = -INFINITY;
  for (int j = 0; j < d; ++j) m = fmax(m, z[c * d + j]);
  double s = 0.0;
  for (int j = 0; j < d; ++j) {
    const double v = z[c * d + j] / m;
    s += v * v;
  }
  const double nrm = sqrt(s);
  for (int j = 0; j < d; ++j) zc[c * d + j] = (z[c * d + j] / m) / nrm;
}

// out[k][j] += sum over the listed cells of R[c][k] * Z[c][j]   (Y = Z_cos R', int_utils.jl:116; Phi_Rk Z_orig', :193)
__global__ void __launch_bounds__(256)
k_hm_wsum(const double* __restrict__ z, const double* __restrict__ r, const int32_t* __restrict__ list, int64_t n_list, int K,
          int d, double* __restrict__ out) {
  __shared__ double zrow[HM_DMAX];
  const int k = threadIdx.x;
  double acc[HM_DMAX];
#pragma unroll
  for (int j = 0; j < HM_DMAX; ++j) acc[j] = 0.0;
  const int64_t per = (n_list + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(lo + per, n_list);
  for (int64_t i = lo; i < hi; ++i) {
    const int64_t c = list ? list[i] : i;
    __syncthreads();
    if (threadIdx.x < d) zrow[threadIdx.x] = z[c * d + threadIdx.x];
    __syncthreads();
    if (k < K) {
      const double w = r[c * K + k];
#pragma unroll
      for (int j = 0; j < HM_DMAX; ++j)
        if (j < d) acc[j] = fma(w, zrow[j], acc[j]);
    }
  }
  if (k < K) {
#pragma unroll
    for (int j = 0; j < HM_DMAX; ++j)
      if (j < d && acc[j] != 0.0) atomicAdd(out + (int64_t)k * d + j, acc[j]);
  }
}

// E[k][b] += sign * Pr[b] * sum_c R[c][k];  O[k][b] += sign * sum_{c in batch b} R[c][k]   over the listed cells
// (int_utils.jl:143-144, :152-153; with sign = +1 on zeroed tables: E and O of init_cluster!, :68-69)
__global__ void __launch_bounds__(256)
k_hm_rsum(const double* __restrict__ r, const int32_t* __restrict__ list, int64_t n_list, const int32_t* __restrict__ batch, int K,
          int B, double sign, const double* __restrict__ pr, double* __restrict__ E, double* __restrict__ O) {
  extern __shared__ double tab[];      // [B][blockDim.x]: column k is owned by thread k
  const int k = threadIdx.x;
  for (int b = 0; b < B; ++b) tab[b * blockDim.x + k] = 0.0;
  const int64_t per = (n_list + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(lo + per, n_list);
  if (k < K) {
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t c = list ? list[i] : i;
      tab[batch[c] * blockDim.x + k] += r[c * K + k];
    }
    double tot = 0.0;
    for (int b = 0; b < B; ++b) tot += tab[b * blockDim.x + k];
    for (int b = 0; b < B; ++b) {
      const double ob = tab[b * blockDim.x + k];
      if (ob != 0.0) atomicAdd(O + (int64_t)k * B + b, sign * ob);
      if (tot != 0.0) atomicAdd(E + (int64_t)k * B + b, sign * tot * pr[b]);
    }
  }
}

// dist[c][k] = 2 (1 - <Y_k, Zcos_c>)  (:62, :119) and, if sd != null, scale_dist = exp(-dist / sigma_k - max_k(...))
// (:128-131); with r != null also R = scale_dist / sum_k (the initial assignment, :63-67)
__global__ void __launch_bounds__(256)
k_hm_dist_scale(const double* __restrict__ zc, const double* __restrict__ Y, const double* __restrict__ sigma, int64_t n, int K,
                int d, double* __restrict__ dist, double* __restrict__ sd, double* __restrict__ r) {
  extern __shared__ double ys[];       // [K][d]
  for (int i = threadIdx.x; i < K * d; i += blockDim.x) ys[i] = Y[i];
  __syncthreads();
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < n; c += nwarp) {
    double m = -INFINITY;
    for (int k = lane; k < K; k += 32) {
      double dot = 0.0;
      for (int j = 0; j < d; ++j) dot = fma(ys[k * d + j], zc[c * d + j], dot);
      const double dd = 2.0 * (1.0 - dot);
      dist[c * K + k] = dd;
      m = fmax(m, -dd / sigma[k]);
    }
    if (!sd) continue;
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
    double s = 0.0;
    for (int k = lane; k < K; k += 32) {
      const double e = exp(-dist[c * K + k] / sigma[k] - m);
      sd[c * K + k] = e;
      s += e;
    }
    if (r) {
      s = warp_sum(s);
      for (int k = lane; k < K; k += 32) r[c * K + k] = sd[c * K + k] / s;
    }
  }
}

// pw[k][b] = ((E + 1) / (O + 1)) ^ theta_b   (:147-149)
__global__ void k_hm_pow_table(const double* __restrict__ E, const double* __restrict__ O, const double* __restrict__ theta, int K,
                               int B, double* __restrict__ pw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * B) pw[i] = pow((E[i] + 1.0) / (O[i] + 1.0), theta[i % B]);
}

// R[c][:] = scale_dist[c][:] .* pw[:][batch_c], normalised to unit sum (:146-151), for the cells of one block
__global__ void __launch_bounds__(256)
k_hm_update_block(const double* __restrict__ sd, const double* __restrict__ pw, const int32_t* __restrict__ list, int64_t n_list,
                  const int32_t* __restrict__ batch, int K, int B, double* __restrict__ r) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n_list; i += nwarp) {
    const int64_t c = list[i];
    const int b = batch[c];
    double s = 0.0;
    for (int k = lane; k < K; k += 32) {
      const double v = sd[c * K + k] * pw[k * B + b];
      r[c * K + k] = v;
      s += fabs(v);                      // optimized_norm is the L1 norm
    }
    s = warp_sum(s);
    for (int k = lane; k < K; k += 32) r[c * K + k] /= s;
  }
}

// the three terms of compute_objective! (:73-86): sum R dist; sum sigma_k R log R; sum R sigma_k theta_b log((O+1)/(E+1))
__global__ void __launch_bounds__(256)
k_hm_objective(const double* __restrict__ r, const double* __restrict__ dist, const double* __restrict__ sigma,
               const double* __restrict__ theta, const double* __restrict__ E, const double* __restrict__ O,
               const int32_t* __restrict__ batch, int64_t n, int K, int B, double* __restrict__ out3) {
  double a = 0.0, e = 0.0, x = 0.0;
  const int64_t total = n * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / K;
    const int k = (int)(i - c * K);
    const int b = batch[c];
    const double rv = r[i];
    a += rv * dist[i];
    const double y = rv * log(rv);
    e += isfinite(y) ? y * sigma[k] : 0.0;                       // safe_entropy (:88-92)
    x += rv * sigma[k] * theta[b] * log((O[k * B + b] + 1.0) / (E[k * B + b] + 1.0));
  }
  a = warp_sum(a);
  e = warp_sum(e);
  x = warp_sum(x);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out3 + 0, a);
    atomicAdd(out3 + 1, e);
    atomicAdd(out3 + 2, x);
  }
}

// Z_corr[c] = Z_orig[c] - sum_k R[c][k] W[k][batch_c + 1][:]   (:191-196; row 0 of W is zeroed, :194)
// then Z_cos[c][j] = Z_corr[c][j] / nj[j]   (:197)
__global__ void __launch_bounds__(256)
k_hm_apply(const double* __restrict__ zo, const double* __restrict__ r, const double* __restrict__ W /*[K][B+1][d]*/,
           const int32_t* __restrict__ batch, const double* __restrict__ nj, int64_t n, int K, int B, int d,
           double* __restrict__ zcorr, double* __restrict__ zcos) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < n; c += nwarp) {
    const int b = batch[c];
    for (int j = lane; j < d; j += 32) {
      double v = zo[c * d + j];
      for (int k = 0; k < K; ++k) v -= r[c * K + k] * W[((int64_t)k * (B + 1) + b + 1) * d + j];   // Z_corr .-= W' Phi_Rk, cluster by cluster
      zcorr[c * d + j] = v;
      zcos[c * d + j] = v / nj[j];
    }
  }
}

__global__ void k_hm_l1_rows(const double* __restrict__ x, int rows, int cols, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double s = 0.0;
  for (int j = 0; j < cols; ++j) s += fabs(x[(int64_t)i * cols + j]);
  out[i] = s;
}
__global__ void k_hm_div_rows(double* __restrict__ x, int rows, int cols, const double* __restrict__ nrm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * cols) x[i] /= nrm[i / cols];
}

// inverse of a small dense matrix by Gauss-Jordan with partial pivoting (inv(x), int_utils.jl:193)
bool invert_small(std::vector<double>& a, int n) {
  std::vector<double> inv((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int col = 0; col < n; ++col) {
    int piv = col;
    for (int r = col + 1; r < n; ++r)
      if (std::fabs(a[(size_t)r * n + col]) > std::fabs(a[(size_t)piv * n + col])) piv = r;
    if (a[(size_t)piv * n + col] == 0.0) return false;
    if (piv != col)
      for (int j = 0; j < n; ++j) {
        std::swap(a[(size_t)piv * n + j], a[(size_t)col * n + j]);
        std::swap(inv[(size_t)piv * n + j], inv[(size_t)col * n + j]);
      }
    const double p = a[(size_t)col * n + col];
    for (int j = 0; j < n; ++j) {
      a[(size_t)col * n + j] /= p;
      inv[(size_t)col * n + j] /= p;
    }
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      const double f = a[(size_t)r * n + col];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) {
        a[(size_t)r * n + j] -= f * a[(size_t)col * n + j];
        inv[(size_t)r * n + j] -= f * inv[(size_t)col * n + j];
      }
    }
  }
  a.swap(inv);
  return true;
}

inline unsigned hm_blocks(int64_t n, int per = 256) { return (unsigned)std::max<int64_t>(1, ceil_div64(n, per)); }

}  // namespace

extern "C" csc_status csc_harmony(csc_ctx* ctx, const double* z_orig, int64_t n, int32_t d, const int32_t* batch, int32_t n_batch,
                                  const double* y0, int32_t K, const double* theta, const double* lamb, const double* sigma,
                                  double block_size, int32_t max_iter_harmony, int32_t max_iter_kmeans, double eps_kmeans,
                                  double eps_harmony, uint64_t seed, double* z_corr_out, double* objective_harmony,
                                  int32_t* n_objective, int32_t* kmeans_rounds, int32_t* n_rounds, double* r_out) {
  if (!ctx || !z_orig || !batch || !y0 || !theta || !lamb || !sigma || !z_corr_out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  const int B = n_batch;
  CSC_REQUIRE(ctx, n >= 2 && n < ((int64_t)1 << 31), "csc_harmony: bad cell count");
  CSC_REQUIRE(ctx, d >= 1 && d <= HM_DMAX, "csc_harmony: %d dimensions (max %d)", d, HM_DMAX);
  CSC_REQUIRE(ctx, K >= 1 && K <= 256, "csc_harmony: %d clusters (max 256)", K);
  CSC_REQUIRE(ctx, B >= 1 && B <= 64, "csc_harmony: %d batches (max 64)", B);
  CSC_REQUIRE(ctx, n >= d, "csc_harmony: fewer cells than dimensions");
  CSC_REQUIRE(ctx, block_size > 0.0 && block_size <= 1.0, "csc_harmony: block_size must be in (0, 1]");
  std::vector<int64_t> nb((size_t)B, 0);
  for (int64_t c = 0; c < n; ++c) {
    CSC_REQUIRE(ctx, batch[c] >= 0 && batch[c] < B, "csc_harmony: batch id %d of cell %lld outside [0,%d)", batch[c], (long long)c, B);
    ++nb[batch[c]];
  }
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_HARMONY);
  cudaStream_t s = ctx->stream;
  // per-batch cell lists (ascending cell id) for the ridge sums
  std::vector<int64_t> boff((size_t)B + 1, 0);
  for (int b = 0; b < B; ++b) boff[b + 1] = boff[b] + nb[b];
  std::vector<int32_t> blist((size_t)n);
  {
    std::vector<int64_t> fill(boff.begin(), boff.end() - 1);
    for (int64_t c = 0; c < n; ++c) blist[fill[batch[c]]++] = (int32_t)c;
  }
  std::vector<double> pr((size_t)B);
  for (int b = 0; b < B; ++b) pr[b] = (double)nb[b] / (double)n;

  DevPtr dZo, dZc, dZr, dR, dDist, dSd, dY, dBatch, dBl, dE, dO, dPw, dTh, dSg, dPr, dObj, dT, dW, dNj, dNk, dKeys, dKeys2, dIds, dPerm;
  const size_t nd = (size_t)n * d * 8, nk = (size_t)n * K * 8;
  CSC_TRY(dev_alloc(ctx, nd, &dZo));
  CSC_TRY(dev_alloc(ctx, nd, &dZc));
  CSC_TRY(dev_alloc(ctx, nd, &dZr));
  CSC_TRY(dev_alloc(ctx, nk, &dR));
  CSC_TRY(dev_alloc(ctx, nk, &dDist));
  CSC_TRY(dev_alloc(ctx, nk, &dSd));
  CSC_TRY(dev_alloc(ctx, (size_t)K * d * 8, &dY));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &dBatch));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &dBl));
  CSC_TRY(dev_alloc(ctx, (size_t)K * B * 8, &dE, true));
  CSC_TRY(dev_alloc(ctx, (size_t)K * B * 8, &dO, true));
  CSC_TRY(dev_alloc(ctx, (size_t)K * B * 8, &dPw));
  CSC_TRY(dev_alloc(ctx, (size_t)B * 8, &dTh));
  CSC_TRY(dev_alloc(ctx, (size_t)K * 8, &dSg));
  CSC_TRY(dev_alloc(ctx, (size_t)B * 8, &dPr));
  CSC_TRY(dev_alloc(ctx, 3 * 8, &dObj));
  CSC_TRY(dev_alloc(ctx, (size_t)K * B * d * 8, &dT));
  CSC_TRY(dev_alloc(ctx, (size_t)K * (B + 1) * d * 8, &dW));
  CSC_TRY(dev_alloc(ctx, (size_t)HM_DMAX * 8, &dNj));
  CSC_TRY(dev_alloc(ctx, (size_t)K * 8, &dNk));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 8, &dKeys));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 8, &dKeys2));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &dIds));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &dPerm));
  double *Zo = (double*)dZo->p, *Zc = (double*)dZc->p, *Zr = (double*)dZr->p, *R = (double*)dR->p, *Dist = (double*)dDist->p,
         *Sd = (double*)dSd->p, *Y = (double*)dY->p, *E = (double*)dE->p, *O = (double*)dO->p, *Pw = (double*)dPw->p,
         *Th = (double*)dTh->p, *Sg = (double*)dSg->p, *Pr = (double*)dPr->p, *Obj = (double*)dObj->p, *T = (double*)dT->p,
         *W = (double*)dW->p, *Nj = (double*)dNj->p, *Nk = (double*)dNk->p;
  const int32_t* Bt = (const int32_t*)dBatch->p;
  const int32_t* Bl = (const int32_t*)dBl->p;
  int32_t* Perm = (int32_t*)dPerm->p;
  CSC_CUDA(ctx, cudaMemcpyAsync(Zo, z_orig, nd, cudaMemcpyHostToDevice, s));
  CSC_CUDA(ctx, cudaMemcpyAsync(dBatch->p, batch, (size_t)n * 4, cudaMemcpyHostToDevice, s));
  CSC_CUDA(ctx, cudaMemcpyAsync(dBl->p, blist.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
  CSC_CUDA(ctx, cudaMemcpyAsync(Y, y0, (size_t)K * d * 8, cudaMemcpyHostToDevice, s));
  CSC_CUDA(ctx, cudaMemcpyAsync(Th, theta, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  CSC_CUDA(ctx, cudaMemcpyAsync(Sg, sigma, (size_t)K * 8, cudaMemcpyHostToDevice, s));
  CSC_CUDA(ctx, cudaMemcpyAsync(Pr, pr.data(), (size_t)B * 8, cudaMemcpyHostToDevice, s));

  const int kthreads = ((K + 31) / 32) * 32;
  const int grid_red = ctx->sm_count * 4;
  const int grid_warp = ctx->sm_count * 8;
  const size_t smem_y = (size_t)K * d * 8;
  if (smem_y > 48 * 1024)
    CSC_CUDA(ctx, cudaFuncSetAttribute(k_hm_dist_scale, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_y));
  const size_t smem_rs = (size_t)B * kthreads * 8;
  if (smem_rs > 48 * 1024)
    CSC_CUDA(ctx, cudaFuncSetAttribute(k_hm_rsum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rs));

  auto normalise_Y = [&]() -> csc_status {      // Y ./ optimized_norm(Y)': every centroid to unit L1 length (:60-61, :117-118)
    k_hm_l1_rows<<<hm_blocks(K), 256, 0, s>>>(Y, K, d, Nk);
    CSC_LAUNCHED(ctx);
    k_hm_div_rows<<<hm_blocks((int64_t)K * d), 256, 0, s>>>(Y, K, d, Nk);
    CSC_LAUNCHED(ctx);
    return CSC_OK;
  };
  std::vector<double> obj_kmeans, obj_harmony;
  std::vector<int32_t> rounds;
  auto objective = [&]() -> csc_status {
    CSC_CUDA(ctx, cudaMemsetAsync(Obj, 0, 24, s));
    k_hm_objective<<<grid_warp, 256, 0, s>>>(R, Dist, Sg, Th, E, O, Bt, n, K, B, Obj);
    CSC_LAUNCHED(ctx);
    double h[3];
    CSC_CUDA(ctx, cudaMemcpyAsync(h, Obj, 24, cudaMemcpyDeviceToHost, s));
    CSC_CUDA(ctx, cudaStreamSynchronize(s));
    obj_kmeans.push_back(h[0] + h[1] + h[2]);
    return CSC_OK;
  };

  // ---- init (int_objects.jl:199-204, init_cluster! int_utils.jl:58-71) -----------------------
  k_hm_zcos_init<<<hm_blocks(n), 256, 0, s>>>(Zo, n, d, Zc);
  CSC_LAUNCHED(ctx);
  CSC_TRY(normalise_Y());
  k_hm_dist_scale<<<grid_warp, 256, smem_y, s>>>(Zc, Y, Sg, n, K, d, Dist, Sd, R);
  CSC_LAUNCHED(ctx);
  k_hm_rsum<<<grid_red, kthreads, smem_rs, s>>>(R, nullptr, n, Bt, K, B, 1.0, Pr, E, O);
  CSC_LAUNCHED(ctx);
  CSC_TRY(objective());
  obj_harmony.push_back(obj_kmeans.back());

  const int n_blocks = (int)std::ceil(1.0 / block_size);
  const int64_t per_block = n / n_blocks;
  uint64_t round = 0;
  const int window = 3;
  for (int it = 0; it < max_iter_harmony; ++it) {
    // ---- cluster! (:112-132) ----
    int j_last = 0;
    for (int i = 1; i <= max_iter_kmeans; ++i) {
      CSC_CUDA(ctx, cudaMemsetAsync(Y, 0, (size_t)K * d * 8, s));
      k_hm_wsum<<<grid_red, kthreads, 0, s>>>(Zc, R, nullptr, n, K, d, Y);
      CSC_LAUNCHED(ctx);
      CSC_TRY(normalise_Y());
      k_hm_dist_scale<<<grid_warp, 256, smem_y, s>>>(Zc, Y, Sg, n, K, d, Dist, Sd, nullptr);
      CSC_LAUNCHED(ctx);
      // update_R! (:134-155): the round's random order, cut into n_blocks pieces of floor(N / n_blocks) cells
      {
        k_hm_perm_keys<<<hm_blocks(n), 256, 0, s>>>(seed, round++, n, (unsigned long long*)dKeys->p, (int32_t*)dIds->p);
        CSC_LAUNCHED(ctx);
        size_t tb = 0;
        CSC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tb, (const unsigned long long*)dKeys->p, (unsigned long long*)dKeys2->p,
                                                      (const int32_t*)dIds->p, Perm, n, 0, 64, s));
        DevPtr tmp;
        CSC_TRY(dev_alloc(ctx, tb, &tmp));
        CSC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp->p, tb, (const unsigned long long*)dKeys->p, (unsigned long long*)dKeys2->p,
                                                      (const int32_t*)dIds->p, Perm, n, 0, 64, s));
        ctx->launches += 9;
      }
      for (int b = 0; b < n_blocks; ++b) {
        const int64_t lo = per_block * b, hi = std::min<int64_t>(per_block * (b + 1), n);
        if (hi <= lo) continue;
        const int32_t* list = Perm + lo;
        const int g_red = (int)std::min<int64_t>(grid_red, ceil_div64(hi - lo, 8));
        k_hm_rsum<<<g_red, kthreads, smem_rs, s>>>(R, list, hi - lo, Bt, K, B, -1.0, Pr, E, O);
        CSC_LAUNCHED(ctx);
        k_hm_pow_table<<<hm_blocks((int64_t)K * B), 256, 0, s>>>(E, O, Th, K, B, Pw);
        CSC_LAUNCHED(ctx);
        k_hm_update_block<<<(unsigned)std::min<int64_t>(grid_warp, ceil_div64(hi - lo, 8)), 256, 0, s>>>(Sd, Pw, list, hi - lo, Bt, K, B, R);
        CSC_LAUNCHED(ctx);
        k_hm_rsum<<<g_red, kthreads, smem_rs, s>>>(R, list, hi - lo, Bt, K, B, 1.0, Pr, E, O);
        CSC_LAUNCHED(ctx);
      }
      CSC_TRY(objective());
      if (i > window) {
        // check_convergence(0) (:160-170): two-iteration windows of objective_kmeans
        const size_t okl = obj_kmeans.size();
        double o_old = 0.0, o_new = 0.0;
        for (int w = 1; w <= window - 1; ++w) {
          o_old += obj_kmeans[okl - 3 - w];
          o_new += obj_kmeans[okl - 2 - w];
        }
        if (std::fabs(o_old - o_new) / std::fabs(o_old) < eps_kmeans) break;
      }
      j_last = i;
    }
    rounds.push_back(j_last);
    obj_harmony.push_back(obj_kmeans.back());
    // ---- moe_correct_ridge (:188-199) ----
    {
      CSC_CUDA(ctx, cudaMemsetAsync(T, 0, (size_t)K * B * d * 8, s));
      for (int b = 0; b < B; ++b) {
        if (nb[b] == 0) continue;
        k_hm_wsum<<<(unsigned)std::min<int64_t>(grid_red, ceil_div64(nb[b], 8)), kthreads, 0, s>>>(Zo, R, Bl + boff[b], nb[b], K, d,
                                                                                              T + (size_t)b * K * d);
        CSC_LAUNCHED(ctx);
      }
      std::vector<double> hT((size_t)K * B * d), hO((size_t)K * B), hW((size_t)K * (B + 1) * d, 0.0);
      CSC_CUDA(ctx, cudaMemcpyAsync(hT.data(), T, hT.size() * 8, cudaMemcpyDeviceToHost, s));
      CSC_CUDA(ctx, cudaMemcpyAsync(hO.data(), O, hO.size() * 8, cudaMemcpyDeviceToHost, s));
      CSC_CUDA(ctx, cudaStreamSynchronize(s));
      const int m = B + 1;
      for (int k = 0; k < K; ++k) {
        // x = Phi_Rk Phi_moe' + lamb: [sum_c R_kc, per-batch sums; per-batch sums, diag(per-batch sums)] + diag(0, lamb)
        std::vector<double> x((size_t)m * m, 0.0);
        double tot = 0.0;
        for (int b = 0; b < B; ++b) {
          const double ob = hO[(size_t)k * B + b];
          tot += ob;
          x[(size_t)0 * m + b + 1] = ob;
          x[(size_t)(b + 1) * m + 0] = ob;
          x[(size_t)(b + 1) * m + b + 1] = ob + lamb[b];
        }
        x[0] = tot;
        if (!invert_small(x, m)) return csc_fail(ctx, CSC_ERR_NUMERIC, "csc_harmony: singular ridge system for cluster %d", k);
        // rhs = Phi_Rk Z_orig': row 0 = sum over the batches, row b + 1 = T[b][k][:]
        for (int j = 0; j < d; ++j) {
          double r0 = 0.0;
          for (int b = 0; b < B; ++b) r0 += hT[((size_t)b * K + k) * d + j];
          for (int row = 1; row < m; ++row) {          // W[1, :] .= 0 (:194): row 0 is never used
            double v = x[(size_t)row * m + 0] * r0;
            for (int b = 0; b < B; ++b) v += x[(size_t)row * m + b + 1] * hT[((size_t)b * K + k) * d + j];
            hW[((size_t)k * m + row) * d + j] = v;
          }
        }
      }
      CSC_CUDA(ctx, cudaMemcpyAsync(W, hW.data(), hW.size() * 8, cudaMemcpyHostToDevice, s));
      // norms of the old Z_cos of the first d CELLS (:197)
      k_hm_l1_rows<<<1, 64, 0, s>>>(Zc, d, d, Nj);
      CSC_LAUNCHED(ctx);
      k_hm_apply<<<grid_warp, 256, 0, s>>>(Zo, R, W, Bt, Nj, n, K, B, d, Zr, Zc);
      CSC_LAUNCHED(ctx);
      CSC_CUDA(ctx, cudaStreamSynchronize(s));
    }
    // check_convergence(1) (:171-179)
    const double o_old = obj_harmony[obj_harmony.size() - 2], o_new = obj_harmony.back();
    if ((o_old - o_new) / std::fabs(o_old) < eps_harmony) break;
  }
  if (rounds.empty()) CSC_CUDA(ctx, cudaMemcpyAsync(Zr, Zo, nd, cudaMemcpyDeviceToDevice, s));   // max_iter_harmony == 0: Z_corr = Z_orig
  CSC_TRY(d2h_copy(ctx, z_corr_out, Zr, nd));
  if (r_out) CSC_TRY(d2h_copy(ctx, r_out, R, nk));
  if (objective_harmony && n_objective) {
    const int cap = *n_objective;
    *n_objective = (int32_t)std::min<size_t>(obj_harmony.size(), (size_t)cap);
    for (int i = 0; i < *n_objective; ++i) objective_harmony[i] = obj_harmony[i];
  }
  if (kmeans_rounds && n_rounds) {
    const int cap = *n_rounds;
    *n_rounds = (int32_t)std::min<size_t>(rounds.size(), (size_t)cap);
    for (int i = 0; i < *n_rounds; ++i) kmeans_rounds[i] = rounds[i];
  }
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------
// cluster_kmeans (int_utils.jl:49-56): Clustering.kmeans(data, K; maxiter = 100, tol = 1e-4) on the unit-length cells.
// Lloyd iterations in fp64; the seeding (Clustering.jl's k-means++ under Julia's RNG) is replaced by the K cells the
// caller names (run_harmony picks them with the counter-based hash, the oracle does the same).
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
k_km_assign(const double* __restrict__ x, const double* __restrict__ cen, int64_t n, int K, int d, int32_t* __restrict__ label,
            double* __restrict__ sums, double* __restrict__ counts, double* __restrict__ objective) {
  extern __shared__ double cs_[];      // [K][d]
  for (int i = threadIdx.x; i < K * d; i += blockDim.x) cs_[i] = cen[i];
  __syncthreads();
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double obj = 0.0;
  for (int64_t c = warp; c < n; c += nwarp) {
    double best = INFINITY;
    int bk = INT_MAX;
    for (int k = lane; k < K; k += 32) {
      double s = 0.0;
      for (int j = 0; j < d; ++j) {
        const double t = x[c * d + j] - cs_[k * d + j];
        s = fma(t, t, s);
      }
      if (s < best) {
        best = s;
        bk = k;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {       // smallest distance, ties to the lower cluster id
      const double ob = __shfl_xor_sync(FULL, best, o);
      const int ok = __shfl_xor_sync(FULL, bk, o);
      if (ob < best || (ob == best && ok < bk)) {
        best = ob;
        bk = ok;
      }
    }
    if (lane == 0) {
      label[c] = bk;
      atomicAdd(counts + bk, 1.0);
      obj += best;
    }
    for (int j = lane; j < d; j += 32) atomicAdd(sums + (int64_t)bk * d + j, x[c * d + j]);
  }
  if (lane == 0 && obj != 0.0) atomicAdd(objective, obj);
}
__global__ void k_km_update(const double* __restrict__ sums, const double* __restrict__ counts, int K, int d, double* __restrict__ cen) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * d) return;
  const double c = counts[i / d];
  if (c > 0.0) cen[i] = sums[i] / c;        // an empty cluster keeps its centre
}
__global__ void k_km_gather(const double* __restrict__ x, const int64_t* __restrict__ idx, int K, int d, double* __restrict__ cen) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * d) cen[i] = x[idx[i / d] * d + i % d];
}
}  // namespace

// x: host fp64 [n x d] (unit_rows != 0: the Z_cos transform of csc_harmony is applied first, so that the centres are
// those of the matrix harmony clusters); init_idx[K]: the cells whose rows seed the centres
extern "C" csc_status csc_kmeans(csc_ctx* ctx, const double* x, int64_t n, int32_t d, int32_t K, const int64_t* init_idx,
                                 int32_t max_iter, double tol, int32_t unit_rows, double* centers_out, int32_t* labels_out,
                                 double* objective_out, int32_t* iters_out) {
  if (!ctx || !x || !init_idx || !centers_out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, n >= 1 && d >= 1 && d <= HM_DMAX && K >= 1 && K <= 256 && K <= n, "csc_kmeans: bad sizes n=%lld d=%d K=%d",
              (long long)n, d, K);
  for (int k = 0; k < K; ++k) CSC_REQUIRE(ctx, init_idx[k] >= 0 && init_idx[k] < n, "csc_kmeans: init_idx[%d] out of range", k);
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_HARMONY);
  cudaStream_t s = ctx->stream;
  DevPtr dX, dXu, dCen, dIdx, dLab, dAcc;
  CSC_TRY(dev_alloc(ctx, (size_t)n * d * 8, &dX));
  CSC_TRY(dev_alloc(ctx, (size_t)K * d * 8, &dCen));
  CSC_TRY(dev_alloc(ctx, (size_t)K * 8, &dIdx));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &dLab));
  CSC_TRY(dev_alloc(ctx, ((size_t)K * d + K + 1) * 8, &dAcc));
  CSC_CUDA(ctx, cudaMemcpyAsync(dX->p, x, (size_t)n * d * 8, cudaMemcpyHostToDevice, s));
  CSC_CUDA(ctx, cudaMemcpyAsync(dIdx->p, init_idx, (size_t)K * 8, cudaMemcpyHostToDevice, s));
  double* X = (double*)dX->p;
  if (unit_rows) {
    CSC_TRY(dev_alloc(ctx, (size_t)n * d * 8, &dXu));
    k_hm_zcos_init<<<hm_blocks(n), 256, 0, s>>>(X, n, d, (double*)dXu->p);
    CSC_LAUNCHED(ctx);
    X = (double*)dXu->p;
  }
  double* cen = (double*)dCen->p;
  double* sums = (double*)dAcc->p;
  double* counts = sums + (size_t)K * d;
  double* obj = counts + K;
  k_km_gather<<<hm_blocks((int64_t)K * d), 256, 0, s>>>(X, (const int64_t*)dIdx->p, K, d, cen);
  CSC_LAUNCHED(ctx);
  const size_t smem = (size_t)K * d * 8;
  if (smem > 48 * 1024) CSC_CUDA(ctx, cudaFuncSetAttribute(k_km_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  double prev = INFINITY, cur = INFINITY;
  int it = 0;
  for (it = 1; it <= std::max(1, max_iter); ++it) {
    CSC_CUDA(ctx, cudaMemsetAsync(dAcc->p, 0, ((size_t)K * d + K + 1) * 8, s));
    k_km_assign<<<ctx->sm_count * 8, 256, smem, s>>>(X, cen, n, K, d, (int32_t*)dLab->p, sums, counts, obj);
    CSC_LAUNCHED(ctx);
    CSC_CUDA(ctx, cudaMemcpyAsync(&cur, obj, 8, cudaMemcpyDeviceToHost, s));
    CSC_CUDA(ctx, cudaStreamSynchronize(s));
    // Clustering.jl: converged when the objective (of the assignment to the current centres) changes by less than tol
    if (std::fabs(cur - prev) < tol) break;
    prev = cur;
    k_km_update<<<hm_blocks((int64_t)K * d), 256, 0, s>>>(sums, counts, K, d, cen);
    CSC_LAUNCHED(ctx);
  }
  CSC_TRY(d2h_copy(ctx, centers_out, cen, (size_t)K * d * 8));
  if (labels_out) CSC_TRY(d2h_copy(ctx, labels_out, dLab->p, (size_t)n * 4));
  if (objective_out) *objective_out = cur;
  if (iters_out) *iters_out = std::min(it, std::max(1, max_iter));
  return CSC_OK;
  CSC_GUARD_END(ctx)
}
