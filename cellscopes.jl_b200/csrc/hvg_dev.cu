// find_variable_genes after the per-gene moments (src/scrna/processing.jl:140-155) ON THE DEVICE: the same algorithm as
// launch_hvg_host (hvg.cu) -- Loess.jl 0.6.3 restated: kd-tree vertices, one local quadratic fit per vertex over the q
// nearest points (tricube weights, column-pivoted Householder QR), cubic Hermite blending, 10^pred, the ratio and the
// stable descending rank -- as a handful of small kernels.  The host version costs 3.1-3.7 ms per step at 33k genes
// whatever the number of GPUs (thread fork / join, five passes of a few hundred microseconds each), which at 8 GPUs was a
// tenth of the step; here it is two radix sorts of 33k keys, one CTA per vertex for the fits and three thread-per-gene
// kernels.  Same operation order per element as the host code and the oracle (oracle/cellscopes_oracle.py::loess_fit);
// the sums of the fits run in a different association (block reductions), which moves the fitted values by ~1e-15
// relative -- nine orders of magnitude inside the margin at the rank boundary of the benchmark matrices.
#include <cub/cub.cuh>

#include <cmath>
#include <limits>

#include "common.cuh"

namespace {

constexpr int HV_T = 512;            // threads of a vertex CTA
constexpr int HV_MAX_VERTS = 512;    // vertices the device path takes (span >= ~0.02); beyond: host path

// order-preserving integer image of a double; -0.0 and +0.0 get the same image (they compare equal)
__device__ __forceinline__ unsigned long long hv_image(double d) {
  d += 0.0;
  const unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// moments and their logarithms (processing.jl:140-146); genes without variance are counted (the caller rejects them)
__global__ void k_hvg_moments(const double* __restrict__ s1s2, int64_t n_genes, double n_cells, double* __restrict__ mu,
                              double* __restrict__ var, double* __restrict__ lx, double* __restrict__ ly,
                              unsigned long long* __restrict__ key, uint32_t* __restrict__ idx, int* __restrict__ bad) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genes) return;
  const double s1 = s1s2[g], s2 = s1s2[n_genes + g];
  const double m = s1 / n_cells;                                   // exact-moment rule, SURVEY A.3-7
  const double v = (s2 - s1 * s1 / n_cells) / (n_cells - 1.0);
  const bool keep = v > 0.0;                                       // filter(:variance => >(0.0)) processing.jl:143
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const double x = keep ? log10(m) : nan, y = keep ? log10(v) : nan;
  mu[g] = m;
  var[g] = v;
  lx[g] = x;
  ly[g] = y;
  key[g] = hv_image(x);
  idx[g] = (uint32_t)g;
  if (!keep) {
    atomicAdd(bad, 1);
    atomicMin(bad + 1, (int)g);
  }
}

__global__ void k_hvg_gather(const double* __restrict__ lx, const double* __restrict__ ly, const uint32_t* __restrict__ sidx,
                             int64_t n, double* __restrict__ xs, double* __restrict__ ys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = sidx[i];
  xs[i] = lx[s];
  ys[i] = ly[s];
}

// kd-tree vertices of Loess.jl (hvg.cu::kd_vertices): both ends plus the medians of the recursive halving down to
// `cutoff` points, sorted, duplicates removed.  One thread: an in-order walk of the recursion emits the medians in
// ascending order (the points are sorted), so "sort + unique" is a comparison with the previous value.
__global__ void k_hvg_vertices(const double* __restrict__ xs, int64_t n, int64_t cutoff, double* __restrict__ verts, int max_verts,
                               int* __restrict__ nv_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int nv = 0;
  bool overflow = false;
  auto emit = [&](double v) {
    if (nv > 0 && verts[nv - 1] == v) return;
    if (nv >= max_verts) {
      overflow = true;
      return;
    }
    verts[nv++] = v;
  };
  emit(xs[0]);
  // frames (lo, hi, stage): stage 0 = descend left first, 1 = emit the median and descend right
  int64_t f_lo[64], f_hi[64];
  int f_st[64];
  int sp = 0;
  f_lo[0] = 0;
  f_hi[0] = n;
  f_st[0] = 0;
  sp = 1;
  while (sp > 0 && !overflow) {
    const int64_t lo = f_lo[sp - 1], hi = f_hi[sp - 1];
    const int64_t m = hi - lo;
    if (m <= cutoff || xs[hi - 1] - xs[lo] <= 0.0) {
      --sp;
      continue;
    }
    const int64_t mid = m / 2;
    if (f_st[sp - 1] == 0) {
      f_st[sp - 1] = 1;
      if (sp >= 64) {
        overflow = true;
        break;
      }
      f_lo[sp] = lo;
      f_hi[sp] = lo + mid;
      f_st[sp] = 0;
      ++sp;
    } else {
      const double med = (m % 2 == 1) ? xs[lo + mid] : (xs[lo + mid - 1] + xs[lo + mid]) / 2;
      emit(med);
      // replace this frame by its right half
      f_lo[sp - 1] = lo + mid;
      f_hi[sp - 1] = hi;
      f_st[sp - 1] = 0;
    }
  }
  emit(xs[n - 1]);
  *nv_out = overflow ? -1 : nv;
}

// sum of NV values per thread over the CTA, result to every thread (fixed tree: deterministic)
template <int NV>
__device__ __forceinline__ void block_sum(double (&x)[NV], double* s_red /* [HV_T / 32][NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; ++j)
    for (int o = 16; o > 0; o >>= 1) x[j] += __shfl_xor_sync(0xffffffffu, x[j], o);
  __syncthreads();                                   // s_red may still be read from the previous reduction
  if (lane == 0)
    for (int j = 0; j < NV; ++j) s_red[warp * NV + j] = x[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double t = 0.0;
    for (int w = 0; w < HV_T / 32; ++w) t += s_red[w * NV + j];
    x[j] = t;
  }
}

// One CTA per vertex: the q nearest points (ties of the q-th distance towards the lower gene index, like the stable
// argsort of |x - v| in the oracle), the weighted system [w, w xc, w xc^2 | w y] in global scratch, Householder QR with
// column pivoting (Julia: qr(us, ColumnNorm()) \ vs), value and slope at the vertex.
__global__ void __launch_bounds__(HV_T)
k_hvg_fit(const double* __restrict__ xs, const double* __restrict__ ys, const uint32_t* __restrict__ sidx, int64_t n, int64_t q,
          const double* __restrict__ verts, const int* __restrict__ nv_ptr, double* __restrict__ scratch, double* __restrict__ val,
          double* __restrict__ slope) {
  __shared__ double s_red[(HV_T / 32) * 5];
  const int vi = blockIdx.x;
  if (vi >= *nv_ptr) return;
  const double v = verts[vi];
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  // every thread computes the same window (a few dozen binary-search steps over cached data)
  int64_t R0;
  {
    int64_t a = 0, b = n;                               // lower_bound(xs, v)
    while (a < b) {
      const int64_t m = (a + b) >> 1;
      if (xs[m] < v) a = m + 1;
      else b = m;
    }
    R0 = a;
  }
  const int64_t nL = R0, nR = n - R0;
  auto dl = [&](int64_t i) { return i >= 1 && i <= nL ? fabs(xs[R0 - i] - v) : (i < 1 ? -inf : inf); };       // i-th to the left
  auto dr = [&](int64_t j) { return j >= 1 && j <= nR ? fabs(xs[R0 + j - 1] - v) : (j < 1 ? -inf : inf); };   // j-th to the right
  int64_t lo_a = max((int64_t)0, q - nR), hi_a = min(q, nL);
  while (lo_a < hi_a) {
    const int64_t m = (lo_a + hi_a) / 2;
    if (dl(m + 1) < dr(q - m)) lo_a = m + 1;
    else hi_a = m;
  }
  const double dq = fmax(dl(lo_a), dr(q - lo_a));       // the q-th smallest distance
  auto count_l = [&](bool strict) {
    int64_t lo = 0, hi = nL;
    while (lo < hi) {
      const int64_t m = (lo + hi) / 2;
      const double d = dl(m + 1);
      if (strict ? d < dq : d <= dq) lo = m + 1;
      else hi = m;
    }
    return lo;
  };
  auto count_r = [&](bool strict) {
    int64_t lo = 0, hi = nR;
    while (lo < hi) {
      const int64_t m = (lo + hi) / 2;
      const double d = dr(m + 1);
      if (strict ? d < dq : d <= dq) lo = m + 1;
      else hi = m;
    }
    return lo;
  };
  const int64_t aL = count_l(true), aLe = count_l(false);
  const int64_t bR = count_r(true), bRe = count_r(false);
  const int64_t need = q - aL - bR;                     // points of the tied rim that are taken
  // the rim: left run = positions [R0 - aLe, R0 - aL), right run = [R0 + bR, R0 + bRe); equal x inside each, hence
  // ascending gene index (the sort was by (x, index)).  The `need` lowest gene indices of their union: merge path.
  const int64_t Lb = R0 - aLe, nLt = aLe - aL, Rb = R0 + bR, nRt = bRe - bR;
  int64_t tl;
  {
    int64_t lo = max((int64_t)0, need - nRt), hi = min(need, nLt);
    while (lo < hi) {
      const int64_t m = (lo + hi) / 2;
      if (sidx[Lb + m] < sidx[Rb + need - m - 1]) lo = m + 1;
      else hi = m;
    }
    tl = lo;
  }
  const int64_t n_inner = aL + bR;
  const double dmax = dq == 0.0 ? 1.0 : dq;
  double* a0 = scratch + (size_t)vi * 4 * (size_t)q;
  double* col[3] = {a0, a0 + q, a0 + 2 * q};
  double* b = a0 + 3 * q;
  for (int64_t r = threadIdx.x; r < q; r += HV_T) {
    int64_t pos;
    if (r < n_inner) pos = R0 - aL + r;
    else if (r - n_inner < tl) pos = Lb + (r - n_inner);
    else pos = Rb + (r - n_inner - tl);
    const double xc = xs[pos] - v;
    const double u = fabs(xc) / dmax;
    const double u3 = u * u * u;
    const double t = 1.0 - u3;
    const double w = sqrt(t * t * t);
    col[0][r] = w;
    col[1][r] = w * xc;
    col[2][r] = w * xc * xc;
    b[r] = ys[pos] * w;
  }
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double red[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int64_t i = k + threadIdx.x; i < q; i += HV_T)
      for (int j = k; j < 3; ++j) red[j] += col[j][i] * col[j][i];
    block_sum<5>(red, s_red);                           // also orders the writes above / of the previous step
    int p = k;
    double best = -1.0;
    for (int j = k; j < 3; ++j)
      if (red[j] > best) {
        best = red[j];
        p = j;
      }
    if (p != k) {
      double* tc = col[k];
      col[k] = col[p];
      col[p] = tc;
      const int tp = perm[k];
      perm[k] = perm[p];
      perm[p] = tp;
    }
    double alpha = sqrt(best);
    if (alpha == 0.0) continue;
    const double* x = col[k];
    if (x[k] > 0) alpha = -alpha;
    // v_i = x_i - alpha [i == k]; pass 1: v.v, v.col_j (j >= k), v.b
    double d5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};           // [0] v.v, [1..3] v.col_j, [4] v.b
    for (int64_t i = k + threadIdx.x; i < q; i += HV_T) {
      const double t = x[i] - (i == k ? alpha : 0.0);
      d5[0] += t * t;
      for (int j = k; j < 3; ++j) d5[1 + j] += t * col[j][i];
      d5[4] += t * b[i];
    }
    block_sum<5>(d5, s_red);
    const double vnorm2 = d5[0];
    if (vnorm2 == 0.0) continue;
    double sc[3] = {0.0, 0.0, 0.0};
    for (int j = k; j < 3; ++j) sc[j] = 2.0 * d5[1 + j] / vnorm2;
    const double sb = 2.0 * d5[4] / vnorm2;
    // pass 2: the reflections (each thread updates the rows it read in pass 1)
    for (int64_t i = k + threadIdx.x; i < q; i += HV_T) {
      const double t = x[i] - (i == k ? alpha : 0.0);
      for (int j = k; j < 3; ++j) col[j][i] -= sc[j] * t;
      b[i] -= sb * t;
    }
    __syncthreads();                                     // x[k] of the next step and the back substitution read other threads' rows
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double cp[3] = {0.0, 0.0, 0.0};
    for (int k = 2; k >= 0; --k) {
      if (k >= q) continue;                              // fewer rows than columns: the coefficient stays 0
      double t = b[k];
      for (int j = k + 1; j < 3; ++j) t -= col[j][k] * cp[j];
      const double d = col[k][k];
      cp[k] = d != 0.0 ? t / d : 0.0;
    }
    double coef[3];
    for (int k = 0; k < 3; ++k) coef[perm[k]] = cp[k];
    val[vi] = coef[0];
    slope[vi] = coef[1];
  }
}

// Loess.jl predict at every gene's own x (cubic Hermite between the enclosing vertices), 10^pred, the ratio, and the key
// of the final order (descending variance_standardized = ascending image of its negative)
__global__ void k_hvg_predict(const double* __restrict__ lx, const double* __restrict__ var, int64_t n, const double* __restrict__ verts,
                              const int* __restrict__ nv_ptr, const double* __restrict__ val, const double* __restrict__ slope,
                              double* __restrict__ ve, double* __restrict__ vs, unsigned long long* __restrict__ key,
                              uint32_t* __restrict__ idx) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int nv = *nv_ptr;
  const double z = lx[g];
  int a = 0, bnd = nv;                                   // lower_bound(verts, z)
  while (a < bnd) {
    const int m = (a + bnd) >> 1;
    if (verts[m] < z) a = m + 1;
    else bnd = m;
  }
  int hi = a;
  if (hi < 1) hi = 1;
  if (hi > nv - 1) hi = nv - 1;
  double pred;
  if (nv == 1) {
    pred = val[0];
  } else {
    const int lo = hi - 1;
    const double v1 = verts[lo], v2 = verts[hi];
    if (z == v1) {
      pred = val[lo];
    } else if (z == v2) {
      pred = val[hi];
    } else {
      const double h = v2 - v1;
      const double t = (z - v1) / h;
      const double omt = 1.0 - t;
      const double h00 = (1 + 2 * t) * omt * omt;
      const double h10 = t * omt * omt;
      const double h01 = t * t * (3 - 2 * t);
      const double h11 = t * t * (t - 1);
      pred = h00 * val[lo] + h10 * h * slope[lo] + h01 * val[hi] + h11 * h * slope[hi];
    }
  }
  const double e = pow(10.0, pred);
  const double s = var[g] / e;
  ve[g] = e;
  vs[g] = s;
  key[g] = hv_image(-s);
  idx[g] = (uint32_t)g;
}

__global__ void k_hvg_order64(const uint32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int64_t)in[i];
}

csc_status sort_pairs(csc_ctx* ctx, const unsigned long long* k_in, unsigned long long* k_out, const uint32_t* v_in, uint32_t* v_out,
                      int64_t n) {
  size_t tb = 0;
  CSC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, v_in, v_out, (int)n, 0, 64, ctx->stream));
  DevPtr tmp;
  CSC_TRY(dev_alloc(ctx, tb, &tmp));
  CSC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp->p, tb, k_in, k_out, v_in, v_out, (int)n, 0, 64, ctx->stream));
  ctx->launches += 9;
  return CSC_OK;
}

}  // namespace

// true: the device path applies to this problem (else the caller takes launch_hvg_host)
bool hvg_device_applies(int64_t n_genes, double span) {
  const char* e = getenv("CSC_HVG_HOST");                 // read per call: a test switches between the two versions
  const bool off = e && atoi(e) != 0;
  if (off || n_genes < 1024 || n_genes >= ((int64_t)1 << 31)) return false;
  // vertices: both ends + one median per internal node of the halving down to ceil(0.2 span n) points
  const double leaves = 2.0 / (0.2 * span);
  return leaves + 2.0 <= (double)HV_MAX_VERTS;
}

// d_s1s2: the exact moment table fp64[2 * n_genes] on the device.  Host outputs as launch_hvg_host.
csc_status launch_hvg_device(csc_ctx* ctx, const double* d_s1s2, int64_t n_genes, int64_t n_cells_total, double span, double* mean,
                             double* variance, double* var_exp, double* var_std, int64_t* order) {
  const int64_t n = n_genes;
  const double cell = 0.2;
  const int64_t q = std::max<int64_t>(1, (int64_t)std::floor(span * (double)n));
  const int64_t cutoff = (int64_t)std::ceil(cell * span * (double)n);
  DevPtr tab, keys, idx, misc, scratch;
  // fp64 tables: mu, var, lx, ly, xs, ys, ve, vs (8 x n); verts, val, slope (3 x HV_MAX_VERTS)
  CSC_TRY(dev_alloc(ctx, ((size_t)8 * n + 3 * HV_MAX_VERTS) * 8, &tab));
  CSC_TRY(dev_alloc(ctx, (size_t)2 * n * 8, &keys));
  CSC_TRY(dev_alloc(ctx, (size_t)2 * n * 4 + (size_t)n * 8, &idx));
  CSC_TRY(dev_alloc(ctx, 16, &misc));
  double* mu = (double*)tab->p;
  double *var = mu + n, *lx = var + n, *ly = lx + n, *xs = ly + n, *ys = xs + n, *ve = ys + n, *vs = ve + n;
  double *verts = vs + n, *val = verts + HV_MAX_VERTS, *slope = val + HV_MAX_VERTS;
  unsigned long long *k0 = (unsigned long long*)keys->p, *k1 = k0 + n;
  int64_t* ord64 = (int64_t*)idx->p;                      // 8-byte aligned part first
  uint32_t *i0 = (uint32_t*)(ord64 + n), *i1 = i0 + n;
  int* d_misc = (int*)misc->p;                            // [0] genes without variance, [1] the first of them, [2] vertices
  const int h_init[4] = {0, std::numeric_limits<int>::max(), 0, 0};
  CSC_CUDA(ctx, cudaMemcpyAsync(d_misc, h_init, 16, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned gb = (unsigned)ceil_div64(n, 256);
  k_hvg_moments<<<gb, 256, 0, ctx->stream>>>(d_s1s2, n, (double)n_cells_total, mu, var, lx, ly, k0, i0, d_misc);
  CSC_LAUNCHED(ctx);
  CSC_TRY(sort_pairs(ctx, k0, k1, i0, i1, n));            // i1 = the stable order by (x, index)
  k_hvg_gather<<<gb, 256, 0, ctx->stream>>>(lx, ly, i1, n, xs, ys);
  CSC_LAUNCHED(ctx);
  k_hvg_vertices<<<1, 32, 0, ctx->stream>>>(xs, n, cutoff, verts, HV_MAX_VERTS, d_misc + 2);
  CSC_LAUNCHED(ctx);
  // the fits: scratch for the widest possible vertex count
  const int max_nv = (int)std::min<double>((double)HV_MAX_VERTS, 2.0 / (cell * span) + 2.0);
  CSC_TRY(dev_alloc(ctx, (size_t)max_nv * 4 * (size_t)q * 8, &scratch));
  k_hvg_fit<<<max_nv, HV_T, 0, ctx->stream>>>(xs, ys, i1, n, q, verts, d_misc + 2, (double*)scratch->p, val, slope);
  CSC_LAUNCHED(ctx);
  k_hvg_predict<<<gb, 256, 0, ctx->stream>>>(lx, var, n, verts, d_misc + 2, val, slope, ve, vs, k0, i0);
  CSC_LAUNCHED(ctx);
  if (order) {
    CSC_TRY(sort_pairs(ctx, k0, k1, i0, i1, n));          // descending variance_standardized, ties by ascending gene index
    k_hvg_order64<<<gb, 256, 0, ctx->stream>>>(i1, n, ord64);
    CSC_LAUNCHED(ctx);
  }
  int h_misc[4] = {0, 0, 0, 0};
  CSC_CUDA(ctx, cudaMemcpyAsync(h_misc, d_misc, 16, cudaMemcpyDeviceToHost, ctx->stream));
  if (mean) CSC_CUDA(ctx, cudaMemcpyAsync(mean, mu, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (variance) CSC_CUDA(ctx, cudaMemcpyAsync(variance, var, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (var_exp) CSC_CUDA(ctx, cudaMemcpyAsync(var_exp, ve, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (var_std) CSC_CUDA(ctx, cudaMemcpyAsync(var_std, vs, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (order) CSC_CUDA(ctx, cudaMemcpyAsync(order, ord64, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // The reference requires every gene to have variance > 0 (see launch_hvg_host)
  if (h_misc[0] != 0)
    return csc_fail(ctx, CSC_ERR_ARG,
                    "find_variable_genes: %lld of %lld genes have variance <= 0 (first: gene %lld); the reference requires "
                    "every gene to vary (processing.jl:143,151) -- its object constructor drops all-zero genes first "
                    "(subset_matrix, utils.jl:313-320)",
                    (long long)h_misc[0], (long long)n_genes, (long long)h_misc[1]);
  if (h_misc[2] <= 0) return csc_fail(ctx, CSC_ERR_NUMERIC, "find_variable_genes: the Loess vertex table overflowed (span %.4g)", span);
  return CSC_OK;
}
