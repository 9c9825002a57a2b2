// run_pca (src/scrna/processing.jl:164-193; MultivariateStats fit(PCA; method=:svd)).
//
// The reference takes a dense LAPACK SVD of the centred H x N matrix.  Here the same
// decomposition is reached through the H x H Gram matrix (the only PCA
// formulation that never re-reads Z more than once and that is exact):
//   colsum = Z' 1,  G = Z' Z                       (one pass over the cell-major Z; sharded by cells,
//                                                    partials are summed over ranks by the host)
//   Gc = G - colsum colsum' / N                    (= (Z - 1 m')' (Z - 1 m'))
//   top-k eigenpairs of Gc in fp64 on the device   (block subspace iteration + Rayleigh-Ritz,
//                                                    small dense eigenproblems by parallel Jacobi)
//   principalvars = lambda / (N-1), tvar = trace(Gc)/(N-1), embedding = (Z - 1 m') U_k
#include <cmath>

#include "common.cuh"

// ------------------------------------------------------------------------------------------
// column sums of the cell-major dense matrix (fp64 accumulation)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_colsum(const float* __restrict__ z, int64_t n_rows, int32_t n_cols, int64_t rows_per_cta, double* __restrict__ out) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_cols) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(r0 + rows_per_cta, n_rows);
  double s = 0.0;
  const float* p = z + r0 * n_cols + h;
#pragma unroll 4
  for (int64_t r = r0; r < r1; ++r, p += n_cols) s += (double)__ldg(p);
  atomicAdd(out + h, s);
}

// ------------------------------------------------------------------------------------------
// Gram v1 (CUDA cores): G[a][b] += sum_c Z[c][a] Z[c][b] over a chunk of cells.
// 128x128 output tile per CTA, 8x8 per thread, fp32 FMA over the chunk, fp64 accumulation
// across chunks (atomics into the fp64 H x H table).  Only tile pairs ta <= tb are computed.
// ------------------------------------------------------------------------------------------
#define GR_BM 128
#define GR_BK 16
__global__ void __launch_bounds__(256)
k_gram_simt(const float* __restrict__ z, int64_t n_rows, int32_t H, int64_t rows_per_chunk, const int2* __restrict__ pairs,
            double* __restrict__ gram) {
  __shared__ __align__(16) float As[2][GR_BK][GR_BM];
  __shared__ __align__(16) float Bs[2][GR_BK][GR_BM];
  const int2 pr = pairs[blockIdx.x];
  const int a0 = pr.x * GR_BM, b0 = pr.y * GR_BM;
  const bool diag = (pr.x == pr.y);
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = min(r0 + rows_per_chunk, n_rows);
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const bool vec_ok = ((H & 3) == 0);
  // loader mapping: 16 rows x 32 float4 per tile; thread -> (row = t/32 [+8], col4 = t%32)
  const int lrow = t >> 5, lc4 = t & 31;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto load_tile = [&](int64_t row, int col0, float4& v) {
    v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < r1) {
      const int col = col0 + lc4 * 4;
      const float* p = z + row * (int64_t)H + col;
      if (vec_ok && col + 3 < H) {
        v = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        if (col < H) v.x = __ldg(p);
        if (col + 1 < H) v.y = __ldg(p + 1);
        if (col + 2 < H) v.z = __ldg(p + 2);
        if (col + 3 < H) v.w = __ldg(p + 3);
      }
    }
  };
  float4 ra[2], rb[2];
  const int64_t nk = ceil_div64(r1 - r0, GR_BK);
  if (nk <= 0) return;
  load_tile(r0 + lrow, a0, ra[0]);
  load_tile(r0 + lrow + 8, a0, ra[1]);
  if (!diag) {
    load_tile(r0 + lrow, b0, rb[0]);
    load_tile(r0 + lrow + 8, b0, rb[1]);
  }
  int buf = 0;
  *reinterpret_cast<float4*>(&As[0][lrow][lc4 * 4]) = ra[0];
  *reinterpret_cast<float4*>(&As[0][lrow + 8][lc4 * 4]) = ra[1];
  if (!diag) {
    *reinterpret_cast<float4*>(&Bs[0][lrow][lc4 * 4]) = rb[0];
    *reinterpret_cast<float4*>(&Bs[0][lrow + 8][lc4 * 4]) = rb[1];
  }
  __syncthreads();
  for (int64_t kb = 0; kb < nk; ++kb) {
    const int64_t rn = r0 + (kb + 1) * GR_BK;
    if (kb + 1 < nk) {
      load_tile(rn + lrow, a0, ra[0]);
      load_tile(rn + lrow + 8, a0, ra[1]);
      if (!diag) {
        load_tile(rn + lrow, b0, rb[0]);
        load_tile(rn + lrow + 8, b0, rb[1]);
      }
    }
    const float(*Ac)[GR_BM] = As[buf];
    const float(*Bc)[GR_BM] = diag ? As[buf] : Bs[buf];
#pragma unroll
    for (int kk = 0; kk < GR_BK; ++kk) {
      const float4 a_lo = *reinterpret_cast<const float4*>(&Ac[kk][ty * 4]);
      const float4 a_hi = *reinterpret_cast<const float4*>(&Ac[kk][64 + ty * 4]);
      const float4 b_lo = *reinterpret_cast<const float4*>(&Bc[kk][tx * 4]);
      const float4 b_hi = *reinterpret_cast<const float4*>(&Bc[kk][64 + tx * 4]);
      const float av[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
      const float bv[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      buf ^= 1;
      *reinterpret_cast<float4*>(&As[buf][lrow][lc4 * 4]) = ra[0];
      *reinterpret_cast<float4*>(&As[buf][lrow + 8][lc4 * 4]) = ra[1];
      if (!diag) {
        *reinterpret_cast<float4*>(&Bs[buf][lrow][lc4 * 4]) = rb[0];
        *reinterpret_cast<float4*>(&Bs[buf][lrow + 8][lc4 * 4]) = rb[1];
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int a = a0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (a >= H) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (b >= H) continue;
      const double v = (double)acc[i][j];
      atomicAdd(&gram[(int64_t)a * H + b], v);
      if (!diag) atomicAdd(&gram[(int64_t)b * H + a], v);
    }
  }
}

extern "C" csc_status csc_pca_partial(csc_ctx* ctx, const csc_dense* scaled, csc_vec* colsum, csc_vec* gram) {
  if (!ctx || !scaled || !colsum || !gram) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  const int64_t H = scaled->n_cols, N = scaled->n_rows;
  CSC_REQUIRE(ctx, H > 0 && H < 65536, "csc_pca_partial: H=%lld unsupported", (long long)H);
  CSC_REQUIRE(ctx, colsum->dtype == CSC_DT_F64 && colsum->n >= H, "csc_pca_partial: colsum must be fp64[H]");
  CSC_REQUIRE(ctx, gram->dtype == CSC_DT_F64 && gram->n >= H * H, "csc_pca_partial: gram must be fp64[H*H]");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_PCA_GRAM);
  if (N == 0) return CSC_OK;
  {
    const int64_t rows_per_cta = 2048;
    dim3 grid((unsigned)ceil_div64(H, 256), (unsigned)ceil_div64(N, rows_per_cta));
    k_colsum<<<grid, 256, 0, ctx->stream>>>(scaled->p(), N, (int32_t)H, rows_per_cta, colsum->as<double>());
    CSC_LAUNCHED(ctx);
  }
  {
    const int nt = (int)ceil_div64(H, GR_BM);
    std::vector<int2> pairs;
    for (int a = 0; a < nt; ++a)
      for (int b = a; b < nt; ++b) pairs.push_back(make_int2(a, b));
    DevPtr d_pairs;
    CSC_TRY(dev_alloc(ctx, pairs.size() * sizeof(int2), &d_pairs));
    CSC_CUDA(ctx, cudaMemcpyAsync(d_pairs->p, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t rows_per_chunk = 4096;
    dim3 grid((unsigned)pairs.size(), (unsigned)ceil_div64(N, rows_per_chunk));
    k_gram_simt<<<grid, 256, 0, ctx->stream>>>(scaled->p(), N, (int32_t)H, rows_per_chunk, (const int2*)d_pairs->p,
                                                gram->as<double>());
    CSC_LAUNCHED(ctx);
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------
// fp64 eigensolver pieces
// ------------------------------------------------------------------------------------------
// Gc = sym(G) - s s' / N
__global__ void k_center_gram(const double* __restrict__ g, const double* __restrict__ s, int32_t H, double inv_n,
                              double* __restrict__ gc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)H * H) return;
  const int a = (int)(idx / H), b = (int)(idx % H);
  gc[idx] = 0.5 * (g[idx] + g[(int64_t)b * H + a]) - s[a] * s[b] * inv_n;
}

// C[M x N] = op(A) * B, row-major.  TA = false: A is [M x K]; TA = true: A is [K x M] (C = A' B).
// 64x64 tile, 256 threads, 4x4 per thread.
template <bool TA>
__global__ void __launch_bounds__(256)
k_dgemm(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int M, int N, int K, int lda,
        int ldb, int ldc) {
  __shared__ double As[16][64 + 1];
  __shared__ double Bs[16][64 + 1];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  double acc[4][4] = {{0}};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = t; i < 16 * 64; i += 256) {
      int kk, mm;
      if (TA) {
        kk = i / 64;
        mm = i % 64;  // A[k][m] contiguous in m
        const int gk = k0 + kk, gm = m0 + mm;
        As[kk][mm] = (gk < K && gm < M) ? A[(int64_t)gk * lda + gm] : 0.0;
      } else {
        mm = i / 16;
        kk = i % 16;  // A[m][k] contiguous in k
        const int gk = k0 + kk, gm = m0 + mm;
        As[kk][mm] = (gk < K && gm < M) ? A[(int64_t)gm * lda + gk] : 0.0;
      }
      const int bk = i / 64, bn = i % 64;
      const int gk2 = k0 + bk, gn = n0 + bn;
      Bs[bk][bn] = (gk2 < K && gn < N) ? B[(int64_t)gk2 * ldb + gn] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < N) C[(int64_t)m * ldc + n] = acc[i][j];
    }
  }
}

static csc_status dgemm(csc_ctx* ctx, bool ta, const double* A, const double* B, double* C, int M, int N, int K, int lda,
                        int ldb, int ldc) {
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
  if (ta) k_dgemm<true><<<grid, 256, 0, ctx->stream>>>(A, B, C, M, N, K, lda, ldb, ldc);
  else k_dgemm<false><<<grid, 256, 0, ctx->stream>>>(A, B, C, M, N, K, lda, ldb, ldc);
  CSC_LAUNCHED(ctx);
  return CSC_OK;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// deterministic Gaussian start block
__global__ void k_init_block(double* __restrict__ q, int64_t n, uint64_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t h = mix64(mix64(seed) ^ (uint64_t)i);
  const double u1 = ((double)(h >> 32) + 0.5) * (1.0 / 4294967296.0);
  const double u2 = ((double)(h & 0xffffffffull) + 0.5) * (1.0 / 4294967296.0);
  q[i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// upper Cholesky S = R'R of a b x b matrix (single CTA), in place; rinv_diag[j] = 1/R_jj (0 if the pivot
// vanished: that direction is dropped instead of poisoning the block)
__global__ void __launch_bounds__(256)
k_cholesky(double* __restrict__ S, int b, double* __restrict__ rinv_diag) {
  __shared__ double s_piv;
  __shared__ double s_scale;
  const int t = threadIdx.x;
  if (t == 0) {
    double mx = 0.0;
    for (int i = 0; i < b; ++i) mx = fmax(mx, fabs(S[(int64_t)i * b + i]));
    s_scale = mx;
  }
  __syncthreads();
  const double tiny = s_scale * 1e-14;
  for (int j = 0; j < b; ++j) {
    if (t == 0) {
      const double d = S[(int64_t)j * b + j];
      if (d > tiny) {
        const double r = sqrt(d);
        S[(int64_t)j * b + j] = r;
        s_piv = 1.0 / r;
      } else {
        S[(int64_t)j * b + j] = 0.0;
        s_piv = 0.0;
      }
      rinv_diag[j] = s_piv;
    }
    __syncthreads();
    const double inv = s_piv;
    for (int c = j + 1 + t; c < b; c += blockDim.x) S[(int64_t)j * b + c] *= inv;  // row j of R
    __syncthreads();
    // trailing update: S[r][c] -= R[j][r] R[j][c] for r,c > j (upper part)
    const int m = b - j - 1;
    for (int idx = t; idx < m * m; idx += blockDim.x) {
      const int r = j + 1 + idx / m, c = j + 1 + idx % m;
      if (c >= r) S[(int64_t)r * b + c] -= S[(int64_t)j * b + r] * S[(int64_t)j * b + c];
    }
    __syncthreads();
  }
}

// Q = Y R^{-1}: one thread per row of Y (forward substitution over the columns), R upper b x b
__global__ void __launch_bounds__(128)
k_trsm_right(const double* __restrict__ Y, const double* __restrict__ R, const double* __restrict__ rinv_diag, int H, int b,
             double* __restrict__ Q) {
  extern __shared__ double sR[];  // b*b
  for (int i = threadIdx.x; i < b * b; i += blockDim.x) sR[i] = R[i];
  __syncthreads();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= H) return;
  const double* y = Y + (int64_t)row * b;
  double* q = Q + (int64_t)row * b;
  for (int j = 0; j < b; ++j) {
    double s = y[j];
    for (int i = 0; i < j; ++i) s -= q[i] * sR[i * b + j];
    q[j] = s * rinv_diag[j];
  }
}

// Parallel cyclic Jacobi for a symmetric n x n matrix (single CTA): A -> diag, V = eigenvectors (columns).
// Round-robin ordering: n/2 disjoint rotations per round applied simultaneously.
// On exit eval[i] (descending) and Vs = V with columns permuted accordingly.
__global__ void __launch_bounds__(1024)
k_jacobi(double* __restrict__ A, double* __restrict__ V, int n, double* __restrict__ eval, double* __restrict__ Vs,
         int max_sweeps, int* __restrict__ sweeps_out) {
  extern __shared__ double sh[];
  const int m = (n + 1) & ~1;
  double* cs = sh;            // [m/2] cos
  double* sn = sh + m / 2;    // [m/2] sin
  int* pp = (int*)(sh + m);   // [m/2]
  int* qq = pp + m / 2;       // [m/2]
  __shared__ double s_off, s_diag;
  __shared__ int s_done;
  const int t = threadIdx.x, nt = blockDim.x;
  for (int i = t; i < n * n; i += nt) V[i] = ((i / n) == (i % n)) ? 1.0 : 0.0;
  __syncthreads();
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    // convergence: off-diagonal Frobenius norm against the diagonal
    if (t == 0) {
      s_off = 0.0;
      s_diag = 0.0;
    }
    __syncthreads();
    double off = 0.0, dg = 0.0;
    for (int i = t; i < n * n; i += nt) {
      const double v = A[i];
      if ((i / n) == (i % n)) dg += v * v;
      else off += v * v;
    }
    off = warp_sum(off);
    dg = warp_sum(dg);
    if ((t & 31) == 0) {
      atomicAdd(&s_off, off);
      atomicAdd(&s_diag, dg);
    }
    __syncthreads();
    if (t == 0) s_done = (s_off <= 1e-26 * (s_diag + s_off)) || (s_off == 0.0);
    __syncthreads();
    if (s_done) break;
    for (int r = 0; r < m - 1; ++r) {
      // pairs of this round
      for (int i = t; i < m / 2; i += nt) {
        int p, q;
        if (i == 0) {
          p = m - 1;
          q = r;
        } else {
          p = (r + i) % (m - 1);
          q = (r - i + (m - 1)) % (m - 1);
        }
        if (p > q) {
          const int tmp = p;
          p = q;
          q = tmp;
        }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[(int64_t)p * n + q];
          if (apq != 0.0) {
            const double app = A[(int64_t)p * n + p], aqq = A[(int64_t)q * n + q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(tt * tt + 1.0);
            s = tt * c;
          }
        }
        cs[i] = c;
        sn[i] = s;
        pp[i] = p;
        qq[i] = q;
      }
      __syncthreads();
      // rows: A <- J' A
      for (int idx = t; idx < (m / 2) * n; idx += nt) {
        const int i = idx / n, j = idx % n;
        const int p = pp[i], q = qq[i];
        if (q >= n) continue;
        const double c = cs[i], s = sn[i];
        const double apj = A[(int64_t)p * n + j], aqj = A[(int64_t)q * n + j];
        A[(int64_t)p * n + j] = c * apj - s * aqj;
        A[(int64_t)q * n + j] = s * apj + c * aqj;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int idx = t; idx < (m / 2) * n; idx += nt) {
        const int i = idx % (m / 2), row = idx / (m / 2);
        const int p = pp[i], q = qq[i];
        if (q >= n) continue;
        const double c = cs[i], s = sn[i];
        const double aip = A[(int64_t)row * n + p], aiq = A[(int64_t)row * n + q];
        A[(int64_t)row * n + p] = c * aip - s * aiq;
        A[(int64_t)row * n + q] = s * aip + c * aiq;
        const double vip = V[(int64_t)row * n + p], viq = V[(int64_t)row * n + q];
        V[(int64_t)row * n + p] = c * vip - s * viq;
        V[(int64_t)row * n + q] = s * vip + c * viq;
      }
      __syncthreads();
    }
  }
  if (t == 0 && sweeps_out) *sweeps_out = sweep;
  // sort eigenvalues descending (rank by counting; ties by index), permute the columns of V
  for (int i = t; i < n; i += nt) {
    const double di = A[(int64_t)i * n + i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double dj = A[(int64_t)j * n + j];
      rank += (dj > di) || (dj == di && j < i);
    }
    eval[rank] = di;
    for (int row = 0; row < n; ++row) Vs[(int64_t)row * n + rank] = V[(int64_t)row * n + i];
  }
}

// residual norms ||Yn_i - theta_i Qn_i|| for the first k columns
__global__ void k_resid(const double* __restrict__ Yn, const double* __restrict__ Qn, const double* __restrict__ theta, int H,
                        int b, int k, double* __restrict__ out) {
  const int i = blockIdx.x;
  if (i >= k) return;
  double s = 0.0;
  for (int r = threadIdx.x; r < H; r += blockDim.x) {
    const double d = Yn[(int64_t)r * b + i] - theta[i] * Qn[(int64_t)r * b + i];
    s += d * d;
  }
  __shared__ double red[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[i] = sqrt(v);
  }
}

static csc_status jacobi(csc_ctx* ctx, double* A, double* V, int n, double* eval, double* Vs, int* d_sweeps) {
  const int m = (n + 1) & ~1;
  size_t smem = (size_t)m * 8 + (size_t)m * 4 + 64;
  k_jacobi<<<1, 1024, smem, ctx->stream>>>(A, V, n, eval, Vs, 60, d_sweeps);
  CSC_LAUNCHED(ctx);
  return CSC_OK;
}

static csc_status chol_qr(csc_ctx* ctx, const double* Y, double* Q, double* S, double* rinv, int H, int b) {
  CSC_TRY(dgemm(ctx, true, Y, Y, S, b, b, H, b, b, b));  // S = Y'Y
  k_cholesky<<<1, 256, 0, ctx->stream>>>(S, b, rinv);
  CSC_LAUNCHED(ctx);
  size_t smem = (size_t)b * b * 8;
  if (smem > 48 * 1024) CSC_CUDA(ctx, cudaFuncSetAttribute(k_trsm_right, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_trsm_right<<<(unsigned)((H + 127) / 128), 128, smem, ctx->stream>>>(Y, S, rinv, H, b, Q);
  CSC_LAUNCHED(ctx);
  return CSC_OK;
}

extern "C" csc_status csc_pca_solve(csc_ctx* ctx, const csc_vec* colsum, const csc_vec* gram, int64_t n_feat,
                                    int64_t n_cells_total, int32_t maxoutdim, double tol, int32_t max_iter, csc_pca** out,
                                    int32_t* iters_out, double* resid_out) {
  if (!ctx || !colsum || !gram || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  const int H = (int)n_feat;
  CSC_REQUIRE(ctx, n_feat > 0 && n_feat < 65536, "csc_pca_solve: bad n_feat");
  CSC_REQUIRE(ctx, colsum->dtype == CSC_DT_F64 && colsum->n >= n_feat, "csc_pca_solve: colsum must be fp64[H]");
  CSC_REQUIRE(ctx, gram->dtype == CSC_DT_F64 && gram->n >= n_feat * n_feat, "csc_pca_solve: gram must be fp64[H*H]");
  CSC_REQUIRE(ctx, n_cells_total >= 2, "csc_pca_solve: need at least 2 cells");
  CSC_REQUIRE(ctx, maxoutdim >= 1, "csc_pca_solve: maxoutdim must be >= 1");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_PCA_SOLVE);
  if (tol <= 0) tol = 1e-9;
  if (max_iter <= 0) max_iter = 500;
  // MultivariateStats: outdim <= min(H, N) (rank is at most N-1 after centring)
  const int k = (int)std::min<int64_t>(maxoutdim, std::min<int64_t>(n_feat, n_cells_total));
  const double N = (double)n_cells_total;

  DevPtr gc;
  CSC_TRY(dev_alloc(ctx, (size_t)H * H * 8, &gc));
  double* Gc = (double*)gc->p;
  k_center_gram<<<(unsigned)ceil_div64((int64_t)H * H, 256), 256, 0, ctx->stream>>>(gram->as<double>(), colsum->as<double>(),
                                                                                      H, 1.0 / N, Gc);
  CSC_LAUNCHED(ctx);
  std::vector<double> diag((size_t)H), hsum((size_t)H);
  CSC_CUDA(ctx, cudaMemcpy2DAsync(diag.data(), 8, Gc, (size_t)(H + 1) * 8, 8, H, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaMemcpyAsync(hsum.data(), colsum->mem->p, (size_t)H * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double trace = 0.0;
  for (double d : diag) trace += d;

  std::vector<double> evals((size_t)k), U((size_t)H * k);
  int iters = 0;
  double resid = 0.0;
  const int direct_max = 160;
  if (H <= direct_max) {
    // small problem: full Jacobi eigendecomposition of Gc
    DevPtr v, vs, ev, sw;
    CSC_TRY(dev_alloc(ctx, (size_t)H * H * 8, &v));
    CSC_TRY(dev_alloc(ctx, (size_t)H * H * 8, &vs));
    CSC_TRY(dev_alloc(ctx, (size_t)H * 8, &ev));
    CSC_TRY(dev_alloc(ctx, 4, &sw));
    CSC_TRY(jacobi(ctx, Gc, (double*)v->p, H, (double*)ev->p, (double*)vs->p, (int*)sw->p));
    std::vector<double> hv((size_t)H * H), he((size_t)H);
    CSC_CUDA(ctx, cudaMemcpyAsync(hv.data(), vs->p, hv.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaMemcpyAsync(he.data(), ev->p, he.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < k; ++j) {
      evals[j] = he[j];
      for (int r = 0; r < H; ++r) U[(size_t)j * H + r] = hv[(size_t)r * H + j];
    }
  } else {
    int b = std::min(H, std::max(64, ((2 * k + 28 + 31) / 32) * 32));
    b = std::min(b, direct_max);
    CSC_REQUIRE(ctx, k <= b, "csc_pca_solve: maxoutdim=%d too large for the block solver (max %d)", k, b);
    DevPtr q, y, qn, yn, s, w, ws, bb, th, rinv, res, sw;
    CSC_TRY(dev_alloc(ctx, (size_t)H * b * 8, &q));
    CSC_TRY(dev_alloc(ctx, (size_t)H * b * 8, &y));
    CSC_TRY(dev_alloc(ctx, (size_t)H * b * 8, &qn));
    CSC_TRY(dev_alloc(ctx, (size_t)H * b * 8, &yn));
    CSC_TRY(dev_alloc(ctx, (size_t)b * b * 8, &s));
    CSC_TRY(dev_alloc(ctx, (size_t)b * b * 8, &w));
    CSC_TRY(dev_alloc(ctx, (size_t)b * b * 8, &ws));
    CSC_TRY(dev_alloc(ctx, (size_t)b * b * 8, &bb));
    CSC_TRY(dev_alloc(ctx, (size_t)b * 8, &th));
    CSC_TRY(dev_alloc(ctx, (size_t)b * 8, &rinv));
    CSC_TRY(dev_alloc(ctx, (size_t)b * 8, &res));
    CSC_TRY(dev_alloc(ctx, 4, &sw));
    double *Q = (double*)q->p, *Y = (double*)y->p, *Qn = (double*)qn->p, *Yn = (double*)yn->p;
    k_init_block<<<(unsigned)ceil_div64((int64_t)H * b, 256), 256, 0, ctx->stream>>>(Y, (int64_t)H * b, 42);
    CSC_LAUNCHED(ctx);
    CSC_TRY(chol_qr(ctx, Y, Q, (double*)s->p, (double*)rinv->p, H, b));
    std::vector<double> hres((size_t)k), hth((size_t)b);
    int next_check = 8;
    bool converged = false;
    for (iters = 1; iters <= max_iter; ++iters) {
      CSC_TRY(dgemm(ctx, false, Gc, Q, Y, H, b, H, H, b, b));  // Y = Gc Q
      if (iters == next_check || iters == max_iter) {
        CSC_TRY(dgemm(ctx, true, Q, Y, (double*)bb->p, b, b, H, b, b, b));  // B = Q'Y
        CSC_TRY(jacobi(ctx, (double*)bb->p, (double*)w->p, b, (double*)th->p, (double*)ws->p, (int*)sw->p));
        CSC_TRY(dgemm(ctx, false, Q, (double*)ws->p, Qn, H, b, b, b, b, b));  // Qn = Q W
        CSC_TRY(dgemm(ctx, false, Y, (double*)ws->p, Yn, H, b, b, b, b, b));  // Yn = Y W = Gc Qn
        k_resid<<<k, 256, 0, ctx->stream>>>(Yn, Qn, (double*)th->p, H, b, k, (double*)res->p);
        CSC_LAUNCHED(ctx);
        CSC_CUDA(ctx, cudaMemcpyAsync(hres.data(), res->p, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CSC_CUDA(ctx, cudaMemcpyAsync(hth.data(), th->p, (size_t)b * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        resid = 0.0;
        for (int i = 0; i < k; ++i) resid = std::max(resid, hres[i]);
        resid /= std::max(hth[0], 1e-300);
        std::swap(Q, Qn);
        std::swap(Y, Yn);
        if (resid <= tol || iters == max_iter) {
          converged = resid <= tol;
          break;
        }
        next_check = iters + std::max(8, iters / 4);
      }
      // re-orthonormalise the block: CholeskyQR2
      CSC_TRY(chol_qr(ctx, Y, Qn, (double*)s->p, (double*)rinv->p, H, b));
      CSC_TRY(chol_qr(ctx, Qn, Q, (double*)s->p, (double*)rinv->p, H, b));
    }
    (void)converged;
    std::vector<double> hq((size_t)H * b);
    CSC_CUDA(ctx, cudaMemcpyAsync(hq.data(), Q, hq.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < k; ++j) {
      evals[j] = hth[j];
      for (int r = 0; r < H; ++r) U[(size_t)j * H + r] = hq[(size_t)r * b + j];
    }
  }
  // sign convention: the largest-magnitude component of every loading vector is positive
  for (int j = 0; j < k; ++j) {
    double best = 0.0;
    for (int r = 0; r < H; ++r)
      if (std::fabs(U[(size_t)j * H + r]) > std::fabs(best)) best = U[(size_t)j * H + r];
    if (best < 0.0)
      for (int r = 0; r < H; ++r) U[(size_t)j * H + r] = -U[(size_t)j * H + r];
  }
  std::unique_ptr<csc_pca> p(new csc_pca());
  p->ctx = ctx;
  p->n_feat = n_feat;
  p->k = k;
  p->tvar = trace / (N - 1.0);
  p->principalvars.resize((size_t)k);
  for (int j = 0; j < k; ++j) p->principalvars[j] = std::max(evals[j], 0.0) / (N - 1.0);
  p->loadings = U;
  p->mean.resize((size_t)H);
  for (int r = 0; r < H; ++r) p->mean[r] = hsum[r] / N;
  p->kpad = ((k + 15) / 16) * 16;
  std::vector<float> uf((size_t)H * p->kpad, 0.f), pm((size_t)p->kpad, 0.f);
  for (int j = 0; j < k; ++j) {
    double acc = 0.0;
    for (int r = 0; r < H; ++r) {
      uf[(size_t)r * p->kpad + j] = (float)U[(size_t)j * H + r];
      acc += p->mean[r] * U[(size_t)j * H + r];
    }
    pm[j] = (float)acc;
  }
  CSC_TRY(dev_alloc(ctx, uf.size() * 4, &p->d_loadings_f32));
  CSC_TRY(dev_alloc(ctx, pm.size() * 4, &p->d_proj_mean));
  CSC_CUDA(ctx, cudaMemcpyAsync(p->d_loadings_f32->p, uf.data(), uf.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CSC_CUDA(ctx, cudaMemcpyAsync(p->d_proj_mean->p, pm.data(), pm.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (iters_out) *iters_out = iters;
  if (resid_out) *resid_out = resid;
  *out = p.release();
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_pca_info(const csc_pca* p, int64_t* n_feat, int32_t* k, double* principalvars, double* tvar,
                                   double* loadings, double* mean) {
  if (!p) return CSC_ERR_ARG;
  if (n_feat) *n_feat = p->n_feat;
  if (k) *k = p->k;
  if (principalvars) std::copy(p->principalvars.begin(), p->principalvars.end(), principalvars);
  if (tvar) *tvar = p->tvar;
  if (loadings) std::copy(p->loadings.begin(), p->loadings.end(), loadings);
  if (mean) std::copy(p->mean.begin(), p->mean.end(), mean);
  return CSC_OK;
}

extern "C" csc_status csc_pca_free(csc_pca* p) {
  delete p;
  return CSC_OK;
}

// ------------------------------------------------------------------------------------------
// embedding = (Z - 1 m') U_k   (v1, CUDA cores): 64 cells x KP components per CTA
// ------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256)
k_transform(const float* __restrict__ z, int64_t n_rows, int32_t H, const float* __restrict__ U /*[H x KP]*/,
            const float* __restrict__ pm, float* __restrict__ out, int32_t k) {
  constexpr int BK = 32, BM = 64, NJ = KP / 16;
  __shared__ float Zs[BM][BK + 1];
  __shared__ float Us[BK][KP];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  float acc[4][NJ];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
  for (int h0 = 0; h0 < H; h0 += BK) {
    for (int i = t; i < BM * BK; i += 256) {
      const int r = i / BK, c = i % BK;
      const int64_t gr = row0 + r;
      const int gc = h0 + c;
      Zs[r][c] = (gr < n_rows && gc < H) ? __ldcs(z + gr * H + gc) : 0.f;
    }
    for (int i = t; i < BK * KP; i += 256) {
      const int r = i / KP, c = i % KP;
      Us[r][c] = (h0 + r < H) ? __ldg(U + (int64_t)(h0 + r) * KP + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < BK; ++kk) {
      float zv[4], uv[NJ];
#pragma unroll
      for (int i = 0; i < 4; ++i) zv[i] = Zs[ty * 4 + i][kk];
#pragma unroll
      for (int j = 0; j < NJ; ++j) uv[j] = Us[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(zv[i], uv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gr = row0 + ty * 4 + i;
    if (gr >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = tx + 16 * j;
      if (c < k) out[gr * k + c] = acc[i][j] - pm[c];
    }
  }
}

extern "C" csc_status csc_pca_transform(csc_ctx* ctx, const csc_dense* scaled, const csc_pca* p, csc_dense** out) {
  if (!ctx || !scaled || !p || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, scaled->n_cols == p->n_feat, "csc_pca_transform: matrix has %lld features, model %lld",
              (long long)scaled->n_cols, (long long)p->n_feat);
  CSC_REQUIRE(ctx, p->kpad <= 128, "csc_pca_transform: more than 128 components is not supported");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_PCA_TRANSFORM);
  csc_dense* e = nullptr;
  CSC_TRY(csc_dense_alloc(ctx, scaled->n_rows, p->k, &e));
  if (scaled->n_rows > 0) {
    const unsigned grid = (unsigned)ceil_div64(scaled->n_rows, 64);
    const float* U = (const float*)p->d_loadings_f32->p;
    const float* pm = (const float*)p->d_proj_mean->p;
    // the padded loading matrix has leading dimension kpad; pick the template that matches it
#define TR_CASE(KP)                                                                                               \
  case KP:                                                                                                        \
    k_transform<KP><<<grid, 256, 0, ctx->stream>>>(scaled->p(), scaled->n_rows, (int32_t)scaled->n_cols, U, pm, e->p(), p->k); \
    break;
    switch (p->kpad) {
      TR_CASE(16) TR_CASE(32) TR_CASE(48) TR_CASE(64) TR_CASE(80) TR_CASE(96) TR_CASE(112) TR_CASE(128)
      default: break;
    }
#undef TR_CASE
    ctx->launches++;
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    if (err != cudaSuccess) {
      csc_dense_free(e);
      return csc_fail(ctx, CSC_ERR_CUDA, "csc_pca_transform: %s", cudaGetErrorString(err));
    }
  }
  *out = e;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}
