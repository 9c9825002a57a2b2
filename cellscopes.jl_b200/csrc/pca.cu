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
#include <string>
#include <cmath>
#include <cuda_pipeline.h>

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

csc_status gram_tc(csc_ctx* ctx, const float* z, int64_t n_rows, int H, double* colsum, double* gram);

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
  // v2: tcgen05 3xTF32 (gram_tc.cu).  CSC_GRAM_IMPL=simt selects the CUDA-core kernel (kept for A/B runs).
  const char* impl = getenv("CSC_GRAM_IMPL");
  if (!(impl && strcmp(impl, "simt") == 0))
    return gram_tc(ctx, scaled->p(), N, (int)H, colsum->as<double>(), gram->as<double>());
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

// C[M x N] = (alpha * op(A) * B + g1 * X1 + g2 * X2) * colscale, row-major, fp64.
//   TA = false: A is [M x K]; TA = true: A is [K x M] (C = A' B).  X1 / X2 / colscale are optional ([M x N],
//   leading dimension ldc; X2 may alias C: every element is read and written by the same thread).
//   blockIdx.z batches over K-chunks of `kchunk` rows (A, B advance by kchunk rows, C by M*ldc): the small
//   Gram products of the block solver are split this way and summed by k_reduce_parts, which keeps them
//   deterministic and fills the machine (a b x b output alone is 4 CTAs).
//   BM x 64 tile, 256 threads, (BM/16) x 4 per thread.
template <bool TA, int BM>
__global__ void __launch_bounds__(256)
k_dgemm(const double* __restrict__ A, const double* __restrict__ B, double* C, int M, int N, int K, int lda, int ldb, int ldc,
        double alpha, double g1, const double* __restrict__ X1, double g2, const double* X2,
        const double* __restrict__ colscale, int kchunk) {
  constexpr int TM = BM / 16;
  __shared__ double As[16][BM + 1];
  __shared__ double Bs[16][64 + 1];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * 64;
  int kbeg = 0, kend = K;
  if (kchunk > 0) {
    kbeg = blockIdx.z * kchunk;
    kend = min(K, kbeg + kchunk);
    C += (int64_t)blockIdx.z * M * ldc;
  }
  double acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
    for (int i = t; i < 16 * BM; i += 256) {
      int kk, mm;
      if (TA) {
        kk = i / BM;
        mm = i % BM;  // A[k][m] contiguous in m
      } else {
        mm = i / 16;
        kk = i % 16;  // A[m][k] contiguous in k
      }
      const int gk = k0 + kk, gm = m0 + mm;
      double v = 0.0;
      if (gk < kend && gm < M) v = TA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
      As[kk][mm] = v;
    }
    for (int i = t; i < 16 * 64; i += 256) {
      const int bk = i / 64, bn = i % 64;
      const int gk = k0 + bk, gn = n0 + bn;
      Bs[bk][bn] = (gk < kend && gn < N) ? B[(int64_t)gk * ldb + gn] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[TM], b[4];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      const int64_t o = (int64_t)m * ldc + n;
      double v = alpha * acc[i][j];
      if (X1) v = fma(g1, X1[o], v);
      if (X2) v = fma(g2, X2[o], v);
      if (colscale) v *= colscale[n];
      C[o] = v;
    }
  }
}

// C[M x N] = alpha * A[M x K] * B[K x N] + g1 * X1 + g2 * X2 for the tall products of the block solver
// (Gc Q, Y R^-1): 32 x 64 tile, 128 threads, 4 x 4 per thread, 16-byte global and shared accesses, the next
// K-tile prefetched into registers while the current one is multiplied out of double-buffered shared memory
// (one barrier per K-step).  Needs even lda/ldb/ldc/K/N and 16-byte aligned bases; dgemm() checks.
__global__ void __launch_bounds__(128)
k_dgemm_nn(const double* __restrict__ A, const double* __restrict__ B, double* C, int M, int N, int K, int lda, int ldb,
           int ldc, double alpha, double g1, const double* __restrict__ X1, double g2, const double* X2, int kchunk,
           double* __restrict__ P) {
  __shared__ __align__(16) double As[2][16][32 + 2];   // [kk][m]
  __shared__ __align__(16) double Bs[2][16][64 + 2];   // [kk][n]
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 64;
  double2 ra[2], rb[4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = t + 128 * u, row = idx >> 3, kc = (idx & 7) * 2;
      const int gm = m0 + row, gk = k0 + kc;
      ra[u] = (gm < M && gk < K) ? __ldg(reinterpret_cast<const double2*>(A + (int64_t)gm * lda + gk)) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = t + 128 * u, kk = idx >> 5, nc = (idx & 31) * 2;
      const int gk = k0 + kk, gn = n0 + nc;
      rb[u] = (gk < K && gn < N) ? *reinterpret_cast<const double2*>(B + (int64_t)gk * ldb + gn) : make_double2(0.0, 0.0);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = t + 128 * u, row = idx >> 3, kc = (idx & 7) * 2;
      As[buf][kc][row] = ra[u].x;
      As[buf][kc + 1][row] = ra[u].y;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = t + 128 * u, kk = idx >> 5, nc = (idx & 31) * 2;
      *reinterpret_cast<double2*>(&Bs[buf][kk][nc]) = rb[u];
    }
  };
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  // split-K: blockIdx.z owns K range [kbeg, K) with K clipped to its chunk; the raw partial goes to P[z]
  const int kbeg = kchunk > 0 ? blockIdx.z * kchunk : 0;
  if (kchunk > 0) K = min(K, kbeg + kchunk);
  load_tiles(kbeg);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < K; k0 += 16) {
    const bool more = k0 + 16 < K;
    if (more) load_tiles(k0 + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      // a thread owns rows {2ty, 2ty+1, 16+2ty, 17+2ty} and columns {2tx, 2tx+1, 32+2tx, 33+2tx}: the 16 lanes of a
      // half-warp then read 256 contiguous bytes per LDS.128 (the blocked 4-in-a-row ownership was a 5-way conflict)
      const double2 a01 = *reinterpret_cast<const double2*>(&As[buf][kk][ty * 2]);
      const double2 a23 = *reinterpret_cast<const double2*>(&As[buf][kk][16 + ty * 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&Bs[buf][kk][tx * 2]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bs[buf][kk][32 + tx * 2]);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y};
      const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + (i < 2 ? ty * 2 + i : 16 + ty * 2 + (i - 2));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
      const int n = n0 + (j < 2 ? tx * 2 : 32 + tx * 2);
      if (n >= N) continue;
      if (kchunk > 0) {
        *reinterpret_cast<double2*>(P + ((int64_t)blockIdx.z * M + m) * N + n) = make_double2(acc[i][j], acc[i][j + 1]);
        continue;
      }
      const int64_t o = (int64_t)m * ldc + n;
      double2 v = make_double2(alpha * acc[i][j], alpha * acc[i][j + 1]);
      if (X1) {
        const double2 x = *reinterpret_cast<const double2*>(X1 + o);
        v.x = fma(g1, x.x, v.x);
        v.y = fma(g1, x.y, v.y);
      }
      if (X2) {
        const double2 x = *reinterpret_cast<const double2*>(X2 + o);
        v.x = fma(g2, x.x, v.x);
        v.y = fma(g2, x.y, v.y);
      }
      *reinterpret_cast<double2*>(C + o) = v;
    }
  }
}

// The same product on the fp64 tensor cores (mma.sync m8n8k4): a warp owns a 32 x 32 block of C as 16 DMMA tiles and
// needs one shared-memory load per operand fragment for 256 multiply-adds of the warp, eight times less shared-memory
// traffic than the register-blocked CUDA-core kernel, which ncu showed waiting on shared memory (fp64 pipe 23 %).
// CTA = 4 warps = 64 x 64 tile; K in steps of 16 through double-buffered shared memory; split-K over blockIdx.z
// (kchunk > 0: raw partial to P[z]).  Same argument conventions and alignment requirements as k_dgemm_nn.
__global__ void __launch_bounds__(128)
k_dgemm_mma(const double* __restrict__ A, const double* __restrict__ B, double* C, int M, int N, int K, int lda, int ldb,
            int ldc, double alpha, double g1, const double* __restrict__ X1, double g2, const double* X2, int kchunk,
            double* __restrict__ P) {
  constexpr int S = 72;                                 // B row stride (doubles): 72 mod 16 == 8 -> fragment loads at 2 wavefronts
  constexpr int SA = 18;                                // A row stride: tile kept [m][k], stores and fragment loads conflict-free
  __shared__ __align__(16) double As[2][64][SA];        // [m][kk]
  __shared__ __align__(16) double Bs[2][16][S];         // [kk][n]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kbeg = kchunk > 0 ? blockIdx.z * kchunk : 0;
  if (kchunk > 0) K = min(K, kbeg + kchunk);
  double2 ra[4], rb[4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {                       // A: 64 rows x 16 k = 512 double2
      const int idx = t + 128 * u, row = idx >> 3, kc = (idx & 7) * 2;
      const int gm = m0 + row, gk = k0 + kc;
      ra[u] = (gm < M && gk < K) ? __ldg(reinterpret_cast<const double2*>(A + (int64_t)gm * lda + gk)) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {                       // B: 16 k x 64 cols = 512 double2
      const int idx = t + 128 * u, kk = idx >> 5, nc = (idx & 31) * 2;
      const int gk = k0 + kk, gn = n0 + nc;
      rb[u] = (gk < K && gn < N) ? *reinterpret_cast<const double2*>(B + (int64_t)gk * ldb + gn) : make_double2(0.0, 0.0);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = t + 128 * u, row = idx >> 3, kc = (idx & 7) * 2;
      *reinterpret_cast<double2*>(&As[buf][row][kc]) = ra[u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = t + 128 * u, kk = idx >> 5, nc = (idx & 31) * 2;
      *reinterpret_cast<double2*>(&Bs[buf][kk][nc]) = rb[u];
    }
  };
  double acc[4][4][2];                                  // [m-block][n-block][2 columns of the lane]
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int fr = lane >> 2, fc = lane & 3;              // fragment coordinates of the lane
  load_tiles(kbeg);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < K; k0 += 16) {
    const bool more = k0 + 16 < K;
    if (more) load_tiles(k0 + 16);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[buf][wm + i * 8 + fr][ks * 4 + fc];     // A[m = fr][k = fc]
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[buf][ks * 4 + fc][wn + j * 8 + fr];     // B[k = fc][n = fr]
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(af[i]), "d"(bf[j]));
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
  // C fragment: lane holds C[row = fr][cols 2 fc, 2 fc + 1] of every 8 x 8 tile
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + wm + i * 8 + fr;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + wn + j * 8 + 2 * fc;
      if (n >= N) continue;
      if (kchunk > 0) {
        *reinterpret_cast<double2*>(P + ((int64_t)blockIdx.z * M + m) * N + n) = make_double2(acc[i][j][0], acc[i][j][1]);
        continue;
      }
      const int64_t o = (int64_t)m * ldc + n;
      double2 v = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
      if (X1) {
        const double2 x = *reinterpret_cast<const double2*>(X1 + o);
        v.x = fma(g1, x.x, v.x);
        v.y = fma(g1, x.y, v.y);
      }
      if (X2) {
        const double2 x = *reinterpret_cast<const double2*>(X2 + o);
        v.x = fma(g2, x.x, v.x);
        v.y = fma(g2, x.y, v.y);
      }
      *reinterpret_cast<double2*>(C + o) = v;
    }
  }
}

// C = alpha * sum_z P[z] + g1 * X1 + g2 * X2 (the epilogue of a split-K k_dgemm_nn; fixed summation order)
__global__ void __launch_bounds__(256)
k_dgemm_combine(const double* __restrict__ P, int nz, double* C, int M, int N, int ldc, double alpha, double g1,
                const double* __restrict__ X1, double g2, const double* X2) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i >= (int64_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  double2 s = make_double2(0.0, 0.0);
  for (int z = 0; z < nz; ++z) {
    const double2 p = *reinterpret_cast<const double2*>(P + (int64_t)z * M * N + i);
    s.x += p.x;
    s.y += p.y;
  }
  const int64_t o = (int64_t)m * ldc + n;
  double2 v = make_double2(alpha * s.x, alpha * s.y);
  if (X1) {
    const double2 x = *reinterpret_cast<const double2*>(X1 + o);
    v.x = fma(g1, x.x, v.x);
    v.y = fma(g1, x.y, v.y);
  }
  if (X2) {
    const double2 x = *reinterpret_cast<const double2*>(X2 + o);
    v.x = fma(g2, x.x, v.x);
    v.y = fma(g2, x.y, v.y);
  }
  *reinterpret_cast<double2*>(C + o) = v;
}

// Optional per-kernel-class timing of the solver (CSC_PCA_PROF=1): CUDA event pairs around every helper, summed by
// label and printed to stderr when the solve ends.  Off by default (no events are created).
struct SolveProf {
  struct Rec { const char* label; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  bool on = false;
  static SolveProf& get() { static SolveProf p; return p; }
  void begin() { const char* e = getenv("CSC_PCA_PROF"); on = e && e[0] == '1'; recs.clear(); }
  void report(cudaStream_t st) {
    if (!on) return;
    cudaStreamSynchronize(st);
    std::vector<std::pair<std::string, std::pair<double, int>>> agg;
    for (auto& r : recs) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.a, r.b);
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
      bool found = false;
      for (auto& a : agg)
        if (a.first == r.label) { a.second.first += ms; a.second.second++; found = true; }
      if (!found) agg.push_back({r.label, {ms, 1}});
    }
    fprintf(stderr, "[csc pca prof]");
    for (auto& a : agg) fprintf(stderr, " %s %.3f ms /%d;", a.first.c_str(), a.second.first, a.second.second);
    fprintf(stderr, "\n");
    recs.clear();
  }
};
struct ProfScope {
  cudaStream_t st;
  int idx = -1;
  ProfScope(csc_ctx* ctx, const char* label) : st(ctx->stream) {
    SolveProf& p = SolveProf::get();
    if (!p.on) return;
    SolveProf::Rec r{label, nullptr, nullptr};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    idx = (int)p.recs.size();
    p.recs.push_back(r);
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(SolveProf::get().recs[idx].b, st);
  }
};

struct GemmEpi {
  double alpha = 1.0;
  double g1 = 0.0;
  const double* X1 = nullptr;
  double g2 = 0.0;
  const double* X2 = nullptr;
  const double* colscale = nullptr;
};

static csc_status dgemm(csc_ctx* ctx, bool ta, const double* A, const double* B, double* C, int M, int N, int K, int lda,
                        int ldb, int ldc, const GemmEpi& e = GemmEpi(), int kchunk = 0) {
  ProfScope ps(ctx, ta ? "gemm_tn" : (K > 256 ? "gemm_big" : "gemm_small"));
  const int nz = kchunk > 0 ? (K + kchunk - 1) / kchunk : 1;
  auto al16 = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15) == 0; };
  if (!ta && kchunk == 0 && !e.colscale && ((lda | ldb | ldc | K | N) & 1) == 0 && al16(A) && al16(B) && al16(C) &&
      al16(e.X1) && al16(e.X2)) {
    // 64 x 64 tiles on the fp64 tensor cores; a grid below two CTAs per SM leaves the pipes waiting on one warp per
    // scheduler, so K is split over blockIdx.z and the partials are summed in a second small kernel (fixed order)
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
    const int n_cta = (int)(grid.x * grid.y);
    int nsplit = K >= 512 ? std::min(8, (2 * ctx->sm_count) / std::max(1, n_cta)) : 1;
    if (nsplit > 1) {
      const int kc = ((K + nsplit - 1) / nsplit + 15) / 16 * 16;
      nsplit = (K + kc - 1) / kc;
      const size_t need = (size_t)nsplit * M * N * 8;
      if (!ctx->gemm_ws || ctx->gemm_ws->bytes < need) CSC_TRY(dev_alloc(ctx, need, &ctx->gemm_ws));
      grid.z = (unsigned)nsplit;
      k_dgemm_mma<<<grid, 128, 0, ctx->stream>>>(A, B, C, M, N, K, lda, ldb, ldc, e.alpha, e.g1, e.X1, e.g2, e.X2, kc,
                                                 (double*)ctx->gemm_ws->p);
      CSC_LAUNCHED(ctx);
      k_dgemm_combine<<<(unsigned)(((int64_t)M * N / 2 + 255) / 256), 256, 0, ctx->stream>>>(
          (const double*)ctx->gemm_ws->p, nsplit, C, M, N, ldc, e.alpha, e.g1, e.X1, e.g2, e.X2);
      CSC_LAUNCHED(ctx);
      return CSC_OK;
    }
    k_dgemm_mma<<<grid, 128, 0, ctx->stream>>>(A, B, C, M, N, K, lda, ldb, ldc, e.alpha, e.g1, e.X1, e.g2, e.X2, 0, nullptr);
    CSC_LAUNCHED(ctx);
    return CSC_OK;
  }
  // tall outputs with few column tiles: halve the row tile so that the grid covers the SMs
  const bool small_tile = !ta && ((N + 63) / 64) * ((M + 63) / 64) < ctx->sm_count;
  if (small_tile) {
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 31) / 32), (unsigned)nz);
    k_dgemm<false, 32><<<grid, 256, 0, ctx->stream>>>(A, B, C, M, N, K, lda, ldb, ldc, e.alpha, e.g1, e.X1, e.g2, e.X2,
                                                       e.colscale, kchunk);
  } else {
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64), (unsigned)nz);
    if (ta)
      k_dgemm<true, 64><<<grid, 256, 0, ctx->stream>>>(A, B, C, M, N, K, lda, ldb, ldc, e.alpha, e.g1, e.X1, e.g2, e.X2,
                                                        e.colscale, kchunk);
    else
      k_dgemm<false, 64><<<grid, 256, 0, ctx->stream>>>(A, B, C, M, N, K, lda, ldb, ldc, e.alpha, e.g1, e.X1, e.g2, e.X2,
                                                         e.colscale, kchunk);
  }
  CSC_LAUNCHED(ctx);
  return CSC_OK;
}

// S[i][j] = sum_z parts[z][i][j]  (fixed order: deterministic); sym: S = (S + S')/2
__global__ void k_reduce_parts(const double* __restrict__ parts, int nz, int b, int sym, double* __restrict__ S) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= b * b) return;
  const int i = idx / b, j = idx % b;
  double s = 0.0, st = 0.0;
  for (int z = 0; z < nz; ++z) {
    s += parts[(int64_t)z * b * b + idx];
    if (sym) st += parts[(int64_t)z * b * b + (int64_t)j * b + i];
  }
  S[idx] = sym ? 0.5 * (s + st) : s;
}

// out = (a * X + c * Y) * colscale
__global__ void k_axpby_cols(const double* X, const double* Y, double a, double c,
                             const double* __restrict__ colscale, int64_t n, int b, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = a * X[i] + c * Y[i];
  if (colscale) v *= colscale[i % b];
  out[i] = v;
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

// Upper Cholesky S = R'R of a b x b matrix followed by the in-place inverse of R, all in shared memory
// (single CTA, b <= 160).  Rinv (upper triangular, zeros below) is written to global memory, so that
// Q = Y R^{-1} becomes a plain matrix product.  A pivot that has lost all its digits against the column's
// own norm (S_jj) drops that direction (row/column of zeros) instead of poisoning the block.
__global__ void __launch_bounds__(1024)
k_chol_inv(const double* __restrict__ S, int b, double* __restrict__ Rinv) {
  extern __shared__ double sm[];
  const int ld = b + 1;
  double* A = sm;             // [b][ld]
  double* rs = sm + b * ld;   // [b] 1/sqrt(pivot)
  double* d0 = rs + b;        // [b] original diagonal
  const int t = threadIdx.x, nt = blockDim.x;
  const int lane = t & 31, warp = t >> 5, nwarp = nt >> 5;
  for (int i = t; i < b * b; i += nt) A[(i / b) * ld + (i % b)] = S[i];
  __syncthreads();
  for (int i = t; i < b; i += nt) d0[i] = fabs(A[i * ld + i]);
  __syncthreads();
  const double rel = 4.0 * (double)b * 1.1e-16;
  // right-looking Cholesky on the unscaled rows: S[r][c] -= S[j][r] S[j][c] / d_j; rows are scaled at the end
  for (int j = 0; j < b; ++j) {
    const double d = A[j * ld + j];
    const bool ok = d > rel * d0[j] && d > 0.0;
    const double invd = ok ? __drcp_rn(d) : 0.0;      // every thread needs it: the short reciprocal, not a full division
    if (t == 0) rs[j] = ok ? rsqrt(d) : 0.0;
    for (int r = j + 1 + warp; r < b; r += nwarp) {
      const double f = A[j * ld + r] * invd;
      for (int c = r + lane; c < b; c += 32) A[r * ld + c] -= f * A[j * ld + c];
    }
    __syncthreads();
  }
  for (int r = warp; r < b; r += nwarp) {
    const double f = rs[r];
    for (int c = r + lane; c < b; c += 32) A[r * ld + c] *= f;   // R (diagonal: d * rsqrt(d) = sqrt(d))
  }
  __syncthreads();
  // inverse of the upper triangle by back substitution, R X = I, one column per group of 8 lanes (columns are
  // independent: no block barrier).  X' is stored in the unused lower triangle, X[l][c] at A[c][l]; the
  // diagonal of R is read from a copy.
  double* rd = d0 + b;
  for (int i = t; i < b; i += nt) {
    const double rii = A[i * ld + i];
    rd[i] = rii != 0.0 ? 1.0 / rii : 0.0;      // reciprocal diagonal: the substitution below multiplies
  }
  __syncthreads();
  {
    const int grp = t >> 3, gl = t & 7;
    const unsigned gmask = 0xFFu << ((t & 31) & ~7);
    for (int c = grp; c < b; c += nt >> 3) {
      for (int i = c; i >= 0; --i) {
        double acc = 0.0;
        for (int l = i + 1 + gl; l <= c; l += 8) acc = fma(A[i * ld + l], A[c * ld + l], acc);
        acc += __shfl_xor_sync(gmask, acc, 1);
        acc += __shfl_xor_sync(gmask, acc, 2);
        acc += __shfl_xor_sync(gmask, acc, 4);
        __syncwarp(gmask);
        if (gl == 0) A[c * ld + i] = ((i == c ? 1.0 : 0.0) - acc) * rd[i];
        __syncwarp(gmask);
      }
    }
  }
  __syncthreads();
  for (int idx = t; idx < b * b; idx += nt) {
    const int r = idx / b, c = idx % b;
    Rinv[idx] = c >= r ? A[c * ld + r] : 0.0;
  }
}

// Panel-blocked variant of k_chol_inv for blocks whose workspace fits (b <= ~144): R'R = S by panels of CHOL_NB rows
// (diagonal block in registers of one warp, panel rows by forward substitution, trailing update by all threads),
// then inv(R) by inverting the CHOL_NB x CHOL_NB diagonal blocks and doubling: X_PQ = -X_PP R_PQ X_QQ.  Same drop rule and
// output as k_chol_inv; 3 block barriers per panel instead of one per column (b = 128: 96 us instead of 210 us).
// X = inv(R) is kept transposed in the strict lower triangle (X[i][c] at A[c][i], i < c) with its diagonal in rd[].
constexpr int CHOL_NB = 16;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// One warp factors the CHOL_NB x CHOL_NB diagonal block at (j0, j0) in registers (lane c = column c of the block) and writes
// R11 (upper, scaled) back plus rd = 1/R_jj.  Unscaled right-looking steps: a[r] -= (A[j][r] / d_j) A[j][c]; the
// row scalings 1/sqrt(d_j) are off the dependency chain.
__device__ __forceinline__ void chol_diag_block(double* A, int ld, int j0, int nbk, const double* d0, double rel,
                                                double* rd, int lane) {
  double a[CHOL_NB], rs[CHOL_NB];
  const int c = lane & (CHOL_NB - 1);
#pragma unroll
  for (int r = 0; r < CHOL_NB; ++r) a[r] = (r <= c && c < nbk && r < nbk) ? A[(j0 + r) * ld + j0 + c] : (r == c ? 1.0 : 0.0);
#pragma unroll
  for (int j = 0; j < CHOL_NB; ++j) {
    const double d = shfl_d(a[j], j);
    const double dj0 = j < nbk ? d0[j0 + j] : 1.0;
    const bool ok = d > rel * dj0 && d > 0.0;
    const double invd = ok ? __drcp_rn(d) : 0.0;
    rs[j] = ok ? rsqrt(d) : 0.0;
    const double f = a[j] * invd;          // A[j][c] / d
#pragma unroll
    for (int r = j + 1; r < CHOL_NB; ++r) a[r] = fma(-shfl_d(a[j], r), f, a[r]);
  }
  if (lane < CHOL_NB && c < nbk) {
#pragma unroll
    for (int r = 0; r < CHOL_NB; ++r)
      if (r <= c) A[(j0 + r) * ld + j0 + c] = a[r] * rs[r];
    rd[j0 + c] = rs[c];
  }
}

// One warp inverts the CHOL_NB x CHOL_NB upper-triangular diagonal block at (j0, j0): lane c owns column c of X, the rows of
// R come by broadcast shuffles.  X goes to the strict lower triangle (transposed).
__device__ __forceinline__ void inv_diag_block(double* A, int ld, int j0, int nbk, const double* rd, int lane) {
  double rcol[CHOL_NB], x[CHOL_NB];
  const int c = lane & (CHOL_NB - 1);
#pragma unroll
  for (int r = 0; r < CHOL_NB; ++r) rcol[r] = (r < c && c < nbk) ? A[(j0 + r) * ld + j0 + c] : 0.0;   // R[r][c], strict upper
  const double rdc = c < nbk ? rd[j0 + c] : 0.0;
#pragma unroll
  for (int i = CHOL_NB - 1; i >= 0; --i) {
    double acc = 0.0;
#pragma unroll
    for (int l = i + 1; l < CHOL_NB; ++l) acc = fma(shfl_d(rcol[i], l), x[l], acc);      // R[i][l] x[l]
    const double rdi = shfl_d(rdc, i);
    x[i] = (i == c) ? rdc : (i < c ? -acc * rdi : 0.0);
  }
  if (lane < CHOL_NB && c < nbk) {
#pragma unroll
    for (int i = 0; i < CHOL_NB; ++i)
      if (i < c) A[(j0 + c) * ld + j0 + i] = x[i];
  }
}

__global__ void __launch_bounds__(512)
k_chol_inv_blocked(const double* __restrict__ S, int b, double* __restrict__ Rinv, int tld) {
  extern __shared__ double sm[];
  const int ld = b + 1;
  double* A = sm;              // [b][ld]
  double* rd = A + b * ld;     // [b] 1 / R_ii (0 for a dropped column)
  double* d0 = rd + b;         // [b] original diagonal
  double* T = d0 + b;          // temp of the inverse levels, [s][tld]
  const int t = threadIdx.x, nt = blockDim.x;
  const int lane = t & 31, warp = t >> 5, nwarp = nt >> 5;
  for (int i = t; i < b * b; i += nt) A[(i / b) * ld + (i % b)] = S[i];
  __syncthreads();
  for (int i = t; i < b; i += nt) d0[i] = fabs(A[i * ld + i]);
  __syncthreads();
  const double rel = 4.0 * (double)b * 1.1e-16;
  for (int j0 = 0; j0 < b; j0 += CHOL_NB) {
    const int j1 = min(j0 + CHOL_NB, b);
    if (warp == 0) chol_diag_block(A, ld, j0, j1 - j0, d0, rel, rd, lane);
    __syncthreads();
    // the last warp inverts the finished diagonal block while the others solve the panel rows:
    // R12 = R11^-T A12, one thread per trailing column
    if (warp == nwarp - 1) {
      inv_diag_block(A, ld, j0, j1 - j0, rd, lane);
    } else {
      for (int c = j1 + t; c < b; c += nt - 32) {
        double v[CHOL_NB];
#pragma unroll
        for (int i = 0; i < CHOL_NB; ++i) {
          if (j0 + i < j1) {
            double acc = A[(j0 + i) * ld + c];
#pragma unroll
            for (int l = 0; l < i; ++l) acc = fma(-A[(j0 + l) * ld + j0 + i], v[l], acc);
            v[i] = acc * rd[j0 + i];
            A[(j0 + i) * ld + c] = v[i];
          }
        }
      }
    }
    __syncthreads();
    // trailing update A22 -= R12' R12 on the upper triangle: 2 rows x 2 columns (c, c + mh) per thread
    const int m = b - j1;
    if (m > 0) {
      // 4 rows x 2 columns (c, c + mh) per thread
      const int mh = (m + 1) >> 1, mq = (m + 3) >> 2;
      for (int idx = t; idx < mq * mh; idx += nt) {
        const int tr = idx / mh, tc = idx % mh;
        const int r0 = j1 + 4 * tr, c0 = j1 + tc, c1 = c0 + mh;
        if (c1 < r0) continue;         // whole tile strictly below the diagonal
        const bool hc1 = c1 < b;
        const int c1c = hc1 ? c1 : c0;
        int rr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) rr[k] = min(r0 + k, b - 1);
        double acc[4][2] = {};
#pragma unroll 4
        for (int l = j0; l < j1; ++l) {
          const double y0 = A[l * ld + c0], y1 = A[l * ld + c1c];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const double x = A[l * ld + rr[k]];
            acc[k][0] = fma(x, y0, acc[k][0]);
            acc[k][1] = fma(x, y1, acc[k][1]);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = r0 + k;
          if (r < b) {
            if (c0 >= r) A[r * ld + c0] -= acc[k][0];
            if (hc1 && c1 >= r) A[r * ld + c1] -= acc[k][1];
          }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // doubling: X_PQ = -X_PP (R_PQ X_QQ) for neighbouring blocks P = [p0, p0+s), Q = [p0+s, min(p0+2s, b))
  for (int s = CHOL_NB; s < b; s <<= 1) {
    const int npair = (b + 2 * s - 1) / (2 * s);
    const int s4 = s >> 2;
    // T[pair][i - p0][c - q0] = sum_{l in Q, l <= c} R[i][l] X[l][c]; a thread: rows i, i + s/4, i + 2s/4, i + 3s/4
    for (int idx = t; idx < npair * s4 * s; idx += nt) {
      const int pr = idx / (s4 * s), rem = idx % (s4 * s);
      const int p0 = pr * 2 * s, q0 = p0 + s;
      const int i = p0 + rem / s, c = q0 + rem % s;
      if (c >= b) continue;
      const double xcc = rd[c];
      double a0 = A[i * ld + c] * xcc, a1 = A[(i + s4) * ld + c] * xcc, a2 = A[(i + 2 * s4) * ld + c] * xcc,
             a3 = A[(i + 3 * s4) * ld + c] * xcc;
#pragma unroll 4
      for (int l = q0; l < c; ++l) {
        const double x = A[c * ld + l];
        a0 = fma(A[i * ld + l], x, a0);
        a1 = fma(A[(i + s4) * ld + l], x, a1);
        a2 = fma(A[(i + 2 * s4) * ld + l], x, a2);
        a3 = fma(A[(i + 3 * s4) * ld + l], x, a3);
      }
      double* Tp = T + (pr * s + (i - p0)) * tld + (c - q0);
      Tp[0] = a0; Tp[s4 * tld] = a1; Tp[2 * s4 * tld] = a2; Tp[3 * s4 * tld] = a3;
    }
    __syncthreads();
    // X[i][c] = -sum_{l in P, l >= i} X[i][l] T[l][c]; a thread: columns c, c + s/4, c + 2s/4, c + 3s/4
    for (int idx = t; idx < npair * s * s4; idx += nt) {
      const int pr = idx / (s * s4), rem = idx % (s * s4);
      const int p0 = pr * 2 * s, q0 = p0 + s;
      const int i = p0 + rem / s4, cc = rem % s4;
      if (q0 + cc >= b) continue;
      const double* Tp = T + (pr * s + (i - p0)) * tld + cc;
      const double xii = rd[i];
      double a0 = xii * Tp[0], a1 = xii * Tp[s4], a2 = xii * Tp[2 * s4], a3 = xii * Tp[3 * s4];
#pragma unroll 4
      for (int l = i + 1; l < q0; ++l) {
        const double x = A[l * ld + i];
        const double* Tl = T + (pr * s + (l - p0)) * tld + cc;
        a0 = fma(x, Tl[0], a0); a1 = fma(x, Tl[s4], a1); a2 = fma(x, Tl[2 * s4], a2); a3 = fma(x, Tl[3 * s4], a3);
      }
      const int c = q0 + cc;
      A[c * ld + i] = -a0;
      if (c + s4 < b) A[(c + s4) * ld + i] = -a1;
      if (c + 2 * s4 < b) A[(c + 2 * s4) * ld + i] = -a2;
      if (c + 3 * s4 < b) A[(c + 3 * s4) * ld + i] = -a3;
    }
    __syncthreads();
  }
  for (int idx = t; idx < b * b; idx += nt) {
    const int r = idx / b, c = idx % b;
    Rinv[idx] = c > r ? A[c * ld + r] : (c == r ? rd[r] : 0.0);
  }
}

// round-robin pair (p < q) number i of round r for m (even) players; q == m - 1 may be the dummy index
__device__ __forceinline__ void jacobi_pair(int m, int r, int i, int& p, int& q) {
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
}

// Parallel cyclic Jacobi for a symmetric n x n matrix (single CTA, matrix resident in shared memory, n <= 160).
// Round-robin ordering: n/2 disjoint rotations per round; the two-sided update J' A J is applied as independent
// 2x2 blocks (pair i, pair j), so a round needs two barriers and touches shared memory only.  The eigenvectors
// are NOT accumulated here: every rotation (c, s) is appended to `rot` and k_apply_rot replays the list on
// whatever tall matrix needs X <- X V (the Ritz step applies it to [Q; Gc Q] directly, thousands of rows in
// parallel).  On exit eval[] is descending and rank[i] = position of original column i in that order.
__global__ void __launch_bounds__(1024)
k_jacobi(const double* __restrict__ Ain, int n, double2* __restrict__ rot, uchar2* __restrict__ rot_pq,
         double* __restrict__ eval, int* __restrict__ rank_out, int max_sweeps, double off_tol2, int* __restrict__ sweeps_out) {
  extern __shared__ double sm[];
  const int ld = n;                // a warp works along one row with scattered columns: no padding needed
  const int m = (n + 1) & ~1;
  const int hp = m / 2;
  double* A = sm;                  // [n][ld]
  double* cs = sm + n * ld;        // [hp]
  double* sn = cs + hp;            // [hp]
  int* pp = (int*)(sn + hp);       // [hp]
  int* qq = pp + hp;               // [hp]
  __shared__ double s_off, s_diag;
  const int t = threadIdx.x, nt = blockDim.x;
  for (int i = t; i < n * n; i += nt) {
    const int r = i / n, c = i % n;
    A[r * ld + c] = 0.5 * (Ain[i] + Ain[c * n + r]);
  }
  // the pair schedule of one sweep (the same for every sweep), for k_apply_rot
  for (int idx = t; idx < (m - 1) * hp; idx += nt) {
    int p, q;
    jacobi_pair(m, idx / hp, idx % hp, p, q);
    rot_pq[idx] = make_uchar2((unsigned char)p, (unsigned char)q);
  }
  __syncthreads();
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (t == 0) {
      s_off = 0.0;
      s_diag = 0.0;
    }
    __syncthreads();
    double off = 0.0, dg = 0.0;
    for (int r = t >> 5; r < n; r += nt >> 5)
      for (int c = t & 31; c < n; c += 32) {
        const double v = A[r * ld + c];
        if (r == c) dg += v * v;
        else off += v * v;
      }
    off = warp_sum(off);
    dg = warp_sum(dg);
    if ((t & 31) == 0) {
      atomicAdd(&s_off, off);
      atomicAdd(&s_diag, dg);
    }
    __syncthreads();
    const bool done = (s_off <= off_tol2 * (s_diag + s_off)) || (s_off == 0.0);
    __syncthreads();
    if (done) break;
    for (int r = 0; r < m - 1; ++r) {
      // rotation of every pair of this round, from the current matrix
      if (t < hp) {
        int p, q;
        jacobi_pair(m, r, t, p, q);
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = A[p * ld + q];
          if (apq != 0.0) {
            const double app = A[p * ld + p], aqq = A[q * ld + q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = rsqrt(tt * tt + 1.0);
            s = tt * c;
          }
        }
        cs[t] = c;
        sn[t] = s;
        pp[t] = p;
        qq[t] = q;
        rot[((int64_t)sweep * (m - 1) + r) * hp + t] = make_double2(c, s);
      }
      __syncthreads();
      // A <- J' A J, one independent 2x2 block per (pair i, pair j); 16 lanes share a pair i
      for (int i = t >> 4; i < hp; i += nt >> 4) {
        const int pi = pp[i], qi = qq[i];
        const double ci = cs[i], si = sn[i];
        for (int j = t & 15; j < hp; j += 16) {
          const int pj = pp[j], qj = qq[j];
          if (qi >= n || qj >= n) {
            // odd n: the pair that holds the dummy index does not rotate (c = 1, s = 0); its real row/column
            // still takes the other pair's rotation
            if (qi >= n && qj >= n) continue;
            if (qi >= n) {
              const double c = cs[j], s = sn[j];
              const double a0 = A[pi * ld + pj], a1 = A[pi * ld + qj];
              A[pi * ld + pj] = c * a0 - s * a1;
              A[pi * ld + qj] = s * a0 + c * a1;
            } else {
              const double a0 = A[pi * ld + pj], a1 = A[qi * ld + pj];
              A[pi * ld + pj] = ci * a0 - si * a1;
              A[qi * ld + pj] = si * a0 + ci * a1;
            }
            continue;
          }
          const double cj = cs[j], sj = sn[j];
          const double b00 = A[pi * ld + pj], b01 = A[pi * ld + qj], b10 = A[qi * ld + pj], b11 = A[qi * ld + qj];
          const double r00 = ci * b00 - si * b10, r01 = ci * b01 - si * b11;
          const double r10 = si * b00 + ci * b10, r11 = si * b01 + ci * b11;
          double o00 = cj * r00 - sj * r01, o01 = sj * r00 + cj * r01;
          double o10 = cj * r10 - sj * r11, o11 = sj * r10 + cj * r11;
          if (i == j) {
            o01 = 0.0;   // annihilated by construction
            o10 = 0.0;
          }
          A[pi * ld + pj] = o00;
          A[pi * ld + qj] = o01;
          A[qi * ld + pj] = o10;
          A[qi * ld + qj] = o11;
        }
      }
      __syncthreads();
    }
  }
  if (t == 0 && sweeps_out) *sweeps_out = sweep;
  __syncthreads();
  // eigenvalues descending (rank by counting; ties by index)
  for (int i = t; i < n; i += nt) {
    const double di = A[i * ld + i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double dj = A[j * ld + j];
      rank += (dj > di) || (dj == di && j < i);
    }
    eval[rank] = di;
    rank_out[i] = rank;
  }
}

// Xout[:, rank[c]] = (X V)[:, c] where V is the product of the recorded rotations: a CTA keeps APPLY_ROWS rows
// of X in shared memory and replays the list; rows are independent, so the grid is n_rows / APPLY_ROWS CTAs.
// blockIdx.y selects one of two matrices (the Ritz step rotates Q and Gc Q with the same list).
constexpr int APPLY_ROWS = 8;
constexpr int APPLY_CHUNK = 8;    // rounds of rotations staged in shared memory at a time (double-buffered cp.async)
__global__ void __launch_bounds__(APPLY_ROWS * 80)
k_apply_rot(const double* __restrict__ X0, double* __restrict__ O0, const double* __restrict__ X1, double* __restrict__ O1,
            int n_rows, int n, const double2* __restrict__ rot, const uchar2* __restrict__ rot_pq,
            const int* __restrict__ sweeps_ptr, const int* __restrict__ rank) {
  extern __shared__ __align__(16) double sm[];
  const int ld = n + 1;
  const int m = (n + 1) & ~1;
  const int hp = m / 2;
  double2* stage = reinterpret_cast<double2*>(sm);                  // [2][APPLY_CHUNK * hp]
  double* xs = sm + 2 * 2 * APPLY_CHUNK * hp;                       // [APPLY_ROWS][ld]
  uchar2* pq = reinterpret_cast<uchar2*>(xs + APPLY_ROWS * ld);     // [(m-1) * hp] pair schedule of a sweep
  const double* X = blockIdx.y ? X1 : X0;
  double* O = blockIdx.y ? O1 : O0;
  const int row0 = blockIdx.x * APPLY_ROWS;
  const int t = threadIdx.x, nt = blockDim.x;
  const int rounds = *sweeps_ptr * (m - 1);
  const int n_chunks = (rounds + APPLY_CHUNK - 1) / APPLY_CHUNK;
  auto fetch = [&](int chunk, int buf) {
    const int64_t base = (int64_t)chunk * APPLY_CHUNK * hp;
    const int cnt = min(APPLY_CHUNK, rounds - chunk * APPLY_CHUNK) * hp;
    for (int i = t; i < cnt; i += nt) __pipeline_memcpy_async(stage + buf * APPLY_CHUNK * hp + i, rot + base + i, 16);
    __pipeline_commit();
  };
  if (n_chunks > 0) fetch(0, 0);
  for (int i = t; i < (m - 1) * hp; i += nt) pq[i] = rot_pq[i];
  for (int i = t; i < APPLY_ROWS * n; i += nt) {
    const int rr = i / n, c = i - rr * n;
    xs[rr * ld + c] = (row0 + rr < n_rows) ? X[(int64_t)(row0 + rr) * n + c] : 0.0;
  }
  const int rr = t / hp, i = t - rr * hp;     // blockDim.x == APPLY_ROWS * hp
  double* xr = xs + rr * ld;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < n_chunks) {
      fetch(ch + 1, buf ^ 1);
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
    const int g0 = ch * APPLY_CHUNK;
    const int cnt = min(APPLY_CHUNK, rounds - g0);
    int r = g0 % (m - 1);
    const double2* st = stage + buf * APPLY_CHUNK * hp + i;
    for (int g = 0; g < cnt; ++g) {
      const double2 cs = st[g * hp];
      const uchar2 pr = pq[r * hp + i];
      if (pr.y < n) {
        const double xp = xr[pr.x], xq = xr[pr.y];
        xr[pr.x] = cs.x * xp - cs.y * xq;
        xr[pr.y] = cs.y * xp + cs.x * xq;
      }
      if (++r == m - 1) r = 0;
      __syncthreads();
    }
  }
  __syncthreads();
  for (int idx = t; idx < APPLY_ROWS * n; idx += nt) {
    const int r2 = idx / n, c = idx - r2 * n;
    if (row0 + r2 < n_rows) O[(int64_t)(row0 + r2) * n + rank[c]] = xs[r2 * ld + c];
  }
}

// out[j] = || R[:, j] ||_2 of the H x b block R, out[b + j] = B[j][j]; one CTA per column
__global__ void __launch_bounds__(256)
k_block_stats(const double* __restrict__ R, int H, int b, const double* __restrict__ B, double* __restrict__ out) {
  __shared__ double red[8];
  const int j = blockIdx.x;
  double s = 0.0;
  for (int r = threadIdx.x; r < H; r += blockDim.x) {
    const double v = R[(int64_t)r * b + j];
    s += v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      out[j] = sqrt(v);
      out[b + j] = B[(int64_t)j * b + j];
    }
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

// diagonal of ones into a zero-filled n x n matrix
__global__ void k_set_identity(double* __restrict__ a, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[(int64_t)i * n + i] = 1.0;
}

constexpr int PCA_SMALL_MAX = 160;   // largest matrix the single-CTA shared-memory kernels hold

constexpr int JACOBI_MAX_SWEEPS = 40;
// rotation list of a Jacobi run: (c, s) per pair per round, followed by the pair schedule of one sweep
struct RotList {
  DevPtr mem;
  double2* cs = nullptr;
  uchar2* pq = nullptr;
};
static csc_status rot_alloc(csc_ctx* ctx, int n, RotList* r) {
  const int m = (n + 1) & ~1;
  const size_t cs_bytes = (size_t)JACOBI_MAX_SWEEPS * (m - 1) * (m / 2) * sizeof(double2);
  CSC_TRY(dev_alloc(ctx, cs_bytes + (size_t)(m - 1) * (m / 2) * sizeof(uchar2), &r->mem));
  r->cs = (double2*)r->mem->p;
  r->pq = (uchar2*)((char*)r->mem->p + cs_bytes);
  return CSC_OK;
}
// off_tol: stop when ||offdiag||_F <= off_tol * ||A||_F
static csc_status jacobi(csc_ctx* ctx, const double* A, int n, const RotList& rot, double* eval, int* rank, int* d_sweeps,
                         double off_tol = 1e-13) {
  const int hp = ((n + 1) & ~1) / 2;
  const size_t smem = (size_t)n * n * 8 + (size_t)hp * 24 + 64;
  CSC_CUDA(ctx, cudaFuncSetAttribute(k_jacobi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(ctx, "jacobi");
  k_jacobi<<<1, 1024, smem, ctx->stream>>>(A, n, rot.cs, rot.pq, eval, rank, JACOBI_MAX_SWEEPS, off_tol * off_tol, d_sweeps);
  CSC_LAUNCHED(ctx);
  return CSC_OK;
}
// O0 = X0 V (columns in eigenvalue order), optionally O1 = X1 V
static csc_status apply_rot(csc_ctx* ctx, const double* X0, double* O0, const double* X1, double* O1, int n_rows, int n,
                            const RotList& rot, const int* d_sweeps, const int* rank) {
  const int m = (n + 1) & ~1, hp = m / 2;
  const size_t smem = (size_t)APPLY_ROWS * (n + 1) * 8 + (size_t)2 * APPLY_CHUNK * hp * 16 + (size_t)(m - 1) * hp * 2 + 16;
  CSC_CUDA(ctx, cudaFuncSetAttribute(k_apply_rot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((n_rows + APPLY_ROWS - 1) / APPLY_ROWS), X1 ? 2u : 1u);
  ProfScope ps(ctx, "apply_rot");
  k_apply_rot<<<grid, APPLY_ROWS * hp, smem, ctx->stream>>>(X0, O0, X1, O1, n_rows, n, rot.cs, rot.pq, d_sweeps, rank);
  CSC_LAUNCHED(ctx);
  return CSC_OK;
}

// workspace of the block solver
struct BlockWs {
  int H, b, nz, kchunk;
  double *parts, *S, *Rinv;
};

// out[b x b] = X' Y for two H x b blocks (K split over the grid, deterministic reduction)
static csc_status small_gram(csc_ctx* ctx, const BlockWs& w, const double* X, const double* Y, int sym, double* out) {
  CSC_TRY(dgemm(ctx, true, X, Y, w.parts, w.b, w.b, w.H, w.b, w.b, w.b, GemmEpi(), w.kchunk));
  {
    ProfScope ps(ctx, "reduce_parts");
    k_reduce_parts<<<(unsigned)((w.b * w.b + 255) / 256), 256, 0, ctx->stream>>>(w.parts, w.nz, w.b, sym, out);
  }
  CSC_LAUNCHED(ctx);
  return CSC_OK;
}

// Q = Y R^{-1} with R'R = Y'Y
static csc_status chol_qr(csc_ctx* ctx, const BlockWs& w, const double* Y, double* Q) {
  CSC_TRY(small_gram(ctx, w, Y, Y, 1, w.S));
  // workspace of the blocked kernel: the matrix, two vectors and the widest temp of the doubling levels
  int trows = 0, tld = 0;
  for (int sz = CHOL_NB; sz < w.b; sz <<= 1) {
    trows = std::max(trows, ((w.b + 2 * sz - 1) / (2 * sz)) * sz);
    tld = std::max(tld, sz + 1);
  }
  const size_t smem_blk = ((size_t)w.b * (w.b + 1) + 2 * (size_t)w.b + (size_t)trows * tld) * 8;
  if (smem_blk <= (size_t)227 * 1024) {
    CSC_CUDA(ctx, cudaFuncSetAttribute(k_chol_inv_blocked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_blk));
    ProfScope ps(ctx, "chol_inv");
    k_chol_inv_blocked<<<1, 512, smem_blk, ctx->stream>>>(w.S, w.b, w.Rinv, tld);
  } else {
    const size_t smem = (size_t)w.b * (w.b + 1) * 8 + (size_t)w.b * 24;
    CSC_CUDA(ctx, cudaFuncSetAttribute(k_chol_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(ctx, "chol_inv");
    k_chol_inv<<<1, 1024, smem, ctx->stream>>>(w.S, w.b, w.Rinv);
  }
  CSC_LAUNCHED(ctx);
  return dgemm(ctx, false, Y, w.Rinv, Q, w.H, w.b, w.b, w.b, w.b, w.b);
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
  SolveProf::get().begin();
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
  // a NaN anywhere in the scaled matrix (a selected gene with zero variance: 0/0 in scale_object) reaches
  // the Gram diagonal; LAPACK's gesdd fails on such input in the reference, and the iteration below would never converge
  for (int r = 0; r < H; ++r)
    if (!std::isfinite(diag[r]) || !std::isfinite(hsum[r]))
      return csc_fail(ctx, CSC_ERR_NUMERIC, "csc_pca_solve: the Gram matrix is not finite (feature row %d); a selected gene has "
                      "zero variance or NaN values in the scaled matrix", r);

  std::vector<double> evals((size_t)k), U((size_t)H * k);
  int iters = 0;
  double resid = 0.0;
  if (H <= PCA_SMALL_MAX) {
    // small problem: full Jacobi eigendecomposition of Gc; V = I * (rotations)
    DevPtr eye, vs, ev, sw, rk;
    RotList rot;
    CSC_TRY(dev_alloc(ctx, (size_t)H * H * 8, &eye, true));
    CSC_TRY(dev_alloc(ctx, (size_t)H * H * 8, &vs));
    CSC_TRY(dev_alloc(ctx, (size_t)H * 8, &ev));
    CSC_TRY(dev_alloc(ctx, 4, &sw));
    CSC_TRY(dev_alloc(ctx, (size_t)H * 4, &rk));
    CSC_TRY(rot_alloc(ctx, H, &rot));
    k_set_identity<<<(unsigned)((H + 255) / 256), 256, 0, ctx->stream>>>((double*)eye->p, H);
    CSC_LAUNCHED(ctx);
    CSC_TRY(jacobi(ctx, Gc, H, rot, (double*)ev->p, (int*)rk->p, (int*)sw->p));
    CSC_TRY(apply_rot(ctx, (const double*)eye->p, (double*)vs->p, nullptr, nullptr, H, H, rot, (const int*)sw->p,
                      (const int*)rk->p));
    std::vector<double> hv((size_t)H * H), he((size_t)H);
    CSC_CUDA(ctx, cudaMemcpyAsync(hv.data(), vs->p, hv.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaMemcpyAsync(he.data(), ev->p, he.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < k; ++j) {
      evals[j] = he[j];
      for (int r = 0; r < H; ++r) U[(size_t)j * H + r] = hv[(size_t)r * H + j];
    }
  } else {
    // Chebyshev-filtered block subspace iteration with Rayleigh-Ritz:
    //   block of b >= k+14 vectors; two plain power cycles and a first Rayleigh-Ritz step give Ritz values
    //   theta_1 >= ... >= theta_b.  Every following cycle applies the degree-m Chebyshev polynomial that is
    //   bounded by 1 on [0, theta_b] (three-term recurrence, one Gc product per degree, fused epilogue) and
    //   re-orthonormalises by CholeskyQR2.  m is chosen so that the growth T_m(x_1) of the leading direction,
    //   which is what contaminates the guard columns, stays inside the range CholeskyQR2 resolves.  A
    //   Rayleigh-Ritz step (Jacobi on the b x b projection, rotations replayed on [Q; Gc Q]) follows every
    //   second cycle and yields the residuals that stop the iteration.  All arithmetic fp64 on the device;
    //   the host sees only the b Ritz values and k residual norms per Ritz step.
    // Block width: k wanted vectors + 14 guard columns, in steps of 16 (k = 50 -> 64).  The b x b kernels (Cholesky,
    // Jacobi, rotation replay) cost O(b^3) on ONE SM and dominated the solve at b = 128 (5.2 of 7.6 ms); the narrower
    // block needs more filter products (40 instead of 33 on the bench matrix, 217 instead of 88 on a flat VisiumHD-like
    // spectrum) but every product and every small kernel is cheaper: 4.0 instead of 7.6 ms, never slower.
    int b = std::min(H, std::max(64, ((k + 14 + 15) / 16) * 16));
    if (const char* eb = getenv("CSC_PCA_BLOCK")) {      // tuning knob: block width (multiple of 16, >= k + 8)
      const int v = atoi(eb);
      if (v >= k + 8 && v <= H) b = v;
    }
    b = std::min(b, PCA_SMALL_MAX);
    CSC_REQUIRE(ctx, k <= b, "csc_pca_solve: maxoutdim=%d too large for the block solver (max %d)", k, b);
    BlockWs w;
    w.H = H;
    w.b = b;
    w.kchunk = 64;
    w.nz = (H + w.kchunk - 1) / w.kchunk;
    DevPtr buf[4], parts, s, rinv, bm, rk, th, res, sw;
    RotList rot;
    for (int i = 0; i < 4; ++i) CSC_TRY(dev_alloc(ctx, (size_t)H * b * 8, &buf[i]));
    CSC_TRY(dev_alloc(ctx, (size_t)w.nz * b * b * 8, &parts));
    CSC_TRY(dev_alloc(ctx, (size_t)b * b * 8, &s));
    CSC_TRY(dev_alloc(ctx, (size_t)b * b * 8, &rinv));
    CSC_TRY(dev_alloc(ctx, (size_t)b * b * 8, &bm));
    CSC_TRY(rot_alloc(ctx, b, &rot));
    CSC_TRY(dev_alloc(ctx, (size_t)b * 4, &rk));
    CSC_TRY(dev_alloc(ctx, (size_t)b * 8, &th));
    CSC_TRY(dev_alloc(ctx, (size_t)b * 8, &res));
    CSC_TRY(dev_alloc(ctx, 4, &sw));
    w.parts = (double*)parts->p;
    w.S = (double*)s->p;
    w.Rinv = (double*)rinv->p;
    double *Q = (double*)buf[0]->p, *T1 = (double*)buf[1]->p, *T2 = (double*)buf[2]->p, *T3 = (double*)buf[3]->p;
    const int64_t nel = (int64_t)H * b;
    std::vector<double> hres((size_t)k), hth((size_t)b);
    bool have_gq = false;   // T3 == Gc Q

    // Q <- Ritz vectors of span(Q) (Q orthonormal on entry), T3 <- Gc Q; Ritz values and residual to the host
    auto rayleigh_ritz = [&]() -> csc_status {
      CSC_TRY(dgemm(ctx, false, Gc, Q, T1, H, b, H, H, b, b));   // Y = Gc Q
      ++iters;
      CSC_TRY(small_gram(ctx, w, Q, T1, 1, (double*)bm->p));     // B = Q'Y
      // the projected problem is solved two digits beyond the residual tolerance
      CSC_TRY(jacobi(ctx, (double*)bm->p, b, rot, (double*)th->p, (int*)rk->p, (int*)sw->p, std::max(1e-13, 0.01 * tol)));
      // Qn = Q W, Yn = Y W = Gc Qn
      CSC_TRY(apply_rot(ctx, Q, T2, T1, T3, H, b, rot, (const int*)sw->p, (const int*)rk->p));
      k_resid<<<k, 256, 0, ctx->stream>>>(T3, T2, (double*)th->p, H, b, k, (double*)res->p);
      CSC_LAUNCHED(ctx);
      CSC_CUDA(ctx, cudaMemcpyAsync(hres.data(), res->p, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CSC_CUDA(ctx, cudaMemcpyAsync(hth.data(), th->p, (size_t)b * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      resid = 0.0;
      for (int i = 0; i < k; ++i) resid = std::max(resid, hres[i]);
      resid /= std::max(std::fabs(hth[0]), 1e-300);
      std::swap(Q, T2);
      have_gq = true;
      return CSC_OK;
    };

    // start: a Gaussian block (well conditioned as it is), two cycles of three plain products + CholeskyQR2
    k_init_block<<<(unsigned)ceil_div64(nel, 256), 256, 0, ctx->stream>>>(Q, nel, 42);
    CSC_LAUNCHED(ctx);
    for (int c = 0; c < 2; ++c) {
      // growth (lambda_1/lambda_b)^3 stays far inside CholeskyQR2's range
      CSC_TRY(dgemm(ctx, false, Gc, Q, T1, H, b, H, H, b, b));
      CSC_TRY(dgemm(ctx, false, Gc, T1, T2, H, b, H, H, b, b));
      CSC_TRY(dgemm(ctx, false, Gc, T2, T1, H, b, H, H, b, b));
      iters += 3;
      CSC_TRY(chol_qr(ctx, w, T1, T2));
      CSC_TRY(chol_qr(ctx, w, T2, Q));
    }
    // Filter cycles.  CholeskyQR orthogonalises the columns in order, so the block behaves like orthogonal (QR)
    // iteration: its leading k columns span the dominant invariant subspace ever better while the columns
    // themselves stay ordered.  That makes the expensive Rayleigh-Ritz step unnecessary inside the loop:
    //   - the filter interval [0, beta] and the growth bound come from the Rayleigh quotients of the columns
    //     (diag of Q'GcQ: beta = the smallest, theta_1 ~ the largest),
    //   - convergence is read off the block residual (I - QQ') Gc q_j of the columns that can carry weight in a
    //     wanted Ritz vector,
    // both by-products of the first product of every cycle.  One Ritz step at the end turns the subspace into
    // eigenpairs and measures the true per-vector residual; if that misses the tolerance the loop resumes with a
    // tighter subspace target.
    double stop_mult = 10.0;
    DevPtr sst;
    CSC_TRY(dev_alloc(ctx, (size_t)2 * b * 8, &sst));
    std::vector<double> hst((size_t)2 * b);
    while (true) {
      while (iters < max_iter) {
        CSC_TRY(dgemm(ctx, false, Gc, Q, T3, H, b, H, H, b, b));      // T3 = Gc Q
        ++iters;
        CSC_TRY(small_gram(ctx, w, Q, T3, 1, (double*)bm->p));        // B = Q' Gc Q
        {
          GemmEpi ep;                                                   // T1 = (I - Q Q') Gc Q = T3 - Q B
          ep.alpha = -1.0;
          ep.g1 = 1.0;
          ep.X1 = T3;
          CSC_TRY(dgemm(ctx, false, Q, (double*)bm->p, T1, H, b, b, b, b, b, ep));
        }
        k_block_stats<<<b, 256, 0, ctx->stream>>>(T1, H, b, (const double*)bm->p, (double*)sst->p);
        CSC_LAUNCHED(ctx);
        CSC_CUDA(ctx, cudaMemcpyAsync(hst.data(), sst->p, (size_t)2 * b * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        // hst[0:b] = column norms of the block residual, hst[b:2b] = Rayleigh quotients of the columns
        double th1 = 1e-300, rmin = 1e300;
        for (int j = 0; j < b; ++j) {
          th1 = std::max(th1, hst[b + j]);
          rmin = std::min(rmin, hst[b + j]);
        }
        // the columns that can carry weight in a wanted Ritz vector: the k with the largest Rayleigh quotients plus at
        // most 8 more (the partners of a near-degenerate k-th eigenvalue); all of them must have stopped leaving the
        // block.  A relative window on the quotients instead would swallow the whole block on a flat spectrum
        // (0.5 % dense VisiumHD-like counts: lambda_k = 1.03 over a bulk at 1.0) and never report convergence.
        std::vector<double> rq(hst.begin() + b, hst.begin() + 2 * b);
        const int kk = std::min(b, k + 8);
        std::nth_element(rq.begin(), rq.begin() + (kk - 1), rq.end(), std::greater<double>());
        std::nth_element(rq.begin(), rq.begin() + (k - 1), rq.begin() + kk, std::greater<double>());
        // ... and of those only the ones within 5 % of the k-th quotient: a clear gap after lambda_k ends the set there
        const double cut = std::max(rq[kk - 1], 0.95 * rq[k - 1]);
        resid = 0.0;
        for (int j = 0; j < b; ++j)
          if (hst[b + j] >= cut) resid = std::max(resid, hst[j]);
        resid /= th1;
        if (resid <= stop_mult * tol) break;
        const double hs[3] = {0.0, rmin, th1};
        const double beta = std::max(hs[1], 1e-10 * th1);
        const double c = 0.5 * beta, e = 0.5 * beta;
        // A filtered guard column (error O(1)) is contaminated by the leading eigenvectors in proportion to
        // T_m(x1)/T_m(1): keep that growth where the Cholesky pivots still have digits (growth^2 * b * eps << 1).
        // The largest Rayleigh quotient underestimates lambda_1 a little: 5 % head-room.
        const double x1 = std::max((1.05 * th1 - c) / e, 1.0 + 1e-12);
        int m = (int)std::floor(std::log(4e5) / std::acosh(x1));
        m = std::max(1, std::min(m, 24));
        // Y0 = Q, Y1 = (Gc Q - c Q)/e
        double *Y0 = Q, *Y1 = T3;
        k_axpby_cols<<<(unsigned)ceil_div64(nel, 256), 256, 0, ctx->stream>>>(T3, Q, 1.0 / e, -c / e, nullptr, nel, b, T3);
        CSC_LAUNCHED(ctx);
        for (int j = 1; j < m; ++j) {
          // Y2 = (2/e) Gc Y1 - (2c/e) Y1 - Y0, written over Y0
          GemmEpi ep;
          ep.alpha = 2.0 / e;
          ep.g1 = -2.0 * c / e;
          ep.X1 = Y1;
          ep.g2 = -1.0;
          ep.X2 = Y0;
          CSC_TRY(dgemm(ctx, false, Gc, Y1, Y0, H, b, H, H, b, b, ep));
          ++iters;
          std::swap(Y0, Y1);
        }
        // Y1 holds the filtered block; one CholeskyQR pass into the other one of {Q, T3} is enough between cycles
        double* dst = (Y1 == T3) ? Q : T3;
        CSC_TRY(chol_qr(ctx, w, Y1, dst));
        if (dst != Q) std::swap(Q, T3);      // Q now names the (nearly) orthonormal block
      }
      // exact orthonormality for the Ritz step, then eigenpairs and the per-vector residual
      CSC_TRY(chol_qr(ctx, w, Q, T1));
      std::swap(Q, T1);
      CSC_TRY(rayleigh_ritz());
      if (resid <= tol || iters >= max_iter) break;
      stop_mult *= 0.1;
    }
    SolveProf::get().report(ctx->stream);
    if (!(resid <= tol))      // also true for a NaN residual
      return csc_fail(ctx, CSC_ERR_NUMERIC, "csc_pca_solve: the eigensolver did not converge: residual %.3e > tol %.3e after "
                      "%d products (max_iter %d)", resid, tol, iters, max_iter);
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
