// Micro-benchmark for the b x b Cholesky + triangular inverse used by the PCA solver's CholeskyQR steps.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/chol_bench tools/micro/chol_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024)
k_chol_inv_ref(const double* __restrict__ S, int b, double* __restrict__ Rinv, long long* stamps) {
  extern __shared__ double sm[];
  const int ld = b + 1;
  double* A = sm;
  double* rs = sm + b * ld;
  double* d0 = rs + b;
  const int t = threadIdx.x, nt = blockDim.x;
  const int lane = t & 31, warp = t >> 5, nwarp = nt >> 5;
  long long t0 = clock64();
  for (int i = t; i < b * b; i += nt) A[(i / b) * ld + (i % b)] = S[i];
  __syncthreads();
  for (int i = t; i < b; i += nt) d0[i] = fabs(A[i * ld + i]);
  __syncthreads();
  long long t1 = clock64();
  const double rel = 4.0 * (double)b * 1.1e-16;
  for (int j = 0; j < b; ++j) {
    const double d = A[j * ld + j];
    const bool ok = d > rel * d0[j] && d > 0.0;
    const double invd = ok ? __drcp_rn(d) : 0.0;
    if (t == 0) rs[j] = ok ? rsqrt(d) : 0.0;
    for (int r = j + 1 + warp; r < b; r += nwarp) {
      const double f = A[j * ld + r] * invd;
      for (int c = r + lane; c < b; c += 32) A[r * ld + c] -= f * A[j * ld + c];
    }
    __syncthreads();
  }
  long long t2 = clock64();
  for (int r = warp; r < b; r += nwarp) {
    const double f = rs[r];
    for (int c = r + lane; c < b; c += 32) A[r * ld + c] *= f;
  }
  __syncthreads();
  double* rd = d0 + b;
  for (int i = t; i < b; i += nt) {
    const double rii = A[i * ld + i];
    rd[i] = rii != 0.0 ? 1.0 / rii : 0.0;
  }
  __syncthreads();
  long long t3 = clock64();
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
  long long t4 = clock64();
  for (int idx = t; idx < b * b; idx += nt) {
    const int r = idx / b, c = idx % b;
    Rinv[idx] = c >= r ? A[c * ld + r] : 0.0;
  }
  __syncthreads();
  long long t5 = clock64();
  if (t == 0) { stamps[0] = t1 - t0; stamps[1] = t2 - t1; stamps[2] = t3 - t2; stamps[3] = t4 - t3; stamps[4] = t5 - t4; }
}

// Candidate: panel-blocked Cholesky (R'R = S, R upper, panels of NB rows) + recursive-doubling inverse of R.
// X = inv(R) is kept transposed in the strict lower triangle (X[i][c] at A[c][i], i < c) with its diagonal in rd[].
constexpr int NB = 16;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// One warp factors the NB x NB diagonal block at (j0, j0) in registers (lane c = column c of the block) and writes
// R11 (upper, scaled) back plus rd = 1/R_jj.  Unscaled right-looking steps: a[r] -= (A[j][r] / d_j) A[j][c]; the
// row scalings 1/sqrt(d_j) are off the dependency chain.
__device__ __forceinline__ void chol_diag_block(double* A, int ld, int j0, int nbk, const double* d0, double rel,
                                                double* rd, int lane) {
  double a[NB], rs[NB];
  const int c = lane & (NB - 1);
#pragma unroll
  for (int r = 0; r < NB; ++r) a[r] = (r <= c && c < nbk && r < nbk) ? A[(j0 + r) * ld + j0 + c] : (r == c ? 1.0 : 0.0);
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const double d = shfl_d(a[j], j);
    const double dj0 = j < nbk ? d0[j0 + j] : 1.0;
    const bool ok = d > rel * dj0 && d > 0.0;
    const double invd = ok ? __drcp_rn(d) : 0.0;
    rs[j] = ok ? rsqrt(d) : 0.0;
    const double f = a[j] * invd;          // A[j][c] / d
#pragma unroll
    for (int r = j + 1; r < NB; ++r) a[r] = fma(-shfl_d(a[j], r), f, a[r]);
  }
  if (lane < NB && c < nbk) {
#pragma unroll
    for (int r = 0; r < NB; ++r)
      if (r <= c) A[(j0 + r) * ld + j0 + c] = a[r] * rs[r];
    rd[j0 + c] = rs[c];
  }
}

// One warp inverts the NB x NB upper-triangular diagonal block at (j0, j0): lane c owns column c of X, the rows of
// R come by broadcast shuffles.  X goes to the strict lower triangle (transposed).
__device__ __forceinline__ void inv_diag_block(double* A, int ld, int j0, int nbk, const double* rd, int lane) {
  double rcol[NB], x[NB];
  const int c = lane & (NB - 1);
#pragma unroll
  for (int r = 0; r < NB; ++r) rcol[r] = (r < c && c < nbk) ? A[(j0 + r) * ld + j0 + c] : 0.0;   // R[r][c], strict upper
  const double rdc = c < nbk ? rd[j0 + c] : 0.0;
#pragma unroll
  for (int i = NB - 1; i >= 0; --i) {
    double acc = 0.0;
#pragma unroll
    for (int l = i + 1; l < NB; ++l) acc = fma(shfl_d(rcol[i], l), x[l], acc);      // R[i][l] x[l]
    const double rdi = shfl_d(rdc, i);
    x[i] = (i == c) ? rdc : (i < c ? -acc * rdi : 0.0);
  }
  if (lane < NB && c < nbk) {
#pragma unroll
    for (int i = 0; i < NB; ++i)
      if (i < c) A[(j0 + c) * ld + j0 + i] = x[i];
  }
}

__global__ void __launch_bounds__(512)
k_chol_inv_new(const double* __restrict__ S, int b, double* __restrict__ Rinv, int tld, long long* stamps) {
  extern __shared__ double sm[];
  const int ld = b + 1;
  double* A = sm;              // [b][ld]
  double* rd = A + b * ld;     // [b] 1 / R_ii (0 for a dropped column)
  double* d0 = rd + b;         // [b] original diagonal
  double* T = d0 + b;          // temp of the inverse levels, [s][tld]
  const int t = threadIdx.x, nt = blockDim.x;
  const int lane = t & 31, warp = t >> 5, nwarp = nt >> 5;
  long long t0 = clock64();
  for (int i = t; i < b * b; i += nt) A[(i / b) * ld + (i % b)] = S[i];
  __syncthreads();
  for (int i = t; i < b; i += nt) d0[i] = fabs(A[i * ld + i]);
  __syncthreads();
  long long t1 = clock64();
  const double rel = 4.0 * (double)b * 1.1e-16;
  long long pa = 0, pb = 0;
  for (int j0 = 0; j0 < b; j0 += NB) {
    const int j1 = min(j0 + NB, b);
    long long q0 = clock64();
    if (warp == 0) chol_diag_block(A, ld, j0, j1 - j0, d0, rel, rd, lane);
    __syncthreads();
    long long q1 = clock64();
    // the last warp inverts the finished diagonal block while the others solve the panel rows:
    // R12 = R11^-T A12, one thread per trailing column
    if (warp == nwarp - 1) {
      inv_diag_block(A, ld, j0, j1 - j0, rd, lane);
    } else {
      for (int c = j1 + t; c < b; c += nt - 32) {
        double v[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
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
    long long q2 = clock64();
    pa += q1 - q0; pb += q2 - q1;
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
  long long t2 = clock64();
  long long t3 = t2;
  // doubling: X_PQ = -X_PP (R_PQ X_QQ) for neighbouring blocks P = [p0, p0+s), Q = [p0+s, min(p0+2s, b))
  for (int s = NB; s < b; s <<= 1) {
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
  long long t4 = clock64();
  for (int idx = t; idx < b * b; idx += nt) {
    const int r = idx / b, c = idx % b;
    Rinv[idx] = c > r ? A[c * ld + r] : (c == r ? rd[r] : 0.0);
  }
  __syncthreads();
  long long t5 = clock64();
  if (t == 0) { stamps[0] = t1 - t0; stamps[1] = t2 - t1; stamps[2] = pa; stamps[3] = t4 - t3; stamps[4] = pb; }
}

int main(int argc, char** argv) {
  const int b = argc > 1 ? atoi(argv[1]) : 58;
  std::vector<double> M(b * b), S(b * b, 0.0);
  srand(1);
  for (auto& x : M) x = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < b; ++j) {
      double a = 0;
      for (int k = 0; k < b; ++k) a += M[k * b + i] * M[k * b + j];
      S[i * b + j] = a + (i == j ? 0.1 : 0.0);
    }
  const double span = argc > 2 ? atof(argv[2]) : 0.0;   // decades of column scaling (filtered blocks: ~5.6)
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < b; ++j) S[i * b + j] *= pow(10.0, -span * i / b) * pow(10.0, -span * j / b);
  double *dS, *dR; long long* dst;
  cudaMalloc(&dS, b * b * 8); cudaMalloc(&dR, b * b * 8); cudaMalloc(&dst, 64);
  cudaMemcpy(dS, S.data(), b * b * 8, cudaMemcpyHostToDevice);
  std::vector<double> R0(b * b), R1(b * b);
  long long st[5];
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto check = [&](const std::vector<double>& Ri, const char* name, float us) {
    // S * Ri * Ri' should be I  (inv(S) = inv(R) inv(R)')
    double worst = 0;
    std::vector<double> T(b * b, 0.0);
    for (int i = 0; i < b; ++i) for (int j = 0; j < b; ++j) { double a = 0; for (int k = 0; k < b; ++k) a += Ri[i * b + k] * Ri[j * b + k]; T[i * b + j] = a; }
    for (int i = 0; i < b; ++i) for (int j = 0; j < b; ++j) { double a = 0; for (int k = 0; k < b; ++k) a += S[i * b + k] * T[k * b + j]; worst = fmax(worst, fabs(a - (i == j))); }
    printf("%-22s %8.1f us  |S inv - I| %.2e  cycles: load %lld chol %lld mid %lld inv %lld store %lld\n", name, us, worst, st[0], st[1], st[2], st[3], st[4]);
  };
  {
    size_t smem = (size_t)(b * (b + 1) + 3 * b) * 8;
    cudaFuncSetAttribute(k_chol_inv_ref, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_chol_inv_ref<<<1, 1024, smem>>>(dS, b, dR, dst);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(R0.data(), dR, b * b * 8, cudaMemcpyDeviceToHost); cudaMemcpy(st, dst, 40, cudaMemcpyDeviceToHost);
    check(R0, "ref 1024 thr", ms * 1e3f);
  }
  {
    // T holds npair * s rows of tld columns at the widest level
    int trows = 0, tld = 0;
    for (int sz = NB; sz < b; sz <<= 1) {
      const int npair = (b + 2 * sz - 1) / (2 * sz);
      trows = std::max(trows, npair * sz);
      tld = std::max(tld, sz + 1);
    }
    size_t smem = (size_t)(b * (b + 1) + 2 * b + trows * tld) * 8;
    printf("new: smem %zu bytes (T %d x %d)\n", smem, trows, tld);
    cudaFuncSetAttribute(k_chol_inv_new, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_chol_inv_new<<<1, 512, smem>>>(dS, b, dR, tld, dst);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(R1.data(), dR, b * b * 8, cudaMemcpyDeviceToHost); cudaMemcpy(st, dst, 40, cudaMemcpyDeviceToHost);
    check(R1, "new blocked", ms * 1e3f);
    double dmax = 0, rmax = 0; for (int i = 0; i < b * b; ++i) { dmax = fmax(dmax, fabs(R1[i] - R0[i])); rmax = fmax(rmax, fabs(R0[i])); }
    printf("   max |new - ref| %.2e of max |ref| %.2e (cuda err %s)\n", dmax, rmax, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
