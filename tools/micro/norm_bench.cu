// Micro-benchmark for the fused normalise + moments pass (sparse.cu: k_normalize): which part of the per-entry work
// bounds it.  Variants of one warp-per-cell kernel over a synthetic CSC (sorted random genes, small counts):
//   structure: S = scalar two-pass loops (the shipped kernel), V = 16-byte loads, the column held in registers
//   flags    : which shared-memory atomics run, fast or exact logarithm
// and a plain shared-memory atomic rate test (random addresses in a 20k-entry table).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o norm_bench norm_bench.cu && ./norm_bench
#include <cuda_runtime.h>
#include <thrust/device_ptr.h>
#include <thrust/scan.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) {                                                                  \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
      exit(1);                                                                                \
    }                                                                                         \
  } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__global__ void k_count(int64_t N, int G, uint32_t thr, int64_t* cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= N) return;
  int n = 0;
  for (int g = lane; g < G; g += 32) n += hash32((uint32_t)c * 0x9E3779B9u + g * 0x85EBCA6Bu) < thr;
  n = __reduce_add_sync(0xffffffffu, n);
  if (lane == 0) cnt[c] = n;
}
__global__ void k_fill(int64_t N, int G, uint32_t thr, const int64_t* colptr, int32_t* rv, float* val) {
  const int lane = threadIdx.x & 31;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= N) return;
  int64_t o = colptr[c];
  for (int g0 = 0; g0 < G; g0 += 32) {
    const int g = g0 + lane;
    const uint32_t h = hash32((uint32_t)c * 0x9E3779B9u + g * 0x85EBCA6Bu);
    const bool p = g < G && h < thr;
    const unsigned m = __ballot_sync(0xffffffffu, p);
    if (p) {
      const int64_t i = o + __popc(m & ((1u << lane) - 1));
      rv[i] = g;
      const uint32_t r = hash32(h + 77u) & 255u;     // counts: 1 (60 %), 2 (20 %), 3..6, a few large
      val[i] = r < 154 ? 1.f : r < 205 ? 2.f : r < 250 ? (float)(3 + (r & 3)) : (float)(10 + (r & 63));
    }
    o += __popc(m);
  }
}

struct Args {
  const int64_t* colptr;
  const int32_t* rowval;
  const float* raw;
  float* out;
  int64_t n_cells;
  int32_t n_genes;
  double sf;
  double* gstats;
  double* nstats;
  int64_t cells_per_cta;
  float ymax, s1, s2;
  double inv_s1, inv_s2;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

enum { F_S1 = 1, F_E = 2, F_Y = 4, F_FASTLOG = 8, F_WRITE = 16 };

template <int FL>
__device__ __forceinline__ float lognorm(float t) {
  if (FL & F_FASTLOG) {
    // log(1 + t): above 0.5 the rounding of 1 + t costs < 2^-24 / log(1.5) relative; below, the series argument matters
    return t >= 0.5f ? __logf(1.f + t) : log1pf(t);
  }
  return log1pf(t);
}

template <int FL>
__device__ __forceinline__ void entry_update(const Args& a, unsigned* s1tab, unsigned* etab, int* y1tab, unsigned* y2tab, int g,
                                             float x, float y) {
  if (FL & (F_S1 | F_E)) {
    const float xr = rintf(x);
    if (xr == x && x >= 0.f && x <= 65535.f) {
      const unsigned xi = (unsigned)x;
      if (FL & F_S1) atomicAdd(&s1tab[g], xi);
      if (FL & F_E) {
        if (xi <= 255u) {
          if (xi >= 2u) atomicAdd(&etab[g], xi * xi - xi);
        } else {
          const double xd = (double)x;
          atomicAdd(&a.gstats[(int64_t)a.n_genes + g], xd * xd - xd);
        }
      }
    } else {
      const double xd = (double)x;
      atomicAdd(&a.gstats[g], xd);
      atomicAdd(&a.gstats[(int64_t)a.n_genes + g], xd * xd);
    }
  }
  if (FL & F_Y) {
    if (fabsf(y) <= a.ymax) {
      atomicAdd(&y1tab[g], __float2int_rn(y * a.s1));
      atomicAdd(&y2tab[g], __float2uint_rn(y * y * a.s2));
    } else {
      const double yd = (double)y;
      atomicAdd(&a.nstats[g], yd);
      atomicAdd(&a.nstats[(int64_t)a.n_genes + g], yd * yd);
    }
  }
}

template <int FL>
__device__ __forceinline__ void flush_tables(const Args& a, unsigned* s1tab, unsigned* etab, int* y1tab, unsigned* y2tab) {
  const int G = a.n_genes;
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const unsigned v = s1tab[g], ex = etab[g];
    if (v) atomicAdd(&a.gstats[g], (double)v);
    if (v || ex) atomicAdd(&a.gstats[(int64_t)G + g], (double)v + (double)ex);
    const int v1 = y1tab[g];
    const unsigned v2 = y2tab[g];
    if (v1) atomicAdd(&a.nstats[g], (double)v1 * a.inv_s1);
    if (v2) atomicAdd(&a.nstats[(int64_t)G + g], (double)v2 * a.inv_s2);
  }
}

// ---- S: the shipped structure
template <int FL>
__global__ void __launch_bounds__(768, 2) k_norm_S(const Args a) {
  extern __shared__ unsigned int stab[];
  const int G = a.n_genes;
  unsigned* s1tab = stab;
  unsigned* etab = stab + G;
  int* y1tab = (int*)(stab + 2 * G);
  unsigned* y2tab = stab + 3 * G;
  for (int g = threadIdx.x; g < 4 * G; g += blockDim.x) stab[g] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * a.cells_per_cta;
  const int64_t c1 = min(c0 + a.cells_per_cta, a.n_cells);
  for (int64_t c = c0 + warp; c < c1; c += nwarp) {
    const int64_t b = a.colptr[c], e = a.colptr[c + 1];
    double s = 0.0;
#pragma unroll 4
    for (int64_t i = b + lane; i < e; i += 32) s += (double)a.raw[i];
    s = warp_sum(s);
    const float scale = (float)((1.0 / s) * a.sf);
#pragma unroll 4
    for (int64_t i = b + lane; i < e; i += 32) {
      const float x = a.raw[i];
      const float y = lognorm<FL>(x * scale);
      if (FL & F_WRITE) __stcs(a.out + i, y);
      if (FL & (F_S1 | F_E | F_Y)) {
        const int g = __ldcs(a.rowval + i);
        entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g, x, y);
      }
    }
  }
  flush_tables<FL>(a, s1tab, etab, y1tab, y2tab);
}

// ---- V: 16-byte loads from the 4-aligned index below the column start, column kept in registers (<= 128 * NIT entries)
template <int FL, int NIT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_norm_V(const Args a) {
  extern __shared__ unsigned int stab[];
  const int G = a.n_genes;
  unsigned* s1tab = stab;
  unsigned* etab = stab + G;
  int* y1tab = (int*)(stab + 2 * G);
  unsigned* y2tab = stab + 3 * G;
  for (int g = threadIdx.x; g < 4 * G; g += blockDim.x) stab[g] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * a.cells_per_cta;
  const int64_t c1 = min(c0 + a.cells_per_cta, a.n_cells);
  int64_t b = 0, e = 0;
  if (c0 + warp < c1) { b = a.colptr[c0 + warp]; e = a.colptr[c0 + warp + 1]; }
  for (int64_t c = c0 + warp; c < c1; c += nwarp) {
    // next column's bounds early (hides the dependent colptr load)
    int64_t nb = 0, ne = 0;
    if (c + nwarp < c1) { nb = a.colptr[c + nwarp]; ne = a.colptr[c + nwarp + 1]; }
    const int64_t base = (b & ~(int64_t)3) + 4 * lane;
    if (e - (b & ~(int64_t)3) <= 128 * NIT) {
      float4 v[NIT];
      double s = 0.0;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int64_t i = base + 128 * it;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < e) {
          v[it] = __ldcs(reinterpret_cast<const float4*>(a.raw + i));
          if (i < b) { if (i + 0 < b) v[it].x = 0.f; if (i + 1 < b) v[it].y = 0.f; if (i + 2 < b) v[it].z = 0.f; }
          if (i + 3 >= e) { if (i + 1 >= e) v[it].y = 0.f; if (i + 2 >= e) v[it].z = 0.f; v[it].w = 0.f; }
          s += ((double)v[it].x + (double)v[it].y) + ((double)v[it].z + (double)v[it].w);
        }
      }
      s = warp_sum(s);
      const float scale = (float)((1.0 / s) * a.sf);
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int64_t i = base + 128 * it;
        if (i < e) {
          float4 y;
          y.x = lognorm<FL>(v[it].x * scale);
          y.y = lognorm<FL>(v[it].y * scale);
          y.z = lognorm<FL>(v[it].z * scale);
          y.w = lognorm<FL>(v[it].w * scale);
          const bool full = i >= b && i + 3 < e;
          if (FL & F_WRITE) {
            if (full) {
              __stcs(reinterpret_cast<float4*>(a.out + i), y);
            } else {
              if (i + 0 >= b && i + 0 < e) __stcs(a.out + i + 0, y.x);
              if (i + 1 >= b && i + 1 < e) __stcs(a.out + i + 1, y.y);
              if (i + 2 >= b && i + 2 < e) __stcs(a.out + i + 2, y.z);
              if (i + 3 >= b && i + 3 < e) __stcs(a.out + i + 3, y.w);
            }
          }
          if (FL & (F_S1 | F_E | F_Y)) {
            const int4 g = __ldcs(reinterpret_cast<const int4*>(a.rowval + i));
            if (full) {
              entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g.x, v[it].x, y.x);
              entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g.y, v[it].y, y.y);
              entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g.z, v[it].z, y.z);
              entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g.w, v[it].w, y.w);
            } else {
              if (i + 0 >= b && i + 0 < e) entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g.x, v[it].x, y.x);
              if (i + 1 >= b && i + 1 < e) entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g.y, v[it].y, y.y);
              if (i + 2 >= b && i + 2 < e) entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g.z, v[it].z, y.z);
              if (i + 3 >= b && i + 3 < e) entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g.w, v[it].w, y.w);
            }
          }
        }
      }
    } else {
      double s = 0.0;
      for (int64_t i = b + lane; i < e; i += 32) s += (double)a.raw[i];
      s = warp_sum(s);
      const float scale = (float)((1.0 / s) * a.sf);
      for (int64_t i = b + lane; i < e; i += 32) {
        const float x = a.raw[i];
        const float y = lognorm<FL>(x * scale);
        if (FL & F_WRITE) __stcs(a.out + i, y);
        if (FL & (F_S1 | F_E | F_Y)) entry_update<FL>(a, s1tab, etab, y1tab, y2tab, __ldcs(a.rowval + i), x, y);
      }
    }
    b = nb;
    e = ne;
  }
  flush_tables<FL>(a, s1tab, etab, y1tab, y2tab);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// prefetch the lines of one column (values and row indices) into L2: lanes 0-15 the values, 16-31 the indices
__device__ __forceinline__ void prefetch_column(const Args& a, int64_t b, int64_t e, int lane) {
  const char* base = lane < 16 ? reinterpret_cast<const char*>(a.raw) : reinterpret_cast<const char*>(a.rowval);
  const int64_t lo = (b * 4) & ~(int64_t)127, hi = e * 4;
  for (int64_t off = lo + 128 * (lane & 15); off < hi; off += 128 * 16) prefetch_l2(base + off);
}

// ---- SP: scalar structure, loads of a group of UN entries issued before their updates, next column prefetched to L2
template <int FL, int UN, int THREADS, int MINB, bool PF, bool PFROW>
__global__ void __launch_bounds__(THREADS, MINB) k_norm_SP(const Args a) {
  extern __shared__ unsigned int stab[];
  const int G = a.n_genes;
  unsigned* s1tab = stab;
  unsigned* etab = stab + G;
  int* y1tab = (int*)(stab + 2 * G);
  unsigned* y2tab = stab + 3 * G;
  for (int g = threadIdx.x; g < 4 * G; g += blockDim.x) stab[g] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * a.cells_per_cta;
  const int64_t c1 = min(c0 + a.cells_per_cta, a.n_cells);
  int64_t b = 0, e = 0, nb = 0, ne = 0;
  if (c0 + warp < c1) { b = a.colptr[c0 + warp]; e = a.colptr[c0 + warp + 1]; }
  if (c0 + warp + nwarp < c1) { nb = a.colptr[c0 + warp + nwarp]; ne = a.colptr[c0 + warp + nwarp + 1]; }
  for (int64_t c = c0 + warp; c < c1; c += nwarp) {
    int64_t nnb = 0, nne = 0;
    if (c + 2 * nwarp < c1) { nnb = a.colptr[c + 2 * nwarp]; nne = a.colptr[c + 2 * nwarp + 1]; }
    if (PF && nb < ne) prefetch_column(a, nb, ne, lane);
    if (PFROW && !PF) {   // only this column's row indices, during its own first pass
      for (int64_t off = ((b * 4) & ~(int64_t)127) + 128 * lane; off < e * 4; off += 128 * 32)
        prefetch_l2(reinterpret_cast<const char*>(a.rowval) + off);
    }
    double s = 0.0;
#pragma unroll 4
    for (int64_t i = b + lane; i < e; i += 32) s += (double)a.raw[i];
    s = warp_sum(s);
    const float scale = (float)((1.0 / s) * a.sf);
    for (int64_t i0 = b + lane; i0 < e; i0 += 32 * UN) {
      float x[UN];
      int g[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int64_t i = i0 + 32 * u;
        x[u] = i < e ? a.raw[i] : 0.f;
        g[u] = i < e ? __ldcs(a.rowval + i) : 0;
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int64_t i = i0 + 32 * u;
        if (i < e) {
          const float y = lognorm<FL>(x[u] * scale);
          if (FL & F_WRITE) __stcs(a.out + i, y);
          entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g[u], x[u], y);
        }
      }
    }
    b = nb; e = ne; nb = nnb; ne = nne;
  }
  flush_tables<FL>(a, s1tab, etab, y1tab, y2tab);
}

// ---- VP: 16-byte loads of values AND row indices into registers before any update; next column prefetched to L2
template <int FL, int NIT, int THREADS, int MINB, bool PF>
__global__ void __launch_bounds__(THREADS, MINB) k_norm_VP(const Args a) {
  extern __shared__ unsigned int stab[];
  const int G = a.n_genes;
  unsigned* s1tab = stab;
  unsigned* etab = stab + G;
  int* y1tab = (int*)(stab + 2 * G);
  unsigned* y2tab = stab + 3 * G;
  for (int g = threadIdx.x; g < 4 * G; g += blockDim.x) stab[g] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * a.cells_per_cta;
  const int64_t c1 = min(c0 + a.cells_per_cta, a.n_cells);
  int64_t b = 0, e = 0, nb = 0, ne = 0;
  if (c0 + warp < c1) { b = a.colptr[c0 + warp]; e = a.colptr[c0 + warp + 1]; }
  if (c0 + warp + nwarp < c1) { nb = a.colptr[c0 + warp + nwarp]; ne = a.colptr[c0 + warp + nwarp + 1]; }
  for (int64_t c = c0 + warp; c < c1; c += nwarp) {
    int64_t nnb = 0, nne = 0;
    if (c + 2 * nwarp < c1) { nnb = a.colptr[c + 2 * nwarp]; nne = a.colptr[c + 2 * nwarp + 1]; }
    if (PF && nb < ne) prefetch_column(a, nb, ne, lane);
    const int64_t al = b & ~(int64_t)3;
    // pass 1: column sum (16-byte loads; the tail of the previous column and the head of the next are masked)
    double s = 0.0;
    for (int64_t i = al + 4 * lane; i < e; i += 128) {
      float4 v = *reinterpret_cast<const float4*>(a.raw + i);
      if (i < b) { if (i + 0 < b) v.x = 0.f; if (i + 1 < b) v.y = 0.f; if (i + 2 < b) v.z = 0.f; }
      if (i + 3 >= e) { if (i + 1 >= e) v.y = 0.f; if (i + 2 >= e) v.z = 0.f; v.w = 0.f; }
      s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    }
    s = warp_sum(s);
    const float scale = (float)((1.0 / s) * a.sf);
    // pass 2: groups of NIT 16-byte loads of values (L1/L2 hits) and indices, then the updates
    for (int64_t i0 = al + 4 * lane; i0 < e; i0 += 128 * NIT) {
      float4 v[NIT];
      int4 g[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int64_t i = i0 + 128 * it;
        if (i < e) {
          v[it] = *reinterpret_cast<const float4*>(a.raw + i);
          g[it] = __ldcs(reinterpret_cast<const int4*>(a.rowval + i));
        }
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int64_t i = i0 + 128 * it;
        if (i < e) {
          float4 y;
          y.x = lognorm<FL>(v[it].x * scale);
          y.y = lognorm<FL>(v[it].y * scale);
          y.z = lognorm<FL>(v[it].z * scale);
          y.w = lognorm<FL>(v[it].w * scale);
          if (i >= b && i + 3 < e) {
            if (FL & F_WRITE) __stcs(reinterpret_cast<float4*>(a.out + i), y);
            entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g[it].x, v[it].x, y.x);
            entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g[it].y, v[it].y, y.y);
            entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g[it].z, v[it].z, y.z);
            entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g[it].w, v[it].w, y.w);
          } else {
            if (i + 0 >= b && i + 0 < e) { if (FL & F_WRITE) __stcs(a.out + i + 0, y.x); entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g[it].x, v[it].x, y.x); }
            if (i + 1 >= b && i + 1 < e) { if (FL & F_WRITE) __stcs(a.out + i + 1, y.y); entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g[it].y, v[it].y, y.y); }
            if (i + 2 >= b && i + 2 < e) { if (FL & F_WRITE) __stcs(a.out + i + 2, y.z); entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g[it].z, v[it].z, y.z); }
            if (i + 3 >= b && i + 3 < e) { if (FL & F_WRITE) __stcs(a.out + i + 3, y.w); entry_update<FL>(a, s1tab, etab, y1tab, y2tab, g[it].w, v[it].w, y.w); }
          }
        }
      }
    }
    b = nb; e = ne; nb = nnb; ne = nne;
  }
  flush_tables<FL>(a, s1tab, etab, y1tab, y2tab);
}

// ---- raw shared-memory atomic rate: NA atomics per lane per round on random table entries
template <int NA, bool CONFLICT_FREE = false>
__global__ void __launch_bounds__(768, 2) k_atoms(unsigned* sink, int G, int rounds) {
  extern __shared__ unsigned int stab[];
  for (int g = threadIdx.x; g < 4 * G; g += blockDim.x) stab[g] = 0u;
  __syncthreads();
  uint32_t h = hash32(blockIdx.x * 1024u + threadIdx.x);
  for (int r = 0; r < rounds; ++r) {
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      h = h * 1664525u + 1013904223u;
      unsigned a = (h >> 8) % (unsigned)G;
      if (CONFLICT_FREE) a = min((a & ~31u) | (threadIdx.x & 31u), (unsigned)G - 1u);   // lane = bank: no conflicts in a warp
      atomicAdd(&stab[a + (k & 3) * G], h & 3u);
    }
  }
  __syncthreads();
  unsigned acc = 0;
  for (int g = threadIdx.x; g < 4 * G; g += blockDim.x) acc += stab[g];
  if (acc == 0xdeadbeefu) sink[0] = acc;
}

static double checksum(const double* d, int n) {
  std::vector<double> h(n);
  CK(cudaMemcpy(h.data(), d, n * sizeof(double), cudaMemcpyDeviceToHost));
  double s = 0;
  for (int i = 0; i < n; ++i) s += h[i] * (1.0 + (i % 7));
  return s;
}

int main(int argc, char** argv) {
  const int G = argc > 1 ? atoi(argv[1]) : 5000;
  const int64_t N = argc > 2 ? atoll(argv[2]) : 500000;
  const double dens = argc > 3 ? atof(argv[3]) : 0.05;
  int sm = 0;
  CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
  int64_t *cnt, *colptr;
  CK(cudaMalloc(&cnt, (N + 1) * 8));
  CK(cudaMalloc(&colptr, (N + 1) * 8));
  CK(cudaMemset(cnt, 0, (N + 1) * 8));
  const uint32_t thr = (uint32_t)(dens * 4294967296.0);
  k_count<<<(unsigned)((N * 32 + 255) / 256), 256>>>(N, G, thr, cnt);
  thrust::exclusive_scan(thrust::device_pointer_cast(cnt), thrust::device_pointer_cast(cnt) + N + 1,
                         thrust::device_pointer_cast(colptr));
  int64_t nnz = 0;
  CK(cudaMemcpy(&nnz, colptr + N, 8, cudaMemcpyDeviceToHost));
  int32_t* rv;
  float *val, *out;
  CK(cudaMalloc(&rv, nnz * 4 + 512));
  CK(cudaMalloc(&val, nnz * 4 + 512));
  CK(cudaMalloc(&out, nnz * 4 + 512));
  k_fill<<<(unsigned)((N * 32 + 255) / 256), 256>>>(N, G, thr, colptr, rv, val);
  CK(cudaDeviceSynchronize());
  printf("G=%d N=%lld nnz=%lld sm=%d\n", G, (long long)N, (long long)nnz, sm);
  double *gst, *nst;
  CK(cudaMalloc(&gst, 2 * G * 8));
  CK(cudaMalloc(&nst, 2 * G * 8));
  void* flush;
  const size_t flush_bytes = 256u << 20;
  CK(cudaMalloc(&flush, flush_bytes));

  Args a;
  a.colptr = colptr; a.rowval = rv; a.raw = val; a.out = out; a.n_cells = N; a.n_genes = G; a.sf = 1e4;
  a.gstats = gst; a.nstats = nst;
  const double ymax = log(1e4 + 1.0) * 1.0001;
  a.ymax = (float)ymax;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const size_t smem = (size_t)4 * G * 4;
  const double bytes = 12.0 * nnz + 8.0 * N;

  auto run = [&](const char* name, auto kern, int threads, int per_sm) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = (int64_t)sm * per_sm;
    int64_t cpc = (N + grid - 1) / grid;
    grid = (N + cpc - 1) / cpc;
    a.cells_per_cta = cpc;
    const double s1 = 0.98 * 2147483648.0 / ((double)cpc * a.ymax);
    const double s2 = 0.98 * 4294967296.0 / ((double)cpc * (double)a.ymax * a.ymax);
    a.s1 = (float)s1; a.s2 = (float)s2; a.inv_s1 = 1.0 / (double)a.s1; a.inv_s2 = 1.0 / (double)a.s2;
    float best = 1e9f;
    double cs_g = 0, cs_n = 0;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaMemset(gst, 0, 2 * G * 8));
      CK(cudaMemset(nst, 0, 2 * G * 8));
      CK(cudaMemset(flush, rep, flush_bytes));
      CK(cudaEventRecord(e0));
      kern<<<(unsigned)grid, threads, smem>>>(a);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
      cs_g = checksum(gst, 2 * G);
      cs_n = checksum(nst, 2 * G);
    }
    std::vector<float> ho(1024);
    CK(cudaMemcpy(ho.data(), out + nnz / 2, 1024 * 4, cudaMemcpyDeviceToHost));
    double so = 0;
    for (float f : ho) so += f;
    printf("%-34s %7.3f ms  %6.0f GB/s (%.2f)  chk raw %.6e norm %.9e out %.6f\n", name, best, bytes / best / 1e6,
           bytes / best / 1e6 / 6531.0, cs_g, cs_n, so);
    fflush(stdout);
  };
  constexpr int ALL = F_S1 | F_E | F_Y | F_WRITE;
  run("S all (shipped)", k_norm_S<ALL>, 768, 2);
  run("S no atomics", k_norm_S<F_WRITE>, 768, 2);
  constexpr int AF = ALL | F_FASTLOG;
  run("SP un4 768x2 pf", k_norm_SP<ALL, 4, 768, 2, true, false>, 768, 2);
  run("SP un4 768x2 pf fastlog", k_norm_SP<AF, 4, 768, 2, true, false>, 768, 2);
  run("SP un4 768x2 pf noatom", k_norm_SP<F_WRITE, 4, 768, 2, true, false>, 768, 2);
  // atomic rate
  unsigned* sink;
  CK(cudaMalloc(&sink, 4));
  auto atoms = [&](const char* name, auto kern, int na) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int rounds = 2000 / na;
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      kern<<<sm * 2, 768, smem>>>(sink, G, rounds);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    const double n_atom = (double)sm * 2 * 768 * rounds * na;
    printf("%-20s %.3f ms: %.2f atomics/ns chip, %.2f lanes/clk/SM @1.9GHz\n", name, best, n_atom / best / 1e6,
           n_atom / best / 1e6 / sm / 1.9);
  };
  atoms("ATOMS x1/round", k_atoms<1>, 1);
  atoms("ATOMS x4/round", k_atoms<4>, 4);
  atoms("ATOMS x16/round", k_atoms<16>, 16);
  atoms("ATOMS x4 conflict-free", k_atoms<4, true>, 4);
  return 0;
}
