// Locality order for the exact kNN scan.
//
// The brute-force scan is exact in any key order, but its cost is not order-free: a row's running top-L list
// is updated whenever a key beats the current L-th best, which for keys in random order happens ~L ln(N/L)
// times per row -- and list updates, not the tensor-core scores, dominate the scan (measured: 57 of 95 ms at
// 500k cells).  If a row meets its true neighbours FIRST, its threshold is tight from the start and almost
// nothing that follows gets through.  So: cluster the unit rows (spherical k-means, ~N/2048 clusters, a few Lloyd
// passes on a strided sample), sort keys and queries by cluster, and let every query block start its scan a few
// tiles before its own cluster.  Pure reordering: results are identical, ids are mapped back at the end.
#include <cmath>
#include <cub/cub.cuh>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int KO_D = 64;          // prepared row length (floats)
constexpr int KO_CHUNK = 128;     // centroids staged in shared memory at a time (32 KB)

// label[i] = argmax_c <x_i, cent_c>; optionally accumulates the per-cluster sums of the assigned points
// (sample stride `stride`: point i is row i * stride).  The scores run in packed fp16 (HFMA2: rows and centroids are
// unit vectors, 16 products per accumulator half): the labels only steer the ORDER of the scan, which is exact in any
// order, and half the instructions of the fp32 loop (0.63 -> 0.35 ms for the pass over 500k rows).
__global__ void __launch_bounds__(128)
k_km_assign(const float* __restrict__ x, int64_t n, int64_t stride, const float* __restrict__ cent, int K, int32_t* __restrict__ label,
            float* __restrict__ sums, int32_t* __restrict__ counts) {
  __shared__ __align__(16) __half2 sc[KO_CHUNK * KO_D / 2];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  const float4* src = reinterpret_cast<const float4*>(x + (valid ? i * stride : 0) * KO_D);
  __half2 xh[KO_D / 2];
#pragma unroll
  for (int j = 0; j < KO_D / 4; ++j) {
    const float4 v = __ldg(src + j);
    xh[2 * j] = __floats2half2_rn(v.x, v.y);
    xh[2 * j + 1] = __floats2half2_rn(v.z, v.w);
  }
  float best = -2.f;
  int bi = 0;
  for (int c0 = 0; c0 < K; c0 += KO_CHUNK) {
    const int nc = min(KO_CHUNK, K - c0);
    __syncthreads();
    for (int j = threadIdx.x; j < nc * KO_D / 4; j += blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(cent + (int64_t)c0 * KO_D) + j);
      sc[2 * j] = __floats2half2_rn(v.x, v.y);
      sc[2 * j + 1] = __floats2half2_rn(v.z, v.w);
    }
    __syncthreads();
    for (int c = 0; c < nc; ++c) {
      const uint4* cr = reinterpret_cast<const uint4*>(sc + c * (KO_D / 2));   // same address for the whole warp: broadcast
      __half2 a0 = __float2half2_rn(0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
      for (int j = 0; j < KO_D / 8; ++j) {
        const uint4 v = cr[j];
        a0 = __hfma2(xh[4 * j], *reinterpret_cast<const __half2*>(&v.x), a0);
        a1 = __hfma2(xh[4 * j + 1], *reinterpret_cast<const __half2*>(&v.y), a1);
        a2 = __hfma2(xh[4 * j + 2], *reinterpret_cast<const __half2*>(&v.z), a2);
        a3 = __hfma2(xh[4 * j + 3], *reinterpret_cast<const __half2*>(&v.w), a3);
      }
      const float2 f = __half22float2(__hadd2(__hadd2(a0, a1), __hadd2(a2, a3)));
      const float sco = f.x + f.y;
      if (sco > best) {
        best = sco;
        bi = c0 + c;
      }
    }
  }
  if (!valid) return;
  if (label) label[i] = bi;
  if (sums) {
    float* dst = sums + (int64_t)bi * KO_D;
#pragma unroll
    for (int j = 0; j < KO_D / 4; ++j) {
      const float4 v = __ldg(src + j);                     // the fp32 row again (L1/L2): the centroids stay fp32 means
      atomicAdd(dst + 4 * j, v.x);
      atomicAdd(dst + 4 * j + 1, v.y);
      atomicAdd(dst + 4 * j + 2, v.z);
      atomicAdd(dst + 4 * j + 3, v.w);
    }
    atomicAdd(counts + bi, 1);
  }
}

// The label pass over ALL rows on the tensor cores (mma.sync m16n8k16, fp16 operands like k_km_assign, fp32 accumulation):
// 1M rows x 512 centroids x 64 dimensions is a 6.6e10-flop GEMM with an argmax epilogue -- 1.6 ms in the HFMA2 loop above
// (a quarter of the packed-fp16 FMA peak), replicated on every rank of a multi-GPU run because each rank orders all keys.
// A CTA stages every centroid once (fp16, rows padded to 72 halves: the 8 x 4 fragment loads of a warp hit 32 different
// banks) and its 8 warps walk over tiles of 32 rows: two m16 tiles share each B fragment, a lane keeps the running best of
// the four rows it sees (strictly greater in increasing centroid order, lower index on a tie across the quad: the rule of
// the scalar kernel).  Labels only steer the ORDER of the exact scan.
constexpr int KA_PITCH = 72;      // halves per staged centroid
constexpr int KA_MAXK = 1024;     // centroids that fit (144 KB); more clusters (> 2M keys) take the scalar kernel
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(256)
k_km_assign_mma(const float* __restrict__ x, int64_t n, const float* __restrict__ cent, int K, int32_t* __restrict__ label) {
  extern __shared__ __align__(16) unsigned char ka_smem[];
  __half* sc = reinterpret_cast<__half*>(ka_smem);
  for (int j = threadIdx.x; j < K * (KO_D / 4); j += blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(cent) + j);
    const int r = j >> 4, c4 = j & 15;
    *reinterpret_cast<uint2*>(sc + r * KA_PITCH + c4 * 4) = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int64_t n_tiles = ceil_div64(n, 32);
  for (int64_t tile = (int64_t)blockIdx.x * 8 + warp; tile < n_tiles; tile += (int64_t)gridDim.x * 8) {
    const int64_t row0 = tile * 32;
    // A fragments of the two m16 tiles, four k-steps of 16: a0 (row g, k 2q..), a1 (row g+8, k 2q..), a2 / a3 the same at k + 8
    uint32_t a[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int64_t r_lo = row0 + mt * 16 + g, r_hi = r_lo + 8;
      const float2* p_lo = reinterpret_cast<const float2*>(x + (r_lo < n ? r_lo : 0) * KO_D);
      const float2* p_hi = reinterpret_cast<const float2*>(x + (r_hi < n ? r_hi : 0) * KO_D);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float2 l0 = __ldg(p_lo + ks * 8 + q), l1 = __ldg(p_lo + ks * 8 + 4 + q);
        const float2 h0 = __ldg(p_hi + ks * 8 + q), h1 = __ldg(p_hi + ks * 8 + 4 + q);
        a[mt][ks][0] = pack_h2(l0.x, l0.y);
        a[mt][ks][1] = pack_h2(h0.x, h0.y);
        a[mt][ks][2] = pack_h2(l1.x, l1.y);
        a[mt][ks][3] = pack_h2(h1.x, h1.y);
      }
    }
    float best[2][2] = {{-3.0e38f, -3.0e38f}, {-3.0e38f, -3.0e38f}};
    int bi[2][2] = {{0, 0}, {0, 0}};
    for (int t = 0; t < K / 8; ++t) {
      float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      const __half* bp = sc + (t * 8 + g) * KA_PITCH + q * 2;      // B fragment: centroid t*8+g, k = 2q.. and 2q+8..
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(bp + ks * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(bp + ks * 16 + 8);
        mma_16816(c[0], a[0][ks], b0, b1);
        mma_16816(c[1], a[1][ks], b0, b1);
      }
      const int n0 = t * 8 + q * 2;                               // c0, c1: row g, centroids n0, n0+1; c2, c3: row g+8
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (c[mt][0] > best[mt][0]) { best[mt][0] = c[mt][0]; bi[mt][0] = n0; }
        if (c[mt][1] > best[mt][0]) { best[mt][0] = c[mt][1]; bi[mt][0] = n0 + 1; }
        if (c[mt][2] > best[mt][1]) { best[mt][1] = c[mt][2]; bi[mt][1] = n0; }
        if (c[mt][3] > best[mt][1]) { best[mt][1] = c[mt][3]; bi[mt][1] = n0 + 1; }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
          const float os = __shfl_xor_sync(0xffffffffu, best[mt][h], o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi[mt][h], o);
          if (os > best[mt][h] || (os == best[mt][h] && oi < bi[mt][h])) {
            best[mt][h] = os;
            bi[mt][h] = oi;
          }
        }
        const int64_t r = row0 + mt * 16 + h * 8 + g;
        if (q == 0 && r < n) label[r] = bi[mt][h];
      }
    }
  }
}

// centroid = normalised mean of its points; an empty cluster keeps its previous centroid
__global__ void k_km_update(const float* __restrict__ sums, const int32_t* __restrict__ counts, int K, float* __restrict__ cent) {
  const int c = blockIdx.x;
  const int lane = threadIdx.x;     // 32 threads
  if (counts[c] == 0) return;
  const float v0 = sums[(int64_t)c * KO_D + lane], v1 = sums[(int64_t)c * KO_D + 32 + lane];
  float s = v0 * v0 + v1 * v1;
  s = warp_sum(s);
  const float inv = s > 0.f ? rsqrtf(s) : 0.f;
  cent[(int64_t)c * KO_D + lane] = v0 * inv;
  cent[(int64_t)c * KO_D + 32 + lane] = v1 * inv;
}

__global__ void k_km_init(const float* __restrict__ x, int64_t n, int K, float* __restrict__ cent) {
  const int c = blockIdx.x;
  const int64_t row = (int64_t)c * (n / K) + (n / K) / 2;     // evenly strided seeds: deterministic
  for (int j = threadIdx.x; j < KO_D; j += blockDim.x) cent[(int64_t)c * KO_D + j] = x[row * KO_D + j];
}

__global__ void k_iota(int32_t* __restrict__ a, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int32_t)i;
}

// seg_start[c] = first sorted position whose label is >= c (labels_sorted ascending); seg_start[K] = n
__global__ void k_seg_start(const int32_t* __restrict__ labels_sorted, int64_t n, int K, int32_t* __restrict__ seg_start) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int prev = i == 0 ? -1 : labels_sorted[i - 1];
  const int cur = i == n ? K : labels_sorted[i];
  for (int c = prev + 1; c <= cur; ++c) seg_start[c] = (int32_t)i;
}

// dst row r = src row perm[r] (64 floats), and inv[perm[r]] = r
__global__ void k_gather64(const float* __restrict__ src, const int32_t* __restrict__ perm, int64_t n, float* __restrict__ dst,
                           int32_t* __restrict__ inv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 16) return;
  const int64_t r = i >> 4;
  const int c = (int)(i & 15);
  const int32_t s = perm[r];
  reinterpret_cast<float4*>(dst)[r * 16 + c] = __ldg(reinterpret_cast<const float4*>(src) + (int64_t)s * 16 + c);
  if (c == 0 && inv) inv[s] = (int32_t)r;
}

// self_pos[r] = sorted key position of the cell that sorted query row r is
__global__ void k_self_pos(const int32_t* __restrict__ perm_q, const int32_t* __restrict__ inv_k, int64_t off, int64_t n,
                           int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = inv_k[off + perm_q[i]];
}

}  // namespace

// label <- rank[label]: clusters renumbered so that neighbours in the new numbering are neighbours on the sphere
__global__ void k_relabel(int32_t* __restrict__ label, int64_t n, const int32_t* __restrict__ rank) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) label[i] = __ldg(rank + label[i]);
}

// label pass over all rows: tensor-core kernel when the centroids fit in shared memory (CSC_KNN_KM_MMA=0: scalar kernel)
static csc_status km_label_all(csc_ctx* ctx, const float* x, int64_t n, const float* cent, int K, int32_t* label) {
  static const bool use_mma = !(getenv("CSC_KNN_KM_MMA") && atoi(getenv("CSC_KNN_KM_MMA")) == 0);
  if (use_mma && K <= KA_MAXK && K % 8 == 0) {
    const size_t smem = (size_t)K * KA_PITCH * 2;
    CSC_CUDA(ctx, cudaFuncSetAttribute(k_km_assign_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_tiles = ceil_div64(n, 32);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n_tiles, 8), (int64_t)ctx->sm_count));
    k_km_assign_mma<<<grid, 256, smem, ctx->stream>>>(x, n, cent, K, label);
  } else {
    k_km_assign<<<(unsigned)ceil_div64(n, 128), 128, 0, ctx->stream>>>(x, n, 1, cent, K, label, nullptr, nullptr);
  }
  CSC_LAUNCHED(ctx);
  return CSC_OK;
}

static csc_status sort_by_label(csc_ctx* ctx, const int32_t* labels, int64_t n, int K, DevPtr* perm, DevPtr* labels_sorted) {
  DevPtr idx, tmp;
  CSC_TRY(dev_alloc(ctx, (size_t)n * 4, &idx));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 4, perm));
  CSC_TRY(dev_alloc(ctx, (size_t)n * 4, labels_sorted));
  k_iota<<<(unsigned)ceil_div64(n, 256), 256, 0, ctx->stream>>>((int32_t*)idx->p, n);
  CSC_LAUNCHED(ctx);
  int bits = 1;
  while ((1 << bits) < K) ++bits;
  size_t tb = 0;
  CSC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tb, labels, (int32_t*)(*labels_sorted)->p, (const int32_t*)idx->p,
                                                (int32_t*)(*perm)->p, (int)n, 0, bits, ctx->stream));
  CSC_TRY(dev_alloc(ctx, tb, &tmp));
  CSC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp->p, tb, labels, (int32_t*)(*labels_sorted)->p, (const int32_t*)idx->p,
                                                (int32_t*)(*perm)->p, (int)n, 0, bits, ctx->stream));
  ctx->launches += 3;
  return CSC_OK;
}

// Cluster the key rows, sort keys (and queries) by cluster.  Outputs (device):
//   ks / qs        gathered prepared rows in sorted order (qs == ks when `same`)
//   perm_k/perm_q  sorted position -> original row;  self_pos[r] = sorted key position of query row r's own cell
//   start_tile[b]  key tile in the middle of query block b's cluster: its scan fans out from there (blocks of `tq`
//                  rows, tiles of `bn` keys)
csc_status knn_build_order(csc_ctx* ctx, const float* q_prep, int64_t n_q, const float* k_prep, int64_t n_keys, bool same,
                           int64_t query_offset, int tq, int bn, DevPtr* ks, DevPtr* qs, DevPtr* perm_k, DevPtr* perm_q,
                           DevPtr* self_pos, DevPtr* start_tile) {
  int K = 16;
  const char* kd = getenv("CSC_KNN_KDIV");
  const int64_t per = kd ? std::max(64, atoi(kd)) : 2048;
  while (K < 4096 && (int64_t)K * per < n_keys) K *= 2;
  DevPtr cent, sums, counts, lab_k, lab_q, labs_k, labs_q, inv_k, seg;
  CSC_TRY(dev_alloc(ctx, (size_t)K * KO_D * 4, &cent));
  CSC_TRY(dev_alloc(ctx, (size_t)K * KO_D * 4, &sums));
  CSC_TRY(dev_alloc(ctx, (size_t)K * 4, &counts));
  k_km_init<<<K, 64, 0, ctx->stream>>>(k_prep, n_keys, K, (float*)cent->p);
  CSC_LAUNCHED(ctx);
  // Lloyd passes on a strided sample of at most 64k rows
  const int64_t stride = std::max<int64_t>(1, n_keys / 65536);
  const int64_t n_s = n_keys / stride;
  for (int it = 0; it < 3; ++it) {
    CSC_CUDA(ctx, cudaMemsetAsync(sums->p, 0, (size_t)K * KO_D * 4, ctx->stream));
    CSC_CUDA(ctx, cudaMemsetAsync(counts->p, 0, (size_t)K * 4, ctx->stream));
    k_km_assign<<<(unsigned)ceil_div64(n_s, 128), 128, 0, ctx->stream>>>(k_prep, n_s, stride, (const float*)cent->p, K, nullptr,
                                                                      (float*)sums->p, (int32_t*)counts->p);
    CSC_LAUNCHED(ctx);
    k_km_update<<<K, 32, 0, ctx->stream>>>((const float*)sums->p, (const int32_t*)counts->p, K, (float*)cent->p);
    CSC_LAUNCHED(ctx);
  }
  // Number the clusters by the angle of their centroid in the plane of the two leading components (the rows are unit
  // vectors of the PCA embedding): a circular order in which neighbouring numbers are neighbouring directions, so that
  // a scan that fans out from a row's own cluster (knn_f16.cu visits tiles c, c+1, c-1, c+2, ...) meets the similar
  // clusters first and its thresholds settle early.
  DevPtr rank;
  {
    std::vector<float> hc((size_t)K * KO_D);
    CSC_CUDA(ctx, cudaMemcpyAsync(hc.data(), cent->p, hc.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<std::pair<float, int32_t>> ang((size_t)K);
    for (int c = 0; c < K; ++c) ang[c] = std::make_pair(std::atan2(hc[(size_t)c * KO_D + 1], hc[(size_t)c * KO_D]), (int32_t)c);
    std::sort(ang.begin(), ang.end());
    std::vector<int32_t> hr((size_t)K);
    for (int p = 0; p < K; ++p) hr[ang[p].second] = p;
    CSC_TRY(dev_alloc(ctx, (size_t)K * 4, &rank));
    CSC_CUDA(ctx, cudaMemcpyAsync(rank->p, hr.data(), (size_t)K * 4, cudaMemcpyHostToDevice, ctx->stream));
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  CSC_TRY(dev_alloc(ctx, (size_t)n_keys * 4, &lab_k));
  CSC_TRY(km_label_all(ctx, k_prep, n_keys, (const float*)cent->p, K, (int32_t*)lab_k->p));
  k_relabel<<<(unsigned)ceil_div64(n_keys, 256), 256, 0, ctx->stream>>>((int32_t*)lab_k->p, n_keys, (const int32_t*)rank->p);
  CSC_LAUNCHED(ctx);
  CSC_TRY(sort_by_label(ctx, (const int32_t*)lab_k->p, n_keys, K, perm_k, &labs_k));
  CSC_TRY(dev_alloc(ctx, (size_t)n_keys * KO_D * 4, ks));
  CSC_TRY(dev_alloc(ctx, (size_t)n_keys * 4, &inv_k));
  k_gather64<<<(unsigned)ceil_div64(n_keys * 16, 256), 256, 0, ctx->stream>>>(k_prep, (const int32_t*)(*perm_k)->p, n_keys,
                                                                           (float*)(*ks)->p, (int32_t*)inv_k->p);
  CSC_LAUNCHED(ctx);
  CSC_TRY(dev_alloc(ctx, (size_t)(K + 1) * 4, &seg));
  k_seg_start<<<(unsigned)ceil_div64(n_keys + 1, 256), 256, 0, ctx->stream>>>((const int32_t*)labs_k->p, n_keys, K, (int32_t*)seg->p);
  CSC_LAUNCHED(ctx);
  const int32_t* labs_q_ptr = nullptr;
  if (same) {
    *qs = *ks;
    *perm_q = *perm_k;
    labs_q_ptr = (const int32_t*)labs_k->p;
  } else {
    CSC_TRY(dev_alloc(ctx, (size_t)n_q * 4, &lab_q));
    CSC_TRY(km_label_all(ctx, q_prep, n_q, (const float*)cent->p, K, (int32_t*)lab_q->p));
    k_relabel<<<(unsigned)ceil_div64(n_q, 256), 256, 0, ctx->stream>>>((int32_t*)lab_q->p, n_q, (const int32_t*)rank->p);
    CSC_LAUNCHED(ctx);
    CSC_TRY(sort_by_label(ctx, (const int32_t*)lab_q->p, n_q, K, perm_q, &labs_q));
    CSC_TRY(dev_alloc(ctx, (size_t)n_q * KO_D * 4, qs));
    k_gather64<<<(unsigned)ceil_div64(n_q * 16, 256), 256, 0, ctx->stream>>>(q_prep, (const int32_t*)(*perm_q)->p, n_q, (float*)(*qs)->p,
                                                                          nullptr);
    CSC_LAUNCHED(ctx);
    labs_q_ptr = (const int32_t*)labs_q->p;
  }
  // small host step: per query block, where its first row's cluster starts among the keys
  const int64_t n_blocks = ceil_div64(n_q, tq);
  std::vector<int32_t> h_seg((size_t)K + 1), h_lab((size_t)n_blocks), h_start((size_t)n_blocks);
  DevPtr d_first;
  CSC_TRY(dev_alloc(ctx, (size_t)n_blocks * 4, &d_first));
  // first label of every block: strided device->device copy, then one small download
  CSC_CUDA(ctx, cudaMemcpy2DAsync(d_first->p, 4, labs_q_ptr, (size_t)tq * 4, 4, (size_t)n_blocks, cudaMemcpyDeviceToDevice, ctx->stream));
  CSC_CUDA(ctx, cudaMemcpyAsync(h_lab.data(), d_first->p, (size_t)n_blocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaMemcpyAsync(h_seg.data(), seg->p, (size_t)(K + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int64_t bq = 0; bq < n_blocks; ++bq) {
    // the tile in the middle of the block's cluster: the scan fans out from there
    const int64_t t = ((int64_t)h_seg[h_lab[bq]] + (int64_t)h_seg[h_lab[bq] + 1]) / 2 / bn;
    h_start[bq] = (int32_t)std::max<int64_t>(0, t);
  }
  CSC_TRY(dev_alloc(ctx, (size_t)n_blocks * 4, start_tile));
  CSC_CUDA(ctx, cudaMemcpyAsync((*start_tile)->p, h_start.data(), (size_t)n_blocks * 4, cudaMemcpyHostToDevice, ctx->stream));
  // self position of every (sorted) query row among the sorted keys
  CSC_TRY(dev_alloc(ctx, (size_t)n_q * 4, self_pos));
  k_self_pos<<<(unsigned)ceil_div64(n_q, 256), 256, 0, ctx->stream>>>((const int32_t*)(*perm_q)->p, (const int32_t*)inv_k->p,
                                                                     query_offset, n_q, (int32_t*)(*self_pos)->p);
  CSC_LAUNCHED(ctx);
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));    // h_start must outlive its copy
  return CSC_OK;
}
