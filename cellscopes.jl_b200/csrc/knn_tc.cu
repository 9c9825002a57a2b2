// kNN scan v2: distance tiles on the 5th-generation tensor cores (tcgen05), accumulators in TMEM,
// operands staged by TMA, fused with a per-row running top-k so that no N x N matrix reaches HBM.
//
//   score tile [128 queries x 128 keys] = Q K' with 3xTF32 error compensation:
//       x = hi + lo (hi = x rounded to tf32, lo = x - hi, both held in fp32 arrays)
//       Q K' ~= Qhi Khi' + Qhi Klo' + Qlo Khi'          (three kind::tf32 MMA passes into one accumulator)
//   which keeps fp32-level accuracy (dropped terms ~2^-21 |q||k|) at one third of the TF32 rate.
//
// Warp roles (192 threads, 1 CTA per SM, one CTA = 128 query rows against ALL keys):
//   warp 0     TMA producer: query tile once, then key tiles (hi/lo, two 128-byte K chunks each) into a
//              2-stage shared-memory ring, mbarrier full/empty.
//   warp 1     TMEM allocator + MMA issuer: one elected thread issues tcgen05.mma, commits to the
//              "stage empty" and "accumulator full" mbarriers.
//   warps 2-5  epilogue: each thread owns one query row (= one TMEM lane).  tcgen05.ld brings 32
//              scores at a time into registers; a max-tree rejects the chunk with one compare against
//              the row's threshold (kept in a register); survivors go into the row's sorted list in
//              shared memory.  Two accumulator stages in TMEM overlap the epilogue with the next tile.
// After the scan the (k + margin) survivors per row are re-scored exactly and sorted (k_knn_refine).
#include <cuda.h>

#include <cfloat>
#include <climits>

#include "common.cuh"
#include "knn_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TC_TQ = 128;      // query rows per CTA = UMMA M
constexpr int TC_BN = 128;      // keys per tile     = UMMA N
constexpr int TC_DP = 64;       // padded feature dimension (floats): two 128-byte swizzle rows
constexpr int TC_STAGES = 3;    // shared-memory stages of key tiles
constexpr int TC_ACC = 2;       // TMEM accumulator stages
constexpr int TC_CHUNK_BYTES = TC_BN * 128;          // one [128 rows x 32 floats] swizzled block = 16 KB
constexpr int TC_TILE_BYTES = 4 * TC_CHUNK_BYTES;    // hi c0, hi c1, lo c0, lo c1 = 64 KB
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 512;                    // 2 x 128 accumulator columns + 64 (Q hi) + 64 (Q lo)
constexpr int TC_TMEM_A = TC_ACC * TC_BN;            // first column of the query operand

using namespace tc;

struct __align__(64) TcMaps {
  CUtensorMap k_hi, k_lo;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
k_knn_scan_tc(const __grid_constant__ TcMaps maps, const float* __restrict__ q_hi, const float* __restrict__ q_lo,
              int64_t n_q, int64_t n_keys, int64_t query_offset, const int32_t* __restrict__ self_ids, int32_t kl,
              int32_t ksteps, float* __restrict__ out_score, int32_t* __restrict__ out_idx, long long* __restrict__ trace, int64_t trace_t0) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment for the swizzled operand blocks
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sB = base;                                   // TC_STAGES x [hi c0 | hi c1 | lo c0 | lo c1] 64 KB
  uint2* lists = reinterpret_cast<uint2*>(sB + TC_STAGES * TC_TILE_BYTES);   // [128 rows][32] (score, id)
  uint64_t* bars = reinterpret_cast<uint64_t*>(lists + TC_TQ * 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar_full = smem_u32(bars + 0);        // [TC_STAGES]
  const uint32_t bar_empty = smem_u32(bars + 4);       // [TC_STAGES]
  const uint32_t bar_tfull = smem_u32(bars + 8);       // [TC_ACC]
  const uint32_t bar_tempty = smem_u32(bars + 10);     // [TC_ACC]
  const uint32_t bar_a = smem_u32(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q0 = (int64_t)blockIdx.x * TC_TQ;
  const int64_t n_tiles = (n_keys + TC_BN - 1) / TC_BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < TC_ACC; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 128);
    }
    mbar_init(bar_a, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.k_hi)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.k_lo)) : "memory");
    }
    for (int64_t t = 0; t < n_tiles; ++t) {
      const int s = (int)(t % TC_STAGES);
      const uint32_t par = (uint32_t)((t / TC_STAGES) & 1);
      mbar_wait(bar_empty + 8 * s, par ^ 1);
      if (elect_one()) {
        if (trace && blockIdx.x == 0 && lane == 0 && t >= trace_t0 && t < trace_t0 + 64) trace[(t - trace_t0) * 8 + 0] = clock64();
        mbar_expect_tx(bar_full + 8 * s, TC_TILE_BYTES);
        unsigned char* dst = sB + (size_t)s * TC_TILE_BYTES;
        const int row = (int)(t * TC_BN);
        tma_load_2d(smem_u32(dst + 0 * TC_CHUNK_BYTES), &maps.k_hi, bar_full + 8 * s, 0, row);
        tma_load_2d(smem_u32(dst + 1 * TC_CHUNK_BYTES), &maps.k_hi, bar_full + 8 * s, 32, row);
        tma_load_2d(smem_u32(dst + 2 * TC_CHUNK_BYTES), &maps.k_lo, bar_full + 8 * s, 0, row);
        tma_load_2d(smem_u32(dst + 3 * TC_CHUNK_BYTES), &maps.k_lo, bar_full + 8 * s, 32, row);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp runs the loop, one elected lane issues) =================
    {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, both K-major, N=128, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_TQ >> 4) << 24);
      // high word of the shared-memory descriptor: SBO = 1024 B, version 1, SWIZZLE_128B
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      mbar_wait(bar_a, 0);
      tc_fence_after();
      for (int64_t t = 0; t < n_tiles; ++t) {
        const int s = (int)(t % TC_STAGES);
        const int a = (int)(t % TC_ACC);
        mbar_wait(bar_tempty + 8 * a, (uint32_t)(((t / TC_ACC) & 1) ^ 1));
        if (trace && blockIdx.x == 0 && lane == 0 && t >= trace_t0 && t < trace_t0 + 64) trace[(t - trace_t0) * 8 + 1] = clock64();
        mbar_wait(bar_full + 8 * s, (uint32_t)((t / TC_STAGES) & 1));
        if (trace && blockIdx.x == 0 && lane == 0 && t >= trace_t0 && t < trace_t0 + 64) trace[(t - trace_t0) * 8 + 2] = clock64();
        tc_fence_after();
        const uint32_t b_addr = smem_u32(sB + (size_t)s * TC_TILE_BYTES);
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * TC_BN);
        // descriptors differ only in their low word: (address >> 4) | LBO; all offsets are compile-time
        const uint32_t b_lo = ((b_addr >> 4) & 0x3FFF) | (1u << 16);
        if (elect_one()) {
        uint32_t acc = 0;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          // pass 0: Qhi Khi'   pass 1: Qhi Klo'   pass 2: Qlo Khi'
          const uint32_t a_col = tmem_base + TC_TMEM_A + ((pass == 2) ? 64u : 0u);
          const uint32_t b_off = (pass == 1) ? 2 * TC_CHUNK_BYTES : 0;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            if (ks < ksteps) {
              const uint32_t off = (uint32_t)(ks >> 2) * TC_CHUNK_BYTES + (uint32_t)(ks & 3) * 32;   // 8 tf32 = 32 bytes
              const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + ((b_off + off) >> 4));
              tc_mma_tf32_ts(d_tmem, a_col + (uint32_t)ks * 8u, db, idesc, acc);
              acc = 1;
            }
          }
        }
        tc_commit(bar_empty + 8 * s);     // key stage reusable once these MMAs have read it
        tc_commit(bar_tfull + 8 * a);     // accumulator ready for the epilogue
        }
        __syncwarp();
        if (trace && blockIdx.x == 0 && lane == 0 && t >= trace_t0 && t < trace_t0 + 64) trace[(t - trace_t0) * 8 + 3] = clock64();
      }
    }
  } else {
    // ================= epilogue: one thread = one query row =================
    // Row lists live in shared memory as [row][32] (score, id) pairs, sorted descending; lane j of the warp
    // owns entry j of whichever row is being updated, so an insertion is one LDS.64, a ballot, two shuffles
    // and one STS.64 for the whole warp (no divergent per-thread shifting loop).
    const unsigned FULL = 0xffffffffu;
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int64_t qrow = q0 + row;
    const bool qvalid = qrow < n_q;
    // the key that is the query itself (excluded): a row range of the keys, or an explicit id per gathered row
    const int self_i = self_ids ? (qvalid ? self_ids[qrow] : -1) : (int)(query_offset + qrow);
    {
      // stage this thread's query row (hi and lo parts) into the TMEM operand block: lane = row, column = feature
      const uint32_t a_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + TC_TMEM_A;
#pragma unroll 1
      for (int part = 0; part < 4; ++part) {           // hi[0:32], hi[32:64], lo[0:32], lo[32:64]
        const float* src = (part < 2 ? q_hi : q_lo) + (qvalid ? qrow : 0) * TC_DP + (part & 1) * 32;
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 f = qvalid ? __ldg(reinterpret_cast<const float4*>(src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[4 * i] = __float_as_uint(f.x);
          v[4 * i + 1] = __float_as_uint(f.y);
          v[4 * i + 2] = __float_as_uint(f.z);
          v[4 * i + 3] = __float_as_uint(f.w);
        }
        tmem_st32(a_base + (uint32_t)part * 32u, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(bar_a);
    }
    uint2* wl = lists + (size_t)quarter * 32 * 32;      // this warp's 32 rows
    for (int r = 0; r < 32; ++r) wl[r * 32 + lane] = make_uint2(__float_as_uint(-INFINITY), (unsigned)INT_MAX);
    __syncwarp();
    float tau = qvalid ? -INFINITY : INFINITY;
    for (int64_t t = 0; t < n_tiles; ++t) {
      const int a = (int)(t % TC_ACC);
      mbar_wait(bar_tfull + 8 * a, (uint32_t)((t / TC_ACC) & 1));
      tc_fence_after();
      const int64_t key0 = t * TC_BN;
#pragma unroll 1
      for (int c = 0; c < TC_BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * TC_BN + c * 32), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float g[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float m = __uint_as_float(r[q * 8]);
#pragma unroll
          for (int j = 1; j < 8; ++j) m = fmaxf(m, __uint_as_float(r[q * 8 + j]));
          g[q] = m;
        }
        const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
        if (__any_sync(FULL, m > tau)) {
          // Slow path, warp-uniform.  Candidates are detected against a snapshot of the thresholds so that the
          // votes of a group are independent instructions (pipelined); the insertion re-checks against the list.
          const float tsnap = tau;
          const bool h0 = __any_sync(FULL, g[0] > tsnap), h1 = __any_sync(FULL, g[1] > tsnap);
          const bool h2 = __any_sync(FULL, g[2] > tsnap), h3 = __any_sync(FULL, g[3] > tsnap);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool hq = q == 0 ? h0 : q == 1 ? h1 : q == 2 ? h2 : h3;
            if (hq) {
              unsigned bm[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int id = (int)key0 + c * 32 + q * 8 + j;          // the same key for every lane
                bm[j] = __ballot_sync(FULL, __uint_as_float(r[q * 8 + j]) > tsnap && id != self_i && id < (int)n_keys);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                unsigned b = bm[j];
                const int id = (int)key0 + c * 32 + q * 8 + j;
                while (b) {
                  const int src = __ffs(b) - 1;
                  b &= b - 1;
                  const float cv = __shfl_sync(FULL, __uint_as_float(r[q * 8 + j]), src);
                  const uint2 e = wl[src * 32 + lane];
                  float es = __uint_as_float(e.x);
                  int ei = (int)e.y;
                  // entries that rank before the candidate form a prefix of the sorted list
                  const int pos = __popc(__ballot_sync(FULL, es > cv || (es == cv && ei < id)));
                  if (pos < kl) {
                    const float ps = __shfl_up_sync(FULL, es, 1);
                    const int pi = __shfl_up_sync(FULL, ei, 1);
                    if (lane == pos) {
                      es = cv;
                      ei = id;
                    } else if (lane > pos) {
                      es = ps;
                      ei = pi;
                    }
                    wl[src * 32 + lane] = make_uint2(__float_as_uint(es), (unsigned)ei);
                    const float nt = __shfl_sync(FULL, es, kl - 1);
                    if (lane == src) tau = nt;
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * a);
      if (trace && blockIdx.x == 0 && t >= trace_t0 && t < trace_t0 + 64 && lane == 0) trace[(t - trace_t0) * 8 + 4 + (warp & 3)] = clock64();
    }
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
      const int64_t orow = q0 + quarter * 32 + r;
      if (orow < n_q && lane < kl) {
        const uint2 e = wl[r * 32 + lane];
        out_score[orow * kl + lane] = __uint_as_float(e.x);
        out_idx[orow * kl + lane] = (int)e.y;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS)
                 : "memory");
  }
}

// hi/lo split of a prepared operand row: hi = round-to-nearest tf32 (low 13 bits zero), lo = x - hi (exact)
__global__ void k_split_tf32(const float* __restrict__ x, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
  const float hf = __uint_as_float(h);
  hi[i] = hf;
  lo[i] = v - hf;
}

// exact re-scoring of the survivors (fp64 accumulation over the fp32 operands) and final ordering:
// descending score, ties by ascending id.  One thread per query row; kl <= 32.
__global__ void __launch_bounds__(128)
k_knn_refine(const float* __restrict__ q, const float* __restrict__ keys, int64_t n_q, int32_t kl, int32_t k, int32_t d,
             const float* __restrict__ in_score, const int32_t* __restrict__ in_idx, float* __restrict__ out_score,
             int32_t* __restrict__ out_idx) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_q) return;
  float s[32];
  int id[32];
  const float4* qr = reinterpret_cast<const float4*>(q + row * TC_DP);
  for (int j = 0; j < 32; ++j) {
    s[j] = -INFINITY;
    id[j] = INT_MAX;
  }
  int cnt = 0;
  for (int j = 0; j < kl; ++j) {
    const int cand = in_idx[row * kl + j];
    if (cand == INT_MAX) continue;
    const float4* kr = reinterpret_cast<const float4*>(keys + (int64_t)cand * TC_DP);
    double acc = 0.0;
    for (int i = 0; i < (d + 3) / 4; ++i) {
      const float4 a = __ldg(qr + i), b = __ldg(kr + i);
      acc += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
    }
    const float v = (float)acc;
    // insertion into the sorted prefix [0, cnt)
    int p = cnt++;
#pragma unroll 1
    while (p > 0 && (v > s[p - 1] || (v == s[p - 1] && cand < id[p - 1]))) {
      s[p] = s[p - 1];
      id[p] = id[p - 1];
      --p;
    }
    s[p] = v;
    id[p] = cand;
  }
  for (int j = 0; j < k; ++j) {
    out_score[row * k + j] = s[j];
    out_idx[row * k + j] = id[j];
  }
  (void)in_score;
}

csc_status make_map(csc_ctx* ctx, EncodeTiledFn fn, CUtensorMap* map, const float* ptr, int64_t n_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)TC_DP, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)TC_DP * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)TC_BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return csc_fail(ctx, CSC_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return CSC_OK;
}

}  // namespace

// q_prep / k_prep: prepared operand rows [n x 64] fp32 (knn.cu::knn_prepare with DP = 64).
// Writes the k best (score, id) per query, exact scores, sorted.
csc_status knn_scan_tc(csc_ctx* ctx, const float* q_prep, int64_t n_q, const float* k_prep, int64_t n_keys, bool same,
                       int64_t query_offset, const int32_t* self_ids, int32_t k, int32_t d_eff, float* out_score,
                       int32_t* out_idx) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CSC_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
      return csc_fail(ctx, CSC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (EncodeTiledFn)fn;
  }
  const int kl = std::min(32, k + 4);       // a margin of survivors so that 3xTF32 rounding cannot change the exact top-k
  const int ksteps = (d_eff + 7) / 8;
  DevPtr khi, klo, qhi, qlo, cs, ci;
  const int64_t nk_el = n_keys * TC_DP, nq_el = n_q * TC_DP;
  CSC_TRY(dev_alloc(ctx, (size_t)nk_el * 4, &khi));
  CSC_TRY(dev_alloc(ctx, (size_t)nk_el * 4, &klo));
  k_split_tf32<<<(unsigned)ceil_div64(nk_el, 256), 256, 0, ctx->stream>>>(k_prep, nk_el, (float*)khi->p, (float*)klo->p);
  CSC_LAUNCHED(ctx);
  if (same) {
    qhi = khi;
    qlo = klo;
  } else {
    CSC_TRY(dev_alloc(ctx, (size_t)nq_el * 4, &qhi));
    CSC_TRY(dev_alloc(ctx, (size_t)nq_el * 4, &qlo));
    k_split_tf32<<<(unsigned)ceil_div64(nq_el, 256), 256, 0, ctx->stream>>>(q_prep, nq_el, (float*)qhi->p, (float*)qlo->p);
    CSC_LAUNCHED(ctx);
  }
  CSC_TRY(dev_alloc(ctx, (size_t)n_q * kl * 4, &cs));
  CSC_TRY(dev_alloc(ctx, (size_t)n_q * kl * 4, &ci));
  TcMaps maps;
  CSC_TRY(make_map(ctx, encode, &maps.k_hi, (const float*)khi->p, n_keys));
  CSC_TRY(make_map(ctx, encode, &maps.k_lo, (const float*)klo->p, n_keys));
  const size_t smem = 1024 + (size_t)TC_TILE_BYTES * TC_STAGES + (size_t)TC_TQ * 32 * 8 + 256;
  CSC_CUDA(ctx, cudaFuncSetAttribute(k_knn_scan_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DevPtr trace;
  const bool want_trace = getenv("CSC_KNN_TRACE") != nullptr;
  if (want_trace) CSC_TRY(dev_alloc(ctx, 64 * 8 * 8, &trace, true));
  cudaEventRecord(ctx->ev_k0, ctx->stream);
  k_knn_scan_tc<<<(unsigned)ceil_div64(n_q, TC_TQ), TC_THREADS, smem, ctx->stream>>>(maps, (const float*)qhi->p, (const float*)qlo->p, n_q, n_keys, query_offset, self_ids, kl, ksteps,
                                                                                      (float*)cs->p, (int32_t*)ci->p,
                                                                                      want_trace ? (long long*)trace->p : nullptr, want_trace ? atoll(getenv("CSC_KNN_TRACE")) : 0);
  CSC_LAUNCHED(ctx);
  cudaEventRecord(ctx->ev_k1, ctx->stream);
  k_knn_refine<<<(unsigned)ceil_div64(n_q, 128), 128, 0, ctx->stream>>>(q_prep, k_prep, n_q, kl, k, d_eff, (const float*)cs->p,
                                                                         (const int32_t*)ci->p, out_score, out_idx);
  CSC_LAUNCHED(ctx);
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1) == cudaSuccess) ctx->stage_ms[CSC_T_KNN_SCAN] += ms;
  if (want_trace) {
    std::vector<long long> h(64 * 8);
    CSC_CUDA(ctx, cudaMemcpy(h.data(), trace->p, h.size() * 8, cudaMemcpyDeviceToHost));
    const long long t0 = h[0];
    fprintf(stderr, "[knn trace] tile: producer_empty_ok  mma_tempty_ok  mma_full_ok  mma_issued | epilogue done q0..q3 (cycles since first)\n");
    for (int t = 0; t < 40; ++t)
      fprintf(stderr, "[knn trace] %2d: %8lld %8lld %8lld %8lld | %8lld %8lld %8lld %8lld\n", t, h[t * 8 + 0] - t0, h[t * 8 + 1] - t0,
              h[t * 8 + 2] - t0, h[t * 8 + 3] - t0, h[t * 8 + 4] - t0, h[t * 8 + 5] - t0, h[t * 8 + 6] - t0, h[t * 8 + 7] - t0);
  }
  return CSC_OK;
}
