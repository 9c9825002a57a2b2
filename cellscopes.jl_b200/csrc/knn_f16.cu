// kNN scan v3 (cosine): filter on the tensor cores in fp16, prove, refine exactly.
//
//   Exact brute force needs N^2 scores but only ~k + a few survive per row, so the scores do not have to be
//   exact -- the LIST does.  The scan computes every score once in fp16 x fp16 -> fp32 (kind::f16: twice the
//   rate and half the operand bytes of tf32, a third of the MMAs of the 3xTF32 scan), keeps the L = 32 best
//   per row, and the refine step re-scores those 32 exactly (fp64 accumulation over the fp32 rows) and PROVES
//   that nothing outside the list can belong to the top k:
//       rows are unit vectors, fp16 keeps 11 significant bits  =>  |approx - exact| <= eps = 1.0e-3  (rigorous:
//       (2^-10 + 2^-22) * sum |q_i||k_i| <= 9.8e-4, plus fp32 accumulation)
//       every key outside the list has approx <= a_L (the 32nd best approx score), hence exact <= a_L + eps
//       if a_L + eps < e_k (the exact k-th best inside the list) the exact top-k is inside the list.  QED.
//   Rows that fail the test (near-duplicate neighbourhoods, k close to L) are gathered and re-run through the
//   exact 3xTF32 scan (knn_tc.cu), so the result is always the exact one, ties by lower id included.
//
// Kernel layout (320 threads, 1 CTA per SM, one CTA = 256 query rows = two 128-row blocks against ALL keys; the two
// blocks share every key tile, which halves the L2 traffic per score -- the scan is L2-bandwidth bound otherwise):
//   warp 0      TMA producer: key tiles [128 keys x 64 halves] (one 128-byte swizzle row per key), 6-stage ring
//   warp 1      TMEM allocator + MMA issuer: A (queries, both blocks) lives in TMEM (2 x 32 columns of packed halves),
//               D = 3 accumulator stages x 2 blocks x 64 columns; the unit of work is a half tile of 64 keys
//   warps 2-5   epilogue of block 0, warps 6-9 epilogue of block 1: thread = query row (TMEM lane); per half tile
//               2 x tcgen05.ld.x32, a FMNMX3 tree, one compare against the row's threshold; survivors go into the row's
//               sorted list in shared memory (warp-cooperative insertion when few rows have candidates, one lane per
//               row when many do).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cfloat>
#include <climits>

#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;

csc_status knn_scan_tc(csc_ctx* ctx, const float* q_prep, int64_t n_q, const float* k_prep, int64_t n_keys, bool same,
                       int64_t query_offset, const int32_t* self_ids, int32_t k, int32_t d_eff, float* out_score,
                       int32_t* out_idx);

namespace {

constexpr int F_TQ = 256;                 // query rows per CTA: two blocks of UMMA M = 128
constexpr int F_BN = 128;                 // keys per shared-memory tile
constexpr int F_HN = 64;                  // keys per MMA / accumulator stage (half a tile) = UMMA N
constexpr int F_DP = 64;                  // padded feature dimension (halves): one 128-byte swizzle row
constexpr int F_STAGES = 6;
constexpr int F_ACC = 3;                  // TMEM accumulator stages
constexpr int F_TILE_BYTES = F_BN * 128;  // 16 KB
constexpr int F_THREADS = 64 + 256;
constexpr int F_L = 32;                   // candidates kept per row
constexpr int F_TMEM_A = F_ACC * 2 * F_HN; // first column of the query operands (2 blocks x 32 columns)
constexpr float F_SCALE = 256.f;          // operands are scaled so that small components stay normal in fp16
constexpr float F_EPS = 1.0e-3f;          // bound on |approx - exact| for unit rows (see the header)

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// max of 32 scores held as raw bits
__device__ __forceinline__ float max32(const uint32_t (&r)[32]) {
  float m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) m[i] = fmax3(__uint_as_float(r[3 * i]), __uint_as_float(r[3 * i + 1]), __uint_as_float(r[3 * i + 2]));
  m[10] = fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31]));
  const float a = fmax3(m[0], m[1], m[2]), b = fmax3(m[3], m[4], m[5]), c = fmax3(m[6], m[7], m[8]);
  return fmax3(fmax3(a, b, c), m[9], m[10]);
}

// Per-row candidate set: 32 (score, id) slots in shared memory, UNSORTED (the refine kernel orders by exact score
// anyway), slot-major so that lane = row is bank-conflict free: sc[slot * 256 + row], id[slot * 256 + row].
// The only thing the scan needs from a row's set is its minimum (the threshold tau) and where it sits.  The owner
// thread keeps a two-level minimum in registers: the minimum and its slot of each group of 8 slots; a new candidate
// overwrites the global minimum's slot and only that group's minimum is recomputed (8 LDS) -- ~35 instructions per
// insertion, no warp-level operation, so the rows of a warp insert independently (dense phases) and a single row
// costs the same (sparse phases).  Ties of the approximate score are irrelevant: a candidate enters only if it is
// strictly better than tau, so everything outside the set is <= tau, which is all the proof needs.
struct RowState {
  float tau;        // min over the 32 slots
  int slot;         // where that minimum sits
  float gm[4];      // min of slots [8g, 8g+8)
  int gs[4];        // its slot
};

__device__ __forceinline__ void row_insert(RowState& st, float* sc, int* ids, int row, float v, int id) {
  const int s = st.slot;
  sc[s * F_TQ + row] = v;
  ids[s * F_TQ + row] = id;
  const int g = s >> 3;
  float m = sc[(g * 8) * F_TQ + row];
  int ms = g * 8;
#pragma unroll
  for (int u = 1; u < 8; ++u) {
    const float x = sc[(g * 8 + u) * F_TQ + row];
    if (x < m) {
      m = x;
      ms = g * 8 + u;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (q == g) {
      st.gm[q] = m;
      st.gs[q] = ms;
    }
  float t = st.gm[0];
  int ts = st.gs[0];
#pragma unroll
  for (int q = 1; q < 4; ++q)
    if (st.gm[q] < t) {
      t = st.gm[q];
      ts = st.gs[q];
    }
  st.tau = t;
  st.slot = ts;
}

// one 32-key chunk of the slow path: scores to the warp's scratch tile (the candidate loop needs them by a run-time
// index), candidate mask, then every lane inserts its own row's candidates
__device__ __forceinline__ void slow_chunk(const uint32_t (&r)[32], float* scratch, int key_base, int self_i, int n_keys,
                                           float* sc, int* ids, int row, int lane, RowState& st, unsigned long long* cnt) {
  unsigned cm = 0u;
  const float tau0 = st.tau;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    scratch[i * 32 + lane] = __uint_as_float(r[i]);
    cm |= (__uint_as_float(r[i]) > tau0 ? 1u : 0u) << i;
  }
  const long long t0 = cnt ? clock64() : 0;
  int ncand = 0;
  while (cm) {
    const int j = __ffs(cm) - 1;
    cm &= cm - 1;
    const int id = key_base + j;
    const float v = scratch[j * 32 + lane];
    if (v > st.tau && id != self_i && id < n_keys) {
      row_insert(st, sc, ids, row, v, id);
      ++ncand;
    }
  }
  __syncwarp();
  if (cnt) {
    int tot = ncand;
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) {
      atomicAdd(cnt + 3, 1ull);
      atomicAdd(cnt + 7, (unsigned long long)(clock64() - t0));
      atomicAdd(cnt + 9, (unsigned long long)tot);
    }
  }
}

__global__ void __launch_bounds__(F_THREADS, 1)
k_knn_scan_f16(const __grid_constant__ CUtensorMap kmap, const __half* __restrict__ q16, int64_t n_q, int64_t n_keys,
               int64_t query_offset, const int32_t* __restrict__ self_pos, const int32_t* __restrict__ start_tile, int32_t ksteps,
               float* __restrict__ out_thr, int32_t* __restrict__ out_idx, int dbg_mode, unsigned long long* cnt) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment for the swizzled tiles, as an offset so that the pointers stay in the shared window
  unsigned char* sB = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);  // F_STAGES x 16 KB
  float* l_sc = reinterpret_cast<float*>(sB + F_STAGES * F_TILE_BYTES);            // [32 slots][256 rows] scores
  int* l_id = reinterpret_cast<int*>(l_sc + F_L * F_TQ);                           // [32 slots][256 rows] key ids
  float* scratch_all = reinterpret_cast<float*>(l_id + F_L * F_TQ);                // [8 warps][32 keys][32 lanes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch_all + 8 * 32 * 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const uint32_t bar_full = smem_u32(bars + 0);        // [F_STAGES]
  const uint32_t bar_empty = smem_u32(bars + 8);       // [F_STAGES]
  const uint32_t bar_tfull = smem_u32(bars + 16);      // [F_ACC]
  const uint32_t bar_tempty = smem_u32(bars + 19);     // [F_ACC]
  const uint32_t bar_a = smem_u32(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q0 = (int64_t)blockIdx.x * F_TQ;
  const int64_t n_tiles = (n_keys + F_BN - 1) / F_BN;
  const int64_t n_half = 2 * n_tiles;
  // locality order (knn_order.cu): this block starts its scan near its own cluster and wraps around
  const int64_t t_start = start_tile ? (int64_t)start_tile[blockIdx.x] % n_tiles : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < F_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < F_ACC; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 256);
    }
    mbar_init(bar_a, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&kmap)) : "memory");
    for (int64_t t = 0; t < n_tiles; ++t) {
      const int s = (int)(t % F_STAGES);
      mbar_wait(bar_empty + 8 * s, (uint32_t)(((t / F_STAGES) & 1) ^ 1));
      if (elect_one()) {
        mbar_expect_tx(bar_full + 8 * s, F_TILE_BYTES);
        int64_t tile = t + t_start;
        if (tile >= n_tiles) tile -= n_tiles;
        tma_load_2d(smem_u32(sB + (size_t)s * F_TILE_BYTES), &kmap, bar_full + 8 * s, 0, (int)(tile * F_BN));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(F_HN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // K-major SWIZZLE_128B: SBO = 1024 B between 8-row groups, version 1
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    mbar_wait(bar_a, 0);
    tc_fence_after();
    for (int64_t h = 0; h < n_half; ++h) {
      const int64_t t = h >> 1;
      const int half = (int)(h & 1);
      const int s = (int)(t % F_STAGES);
      const int a = (int)(h % F_ACC);
      mbar_wait(bar_tempty + 8 * a, (uint32_t)(((h / F_ACC) & 1) ^ 1));
      if (half == 0) mbar_wait(bar_full + 8 * s, (uint32_t)((t / F_STAGES) & 1));
      tc_fence_after();
      if (elect_one()) {
        // keys [64 half, +64) of the tile: 8 swizzle groups of 1024 bytes further
        const uint32_t b_addr = smem_u32(sB + (size_t)s * F_TILE_BYTES) + (uint32_t)half * (F_HN * 128);
        const uint32_t b_lo = ((b_addr >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const uint32_t d_tmem = tmem_base + (uint32_t)((a * 2 + blk) * F_HN);
          const uint32_t a_tmem = tmem_base + F_TMEM_A + (uint32_t)blk * 32u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < ksteps) {
              // one k-step = 16 halves = 32 bytes of the swizzled row = 8 packed TMEM columns of A
              const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + (uint32_t)(ks * 2));
              tc_mma_f16_ts(d_tmem, a_tmem + (uint32_t)ks * 8u, db, idesc, ks ? 1u : 0u);
            }
          }
        }
        if (half == 1) tc_commit(bar_empty + 8 * s);     // the key stage is reusable once both halves were read
        tc_commit(bar_tfull + 8 * a);
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue: one thread = one query row =================
    const unsigned FULL = 0xffffffffu;
    const int blk = (warp - 2) >> 2;                    // which 128-row block
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = blk * 128 + quarter * 32 + lane;
    const int64_t qrow = q0 + row;
    const bool qvalid = qrow < n_q;
    // the key that is this row itself (excluded): a row range of the keys, or an explicit position per row
    const int self_i = self_pos ? (qvalid ? self_pos[qrow] : -1) : (int)(query_offset + qrow);
    float* scratch = scratch_all + (warp - 2) * 32 * 32;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    {
      // stage this thread's query row into TMEM: lane = row within the block, 32 columns of two packed halves
      uint32_t v[32];
      const uint4* src = reinterpret_cast<const uint4*>(q16 + (qvalid ? qrow : 0) * F_DP);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 f = qvalid ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0u);
        v[4 * i] = f.x;
        v[4 * i + 1] = f.y;
        v[4 * i + 2] = f.z;
        v[4 * i + 3] = f.w;
      }
      tmem_st32(lane_base + F_TMEM_A + (uint32_t)blk * 32u, v);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(bar_a);
    }
    RowState st;
    for (int j = 0; j < F_L; ++j) {
      l_sc[j * F_TQ + row] = -INFINITY;
      l_id[j * F_TQ + row] = INT_MAX;
    }
    st.tau = qvalid ? -INFINITY : INFINITY;             // the row's 32nd best so far (the minimum of its set)
    st.slot = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      st.gm[q] = -INFINITY;
      st.gs[q] = q * 8;
    }
    __syncwarp();
    // (The loop is bound by the TMEM read rate -- every fp32 score has to come out of tensor memory once, ~94 B/clk
    // per SM measured -- so prefetching the next half tile into a second register set bought nothing and spilled.)
    for (int64_t h = 0; h < n_half; ++h) {
      const int a = (int)(h % F_ACC);
      mbar_wait(bar_tfull + 8 * a, (uint32_t)((h / F_ACC) & 1));
      tc_fence_after();
      int64_t tile = (h >> 1) + t_start;
      if (tile >= n_tiles) tile -= n_tiles;
      const int key0 = (int)(tile * F_BN) + (int)(h & 1) * F_HN;
      uint32_t r0[32], r1[32];
      const uint32_t taddr = lane_base + (uint32_t)((a * 2 + blk) * F_HN);
      tmem_ld32(taddr, r0);
      tmem_ld32(taddr + 32, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // the accumulator is in registers: hand the TMEM stage back before looking at the scores
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * a);
      const float m0 = max32(r0), m1 = max32(r1);
      if (dbg_mode != 1 && __any_sync(FULL, fmaxf(m0, m1) > st.tau)) {
        const long long ts = cnt ? clock64() : 0;
        if (__any_sync(FULL, m0 > st.tau)) slow_chunk(r0, scratch, key0, self_i, (int)n_keys, l_sc, l_id, row, lane, st, cnt);
        if (__any_sync(FULL, m1 > st.tau)) slow_chunk(r1, scratch, key0 + 32, self_i, (int)n_keys, l_sc, l_id, row, lane, st, cnt);
        if (cnt && lane == 0) {
          atomicAdd(cnt + 2, 1ull);
          atomicAdd(cnt + 4, (unsigned long long)(clock64() - ts));
        }
      }
    }
    __syncwarp();
    // the row's 32 candidates (unsorted) and a_L = the smallest approximate score among them
    if (qvalid) {
      for (int j = 0; j < F_L; ++j) out_idx[qrow * F_L + j] = l_id[j * F_TQ + row];
      out_thr[qrow] = st.tau * (1.f / (F_SCALE * F_SCALE));    // -inf: fewer than 32 keys
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// prepared fp32 rows [n x 64] -> fp16 rows scaled by 256
__global__ void k_to_f16(const float* __restrict__ x, int64_t n, __half* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = __float2half_rn(x[i] * F_SCALE);
}

// Exact re-scoring of the 32 candidates (fp64 accumulation over the fp32 operands), final ordering (descending
// score, ties by ascending id) and the completeness proof: rows whose exact k-th score does not clear
// a_L + eps are appended to the fallback list.  One thread per query row.
__global__ void __launch_bounds__(128)
k_knn_refine_prove(const float* __restrict__ q, const float* __restrict__ keys, int64_t n_q, int32_t k, int32_t d,
                   const float* __restrict__ thr, const int32_t* __restrict__ in_idx, const int32_t* __restrict__ perm_k,
                   const int32_t* __restrict__ perm_q, float* __restrict__ out_score, int32_t* __restrict__ out_idx,
                   int32_t* __restrict__ fallback_rows, int32_t* __restrict__ n_fallback) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_q) return;
  float s[F_L];
  int id[F_L];
  const float4* qr = reinterpret_cast<const float4*>(q + row * 64);
  int cnt = 0;
  for (int j = 0; j < F_L; ++j) {
    const int pos = in_idx[row * F_L + j];
    if (pos == INT_MAX) continue;
    const int cand = perm_k ? perm_k[pos] : pos;         // original key id: what is reported and what breaks ties
    const float4* kr = reinterpret_cast<const float4*>(keys + (int64_t)pos * 64);
    double acc = 0.0;
    for (int i = 0; i < (d + 3) / 4; ++i) {
      const float4 a = __ldg(qr + i), b = __ldg(kr + i);
      acc += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
    }
    const float v = (float)acc;
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
  const int64_t orow = perm_q ? perm_q[row] : row;
  for (int j = 0; j < k; ++j) {
    out_score[orow * k + j] = j < cnt ? s[j] : -INFINITY;
    out_idx[orow * k + j] = j < cnt ? id[j] : INT_MAX;
  }
  // a_L = -inf when fewer than 32 keys exist: the list holds every key and the proof is trivial
  const float a_l = thr[row];
  if (cnt < k || !(a_l + F_EPS < s[k - 1])) {
    if (a_l != -INFINITY) fallback_rows[atomicAdd(n_fallback, 1)] = (int32_t)row;
  }
}

__global__ void k_gather_rows(const float* __restrict__ src, const int32_t* __restrict__ rows, int32_t n, int64_t row_offset,
                              const int32_t* __restrict__ self_pos, float* __restrict__ dst, int32_t* __restrict__ self_ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * 16) return;
  const int r = (int)(i >> 4), c = (int)(i & 15);
  reinterpret_cast<float4*>(dst)[(int64_t)r * 16 + c] = __ldg(reinterpret_cast<const float4*>(src) + (int64_t)rows[r] * 16 + c);
  if (c == 0) self_ids[r] = self_pos ? self_pos[rows[r]] : (int32_t)(row_offset + rows[r]);
}
// results of the exact re-scan -> their rows; key positions -> original ids, equal scores ordered by original id.
// One thread per re-scanned row (k <= 32).
__global__ void k_scatter_results(const float* __restrict__ s, const int32_t* __restrict__ idx, const int32_t* __restrict__ rows,
                                  int32_t n, int32_t k, const int32_t* __restrict__ perm_k, const int32_t* __restrict__ perm_q,
                                  float* __restrict__ out_score, int32_t* __restrict__ out_idx) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float sc[32];
  int id[32];
  for (int j = 0; j < k; ++j) {
    const float v = s[(int64_t)r * k + j];
    const int pos = idx[(int64_t)r * k + j];
    const int oid = (perm_k && pos != INT_MAX) ? perm_k[pos] : pos;
    int p = j;
    while (p > 0 && sc[p - 1] == v && id[p - 1] > oid) {      // scores arrive sorted: only ties need reordering
      sc[p] = sc[p - 1];
      id[p] = id[p - 1];
      --p;
    }
    sc[p] = v;
    id[p] = oid;
  }
  const int64_t orow = perm_q ? perm_q[rows[r]] : rows[r];
  for (int j = 0; j < k; ++j) {
    out_score[orow * k + j] = sc[j];
    out_idx[orow * k + j] = id[j];
  }
}

csc_status make_map_f16(csc_ctx* ctx, EncodeTiledFn fn, CUtensorMap* map, const __half* ptr, int64_t n_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)F_DP, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)F_DP * 2};
  cuuint32_t box[2] = {(cuuint32_t)F_DP, (cuuint32_t)F_BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return csc_fail(ctx, CSC_ERR_CUDA, "cuTensorMapEncodeTiled (f16 keys) failed with %d", (int)r);
  return CSC_OK;
}

}  // namespace

// Unit-norm prepared rows only (cosine).  Writes the exact k best (score, id) per query, sorted.
csc_status knn_build_order(csc_ctx* ctx, const float* q_prep, int64_t n_q, const float* k_prep, int64_t n_keys, bool same,
                           int64_t query_offset, int tq, int bn, DevPtr* ks, DevPtr* qs, DevPtr* perm_k, DevPtr* perm_q,
                           DevPtr* self_pos, DevPtr* start_tile);

csc_status knn_scan_f16(csc_ctx* ctx, const float* q_prep, int64_t n_q, const float* k_prep, int64_t n_keys, bool same,
                        int64_t query_offset, int32_t k, int32_t d_eff, float* out_score, int32_t* out_idx,
                        int64_t* n_fallback_out) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CSC_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
      return csc_fail(ctx, CSC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (EncodeTiledFn)fn;
  }
  const int ksteps = (d_eff + 15) / 16;
  // locality order (knn_order.cu) once the scan is long enough to pay for the clustering; CSC_KNN_ORDER=0 disables
  DevPtr ks, qs, perm_k, perm_q, self_pos, start_tile;
  const char* ord_env = getenv("CSC_KNN_ORDER");
  const bool ordered = n_keys >= 32768 && n_keys < ((int64_t)1 << 31) && !(ord_env && atoi(ord_env) == 0);
  if (ordered) {
    CSC_TRY(knn_build_order(ctx, q_prep, n_q, k_prep, n_keys, same, query_offset, F_TQ, F_BN, &ks, &qs, &perm_k, &perm_q,
                            &self_pos, &start_tile));
    q_prep = (const float*)qs->p;
    k_prep = (const float*)ks->p;
  }
  const int32_t* d_perm_k = ordered ? (const int32_t*)perm_k->p : nullptr;
  const int32_t* d_perm_q = ordered ? (const int32_t*)perm_q->p : nullptr;
  const int32_t* d_self = ordered ? (const int32_t*)self_pos->p : nullptr;
  const int32_t* d_start = ordered ? (const int32_t*)start_tile->p : nullptr;
  DevPtr k16, q16, thr, cand, fb_rows, fb_count;
  const int64_t nk_el = n_keys * F_DP, nq_el = n_q * F_DP;
  CSC_TRY(dev_alloc(ctx, (size_t)nk_el * 2, &k16));
  k_to_f16<<<(unsigned)ceil_div64(nk_el, 256), 256, 0, ctx->stream>>>(k_prep, nk_el, (__half*)k16->p);
  CSC_LAUNCHED(ctx);
  if (same) {
    q16 = k16;
  } else {
    CSC_TRY(dev_alloc(ctx, (size_t)nq_el * 2, &q16));
    k_to_f16<<<(unsigned)ceil_div64(nq_el, 256), 256, 0, ctx->stream>>>(q_prep, nq_el, (__half*)q16->p);
    CSC_LAUNCHED(ctx);
  }
  CSC_TRY(dev_alloc(ctx, (size_t)n_q * 4, &thr));
  CSC_TRY(dev_alloc(ctx, (size_t)n_q * F_L * 4, &cand));
  CSC_TRY(dev_alloc(ctx, (size_t)n_q * 4, &fb_rows));
  CSC_TRY(dev_alloc(ctx, 4, &fb_count, true));
  DevPtr dbg_cnt;
  if (getenv("CSC_KNN_CNT")) CSC_TRY(dev_alloc(ctx, 128, &dbg_cnt, true));
  CUtensorMap kmap;
  CSC_TRY(make_map_f16(ctx, encode, &kmap, (const __half*)k16->p, n_keys));
  const size_t smem = 1024 + (size_t)F_TILE_BYTES * F_STAGES + (size_t)F_TQ * F_L * 8 + (size_t)8 * 32 * 32 * 4 + 512;
  CSC_CUDA(ctx, cudaFuncSetAttribute(k_knn_scan_f16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEventRecord(ctx->ev_k0, ctx->stream);
  k_knn_scan_f16<<<(unsigned)ceil_div64(n_q, F_TQ), F_THREADS, smem, ctx->stream>>>(
      kmap, (const __half*)q16->p, n_q, n_keys, query_offset, d_self, d_start, ksteps, (float*)thr->p, (int32_t*)cand->p,
      getenv("CSC_KNN_DBG") ? atoi(getenv("CSC_KNN_DBG")) : 0, dbg_cnt ? (unsigned long long*)dbg_cnt->p : nullptr);
  CSC_LAUNCHED(ctx);
  cudaEventRecord(ctx->ev_k1, ctx->stream);
  k_knn_refine_prove<<<(unsigned)ceil_div64(n_q, 128), 128, 0, ctx->stream>>>(
      q_prep, k_prep, n_q, k, d_eff, (const float*)thr->p, (const int32_t*)cand->p, d_perm_k, d_perm_q, out_score, out_idx,
      (int32_t*)fb_rows->p, (int32_t*)fb_count->p);
  CSC_LAUNCHED(ctx);
  int32_t n_fb = 0;
  CSC_CUDA(ctx, cudaMemcpyAsync(&n_fb, fb_count->p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1) == cudaSuccess) ctx->stage_ms[CSC_T_KNN_SCAN] += ms;
  if (n_fallback_out) *n_fallback_out = n_fb;
  if (dbg_cnt) {
    unsigned long long h[11];
    cudaMemcpy(h, dbg_cnt->p, sizeof(h), cudaMemcpyDeviceToHost);
    const double wt = (double)ceil_div64(n_q, F_TQ) * 8 * (double)(2 * ceil_div64(n_keys, F_BN));
    fprintf(stderr, "[knn f16] warp x half-tiles %.3g; slow entries %llu (%.2f%%), cycles in slow path %.3g (%.0f per entry)\n", wt, h[2],
            100.0 * h[2] / wt, (double)h[4], h[2] ? (double)h[4] / h[2] : 0.0);
    fprintf(stderr, "[knn f16] dense chunks %llu: %.3g cycles (%.0f each), %.3g candidates (%.1f each); sparse chunks %llu: %.3g cycles (%.0f each), %.3g candidates; cooperative attempts %llu accepted %llu\n",
            h[5], (double)h[7], h[5] ? (double)h[7] / h[5] : 0.0, (double)h[9], h[5] ? (double)h[9] / h[5] : 0.0, h[6], (double)h[8],
            h[6] ? (double)h[8] / h[6] : 0.0, (double)h[10], h[0], h[1]);
  }
  if (n_fb > 0) {
    // the unproven rows go through the exact 3xTF32 scan
    DevPtr gq, gself, gs, gi;
    CSC_TRY(dev_alloc(ctx, (size_t)n_fb * 64 * 4, &gq));
    CSC_TRY(dev_alloc(ctx, (size_t)n_fb * 4, &gself));
    CSC_TRY(dev_alloc(ctx, (size_t)n_fb * k * 4, &gs));
    CSC_TRY(dev_alloc(ctx, (size_t)n_fb * k * 4, &gi));
    k_gather_rows<<<(unsigned)ceil_div64((int64_t)n_fb * 16, 256), 256, 0, ctx->stream>>>(
        q_prep, (const int32_t*)fb_rows->p, n_fb, query_offset, d_self, (float*)gq->p, (int32_t*)gself->p);
    CSC_LAUNCHED(ctx);
    CSC_TRY(knn_scan_tc(ctx, (const float*)gq->p, n_fb, k_prep, n_keys, false, 0, (const int32_t*)gself->p, k, d_eff,
                        (float*)gs->p, (int32_t*)gi->p));
    k_scatter_results<<<(unsigned)ceil_div64(n_fb, 128), 128, 0, ctx->stream>>>(
        (const float*)gs->p, (const int32_t*)gi->p, (const int32_t*)fb_rows->p, n_fb, k, d_perm_k, d_perm_q, out_score, out_idx);
    CSC_LAUNCHED(ctx);
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return CSC_OK;
}
