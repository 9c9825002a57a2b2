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
// Kernel layout (576 threads, 1 CTA per SM, one CTA = 256 query rows = two 128-row blocks against ALL keys; the two
// blocks share every key tile, which halves the L2 traffic per score -- the scan is L2-bandwidth bound otherwise):
//   warp 0      TMA producer: key tiles [128 keys x 64 halves] (one 128-byte swizzle row per key), 3-stage ring
//   warp 1      TMEM allocator + MMA issuer: A (queries, both blocks) lives in TMEM (2 x 32 columns of packed halves),
//               D = 3 accumulator stages x 2 blocks x 64 columns; the unit of work is a half tile of 64 keys
//   warps 2-17  epilogue, 16 warps = 2 blocks x 2 column parts x 4 TMEM lane quarters: thread = (query row, part);
//               per half tile of 64 keys part p reads columns [32p, 32p+32) with ONE tcgen05.ld.x32, runs a FMNMX3 tree
//               and one compare against its threshold; survivors go into the thread's own 16-slot candidate set in
//               shared memory.  A warp gets ~43 B/clk out of TMEM however many loads it has in flight (x32 = 96 cycles;
//               tools/micro/tmem_ld_bench.cu: 4 warps 171, 8 warps 310, 16 warps 420 B/clk/SM), and it cannot look at
//               scores while it waits for them, so the rate at which a CTA drains its accumulators is set by the
//               number of epilogue warps per scheduler: 4 per scheduler instead of 2 halves the time per half tile.
//               The two parts of a row keep separate sets (no communication in the loop); everything outside both is
//               <= max(tau_0, tau_1), which is the bound the proof uses.
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
constexpr int F_BN = 128;                 // keys per shared-memory tile = UMMA N: one MMA unit is (key tile, block)
constexpr int F_DP = 64;                  // padded feature dimension (halves): one 128-byte swizzle row
constexpr int F_STAGES = 3;               // key tile ring (the MMA loop is unrolled over one turn of it)
constexpr int F_ACC = 3;                  // TMEM accumulators of 128 columns, used round-robin by the units
constexpr int F_TILE_BYTES = F_BN * 128;  // 16 KB
constexpr int F_PAIRS = 8;                // (block, TMEM lane quarter): 32 query rows each
constexpr int F_THREADS = 64 + 2 * F_PAIRS * 32;   // producer, MMA issuer, 8 drain warps, 8 insert warps
constexpr int F_LP = 64;                  // pitch of a row's candidate list in global memory (FL = 32 or 48 of them used)
constexpr int F_QS = 5;                   // parked chunks per pair (slots of 32 keys x 32 rows x 4 bytes)
constexpr int F_TMEM_A = F_ACC * F_BN;    // first column of the query operands (2 blocks x 32 columns)
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

// Candidate sets.  A pair of warps serves 32 query rows, thread = row in both.  The DRAIN warp empties the accumulators
// and only asks "does any row of mine have a score above its threshold in these 32 keys?"; if so it PARKS the 32 x 32
// chunk in the pair's queue in shared memory (8 x STS.128 per lane) and returns to the accumulator stream at once.  The INSERT warp works the queue off concurrently and owns the sets: FL = 32 or 48
// slots per row -- scores in REGISTERS of the row's thread, key ids in shared memory --, unsorted (the refine kernel orders by exact score
// anyway).  Each score carries its slot number in the 5 low mantissa bits, so the set's minimum -- the threshold tau --
// and the slot it sits in come out of one FMNMX3 tree; a new candidate overwrites that slot by predicated moves and the
// tree is redone: ~100 instructions without a dependent memory access.  Ties of the approximate score are irrelevant: a
// candidate enters only if it beats tau, so everything outside the set is <= tau up to the packing perturbation (< 32 ulp
// per value, < 64 ulp accumulated: 7.7e-6 relative), which the reported bound adds back (F_PACK_SLACK).
//   (Inserting inside the drain loop cost ~1800 cycles per chunk with a candidate, and since a TMEM buffer is refilled
//   only after all warps of its block have let go of it, one such warp per unit -- 69 % of the units have one -- held up
//   the tensor pipe for everybody: the scan ran at the pace of the slowest warp.  A warp-cooperative insert (lane = slot,
//   shuffles and warp reductions) is worse still: ~600 cycles per insertion, its collectives form one long dependent chain.)
constexpr float F_PACK_SLACK = 2.0e-5f;   // cosine units; |score| <= 1, 6 slot bits: 64 ulp per value, 128 ulp accumulated < 1.6e-5
constexpr float F_EMPTY = -3.0e38f;       // empty slot (finite: -inf with slot bits would be a NaN)
constexpr uint32_t F_SLOT_MASK = 63u;     // slot number in the 6 low mantissa bits (FL <= 64)

__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// minimum of FL values by a tree of 3-input minima
template <int N>
__device__ __forceinline__ float min_tree(const float (&x)[N]) {
  if constexpr (N == 1) {
    return x[0];
  } else if constexpr (N == 2) {
    return fminf(x[0], x[1]);
  } else {
    constexpr int M = (N + 2) / 3;
    float m[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (3 * i + 2 < N) m[i] = fmin3(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
      else if (3 * i + 1 < N) m[i] = fminf(x[3 * i], x[3 * i + 1]);
      else m[i] = x[3 * i];
    }
    return min_tree<M>(m);
  }
}
// v beats tau: it replaces the set's minimum.  ids = the row's candidate list in GLOBAL memory ([F_LP] ints, written
// through on every insertion: a fire-and-forget 4-byte store that stays in L2 -- round 1 kept the ids in 32 KB of
// shared memory, which now holds a deeper park queue and lets the set grow to 48 slots).
template <int FL>
__device__ __forceinline__ void set_insert(float (&x)[FL], int32_t* ids, float& tau, float v, int id) {
  const uint32_t slot = __float_as_uint(tau) & F_SLOT_MASK;
  const float vp = __uint_as_float((__float_as_uint(v) & ~F_SLOT_MASK) | slot);
#pragma unroll
  for (int q = 0; q < FL; ++q) x[q] = (slot == (uint32_t)q) ? vp : x[q];
  ids[slot] = id;
  tau = min_tree<FL>(x);
}

// slot layout: [8 column quads][32 lanes] float4
constexpr int F_SLOT_BYTES = 8 * 32 * 16;
__device__ __forceinline__ void park_chunk(const uint32_t (&r)[32], unsigned char* slot, int lane) {
  float4* sv = reinterpret_cast<float4*>(slot);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    sv[i * 32 + lane] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                    __uint_as_float(r[4 * i + 3]));
}

template <int KS, int FL>   // k-steps of 16 halves: ceil(d / 16); candidates kept per row
__global__ void __launch_bounds__(F_THREADS, 1)
k_knn_scan_f16(const __grid_constant__ CUtensorMap kmap, const __half* __restrict__ q16, int64_t n_q, int64_t n_keys,
               int64_t query_offset, const int32_t* __restrict__ self_pos, const int32_t* __restrict__ start_tile,
               float* __restrict__ out_thr, int32_t* __restrict__ out_idx, int n_whole, int n_parts,
               float* __restrict__ xtr_thr, int32_t* __restrict__ xtr_idx, int dbg_mode, unsigned long long* cnt) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment for the swizzled tiles, as an offset so that the pointers stay in the shared window
  unsigned char* sB = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);  // F_STAGES x 16 KB
  unsigned char* q_slots = sB + F_STAGES * F_TILE_BYTES;                            // [8 pairs][F_QS] parked chunks
  float* s_tau = reinterpret_cast<float*>(q_slots + F_PAIRS * F_QS * F_SLOT_BYTES);    // [256 rows] thresholds, insert -> drain
  int* q_key = reinterpret_cast<int*>(s_tau + F_TQ);                                  // [8 pairs][F_QS] first key of a parked chunk
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_key + F_PAIRS * F_QS);
  const uint32_t bar_full = smem_u32(bars + 0);        // [F_STAGES]
  const uint32_t bar_empty = smem_u32(bars + 8);       // [F_STAGES]
  const uint32_t bar_tfull = smem_u32(bars + 16);      // [F_ACC]
  const uint32_t bar_tempty = smem_u32(bars + 19);     // [F_ACC]
  const uint32_t bar_a = smem_u32(bars + 22);
  const uint32_t bar_qfull = smem_u32(bars + 24);      // [8 pairs][F_QS]  drain -> insert: slot holds a chunk
  const uint32_t bar_qempty = smem_u32(bars + 24 + F_PAIRS * F_QS);   // insert -> drain: slot may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24 + 2 * F_PAIRS * F_QS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTAs [0, n_whole) scan every key tile for their 256 rows.  The blocks behind them -- the thin last wave of a grid
  // that is not a multiple of the SM count -- are cut into n_parts CTAs each, every part scanning its share of the
  // key tiles with its own candidate sets: list 0 goes to the row's usual place, lists 1.. to the extra arrays, and
  // the refine kernel merges them (the completeness proof takes the largest of the parts' bounds).
  const int xi = (int)blockIdx.x - n_whole;
  const int qblock = xi < 0 ? (int)blockIdx.x : n_whole + xi / n_parts;
  const int part = xi < 0 ? 0 : xi % n_parts;
  const int64_t q0 = (int64_t)qblock * F_TQ;
  const int64_t n_tiles_all = (n_keys + F_BN - 1) / F_BN;
  const int64_t tile_lo = xi < 0 ? 0 : n_tiles_all * part / n_parts;
  const int64_t tile_hi = xi < 0 ? n_tiles_all : n_tiles_all * (part + 1) / n_parts;
  const int64_t n_tiles = tile_hi - tile_lo;            // tiles THIS CTA scans; tile numbers below are relative to tile_lo
  const int key_lo = (int)(tile_lo * F_BN);             // first key of the CTA's range
  // locality order (knn_order.cu): the scan fans out from the middle of the block's own cluster, tiles c, c+1, c-1,
  // c+2, ... (the key order is circular), so that similar keys come early and the thresholds settle; a part starts at
  // the tile of its range that is nearest to the cluster
  int64_t t_start = start_tile ? (int64_t)start_tile[qblock] % n_tiles_all : 0;
  t_start = min(max(t_start, tile_lo), tile_hi - 1) - tile_lo;

  if (threadIdx.x == 0) {
    for (int i = 0; i < F_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < F_ACC; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);                   // the 4 drain warps of the unit's block
    }
    mbar_init(bar_a, 256);
    for (int i = 0; i < F_PAIRS * F_QS; ++i) {
      mbar_init(bar_qfull + 8 * i, 1);
      mbar_init(bar_qempty + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < F_TQ) s_tau[threadIdx.x] = F_EMPTY;     // "no threshold yet": every score is a candidate
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The CTA owns all 512 columns, so the allocation starts at lane 0, column 0: addresses below are compile-time
  // constants (the MMA warp would otherwise move every one of them into a uniform register before each use).
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&kmap)) : "memory");
    for (int64_t t = 0; t < n_tiles; ++t) {
      const int s = (int)(t % F_STAGES);
      mbar_wait(bar_empty + 8 * s, (uint32_t)(((t / F_STAGES) & 1) ^ 1));
      if (elect_one()) {
        mbar_expect_tx(bar_full + 8 * s, F_TILE_BYTES);
        int64_t tile = (t & 1) ? t_start + ((t + 1) >> 1) : t_start - ((t + 1) >> 1);
        if (tile >= n_tiles) tile -= n_tiles;
        if (tile < 0) tile += n_tiles;
        tma_load_2d(smem_u32(sB + (size_t)s * F_TILE_BYTES), &kmap, bar_full + 8 * s, 0, key_lo + (int)(tile * F_BN));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // One unit = (key tile, block): D[128 rows x 128 keys] = A_blk (TMEM) x tile' in ksteps MMAs of K = 16.  Unit
    // u = 2 * tile + blk accumulates into TMEM buffer u % 3; a single in-order issuer keeps the buffer hand-over safe
    // (a block's epilogue cannot be ahead of the other block's commit on the same buffer).
    // This warp's instruction stream is the critical path of the whole kernel: a lone warp retires one dependent
    // instruction every ~4 cycles, and with run-time stage / buffer indices one 64-key step cost ~180 instructions
    // (R2UR moves, index arithmetic, barrier bookkeeping) = ~750 cycles against 256 of tensor work.  Hence N = 128
    // per MMA and a loop unrolled over one turn of the key ring (6 tiles = 12 units = 4 turns of the buffers): every
    // descriptor, TMEM address and barrier parity but one is a compile-time constant.
    // instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(F_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // K-major SWIZZLE_128B: SBO = 1024 B between 8-row groups, version 1
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_lo0 = ((smem_u32(sB) >> 4) & 0x3FFF) | (1u << 16);
    mbar_wait(bar_a, 0);
    tc_fence_after();
    uint32_t ring_par = 0;
    for (int64_t t0 = 0; t0 < n_tiles; t0 += F_STAGES, ring_par ^= 1u) {
#pragma unroll
      for (int j = 0; j < F_STAGES; ++j) {
        if (t0 + j < n_tiles) {
          mbar_wait(bar_full + 8 * j, ring_par);
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            constexpr int kUnitsPerTurn = 2 * F_STAGES;
            static_assert(kUnitsPerTurn % F_ACC == 0 && ((kUnitsPerTurn / F_ACC) & 1) == 0,
                          "buffer parities must repeat every turn of the key ring");
            const int u = 2 * j + blk;
            const int a = u % F_ACC;
            const uint32_t kpar = (uint32_t)((u / F_ACC) & 1);
            mbar_wait(bar_tempty + 8 * a, kpar ^ 1u);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t d_tmem = tmem_base + (uint32_t)(a * F_BN);
              const uint32_t a_tmem = tmem_base + F_TMEM_A + (uint32_t)blk * 32u;
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                // one k-step = 16 halves = 32 bytes of the swizzled row = 8 packed TMEM columns of A
                const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo0 + (uint32_t)(j * (F_TILE_BYTES >> 4) + ks * 2));
                tc_mma_f16_ts(d_tmem, a_tmem + (uint32_t)ks * 8u, db, idesc, ks ? 1u : 0u);
              }
              if (blk == 1) tc_commit(bar_empty + 8 * j);     // the key stage is reusable once both blocks have read it
              tc_commit(bar_tfull + 8 * a);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp < 2 + F_PAIRS) {
    // ================= drain: one thread = one query row, accumulators -> registers -> "any candidate?" =================
    const unsigned FULL = 0xffffffffu;
    const int blk = (warp - 2) >> 2;                    // which 128-row block
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int pair = blk * 4 + quarter;
    const int row = blk * 128 + quarter * 32 + lane;
    const int64_t qrow = q0 + row;
    const bool qvalid = qrow < n_q;
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
    unsigned char* my_slots = q_slots + pair * (F_QS * F_SLOT_BYTES);
    const volatile float* my_tau = s_tau + row;
    int qs = 0;                                         // next queue slot to fill
    uint32_t qpar = 1u;                                 // parity that says "slot qs is free" (fresh barriers pass 1)
    // look at a chunk; if some row has a candidate: wait for the slot, dump, publish
#define F_LOOK(r, kb)                                                                  \
  do {                                                                                 \
    if (__any_sync(FULL, max32(r) > tau) && dbg_mode != 1) {                            \
      const long long tw0 = cnt ? clock64() : 0;                                       \
      mbar_wait(bar_qempty + 8 * (pair * F_QS + qs), qpar);                             \
      if (cnt && lane == 0) {                                                          \
        atomicAdd(cnt + 0, 1ull);                                                      \
        atomicAdd(cnt + 2, (unsigned long long)(clock64() - tw0));                     \
      }                                                                                \
      park_chunk(r, my_slots + qs * F_SLOT_BYTES, lane);                                \
      if (lane == 0) q_key[pair * F_QS + qs] = (kb);                                    \
      __syncwarp();                                                                    \
      if (lane == 0) mbar_arrive(bar_qfull + 8 * (pair * F_QS + qs));                   \
      if (++qs == F_QS) {                                                              \
        qs = 0;                                                                        \
        qpar ^= 1u;                                                                    \
      }                                                                                \
    }                                                                                  \
  } while (0)
    // unit u = 2 * tile + blk lives in buffer u % 3.  The four 32-column chunks of a unit go through two register sets:
    // a chunk is examined while the next one is on its way out of TMEM.
    int a = blk;
    uint32_t kpar = 0;
    for (int64_t t = 0; t < n_tiles; ++t) {
      const float tau = qvalid ? *my_tau : INFINITY;    // may lag behind the insert warp: only ever too low
      mbar_wait(bar_tfull + 8 * a, kpar);
      tc_fence_after();
      int64_t tile = (t & 1) ? t_start + ((t + 1) >> 1) : t_start - ((t + 1) >> 1);
      if (tile >= n_tiles) tile -= n_tiles;
      if (tile < 0) tile += n_tiles;
      const int key0 = key_lo + (int)(tile * F_BN);
      const uint32_t taddr = lane_base + (uint32_t)(a * F_BN);
      uint32_t rA[32], rB[32];
      tmem_ld32(taddr, rA);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_ld32(taddr + 32, rB);
      F_LOOK(rA, key0);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_ld32(taddr + 64, rA);
      F_LOOK(rB, key0 + 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_ld32(taddr + 96, rB);
      F_LOOK(rA, key0 + 64);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // the whole accumulator has been read: hand the TMEM buffer back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * a);
      a += 2;
      if (a >= F_ACC) {
        a -= F_ACC;
        kpar ^= 1u;
      }
      F_LOOK(rB, key0 + 96);
    }
    // end of stream
    mbar_wait(bar_qempty + 8 * (pair * F_QS + qs), qpar);
    if (lane == 0) {
      q_key[pair * F_QS + qs] = -1;
      mbar_arrive(bar_qfull + 8 * (pair * F_QS + qs));
    }
#undef F_LOOK
  } else {
    // ================= insert: one thread = one query row, owns the row's candidate set =================
    const int pair = warp - (2 + F_PAIRS);
    const int row = pair * 32 + lane;                   // pair = blk * 4 + quarter
    const int64_t qrow = q0 + row;
    const bool qvalid = qrow < n_q;
    // the key that is this row itself (excluded): a row range of the keys, or an explicit position per row
    const int self_i = self_pos ? (qvalid ? self_pos[qrow] : -1) : (int)(query_offset + qrow);
    float x[FL];
    // the row's candidate list in global memory: list 0 at the row's place; list p >= 1 of a split block at
    // [(p - 1) * split rows + row within the split blocks]
    const int64_t split_rows = ((int64_t)gridDim.x - n_whole) / max(n_parts, 1) * F_TQ;
    const int64_t xrow = (int64_t)(part - 1) * split_rows + (qrow - (int64_t)n_whole * F_TQ);
    int32_t* ids = part == 0 ? out_idx + (qvalid ? qrow : 0) * F_LP : xtr_idx + (qvalid ? xrow : 0) * F_LP;
#pragma unroll
    for (int j = 0; j < FL; ++j) x[j] = __uint_as_float((__float_as_uint(F_EMPTY) & ~F_SLOT_MASK) | (uint32_t)j);
    if (qvalid) {
      int4* i4 = reinterpret_cast<int4*>(ids);
#pragma unroll
      for (int j = 0; j < F_LP / 4; ++j) i4[j] = make_int4(INT_MAX, INT_MAX, INT_MAX, INT_MAX);
    }
    float tau = min_tree<FL>(x);
    const unsigned char* my_slots = q_slots + pair * (F_QS * F_SLOT_BYTES);
    int qs = 0;
    uint32_t qpar = 0u;
    while (true) {
      mbar_wait(bar_qfull + 8 * (pair * F_QS + qs), qpar);
      const int key_base = q_key[pair * F_QS + qs];
      if (key_base < 0) break;
      const long long tp0 = cnt ? clock64() : 0;
      int n_ins = 0;
      const unsigned char* slot = my_slots + qs * F_SLOT_BYTES;
      const float* sv = reinterpret_cast<const float*>(slot);
      unsigned cm4[4] = {0u, 0u, 0u, 0u};                 // columns that beat tau (four independent partial masks)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = reinterpret_cast<const float4*>(slot)[i * 32 + lane];
        cm4[0] |= (v.x > tau ? 1u : 0u) << (4 * i);
        cm4[1] |= (v.y > tau ? 1u : 0u) << (4 * i + 1);
        cm4[2] |= (v.z > tau ? 1u : 0u) << (4 * i + 2);
        cm4[3] |= (v.w > tau ? 1u : 0u) << (4 * i + 3);
      }
      unsigned cm = (cm4[0] | cm4[1]) | (cm4[2] | cm4[3]);
      while (cm) {
        const int j = __ffs(cm) - 1;
        cm &= cm - 1;
        const int id = key_base + j;
        const float v = sv[((j >> 2) * 32 + lane) * 4 + (j & 3)];
        if (v > tau && id != self_i && id < (int)n_keys && qvalid) {
          set_insert<FL>(x, ids, tau, v, id);
          ++n_ins;
        }
      }
      s_tau[row] = tau;
      __syncwarp();
      if (cnt) {
        for (int o = 16; o > 0; o >>= 1) n_ins += __shfl_xor_sync(0xffffffffu, n_ins, o);
        if (lane == 0) {
          atomicAdd(cnt + 3, (unsigned long long)(clock64() - tp0));
          atomicAdd(cnt + 4, (unsigned long long)n_ins);
        }
      }
      if (lane == 0) mbar_arrive(bar_qempty + 8 * (pair * F_QS + qs));
      if (++qs == F_QS) {
        qs = 0;
        qpar ^= 1u;
      }
    }
    // the candidate ids are in place (written through on insertion); the bound on every score outside the set.
    // A set that never filled up reports -inf: it holds every key (of the part's range)
    if (qvalid)
      (part == 0 ? out_thr + qrow : xtr_thr + xrow)[0] = tau < -1.0e37f ? -INFINITY : tau * (1.f / (F_SCALE * F_SCALE)) + F_PACK_SLACK;
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
// a_L + eps are appended to the fallback list.  One WARP per query row, one lane per candidate, 32 candidates at a
// time (a list of FL = 48 is two sub-lists, the second half empty; the running best 32 are merged like the lists of
// the key-range parts -- k <= 32, so nothing below rank 32 matters): the candidate list is read coalesced, the 32 key
// rows are gathered side by side, and the order comes from a bitonic network over the lanes.  (One thread per row with an insertion sort kept its 32-entry lists in local memory and
// read the candidate table with a 128-byte stride: 2.6 ms at 500k rows.)
// the exact score of a (query, key) pair: fp64 accumulation of the fp32 products (one expression, shared by the refine
// kernel and the direct scan of the unproven rows so that both produce the same bits)
__device__ __forceinline__ float exact_score64(const float4* __restrict__ qr, const float4* __restrict__ kr, int d) {
  double acc = 0.0;
  for (int i = 0; i < (d + 3) / 4; ++i) {
    const float4 a = __ldg(qr + i), b = __ldg(kr + i);
    acc += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
  }
  return (float)acc;
}

__device__ __forceinline__ bool ranks_before(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }

// sorts (v, id) over the 32 lanes: afterwards lane j holds the j-th best
__device__ __forceinline__ void bitonic_sort32(float& v, int& id, int lane, int first_k2) {
  const unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
    if (k2 < first_k2) continue;                          // first_k2 = 32: only the last merge (the input is bitonic)
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      const float ov = __shfl_xor_sync(FULL, v, j);
      const int oid = __shfl_xor_sync(FULL, id, j);
      const bool keep_first = ((lane & j) == 0) == ((lane & k2) == 0);
      const bool other_first = ranks_before(ov, oid, v, id);
      const bool mine_first = ranks_before(v, id, ov, oid);
      if (keep_first ? other_first : mine_first) {
        v = ov;
        id = oid;
      }
    }
  }
}

// Rows from split_row0 on have n_lists candidate lists (the key-range parts of the scan's last wave): list 0 in
// in_idx / thr, list p >= 1 in xtr_idx / xtr_thr at row (p - 1) * split_rows + (row - split_row0).
__global__ void __launch_bounds__(256)
k_knn_refine_prove(const float* __restrict__ q, const float* __restrict__ keys, int64_t n_q, int32_t k, int32_t d,
                   const float* __restrict__ thr, const int32_t* __restrict__ in_idx, int64_t split_row0, int64_t split_rows,
                   int32_t n_lists, int32_t n_sub, const float* __restrict__ xtr_thr, const int32_t* __restrict__ xtr_idx,
                   const int32_t* __restrict__ perm_k, const int32_t* __restrict__ perm_q, float* __restrict__ out_score,
                   int32_t* __restrict__ out_idx, int32_t* __restrict__ fallback_rows, int32_t* __restrict__ n_fallback) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_q) return;
  const int lists = row >= split_row0 ? n_lists : 1;
  const float4* qr = reinterpret_cast<const float4*>(q + row * 64);
  float v = -INFINITY;                                    // running best 32, sorted over the lanes
  int id = INT_MAX;                                       // original key id: what is reported and what breaks ties
  int cnt = 0;
  float a_l = -INFINITY;
  for (int ls = 0; ls < lists * n_sub; ++ls) {
    const int l = ls / n_sub, sub = ls - l * n_sub;
    const int64_t xrow = (int64_t)(l - 1) * split_rows + (row - split_row0);
    const int pos = l == 0 ? in_idx[row * F_LP + sub * 32 + lane] : xtr_idx[xrow * F_LP + sub * 32 + lane];
    a_l = fmaxf(a_l, l == 0 ? thr[row] : xtr_thr[xrow]);
    const bool valid = pos != INT_MAX;
    cnt += __popc(__ballot_sync(FULL, valid));
    float cv = -INFINITY;
    int cid = INT_MAX;
    if (valid) {
      cid = perm_k ? perm_k[pos] : pos;
      cv = exact_score64(qr, reinterpret_cast<const float4*>(keys + (int64_t)pos * 64), d);
    }
    bitonic_sort32(cv, cid, lane, 2);
    if (ls == 0) {
      v = cv;
      id = cid;
    } else {
      // best 32 of two sorted lists: the better of (mine, the other list reversed) per lane is a bitonic sequence
      const float rv = __shfl_sync(FULL, cv, 31 - lane);
      const int rid = __shfl_sync(FULL, cid, 31 - lane);
      if (ranks_before(rv, rid, v, id)) {
        v = rv;
        id = rid;
      }
      bitonic_sort32(v, id, lane, 32);
    }
  }
  const int64_t orow = perm_q ? perm_q[row] : row;
  if (lane < k) {
    out_score[orow * k + lane] = lane < cnt ? v : -INFINITY;
    out_idx[orow * k + lane] = lane < cnt ? id : INT_MAX;
  }
  // a_L = -inf when a list never filled up: it holds every key of its range and the proof is trivial for it
  const float s_k = __shfl_sync(FULL, v, (k - 1) & 31);
  if (lane == 0) {
    if (cnt < k || !(a_l + F_EPS < s_k)) {
      if (a_l != -INFINITY) fallback_rows[atomicAdd(n_fallback, 1)] = (int32_t)row;
    }
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

// ---- a handful of unproven rows: direct exact scan ---------------------------------------------------------------
// The 3xTF32 re-scan works on blocks of 128 query rows, one CTA per block against ALL keys: for the one or two rows a
// 1M-cell run leaves unproven that is a single CTA walking a million keys (7.3 ms of a 190 ms step, ncu r2).  Few rows
// are scanned directly instead: every (row, key) pair gets the exact score of the refine kernel (fp64 accumulation of
// the fp32 products, same expression, hence the same bits), a key whose score reaches the row's current k-th best (the
// exact k-th of its candidate list) is appended to the row's buffer -- that is every member of the true top k plus
// the few keys that tie or beat the list's tail --, and one CTA per row orders the buffer by (score descending,
// original id ascending) and writes the first k distinct keys.  Rows with fewer than k candidates so far, or whose
// buffer overflows, report failure and the caller takes the tensor-core re-scan after all.
constexpr int FB_CAP = 1024;             // buffer entries per row
constexpr int FB_DIRECT_MAX = 512;       // rows up to which the direct scan is used

__global__ void __launch_bounds__(256)
k_knn_fb_scan(const float* __restrict__ q, const float* __restrict__ keys, int64_t n_keys, int32_t d, int32_t k,
              const int32_t* __restrict__ fb_rows, int64_t query_offset, const int32_t* __restrict__ self_pos,
              const int32_t* __restrict__ perm_q, const float* __restrict__ out_score, int32_t* __restrict__ buf_id,
              float* __restrict__ buf_score, int32_t* __restrict__ buf_n) {
  const int r = blockIdx.y;
  const int row = fb_rows[r];
  const int64_t orow = perm_q ? perm_q[row] : row;
  const float s_k = out_score[orow * k + (k - 1)];       // -inf when the list held fewer than k keys: everything qualifies
  const int self_i = self_pos ? self_pos[row] : (int)(query_offset + row);
  const float4* qr = reinterpret_cast<const float4*>(q + (int64_t)row * 64);
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_keys; p += (int64_t)gridDim.x * blockDim.x) {
    if (p == self_i) continue;
    const float s = exact_score64(qr, reinterpret_cast<const float4*>(keys + p * 64), d);
    if (s >= s_k) {
      const int slot = atomicAdd(buf_n + r, 1);
      if (slot < FB_CAP) {
        buf_id[(int64_t)r * FB_CAP + slot] = (int32_t)p;
        buf_score[(int64_t)r * FB_CAP + slot] = s;
      }
    }
  }
}

// one CTA per row: bitonic sort of the buffer by (score descending, original id ascending), first k distinct keys out
__global__ void __launch_bounds__(256)
k_knn_fb_merge(const int32_t* __restrict__ fb_rows, const int32_t* __restrict__ buf_id, const float* __restrict__ buf_score,
               const int32_t* __restrict__ buf_n, int32_t k, const int32_t* __restrict__ perm_k, const int32_t* __restrict__ perm_q,
               float* __restrict__ out_score, int32_t* __restrict__ out_idx, int32_t* __restrict__ failed) {
  __shared__ unsigned long long key[FB_CAP];
  const int r = blockIdx.x;
  const int n = buf_n[r];
  if (n > FB_CAP || n < k) {
    if (threadIdx.x == 0) failed[r] = 1;
    return;
  }
  for (int i = threadIdx.x; i < FB_CAP; i += blockDim.x) {
    unsigned long long v = 0ull;                          // sorts last (descending order)
    if (i < n) {
      const float s = buf_score[(int64_t)r * FB_CAP + i];
      const uint32_t u = __float_as_uint(s);
      const uint32_t o = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      const int pos = buf_id[(int64_t)r * FB_CAP + i];
      const uint32_t oid = (uint32_t)(perm_k ? perm_k[pos] : pos);
      v = ((unsigned long long)o << 32) | (unsigned long long)(0xFFFFFFFFu - oid);
    }
    key[i] = v;
  }
  __syncthreads();
  for (int k2 = 2; k2 <= FB_CAP; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < FB_CAP; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const bool desc = (i & k2) == 0;
          const unsigned long long a = key[i], b = key[l];
          if (desc ? a < b : a > b) {
            key[i] = b;
            key[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const int64_t orow = perm_q ? perm_q[fb_rows[r]] : fb_rows[r];
    int w = 0;
    unsigned long long prev = ~0ull;
    for (int i = 0; i < n && w < k; ++i) {
      if (key[i] == prev) continue;                       // the same key collected twice cannot happen; equal (score, id) can only be it
      prev = key[i];
      const uint32_t o = (uint32_t)(key[i] >> 32);
      out_score[orow * k + w] = __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
      out_idx[orow * k + w] = (int32_t)(0xFFFFFFFFu - (uint32_t)(key[i] & 0xFFFFFFFFull));
      ++w;
    }
    failed[r] = w < k ? 1 : 0;
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
  CSC_TRY(dev_alloc(ctx, (size_t)n_q * F_LP * 4, &cand));
  // The scan runs one CTA per SM.  When the number of 256-row blocks is not a multiple of the SM count, the last wave
  // is thin (1954 blocks on 148 SMs: 13 full waves and 30 CTAs that take as long as a full wave).  Those blocks are
  // cut into n_parts CTAs over the key range instead (see the kernel); CSC_KNN_SPLIT=0 disables.
  const int64_t n_blocks = ceil_div64(n_q, F_TQ);
  const int64_t n_key_tiles = ceil_div64(n_keys, F_BN);
  int64_t n_split_blocks = n_blocks % ctx->sm_count;
  int n_parts = 1;
  if (n_split_blocks > 0 && !(getenv("CSC_KNN_SPLIT") && atoi(getenv("CSC_KNN_SPLIT")) == 0)) {
    n_parts = (int)std::min<int64_t>(std::min<int64_t>(8, ctx->sm_count / n_split_blocks), n_key_tiles / 8);
    if (const char* force = getenv("CSC_KNN_SPLIT")) n_parts = std::max(1, std::min<int>(atoi(force), (int)std::max<int64_t>(1, n_key_tiles)));
  }
  if (n_parts < 2) {
    n_parts = 1;
    n_split_blocks = 0;
  }
  const int64_t n_whole = n_blocks - n_split_blocks;
  const int64_t split_rows = n_split_blocks * F_TQ;
  DevPtr xthr, xcand;
  if (n_parts > 1) {
    CSC_TRY(dev_alloc(ctx, (size_t)(n_parts - 1) * split_rows * 4, &xthr));
    CSC_TRY(dev_alloc(ctx, (size_t)(n_parts - 1) * split_rows * F_LP * 4, &xcand));
  }
  CSC_TRY(dev_alloc(ctx, (size_t)n_q * 4, &fb_rows));
  CSC_TRY(dev_alloc(ctx, 4, &fb_count, true));
  DevPtr dbg_cnt;
  if (getenv("CSC_KNN_CNT")) CSC_TRY(dev_alloc(ctx, 128, &dbg_cnt, true));
  CUtensorMap kmap;
  CSC_TRY(make_map_f16(ctx, encode, &kmap, (const __half*)k16->p, n_keys));
  const size_t smem = 1024 + (size_t)F_TILE_BYTES * F_STAGES + (size_t)F_PAIRS * F_QS * F_SLOT_BYTES + (size_t)F_TQ * 4 + 128 + (24 + 2 * F_PAIRS * F_QS) * 8 + 64;
  const int dbg_mode = getenv("CSC_KNN_DBG") ? atoi(getenv("CSC_KNN_DBG")) : 0;
  cudaEventRecord(ctx->ev_k0, ctx->stream);
  // candidates kept per row: the proof needs e_k (exact k-th best) to clear a_L (approximate L-th best) by eps = 1e-3.
  // L = 32 does that for k <= 24 on all but ~0.05 % of the rows; k = 25..32 (run_clustering's default is 30,
  // processing.jl:309) takes L = 48.  CSC_KNN_L = 32 | 48 forces one.
  int fl = k <= 24 ? 32 : 48;
  if (const char* fe = getenv("CSC_KNN_L")) fl = atoi(fe) >= 48 ? 48 : 32;
#define F_LAUNCH(KS)                                                                                                     \
  do {                                                                                                                   \
    if (fl == 32) F_LAUNCH2(KS, 32);                                                                                     \
    else F_LAUNCH2(KS, 48);                                                                                              \
  } while (0)
#define F_LAUNCH2(KS, FL)                                                                                                \
  do {                                                                                                                   \
    CSC_CUDA(ctx, cudaFuncSetAttribute(k_knn_scan_f16<KS, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    k_knn_scan_f16<KS, FL><<<(unsigned)(n_whole + n_split_blocks * n_parts), F_THREADS, smem, ctx->stream>>>(             \
        kmap, (const __half*)q16->p, n_q, n_keys, query_offset, d_self, d_start, (float*)thr->p, (int32_t*)cand->p,       \
        (int)n_whole, n_parts, xthr ? (float*)xthr->p : nullptr, xcand ? (int32_t*)xcand->p : nullptr, dbg_mode,          \
        dbg_cnt ? (unsigned long long*)dbg_cnt->p : nullptr);                                                             \
  } while (0)
  switch (ksteps) {
    case 1: F_LAUNCH(1); break;
    case 2: F_LAUNCH(2); break;
    case 3: F_LAUNCH(3); break;
    default: F_LAUNCH(4); break;
  }
#undef F_LAUNCH
#undef F_LAUNCH2
  CSC_LAUNCHED(ctx);
  cudaEventRecord(ctx->ev_k1, ctx->stream);
  k_knn_refine_prove<<<(unsigned)ceil_div64(n_q, 8), 256, 0, ctx->stream>>>(
      q_prep, k_prep, n_q, k, d_eff, (const float*)thr->p, (const int32_t*)cand->p, n_whole * F_TQ, split_rows, n_parts,
      fl / 32 + (fl % 32 ? 1 : 0), xthr ? (const float*)xthr->p : nullptr, xcand ? (const int32_t*)xcand->p : nullptr, d_perm_k, d_perm_q, out_score, out_idx,
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
    const double units = (double)ceil_div64(n_q, F_TQ) * F_PAIRS * (double)ceil_div64(n_keys, F_BN);   // (drain warp, key tile)
    fprintf(stderr,
            "[knn f16] %.3g warp-units; parked chunks %llu (%.3f per warp-unit), rows flagged per chunk %.2f, drain cycles "
            "blocked on a full queue %.3g (%.0f per park); insert warps: %.0f cycles per chunk, %.2f insertions per chunk, "
            "%.1f insertions per row\n",
            units, h[0], (double)h[0] / units, h[0] ? (double)h[1] / h[0] : 0.0, (double)h[2], h[0] ? (double)h[2] / h[0] : 0.0,
            h[0] ? (double)h[3] / h[0] : 0.0, h[0] ? (double)h[4] / h[0] : 0.0, (double)h[4] / (double)n_q);
  }
  bool direct_ok = false;
  if (n_fb > 0 && n_fb <= FB_DIRECT_MAX && !(getenv("CSC_KNN_FB_DIRECT") && atoi(getenv("CSC_KNN_FB_DIRECT")) == 0)) {
    // a handful of unproven rows: direct exact scan (see k_knn_fb_scan)
    DevPtr bid, bsc, bn, bfail;
    CSC_TRY(dev_alloc(ctx, (size_t)n_fb * FB_CAP * 4, &bid));
    CSC_TRY(dev_alloc(ctx, (size_t)n_fb * FB_CAP * 4, &bsc));
    CSC_TRY(dev_alloc(ctx, (size_t)n_fb * 4, &bn, true));
    CSC_TRY(dev_alloc(ctx, (size_t)n_fb * 4, &bfail, true));
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n_keys, 256), (int64_t)ctx->sm_count * 8 / std::min(n_fb, 8)));
    k_knn_fb_scan<<<dim3(gx, (unsigned)n_fb), 256, 0, ctx->stream>>>(q_prep, k_prep, n_keys, d_eff, k, (const int32_t*)fb_rows->p,
                                                                     query_offset, d_self, d_perm_q, out_score, (int32_t*)bid->p,
                                                                     (float*)bsc->p, (int32_t*)bn->p);
    CSC_LAUNCHED(ctx);
    k_knn_fb_merge<<<(unsigned)n_fb, 256, 0, ctx->stream>>>((const int32_t*)fb_rows->p, (const int32_t*)bid->p, (const float*)bsc->p,
                                                            (const int32_t*)bn->p, k, d_perm_k, d_perm_q, out_score, out_idx,
                                                            (int32_t*)bfail->p);
    CSC_LAUNCHED(ctx);
    std::vector<int32_t> hfail((size_t)n_fb);
    CSC_CUDA(ctx, cudaMemcpyAsync(hfail.data(), bfail->p, (size_t)n_fb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    direct_ok = true;
    for (int32_t f : hfail) direct_ok = direct_ok && f == 0;
  }
  if (n_fb > 0 && !direct_ok) {
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
