// PCA Gram matrix G += Z'Z on the 5th-generation tensor cores (run_pca, src/scrna/processing.jl:164-193: the
// reference hands the centred H x N matrix to a LAPACK SVD; here the same decomposition goes through the
// H x H Gram matrix).
//
//   Z is the cell-major scaled matrix [N cells][H genes] fp32.  G[a][b] = sum_c Z[c][a] Z[c][b]: the contraction
//   runs over cells, i.e. over the SLOW index of Z, so both MMA operands are "MN-major" (the 128-byte rows that
//   TMA lays down in shared memory run along the gene axis).  tcgen05 takes MN-major operands directly (16-bit:
//   plain 128-byte swizzle; tf32 would need the 32-byte-atomicity variant), so no transpose is ever materialised.
//
//   Split precision: Z = hi + lo with hi = fp16(Z), lo = fp16(Z - hi) (k_split_colsum, fused with the fp64 column
//   sums that the centring needs): 22 significant bits, the same as the tf32 hi / fp32 lo split of 3xTF32.
//   Z'Z ~= hi'hi + hi'lo + lo'hi, three kind::f16 MMA passes into one fp32 TMEM accumulator; products of 11-bit
//   factors are exact in fp32, so nothing is lost that 3xTF32 keeps -- but a K = 16 fp16 MMA does the work of two
//   K = 8 tf32 MMAs in the time of one, and the operands are half as many bytes.  The second point mattered as much
//   as the first: with fp32 operands the kernel pulled 108 GB per launch out of L2 (48 KB per 16 cells and CTA,
//   9.8 TB/s) and was L2-bound at 0.9 of what looked like the TF32 tensor peak.
//   (|Z| <= scale_max = 10: no overflow; values below fp16's normal range lose bits, an ABSOLUTE error <= 3e-8.)
//
//   Accumulation.  The tensor core adds into its fp32 accumulator with truncation, so a long same-sign sum (a
//   diagonal entry, a pair of correlated genes) picks up a relative bias proportional to the number of MMAs it
//   went through: 3e-5 after 4096 cells (measured), far too much for eigenvectors that must hold 1e-4.  The
//   kernel therefore lets the tensor core accumulate only 32 cells (6 MMAs, bias < 1e-6), then the flush warps
//   pull the tile out of TMEM and add it, round-to-nearest, into fp32 REGISTER accumulators (128 x 256 tile =
//   128 registers in each of 256 threads); TMEM is double-buffered so the tensor pipe never waits for that.
//   Every 4096 cells the registers are added into the CTA's private fp64 tile in global memory (plain
//   load-add-store, no atomics); k_gram_reduce sums the splits in a fixed order and mirrors the tiles into the
//   full symmetric fp64 matrix.
//
//   One CTA = one 128 x 256 tile (genes a in [128 ti, +128), b in [256 tj, +256), ti <= 2 tj + 1) over one
//   contiguous range of cells ("split"): 72 tiles x 2 splits = 144 CTAs for H = 2000.
//
// Warp roles (320 threads, 1 CTA per SM):
//   warp 0      TMA producer: [16 cells x 64 genes] fp16 boxes (SWIZZLE_128B, MN-major) of hi and lo into a 6-stage ring
//   warp 1      TMEM allocator + MMA issuer (one elected lane)
//   warps 2-9   flush: tcgen05.ld -> register accumulators -> fp64 tile
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int GT_BM = 128;                            // tile rows  = UMMA M
constexpr int GT_BN = 256;                            // tile cols  = UMMA N
constexpr int GT_BK = 16;                             // cells per pipeline stage
constexpr int GT_STAGES = 6;
constexpr int GT_ATOM_BYTES = GT_BK * 128;            // [16 cells][64 genes] fp16: one 128-byte row per cell
constexpr int GT_PART_BYTES = 6 * GT_ATOM_BYTES;      // hi or lo: A atoms 0-1 (128 genes), B atoms 2-5 (256 genes)
constexpr int GT_STAGE_BYTES = 2 * GT_PART_BYTES;     // hi | lo = 24 KB
constexpr int GT_THREADS = 64 + 256;
constexpr int GT_TC_STAGES = 2;                       // stages (32 cells) accumulated inside the tensor core
constexpr int GT_FLUSH_CHUNKS = 128;                  // chunks (4096 cells) accumulated in fp32 registers
constexpr int GT_TILE_ELEMS = GT_BM * GT_BN;

struct __align__(64) GramMaps {
  CUtensorMap hi_a, lo_a, hi_b, lo_b;     // boxes of 2 atoms (A operand) and of 4 atoms (B operand)
};

__global__ void __launch_bounds__(GT_THREADS, 1)
k_gram_tc(const __grid_constant__ GramMaps maps, const int2* __restrict__ pairs, int n_pairs, int64_t n_cells,
          int64_t cells_per_split, double* __restrict__ part) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* stages = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)GT_STAGES * GT_STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  const uint32_t bar_full = smem_u32(bars + 0);      // [GT_STAGES]
  const uint32_t bar_empty = smem_u32(bars + 8);     // [GT_STAGES]
  const uint32_t bar_tfull = smem_u32(bars + 16);    // [2] TMEM buffer holds a finished chunk
  const uint32_t bar_tempty = smem_u32(bars + 18);   // [2] TMEM buffer drained

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int2 pr = pairs[blockIdx.x];
  // the A genes are a sub-range of the B genes (the tile straddles the diagonal): load B only
  const bool diag = (pr.x >> 1) == pr.y;
  const int64_t c_begin = (int64_t)blockIdx.y * cells_per_split;
  const int64_t c_end = min(n_cells, c_begin + cells_per_split);
  const int64_t n_stages = c_end > c_begin ? (c_end - c_begin + GT_BK - 1) / GT_BK : 0;
  const int64_t n_chunks = (n_stages + GT_TC_STAGES - 1) / GT_TC_STAGES;
  double* out = part + ((size_t)blockIdx.y * n_pairs + blockIdx.x) * (size_t)GT_TILE_ELEMS;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GT_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 256);
    }
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
  // the CTA owns all 512 columns: the allocation starts at column 0 and TMEM addresses are compile-time constants
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.hi_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.lo_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.hi_b)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.lo_b)) : "memory");
    }
    // one box = [atoms][16 cells][64 genes]: the B operand (4 atoms) and, off the diagonal, the A operand (2 atoms) of hi
    // and of lo -- four TMA instructions per stage (a lone warp's instruction stream is what paces these kernels)
    const uint32_t tx_bytes = diag ? 8u * GT_ATOM_BYTES : 12u * GT_ATOM_BYTES;
    for (int64_t st = 0; st < n_stages; ++st) {
      const int s = (int)(st % GT_STAGES);
      mbar_wait(bar_empty + 8 * s, (uint32_t)(((st / GT_STAGES) & 1) ^ 1));
      if (elect_one()) {
        mbar_expect_tx(bar_full + 8 * s, tx_bytes);
        const uint32_t dst = smem_u32(stages + (size_t)s * GT_STAGE_BYTES);
        const int cell = (int)(c_begin + st * GT_BK);
        if (!diag) {
          tma_load_3d(dst, &maps.hi_a, bar_full + 8 * s, 0, cell, pr.x * 2);
          tma_load_3d(dst + GT_PART_BYTES, &maps.lo_a, bar_full + 8 * s, 0, cell, pr.x * 2);
        }
        tma_load_3d(dst + 2u * GT_ATOM_BYTES, &maps.hi_b, bar_full + 8 * s, 0, cell, pr.y * 4);
        tma_load_3d(dst + GT_PART_BYTES + 2u * GT_ATOM_BYTES, &maps.lo_b, bar_full + 8 * s, 0, cell, pr.y * 4);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // instruction descriptor: D = F32, A = B = F16 (format 0), both MN-major (bits 15, 16), N = 256, M = 128; one
    // instruction covers K = 16 cells.  Shared-memory descriptors: 128-byte swizzle (layout type 2), MN-major:
    // LBO = distance between 64-gene atoms, SBO = distance between 8-cell groups (1024 bytes), version 1.
    // (kind::tf32 MN-major operands exist only in the "128-byte swizzle with 32-byte atomicity" layout, type 1, TMA
    // CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, K-atom = 4 rows: with the plain swizzle that MMA silently produced zeros.)
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(GT_BN >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t lbo = (uint32_t)(GT_ATOM_BYTES >> 4) << 16;
    const uint32_t a_atom = diag ? 2u + (uint32_t)(pr.x & 1) * 2u : 0u;
    // The loop is unrolled over two turns of the stage ring (12 stages = 6 chunks = 3 turns of the two TMEM buffers), so
    // that stage, buffer, descriptors and all barrier parities but one are compile-time constants: with run-time indices
    // this warp spent ~950 cycles per stage on ~100 instructions while the tensor pipe needed 384 (pipe 46 % active).
    static_assert(GT_STAGES == 6 && GT_TC_STAGES == 2, "unrolling below assumes a 6-stage ring and 2 stages per chunk");
    const uint32_t sb0 = smem_u32(stages);
    uint32_t gpar = 0;                                     // parity of the group index g: 12 g stages are behind us
    for (int64_t st0 = 0; st0 < n_stages; st0 += 12, gpar ^= 1u) {
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        if (st0 + j < n_stages) {
          const int s = j % GT_STAGES;
          const int buf = (j >> 1) & 1;
          const bool chunk_first = (j & 1) == 0;
          const bool chunk_last = (j & 1) == 1 || st0 + j + 1 == n_stages;
          if (chunk_first) {
            // chunk c = 6 g + j/2 uses buffer c & 1 for the (c >> 1)-th time: parity (3 g + (j >> 2)) & 1
            mbar_wait(bar_tempty + 8 * buf, ((gpar + (uint32_t)(j >> 2)) & 1u) ^ 1u);
            tc_fence_after();
          }
          mbar_wait(bar_full + 8 * s, (uint32_t)((j / GT_STAGES) & 1));
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sb = sb0 + (uint32_t)(s * GT_STAGE_BYTES);
            const uint32_t d_tmem = (uint32_t)buf * GT_BN;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              // pass 0: hi'hi   pass 1: hi'lo   pass 2: lo'hi
              const uint32_t a_addr = sb + (pass == 2 ? GT_PART_BYTES : 0) + a_atom * GT_ATOM_BYTES;
              const uint32_t b_addr = sb + (pass == 1 ? GT_PART_BYTES : 0) + 2u * GT_ATOM_BYTES;
              const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(((a_addr >> 4) & 0x3FFF) | lbo);
              const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(((b_addr >> 4) & 0x3FFF) | lbo);
              tc_mma_f16_ss(d_tmem, da, db, idesc, (chunk_first && pass == 0) ? 0u : 1u);
            }
            tc_commit(bar_empty + 8 * s);
            if (chunk_last) tc_commit(bar_tfull + 8 * buf);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================= flush warps: TMEM -> fp32 registers -> private fp64 tile =================
    const int ew = warp - 2;                     // 0..7
    const int quarter = warp & 3;                // TMEM lane quarter this warp may access
    const int half = ew >> 2;                    // which 128 of the 256 columns
    const int row = quarter * 32 + lane;
    double* orow = out + (size_t)row * GT_BN + half * 128;
    float acc[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) acc[i] = 0.f;
    bool first_flush = true;
    for (int64_t ch = 0; ch < n_chunks; ++ch) {
      const int buf = (int)(ch & 1);
      mbar_wait(bar_tfull + 8 * buf, (uint32_t)((ch >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * GT_BN + half * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c * 32u, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * buf);
      if (((ch + 1) % GT_FLUSH_CHUNKS) == 0 || ch + 1 == n_chunks) {
        double2* dst = reinterpret_cast<double2*>(orow);
        if (first_flush) {
#pragma unroll
          for (int j = 0; j < 64; ++j) dst[j] = make_double2((double)acc[2 * j], (double)acc[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            double2 v = dst[j];
            v.x += (double)acc[2 * j];
            v.y += (double)acc[2 * j + 1];
            dst[j] = v;
          }
        }
        first_flush = false;
#pragma unroll
        for (int i = 0; i < 128; ++i) acc[i] = 0.f;
      }
    }
    if (n_chunks == 0) {
      // an empty cell range still owns a tile: it must read as zero
      for (int i = threadIdx.x - 64; i < GT_TILE_ELEMS; i += 256) out[i] = 0.0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// gram[a][b] (+)= sum over splits of the tile entries; tiles that do not straddle the diagonal are mirrored.  With
// `colsum` the operands carry a column of ones at index H: that column of the extended Gram is Z'1.
__global__ void __launch_bounds__(256)
k_gram_reduce(const double* __restrict__ part, const int2* __restrict__ pairs, int n_pairs, int n_splits, int H,
              double* __restrict__ gram, double* __restrict__ colsum) {
  const int2 pr = pairs[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // entry of the 128 x 256 tile
  const int r = idx / GT_BN, c = idx % GT_BN;
  const int a = pr.x * GT_BM + r, b = pr.y * GT_BN + c;
  if (a >= H || b > H || (b == H && !colsum)) return;
  double s = 0.0;
  for (int z = 0; z < n_splits; ++z) s += part[((size_t)z * n_pairs + blockIdx.y) * (size_t)GT_TILE_ELEMS + idx];
  if (b == H) {
    colsum[a] += s;
    return;
  }
  gram[(int64_t)a * H + b] += s;
  if ((pr.x >> 1) != pr.y) gram[(int64_t)b * H + a] += s;
}

// hi = fp16(z) (round to nearest), lo = fp16(z - hi), written with row pitch Hp; colsum[h] += sum over rows of z (fp64).
// One thread = 4 consecutive columns of a block of rows.
__global__ void __launch_bounds__(256)
k_split_colsum(const float* __restrict__ z, int64_t n_rows, int H, int Hp, int64_t rows_per_cta, __half* __restrict__ hi16,
               __half* __restrict__ lo16, double* __restrict__ colsum) {
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= Hp) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(r0 + rows_per_cta, n_rows);
  const bool vec = (H & 3) == 0;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t r = r0; r < r1; ++r) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const float* src = z + r * (int64_t)H + c0;
    if (c0 >= H) {
      // padding columns up to the next whole atom: zeros
    } else if (vec) {
      const float4 f = __ldcs(reinterpret_cast<const float4*>(src));
      v[0] = f.x;
      v[1] = f.y;
      v[2] = f.z;
      v[3] = f.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + j < H) v[j] = __ldcs(src + j);
    }
    __half h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __float2half_rn(v[j]);
      l[j] = __float2half_rn(v[j] - __half2float(h[j]));
      s[j] += (double)v[j];
    }
    __half2 p0 = __halves2half2(h[0], h[1]), p1 = __halves2half2(h[2], h[3]);
    *reinterpret_cast<uint2*>(hi16 + r * (int64_t)Hp + c0) = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
    p0 = __halves2half2(l[0], l[1]);
    p1 = __halves2half2(l[2], l[3]);
    *reinterpret_cast<uint2*>(lo16 + r * (int64_t)Hp + c0) = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (c0 + j < H) atomicAdd(colsum + c0 + j, s[j]);
}

// 3-D view of a [n_rows][Hp] fp16 matrix (Hp a multiple of 64, zero padded): (gene within atom, cell, atom); a box of
// `atoms` atoms lands in shared memory as [atom][16 cells][64 genes], each 128-byte row swizzled (SWIZZLE_128B)
csc_status gram_make_map16(csc_ctx* ctx, EncodeTiledFn fn, CUtensorMap* map, const __half* ptr, int64_t n_rows, int Hp, int atoms) {
  cuuint64_t dims[3] = {64, (cuuint64_t)n_rows, (cuuint64_t)(Hp / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)Hp * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)GT_BK, (cuuint32_t)atoms};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return csc_fail(ctx, CSC_ERR_CUDA, "cuTensorMapEncodeTiled (gram, fp16) failed with %d", (int)r);
  return CSC_OK;
}

}  // namespace

// gram[H*H] += hi'hi + hi'lo + lo'hi for operands that are split already: hi16 / lo16 are [n_rows][Hp] fp16 with
// Hp = gram_split_pitch(columns), padding columns zero -- except, with `colsum_from_ones`, column H, which holds 1 in
// hi (0 in lo) in every row: colsum[H] += Z'1 then comes out of the same MMAs.
int gram_split_pitch(int H) { return (H + 63) & ~63; }   // whole 64-gene atoms (the 3-D tensor maps step over atoms)

csc_status gram_tc_split(csc_ctx* ctx, const __half* hi16p, const __half* lo16p, int64_t n_rows, int H, int Hp, double* gram,
                         double* colsum_from_ones) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CSC_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
      return csc_fail(ctx, CSC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (EncodeTiledFn)fn;
  }
  DevPtr d_pairs, part;
  const int Hx = colsum_from_ones ? H + 1 : H;     // columns the tiles have to cover
  const int ntm = (H + GT_BM - 1) / GT_BM, ntn = (Hx + GT_BN - 1) / GT_BN;
  std::vector<int2> pairs;
  for (int tj = 0; tj < ntn; ++tj)
    for (int ti = 0; ti < ntm && ti <= 2 * tj + 1; ++ti) pairs.push_back(make_int2(ti, tj));
  const int n_pairs = (int)pairs.size();
  // splits: fill the SMs, but keep at least one register-flush period of cells per split
  const int64_t flush_cells = (int64_t)GT_FLUSH_CHUNKS * GT_TC_STAGES * GT_BK;
  int n_splits = std::max(1, ctx->sm_count / n_pairs);
  n_splits = (int)std::max<int64_t>(1, std::min<int64_t>(n_splits, ceil_div64(n_rows, flush_cells)));
  const int64_t align = (int64_t)GT_TC_STAGES * GT_BK;
  int64_t cells_per_split = ceil_div64(ceil_div64(n_rows, n_splits), align) * align;
  n_splits = (int)ceil_div64(n_rows, cells_per_split);
  CSC_TRY(dev_alloc(ctx, pairs.size() * sizeof(int2), &d_pairs));
  CSC_CUDA(ctx, cudaMemcpyAsync(d_pairs->p, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
  CSC_TRY(dev_alloc(ctx, (size_t)n_pairs * n_splits * GT_TILE_ELEMS * 8, &part));
  GramMaps maps;
  CSC_TRY(gram_make_map16(ctx, encode, &maps.hi_a, hi16p, n_rows, Hp, 2));
  CSC_TRY(gram_make_map16(ctx, encode, &maps.lo_a, lo16p, n_rows, Hp, 2));
  CSC_TRY(gram_make_map16(ctx, encode, &maps.hi_b, hi16p, n_rows, Hp, 4));
  CSC_TRY(gram_make_map16(ctx, encode, &maps.lo_b, lo16p, n_rows, Hp, 4));
  const size_t smem = 1024 + (size_t)GT_STAGES * GT_STAGE_BYTES + 256;
  CSC_CUDA(ctx, cudaFuncSetAttribute(k_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEventRecord(ctx->ev_k0, ctx->stream);
  k_gram_tc<<<dim3((unsigned)n_pairs, (unsigned)n_splits), GT_THREADS, smem, ctx->stream>>>(
      maps, (const int2*)d_pairs->p, n_pairs, n_rows, cells_per_split, (double*)part->p);
  CSC_LAUNCHED(ctx);
  cudaEventRecord(ctx->ev_k1, ctx->stream);
  k_gram_reduce<<<dim3(GT_TILE_ELEMS / 256, (unsigned)n_pairs), 256, 0, ctx->stream>>>(
      (const double*)part->p, (const int2*)d_pairs->p, n_pairs, n_splits, H, gram, colsum_from_ones);
  CSC_LAUNCHED(ctx);
  // the pageable pair table must outlive the async copy
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1) == cudaSuccess) ctx->stage_ms[CSC_T_GRAM_MMA] += ms;
  return CSC_OK;
}

// colsum[H] += column sums of z; gram[H*H] += z'z (full symmetric), both fp64.
csc_status gram_tc(csc_ctx* ctx, const float* z, int64_t n_rows, int H, double* colsum, double* gram) {
  const int Hp = gram_split_pitch(H);
  DevPtr hi16, lo16;
  CSC_TRY(dev_alloc(ctx, (size_t)n_rows * Hp * 2, &hi16));
  CSC_TRY(dev_alloc(ctx, (size_t)n_rows * Hp * 2, &lo16));
  const int64_t rows_per_cta = 256;
  dim3 grid((unsigned)((Hp / 4 + 255) / 256), (unsigned)ceil_div64(n_rows, rows_per_cta));
  k_split_colsum<<<grid, 256, 0, ctx->stream>>>(z, n_rows, H, Hp, rows_per_cta, (__half*)hi16->p, (__half*)lo16->p, colsum);
  CSC_LAUNCHED(ctx);
  return gram_tc_split(ctx, (const __half*)hi16->p, (const __half*)lo16->p, n_rows, H, Hp, gram, nullptr);
}
