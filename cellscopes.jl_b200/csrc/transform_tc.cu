// PCA embedding as a tensor-core GEMM over the split-precision operands the Gram pass already wrote (run_pca's
// transform, processing.jl:178-179):   E[N x k] = Z[N x H] U[H x k] - 1 (m'U),   Z = hi + lo (fp16 halves, cell-major
// [N][pitch], written by k_scale_split for k_gram_tc), U = Uhi + Ulo (fp16, transposed to [64][pitch]).
// Three tcgen05 kind::f16 passes per 64-gene chunk (hi Uhi', hi Ulo', lo Uhi'; lo Ulo' is below 2^-22), fp32 accumulation
// in TMEM over the whole gene axis, K-major operands by TMA (SWIZZLE_128B boxes of 128 cells x 64 genes and 64
// components x 64 genes).  HBM-bound by design: 4 B per (cell, gene) = 8.2 GB at 1M cells x 2048 -> 1.3 ms at the copy
// rate, against the 10.9 ms of the sparse gather version it replaces (k_transform_sparse: 0.05 of HBM, bound by L2
// gathers of the loading rows; ncu r2).  The tensor work (3 x 2 N pitch 64 flop = 0.8 Tflop) needs 0.6 ms.
//   warp 0: TMA producer; warp 1: TMEM allocation + MMA issue; warps 2-5: epilogue (thread = cell: accumulator row out of
//   TMEM, minus m'U, 50 floats to the embedding).  2 stages x 48 KB per CTA, two CTAs per SM so that one CTA's epilogue
//   overlaps the other's loads.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int T_BM = 128;                 // cells per CTA (UMMA M)
constexpr int T_BN = 64;                  // components, padded (UMMA N)
constexpr int T_BK = 64;                  // genes per chunk: one 128-byte swizzle row
constexpr int T_STAGES = 2;
constexpr int T_SEGS = 4;                 // the gene axis is accumulated in 4 TMEM accumulators (64 columns each) that the
                                          // epilogue adds up round-to-nearest: the tensor core's adds into its fp32
                                          // accumulator truncate, and 375 of them in a row cost 6.8e-5 of a component's rms
                                          // at 2000 genes (measured); 94 per segment bring that to the 2e-5 level
constexpr int T_A_BYTES = T_BM * 128;     // 16 KB
constexpr int T_B_BYTES = T_BN * 128;     // 8 KB
constexpr int T_STAGE_BYTES = 2 * T_A_BYTES + 2 * T_B_BYTES;   // hi, lo of both operands: 48 KB
constexpr int T_THREADS = 6 * 32;

// U (fp32 [H][ldu], component j in column j) -> Uhi / Ulo fp16 [64][pitch], zero beyond H and beyond k
__global__ void k_split_loadings(const float* __restrict__ U, int H, int ldu, int k, int pitch, __half* __restrict__ uhi,
                                 __half* __restrict__ ulo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T_BN * pitch) return;
  const int j = i / pitch, h = i - j * pitch;
  const float v = (j < k && h < H) ? U[(int64_t)h * ldu + j] : 0.f;
  const __half a = __float2half_rn(v);
  uhi[i] = a;
  ulo[i] = __float2half_rn(v - __half2float(a));
}

__global__ void __launch_bounds__(T_THREADS, 2)
k_transform_tc(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
               const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, int64_t n_rows,
               int n_chunks, int k, const float* __restrict__ proj_mean, float* __restrict__ out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* sbase = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ uint64_t bars[2 * T_STAGES + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + T_STAGES);
  const uint32_t bar_done = smem_u32(bars + 2 * T_STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * T_BM;

  if (threadIdx.x == 0) {
    for (int i = 0; i < T_STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)(T_SEGS * T_BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c % T_STAGES;
      mbar_wait(bar_empty + 8 * s, (uint32_t)(((c / T_STAGES) & 1) ^ 1));
      if (elect_one()) {
        const uint32_t dst = smem_u32(sbase + (size_t)s * T_STAGE_BYTES);
        mbar_expect_tx(bar_full + 8 * s, T_STAGE_BYTES);
        tma_load_2d(dst, &map_ahi, bar_full + 8 * s, c * T_BK, (int)row0);
        tma_load_2d(dst + T_A_BYTES, &map_alo, bar_full + 8 * s, c * T_BK, (int)row0);
        tma_load_2d(dst + 2 * T_A_BYTES, &map_bhi, bar_full + 8 * s, c * T_BK, 0);
        tma_load_2d(dst + 2 * T_A_BYTES + T_B_BYTES, &map_blo, bar_full + 8 * s, c * T_BK, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // D = F32, A = B = F16, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(T_BN >> 3) << 17) | ((uint32_t)(T_BM >> 4) << 24);
    const int per_seg = (n_chunks + T_SEGS - 1) / T_SEGS;
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c % T_STAGES;
      mbar_wait(bar_full + 8 * s, (uint32_t)((c / T_STAGES) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = smem_u32(sbase + (size_t)s * T_STAGE_BYTES);
        const uint32_t a_lo = a_hi + T_A_BYTES, b_hi = a_hi + 2 * T_A_BYTES, b_lo = b_hi + T_B_BYTES;
        const uint64_t dah = make_desc_sw128(a_hi), dal = make_desc_sw128(a_lo), dbh = make_desc_sw128(b_hi), dbl = make_desc_sw128(b_lo);
#pragma unroll
        for (int ks = 0; ks < T_BK / 16; ++ks) {
          const uint64_t o = (uint64_t)(ks * 2);          // 16 halves = 32 bytes along the swizzled row, in 16-byte units
          const uint32_t d_tmem = tmem_base + (uint32_t)((c / per_seg) * T_BN);
          tc_mma_f16_ss(d_tmem, dah + o, dbh + o, idesc, (c % per_seg != 0 || ks > 0) ? 1u : 0u);
          tc_mma_f16_ss(d_tmem, dah + o, dbl + o, idesc, 1u);
          tc_mma_f16_ss(d_tmem, dal + o, dbh + o, idesc, 1u);
        }
        tc_commit(bar_empty + 8 * s);                     // the stage is reusable once these MMAs have read it
        if (c == n_chunks - 1) tc_commit(bar_done);
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue: thread = cell =================
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int64_t row = row0 + q * 32 + lane;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_segs = (n_chunks + ((n_chunks + T_SEGS - 1) / T_SEGS) - 1) / ((n_chunks + T_SEGS - 1) / T_SEGS);   // segments in use
    float acc[T_BN];
#pragma unroll
    for (int j = 0; j < T_BN; ++j) acc[j] = 0.f;
    for (int sg = 0; sg < n_segs; ++sg) {
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr + (uint32_t)(sg * T_BN), r0);
      tmem_ld32(taddr + (uint32_t)(sg * T_BN) + 32, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        acc[j] += __uint_as_float(r0[j]);
        acc[32 + j] += __uint_as_float(r1[j]);
      }
    }
    if (row < n_rows) {
      float* o = out + row * k;
#pragma unroll
      for (int j = 0; j < T_BN; ++j)
        if (j < k) o[j] = acc[j] - __ldg(proj_mean + j);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(T_SEGS * T_BN)) : "memory");
  }
}

csc_status make_map_2d(csc_ctx* ctx, EncodeTiledFn fn, CUtensorMap* map, const __half* ptr, int64_t n_rows, int pitch, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)T_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return csc_fail(ctx, CSC_ERR_CUDA, "cuTensorMapEncodeTiled (transform operands) failed with %d", (int)r);
  return CSC_OK;
}

}  // namespace

// out: fp32 [n_rows][k]; hi16 / lo16: [n_rows][pitch] (pitch a multiple of 64, column H = 1, zero beyond)
csc_status transform_tc_split(csc_ctx* ctx, const __half* hi16, const __half* lo16, int64_t n_rows, int H, int pitch, const csc_pca* p,
                              float* out) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CSC_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
      return csc_fail(ctx, CSC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (EncodeTiledFn)fn;
  }
  CSC_REQUIRE(ctx, p->k <= T_BN && pitch % T_BK == 0 && H <= pitch, "transform_tc_split: k=%d pitch=%d H=%d", p->k, pitch, H);
  if (n_rows == 0) return CSC_OK;
  DevPtr uhi, ulo;
  CSC_TRY(dev_alloc(ctx, (size_t)T_BN * pitch * 2, &uhi));
  CSC_TRY(dev_alloc(ctx, (size_t)T_BN * pitch * 2, &ulo));
  k_split_loadings<<<(unsigned)ceil_div64((int64_t)T_BN * pitch, 256), 256, 0, ctx->stream>>>(
      (const float*)p->d_loadings_f32->p, H, p->kpad, p->k, pitch, (__half*)uhi->p, (__half*)ulo->p);
  CSC_LAUNCHED(ctx);
  CUtensorMap mah, mal, mbh, mbl;
  CSC_TRY(make_map_2d(ctx, encode, &mah, hi16, n_rows, pitch, T_BM));
  CSC_TRY(make_map_2d(ctx, encode, &mal, lo16, n_rows, pitch, T_BM));
  CSC_TRY(make_map_2d(ctx, encode, &mbh, (const __half*)uhi->p, T_BN, pitch, T_BN));
  CSC_TRY(make_map_2d(ctx, encode, &mbl, (const __half*)ulo->p, T_BN, pitch, T_BN));
  const size_t smem = 1024 + (size_t)T_STAGES * T_STAGE_BYTES;
  CSC_CUDA(ctx, cudaFuncSetAttribute(k_transform_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_transform_tc<<<(unsigned)ceil_div64(n_rows, T_BM), T_THREADS, smem, ctx->stream>>>(mah, mal, mbh, mbl, n_rows, pitch / T_BK, p->k,
                                                                                       (const float*)p->d_proj_mean->p, out);
  CSC_LAUNCHED(ctx);
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CSC_OK;
}

extern "C" csc_status csc_pca_transform_split(csc_ctx* ctx, const csc_split* ops, const csc_pca* p, csc_dense** out) {
  if (!ctx || !ops || !p || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, ops->n_feat == p->n_feat, "csc_pca_transform_split: operands hold %lld features, the model %lld",
              (long long)ops->n_feat, (long long)p->n_feat);
  CSC_REQUIRE(ctx, p->k <= T_BN, "csc_pca_transform_split: more than %d components (use csc_pca_transform_sparse)", T_BN);
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_PCA_TRANSFORM);
  csc_dense* e = nullptr;
  CSC_TRY(csc_dense_alloc(ctx, ops->n_rows, p->k, &e));
  const csc_status rc = transform_tc_split(ctx, (const __half*)ops->hi->p, (const __half*)ops->lo->p, ops->n_rows, (int)ops->n_feat,
                                           ops->pitch, p, e->p());
  if (rc != CSC_OK) {
    csc_dense_free(e);
    return rc;
  }
  *out = e;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_split_free(csc_split* s) {
  delete s;
  return CSC_OK;
}
