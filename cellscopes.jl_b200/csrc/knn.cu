// Exact brute-force kNN in PCA space (the graph construction inside run_clustering,
// src/scrna/processing.jl:195-201 and 263-266; the reference uses approximate NN-descent).
//
// Both metrics are reduced to "largest inner product":
//   cosine     q' = x/|x|,      key' = x/|x|          score = cos,            dist = max(1 - score, 0)
//   euclidean  q' = [2x, -1],   key' = [x, |x|^2]     score = 2 q.x - |x|^2,  dist = sqrt(max(|q|^2 - score, 0))
// so one kernel serves both: no N x N matrix ever reaches HBM, every query keeps a running
// sorted top-k and a register threshold; a candidate costs one compare unless it enters the list.
#include <cfloat>
#include <climits>
#include <cmath>
#include <cuda_pipeline.h>

#include "common.cuh"
#include "knn_common.cuh"

// ------------------------------------------------------------------------------------------
// prepare padded operand rows
// ------------------------------------------------------------------------------------------
// one warp per row: out[row][0..DP) = transformed columns [lo,hi) of x, zero padded; aux[row] = |x|^2
__global__ void __launch_bounds__(256)
k_knn_prep(const float* __restrict__ x, int64_t n_rows, int32_t ld, int32_t lo, int32_t hi, int32_t DP, int metric,
           int is_query, float* __restrict__ out, float* __restrict__ sqnorm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const float* p = x + row * (int64_t)ld;
  const int d = hi - lo;
  double s = 0.0;
  for (int c = lane; c < d; c += 32) {
    const double v = (double)p[lo + c];
    s += v * v;
  }
  s = warp_sum(s);
  float* o = out + row * (int64_t)DP;
  if (metric == CSC_METRIC_COSINE) {
    const double inv = s > 0.0 ? 1.0 / sqrt(s) : 0.0;  // zero vector -> score 0 -> distance 1
    for (int c = lane; c < DP; c += 32) o[c] = c < d ? (float)((double)p[lo + c] * inv) : 0.f;
  } else {
    for (int c = lane; c < DP; c += 32) {
      float v = 0.f;
      if (c < d) v = is_query ? 2.f * p[lo + c] : p[lo + c];
      else if (c == d) v = is_query ? -1.f : (float)s;
      o[c] = v;
    }
  }
  if (lane == 0 && sqnorm) sqnorm[row] = (float)s;
}

// ------------------------------------------------------------------------------------------
// v1 scan kernel (CUDA cores): one thread owns one query row.
//   q row in registers (DP floats), key tiles of 64 rows in shared memory (double-buffered cp.async),
//   inner product by broadcast LDS.128 + FFMA, running top-k per thread in shared memory.
// ------------------------------------------------------------------------------------------
#define KNN_TQ 128
#define KNN_TK 64

template <int DP>
__global__ void __launch_bounds__(KNN_TQ)
k_knn_scan(const float* __restrict__ q, int64_t n_q, const float* __restrict__ keys, int64_t n_keys, int64_t query_offset,
           int32_t k, float* __restrict__ out_score, int32_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* ktile = reinterpret_cast<float*>(smem_raw);                       // [2][KNN_TK][DP]
  float* ls = ktile + 2 * KNN_TK * DP;                                     // [k][KNN_TQ]
  int* li = reinterpret_cast<int*>(ls + (size_t)k * KNN_TQ);               // [k][KNN_TQ]
  const int t = threadIdx.x;
  const int64_t qrow = (int64_t)blockIdx.x * KNN_TQ + t;
  const bool qvalid = qrow < n_q;
  const int64_t self = query_offset + qrow;
  float qr[DP];
  {
    const float4* src = reinterpret_cast<const float4*>(q + (qvalid ? qrow : 0) * DP);
#pragma unroll
    for (int i = 0; i < DP / 4; ++i) {
      const float4 v = __ldg(src + i);
      qr[4 * i] = v.x;
      qr[4 * i + 1] = v.y;
      qr[4 * i + 2] = v.z;
      qr[4 * i + 3] = v.w;
    }
  }
  for (int j = 0; j < k; ++j) {
    ls[j * KNN_TQ + t] = -INFINITY;
    li[j * KNN_TQ + t] = INT_MAX;
  }
  float tau = -INFINITY;
  const int64_t n_tiles = ceil_div64(n_keys, KNN_TK);
  constexpr int V4_PER_TILE = KNN_TK * DP / 4;
  auto issue_tile = [&](int64_t tile, int buf) {
    const int64_t base = tile * KNN_TK;
    float4* dst = reinterpret_cast<float4*>(ktile + (size_t)buf * KNN_TK * DP);
    const float4* src = reinterpret_cast<const float4*>(keys + base * DP);
    const int64_t rem_v4 = (n_keys - base) * (DP / 4);
    const int64_t valid_v4 = rem_v4 < V4_PER_TILE ? rem_v4 : (int64_t)V4_PER_TILE;
    for (int i = t; i < V4_PER_TILE; i += KNN_TQ) {
      if (i < valid_v4) __pipeline_memcpy_async(dst + i, src + i, 16);
      else dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __pipeline_commit();
  };
  issue_tile(0, 0);
  for (int64_t tile = 0; tile < n_tiles; ++tile) {
    const int buf = (int)(tile & 1);
    if (tile + 1 < n_tiles) {
      issue_tile(tile + 1, buf ^ 1);
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
    const float* kt = ktile + (size_t)buf * KNN_TK * DP;
    const int64_t base = tile * KNN_TK;
    const int jmax = (n_keys - base) < KNN_TK ? (int)(n_keys - base) : KNN_TK;
    for (int j = 0; j < jmax; ++j) {
      const float4* kr = reinterpret_cast<const float4*>(kt + j * DP);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int i = 0; i < DP / 4; ++i) {
        const float4 kv = kr[i];  // same address for the whole warp: broadcast
        a0 = fmaf(qr[4 * i], kv.x, a0);
        a1 = fmaf(qr[4 * i + 1], kv.y, a1);
        a2 = fmaf(qr[4 * i + 2], kv.z, a2);
        a3 = fmaf(qr[4 * i + 3], kv.w, a3);
      }
      const float s = (a0 + a1) + (a2 + a3);
      const int64_t id = base + j;
      // keys arrive in ascending id, so an equal score never displaces an earlier (lower) id
      if (s > tau && id != self && qvalid) {
        topk_insert(ls + t, li + t, k, KNN_TQ, s, (int)id);
        tau = ls[(k - 1) * KNN_TQ + t];
      }
    }
    __syncthreads();
  }
  if (qvalid) {
    for (int j = 0; j < k; ++j) {
      out_score[qrow * k + j] = ls[j * KNN_TQ + t];
      out_idx[qrow * k + j] = li[j * KNN_TQ + t];
    }
  }
}

// score -> distance, in place
__global__ void k_knn_finish(float* __restrict__ score, const float* __restrict__ qsq, int64_t n_q, int32_t k, int metric) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_q * k) return;
  const float s = score[i];
  float d;
  if (metric == CSC_METRIC_COSINE) d = fmaxf(1.f - s, 0.f);  // Distances.jl CosineDist: max(1 - cos, 0)
  else d = sqrtf(fmaxf(qsq[i / k] - s, 0.f));
  score[i] = d;
}

csc_status knn_scan_tc(csc_ctx* ctx, const float* q_prep, int64_t n_q, const float* k_prep, int64_t n_keys, bool same,
                       int64_t query_offset, const int32_t* self_ids, int32_t k, int32_t d_eff, float* out_score,
                       int32_t* out_idx);
csc_status knn_scan_f16(csc_ctx* ctx, const float* q_prep, int64_t n_q, const float* k_prep, int64_t n_keys, bool same,
                        int64_t query_offset, int32_t k, int32_t d_eff, float* out_score, int32_t* out_idx,
                        int64_t* n_fallback_out);

static int pad_dim(int d) {
  if (d <= 16) return 16;
  if (d <= 32) return 32;
  if (d <= 64) return 64;
  return -1;
}

csc_status knn_prepare(csc_ctx* ctx, const csc_dense* x, int32_t dim_lo, int32_t dim_hi, int32_t metric, int is_query,
                       int DP, DevPtr* out, DevPtr* sqnorm) {
  CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(x->n_rows, 1) * DP * 4, out));
  if (sqnorm) CSC_TRY(dev_alloc(ctx, (size_t)std::max<int64_t>(x->n_rows, 1) * 4, sqnorm));
  if (x->n_rows == 0) return CSC_OK;
  k_knn_prep<<<(unsigned)ceil_div64(x->n_rows * 32, 256), 256, 0, ctx->stream>>>(
      x->p(), x->n_rows, (int32_t)x->n_cols, dim_lo, dim_hi, DP, metric, is_query, (float*)(*out)->p,
      sqnorm ? (float*)(*sqnorm)->p : nullptr);
  CSC_LAUNCHED(ctx);
  return CSC_OK;
}

template <int DP>
static csc_status launch_scan(csc_ctx* ctx, const float* q, int64_t n_q, const float* keys, int64_t n_keys, int64_t qoff,
                              int32_t k, float* score, int32_t* idx) {
  size_t smem = (size_t)2 * KNN_TK * DP * 4 + (size_t)k * KNN_TQ * 8;
  CSC_REQUIRE(ctx, smem <= (size_t)220 * 1024, "csc_knn: k=%d needs %zu bytes of shared memory", k, smem);
  if (smem > 48 * 1024)
    CSC_CUDA(ctx, cudaFuncSetAttribute(k_knn_scan<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_knn_scan<DP><<<(unsigned)ceil_div64(n_q, KNN_TQ), KNN_TQ, smem, ctx->stream>>>(q, n_q, keys, n_keys, qoff, k, score, idx);
  CSC_LAUNCHED(ctx);
  return CSC_OK;
}

extern "C" csc_status csc_knn(csc_ctx* ctx, const csc_dense* queries, const csc_dense* keys, int64_t query_offset,
                              int32_t dim_lo, int32_t dim_hi, int32_t k, int32_t metric, csc_vec** idx, csc_vec** dist) {
  if (!ctx || !queries || !keys || !idx || !dist) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, metric == CSC_METRIC_COSINE || metric == CSC_METRIC_EUCLIDEAN, "csc_knn: unknown metric %d", metric);
  CSC_REQUIRE(ctx, queries->n_cols == keys->n_cols, "csc_knn: queries have %lld columns, keys %lld",
              (long long)queries->n_cols, (long long)keys->n_cols);
  CSC_REQUIRE(ctx, 0 <= dim_lo && dim_lo < dim_hi && dim_hi <= keys->n_cols, "csc_knn: bad dims [%d,%d) of %lld", dim_lo,
              dim_hi, (long long)keys->n_cols);
  CSC_REQUIRE(ctx, k >= 1 && (int64_t)k <= keys->n_rows - 1, "csc_knn: k=%d must be in [1, n_keys-1=%lld]", k,
              (long long)keys->n_rows - 1);
  CSC_REQUIRE(ctx, keys->n_rows < ((int64_t)1 << 31), "csc_knn: more than 2^31 keys");
  CSC_REQUIRE(ctx, query_offset >= 0 && query_offset + queries->n_rows <= keys->n_rows,
              "csc_knn: queries [%lld,+%lld) are not a row range of the keys", (long long)query_offset,
              (long long)queries->n_rows);
  const int d = dim_hi - dim_lo + (metric == CSC_METRIC_EUCLIDEAN ? 1 : 0);
  // v3 (fp16 filter on tcgen05 + proof + exact refine) serves cosine with k <= 32; v2 (3xTF32 on tcgen05) k <= 32
  // and up to 64 dimensions; the CUDA-core kernel the rest.  CSC_KNN_IMPL = simt | tf32 forces an older path.
  const char* impl = getenv("CSC_KNN_IMPL");
  const bool use_tc = d <= 64 && k <= 32 && !(impl && strcmp(impl, "simt") == 0);
  const bool use_f16 = use_tc && metric == CSC_METRIC_COSINE && !(impl && strcmp(impl, "tf32") == 0);
  const int DP = use_tc ? 64 : pad_dim(d);
  CSC_REQUIRE(ctx, DP > 0, "csc_knn: %d dimensions not supported (max 64, 63 for euclidean)", dim_hi - dim_lo);
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_KNN);
  DevPtr kp, qp, qsq;
  const bool same = (queries == keys) && metric == CSC_METRIC_COSINE;
  CSC_TRY(knn_prepare(ctx, keys, dim_lo, dim_hi, metric, 0, DP, &kp, nullptr));
  if (same) {
    qp = kp;
  } else {
    CSC_TRY(knn_prepare(ctx, queries, dim_lo, dim_hi, metric, 1, DP, &qp, &qsq));
  }
  csc_vec *vi = nullptr, *vd = nullptr;
  CSC_TRY(csc_vec_alloc(ctx, queries->n_rows * k, CSC_DT_I32, &vi));
  csc_status s = csc_vec_alloc(ctx, queries->n_rows * k, CSC_DT_F32, &vd);
  if (s != CSC_OK) {
    csc_vec_free(vi);
    return s;
  }
  if (queries->n_rows > 0) {
    const float* qd = (const float*)qp->p;
    const float* kd = (const float*)kp->p;
    if (use_f16) {
      int64_t n_fb = 0;
      s = knn_scan_f16(ctx, qd, queries->n_rows, kd, keys->n_rows, same, query_offset, k, d, vd->as<float>(),
                       vi->as<int32_t>(), &n_fb);
      ctx->knn_fallback_rows = n_fb;
    } else if (use_tc) {
      s = knn_scan_tc(ctx, qd, queries->n_rows, kd, keys->n_rows, same, query_offset, nullptr, k, d, vd->as<float>(),
                      vi->as<int32_t>());
    } else {
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    switch (DP) {
      case 16: s = launch_scan<16>(ctx, qd, queries->n_rows, kd, keys->n_rows, query_offset, k, vd->as<float>(), vi->as<int32_t>()); break;
      case 32: s = launch_scan<32>(ctx, qd, queries->n_rows, kd, keys->n_rows, query_offset, k, vd->as<float>(), vi->as<int32_t>()); break;
      default: s = launch_scan<64>(ctx, qd, queries->n_rows, kd, keys->n_rows, query_offset, k, vd->as<float>(), vi->as<int32_t>()); break;
    }
    cudaEventRecord(ctx->ev_k1, ctx->stream);
    }
    if (s == CSC_OK) {
      k_knn_finish<<<(unsigned)ceil_div64(queries->n_rows * k, 256), 256, 0, ctx->stream>>>(
          vd->as<float>(), qsq ? (const float*)qsq->p : nullptr, queries->n_rows, k, metric);
      ctx->launches++;
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) s = csc_fail(ctx, CSC_ERR_CUDA, "csc_knn: %s", cudaGetErrorString(e));
      float ms = 0.f;
      if (!use_tc && e == cudaSuccess && cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1) == cudaSuccess)
        ctx->stage_ms[CSC_T_KNN_SCAN] += ms;
    }
    if (s != CSC_OK) {
      csc_vec_free(vi);
      csc_vec_free(vd);
      return s;
    }
  }
  *idx = vi;
  *dist = vd;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_knn_last_fallback_rows(csc_ctx* ctx, int64_t* n_rows) {
  if (!ctx || !n_rows) return CSC_ERR_ARG;
  *n_rows = ctx->knn_fallback_rows;
  return CSC_OK;
}
