// Internal definitions shared by the translation units of libcellscopes_cuda.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <map>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "cellscopes_cuda.h"

// ------------------------------------------------------------------------------------------
// context and handles
// ------------------------------------------------------------------------------------------
// Caching block allocator of a context.  Every stage allocates work buffers of the same sizes on every call
// (a step of the path allocates ~60 buffers, some of them gigabytes); handing those to the driver each time,
// even through the stream-ordered pool, cost tens of milliseconds of host time per step (measured).  Freed
// blocks are kept in a size-ordered free list instead and reused when a request fits within 25 %; they go back
// to the driver at csc_destroy or when an allocation fails.  Reuse is safe without events because a context
// issues all of its work on one stream: a block is freed by the host after the last kernel that uses it was
// enqueued, and whoever gets it next enqueues later on the same stream.
struct BlockCache {
  std::multimap<size_t, void*> free_blocks;
  size_t cached_bytes = 0;
  void* take(size_t bytes, size_t* block_bytes) {
    auto it = free_blocks.lower_bound(bytes);
    if (it == free_blocks.end() || it->first > bytes + bytes / 4 + (1u << 20)) return nullptr;
    void* p = it->second;
    *block_bytes = it->first;
    cached_bytes -= it->first;
    free_blocks.erase(it);
    return p;
  }
  void give(size_t bytes, void* p) {
    free_blocks.emplace(bytes, p);
    cached_bytes += bytes;
  }
  void release_all() {
    for (auto& kv : free_blocks) cudaFree(kv.second);
    free_blocks.clear();
    cached_bytes = 0;
  }
};

struct DevMem;
struct csc_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  cudaEvent_t ev_timer0 = nullptr, ev_timer1 = nullptr;
  cudaEvent_t ev_stage0 = nullptr, ev_stage1 = nullptr;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // around one kernel of interest inside a stage
  double stage_ms[CSC_T_COUNT] = {0};
  int64_t launches = 0;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  std::shared_ptr<bool> alive;  // cleared by csc_destroy: late frees then fall back to cudaFree
  BlockCache cache;
  // page-locked host blocks handed to callers (csc_host_alloc): live blocks by pointer, free blocks by size
  std::map<void*, size_t> host_live;
  std::multimap<size_t, void*> host_free;
  // pinned bounce buffers for device->host copies into pageable memory (see d2h_copy)
  void* bounce[2] = {nullptr, nullptr};
  // page-locked slots of the host lane of the two-lane upload (csc_sparse_upload), grown on demand
  void* up_stage = nullptr;
  size_t up_stage_bytes = 0;
  cudaEvent_t ev_bounce[2] = {nullptr, nullptr};
  std::shared_ptr<DevMem> gemm_ws;   // split-K partials of the fp64 GEMM (grow-only)
  int64_t knn_fallback_rows = 0;   // rows of the last csc_knn call that needed the exact re-scan
  // multi-GPU (comm.cu): the NCCL communicator of this context's rank, collectives run on `stream`
  void* comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  int trace = 0;                // CSC_LOG_LEVEL >= 2: print host-side checkpoints of the stages
  double trace_t0 = 0.0;
};
void csc_trace(csc_ctx* ctx, const char* label);
void comm_release(csc_ctx* ctx);   // comm.cu
#define CSC_TRACE(ctx, label)                 \
  do {                                        \
    if ((ctx)->trace) csc_trace((ctx), label); \
  } while (0)

// device allocation with shared ownership (a normalised matrix shares colptr/rowval with the raw one)
struct DevMem {
  void* p = nullptr;
  size_t bytes = 0;       // requested size
  size_t block_bytes = 0; // size of the underlying block (the cache key)
  int device = 0;
  bool owned = true;
  csc_ctx* ctx = nullptr;
  std::shared_ptr<bool> ctx_alive;
  ~DevMem() {
    if (p && owned) {
      if (ctx_alive && *ctx_alive) {
        ctx->cache.give(block_bytes, p);
      } else {
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != device) cudaSetDevice(device);
        cudaFree(p);
        if (cur != device) cudaSetDevice(cur);
      }
    }
  }
};
typedef std::shared_ptr<DevMem> DevPtr;

struct csc_sparse {
  csc_ctx* ctx;
  int64_t n_genes, n_cells, nnz;
  DevPtr colptr;  // int64 [n_cells+1], 0-based
  DevPtr rowval;  // int32 [nnz], 0-based
  DevPtr val;     // float [nnz]
  // upper bound on |value| when one is known (set by csc_normalize, inherited by csc_sparse_select_rows), else +inf:
  // k_scale_stats_fx derives the scale of its fixed-point accumulators from it
  double vmax_bound = std::numeric_limits<double>::infinity();
  const int64_t* cp() const { return (const int64_t*)colptr->p; }
  const int32_t* rv() const { return (const int32_t*)rowval->p; }
  float* v() const { return (float*)val->p; }
};

struct csc_dense {
  csc_ctx* ctx;
  int64_t n_rows, n_cols;  // row-major, rows = cells
  DevPtr mem;
  float* p() const { return (float*)mem->p; }
};

struct csc_vec {
  csc_ctx* ctx;
  int64_t n;
  int32_t dtype;
  DevPtr mem;
  template <typename T>
  T* as() const { return (T*)mem->p; }
};

struct csc_pca {
  csc_ctx* ctx;
  int64_t n_feat;
  int32_t k;
  std::vector<double> principalvars, loadings, mean;  // loadings column-major H x k
  double tvar;
  DevPtr d_loadings_f32;  // [n_feat x kpad] row-major fp32 (kpad = k rounded up to 16)
  DevPtr d_proj_mean;     // fp32 [kpad]: m' U
  int32_t kpad;
};

// split-precision operands of the scaled matrix (written by csc_scale_gram_partial): hi = fp16(Z), lo = fp16(Z - hi),
// both [n_rows][pitch] with column n_feat = 1 and zeros up to the pitch
struct csc_split {
  csc_ctx* ctx;
  int64_t n_rows, n_feat;
  int32_t pitch;
  DevPtr hi, lo;
};

struct csc_graph {
  csc_ctx* ctx;
  int64_t n_rows, n_cols, nnz;
  int64_t col_lo;
  DevPtr colptr;  // int64 [n_cols+1]
  DevPtr rowval;  // int32 [nnz]
  DevPtr val;     // double [nnz]
};

// ------------------------------------------------------------------------------------------
// error plumbing: no exception or abort ever crosses the C ABI
// ------------------------------------------------------------------------------------------
csc_status csc_fail(csc_ctx* ctx, csc_status code, const char* fmt, ...);

#define CSC_CUDA(ctx, call)                                                                      \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return csc_fail((ctx), e__ == cudaErrorMemoryAllocation ? CSC_ERR_OOM : CSC_ERR_CUDA,      \
                      "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));     \
  } while (0)

#define CSC_TRY(expr)                 \
  do {                                \
    csc_status s__ = (expr);          \
    if (s__ != CSC_OK) return s__;    \
  } while (0)

#define CSC_REQUIRE(ctx, cond, ...) \
  do {                              \
    if (!(cond)) return csc_fail((ctx), CSC_ERR_ARG, __VA_ARGS__); \
  } while (0)

// after a kernel launch: count it and surface launch errors
#define CSC_LAUNCHED(ctx)                     \
  do {                                        \
    (ctx)->launches++;                        \
    CSC_CUDA((ctx), cudaGetLastError());      \
  } while (0)

csc_status dev_alloc(csc_ctx* ctx, size_t bytes, DevPtr* out, bool zero = false);
// device -> caller-owned (pageable) host memory through double-buffered pinned staging: the DMA of chunk i+1
// overlaps the host memcpy of chunk i (about twice the rate of a plain cudaMemcpy into pageable memory)
csc_status d2h_copy(csc_ctx* ctx, void* host, const void* dev, size_t bytes);

// Wrap a C-ABI body so that C++ exceptions (std::bad_alloc from the host containers) become codes.
#define CSC_GUARD_BEGIN try {
#define CSC_GUARD_END(ctx)                                                    \
  }                                                                           \
  catch (const std::bad_alloc&) {                                             \
    return csc_fail((ctx), CSC_ERR_OOM, "host allocation failed");            \
  }                                                                           \
  catch (const std::exception& ex) {                                          \
    return csc_fail((ctx), CSC_ERR_INTERNAL, "exception: %s", ex.what());     \
  }                                                                           \
  catch (...) {                                                               \
    return csc_fail((ctx), CSC_ERR_INTERNAL, "unknown exception");            \
  }

// stage timing: CUDA events on the launching stream, accumulated per stage slot
struct StageTimer {
  csc_ctx* ctx;
  int slot;
  StageTimer(csc_ctx* c, int s) : ctx(c), slot(s) { cudaEventRecord(ctx->ev_stage0, ctx->stream); }
  ~StageTimer() {
    cudaEventRecord(ctx->ev_stage1, ctx->stream);
    if (cudaEventSynchronize(ctx->ev_stage1) == cudaSuccess) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ctx->ev_stage0, ctx->ev_stage1) == cudaSuccess) ctx->stage_ms[slot] += ms;
    }
  }
};

__host__ __device__ static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// streaming (read-once) loads: keep them out of L1
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }
#endif

// stage launchers implemented in the other translation units ------------------------------------
// hvg_dev.cu: the same finish on the device (default for >= 1024 genes; CSC_HVG_HOST=1 keeps the host version)
bool hvg_device_applies(int64_t n_genes, double span);
csc_status launch_hvg_device(csc_ctx* ctx, const double* d_s1s2, int64_t n_genes, int64_t n_cells_total, double span, double* mean,
                             double* variance, double* var_exp, double* var_std, int64_t* order);
csc_status launch_hvg_host(csc_ctx* ctx, const double* s1s2, int64_t n_genes, int64_t n_cells_total,
                           int64_t n_features, double span, double* mean, double* variance, double* var_exp,
                           double* var_std, int64_t* order, bool collective = false);
