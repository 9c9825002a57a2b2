// Lifecycle, buffers, upload/download: the plumbing half of the C ABI (include/cellscopes_cuda.h).
#include "common.cuh"

#include <chrono>

static thread_local std::string g_create_err;

csc_status csc_fail(csc_ctx* ctx, csc_status code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  try {
    if (ctx) ctx->err = buf;
    else g_create_err = buf;
  } catch (...) {
  }
  return code;
}

void csc_trace(csc_ctx* ctx, const char* label) {
  cudaStreamSynchronize(ctx->stream);
  const double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  fprintf(stderr, "[csc trace] %-28s +%9.3f ms\n", label, ctx->trace_t0 > 0 ? now - ctx->trace_t0 : 0.0);
  ctx->trace_t0 = now;
}

csc_status dev_alloc(csc_ctx* ctx, size_t bytes, DevPtr* out, bool zero) {
  DevPtr m;
  try {
    m = std::make_shared<DevMem>();
  } catch (...) {
    return csc_fail(ctx, CSC_ERR_OOM, "host allocation failed");
  }
  m->device = ctx->device;
  m->bytes = bytes;
  m->ctx = ctx;
  m->ctx_alive = ctx->alive;
  // block sizes: 512-byte granules below 1 MiB, 2 MiB granules above (so that near-equal requests share blocks)
  size_t req = bytes ? bytes : 16;
  req = req < ((size_t)1 << 20) ? (req + 511) & ~(size_t)511 : (req + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  const auto t0 = std::chrono::steady_clock::now();
  m->p = ctx->cache.take(req, &m->block_bytes);   // a cached block may be larger than the request
  cudaError_t e = cudaSuccess;
  if (!m->p) {
    e = cudaMalloc(&m->p, req);
    if (e == cudaErrorMemoryAllocation) {
      // give the cached blocks back to the driver and retry once
      cudaGetLastError();
      cudaStreamSynchronize(ctx->stream);
      ctx->cache.release_all();
      e = cudaMalloc(&m->p, req);
    }
    m->block_bytes = req;
  }
  ctx->stage_ms[CSC_T_ALLOC_HOST] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (e != cudaSuccess) {
    m->p = nullptr;
    cudaGetLastError();
    return csc_fail(ctx, CSC_ERR_OOM, "cudaMalloc(%zu bytes) failed: %s", req, cudaGetErrorString(e));
  }
  if (zero) CSC_CUDA(ctx, cudaMemsetAsync(m->p, 0, req, ctx->stream));
  *out = m;
  return CSC_OK;
}

static constexpr size_t BOUNCE_BYTES = (size_t)16 << 20;

// is [p, p+bytes) inside a block this context handed out through csc_host_alloc?
static bool is_ctx_pinned(csc_ctx* ctx, const void* p, size_t bytes) {
  auto it = ctx->host_live.upper_bound(const_cast<void*>(p));
  if (it == ctx->host_live.begin()) return false;
  --it;
  const char* b = (const char*)it->first;
  return (const char*)p >= b && (const char*)p + bytes <= b + it->second;
}

csc_status d2h_copy(csc_ctx* ctx, void* host, const void* dev, size_t bytes) {
  if (bytes == 0) return CSC_OK;
  if (bytes <= ((size_t)1 << 20) || is_ctx_pinned(ctx, host, bytes)) {
    CSC_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CSC_OK;
  }
  for (int b = 0; b < 2; ++b) {
    if (!ctx->bounce[b]) {
      CSC_CUDA(ctx, cudaHostAlloc(&ctx->bounce[b], BOUNCE_BYTES, cudaHostAllocDefault));
      CSC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_bounce[b], cudaEventDisableTiming));
    }
  }
  const size_t n_chunks = (bytes + BOUNCE_BYTES - 1) / BOUNCE_BYTES;
  for (size_t i = 0; i <= n_chunks; ++i) {
    const int b = (int)(i & 1);
    if (i < n_chunks) {
      const size_t off = i * BOUNCE_BYTES, len = std::min(BOUNCE_BYTES, bytes - off);
      CSC_CUDA(ctx, cudaMemcpyAsync(ctx->bounce[b], (const char*)dev + off, len, cudaMemcpyDeviceToHost, ctx->stream));
      CSC_CUDA(ctx, cudaEventRecord(ctx->ev_bounce[b], ctx->stream));
    }
    if (i >= 1) {
      const size_t off = (i - 1) * BOUNCE_BYTES, len = std::min(BOUNCE_BYTES, bytes - off);
      CSC_CUDA(ctx, cudaEventSynchronize(ctx->ev_bounce[b ^ 1]));
      memcpy((char*)host + off, ctx->bounce[b ^ 1], len);
    }
  }
  return CSC_OK;
}

extern "C" {

int32_t csc_abi_version(void) { return CSC_ABI_VERSION; }

csc_status csc_host_alloc(csc_ctx* ctx, int64_t bytes, void** out) {
  if (!ctx || !out || bytes < 0) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t req = std::max<size_t>((size_t)bytes, 64);
  req = (req + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);      // 1 MiB granules
  void* p = nullptr;
  size_t got = req;
  auto it = ctx->host_free.lower_bound(req);
  if (it != ctx->host_free.end() && it->first <= req + req / 4 + ((size_t)1 << 20)) {
    p = it->second;
    got = it->first;
    ctx->host_free.erase(it);
  } else {
    cudaError_t e = cudaHostAlloc(&p, req, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (auto& kv : ctx->host_free) cudaFreeHost(kv.second);
      ctx->host_free.clear();
      e = cudaHostAlloc(&p, req, cudaHostAllocDefault);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      return csc_fail(ctx, CSC_ERR_OOM, "cudaHostAlloc(%zu bytes) failed: %s", req, cudaGetErrorString(e));
    }
  }
  ctx->host_live[p] = got;
  *out = p;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

csc_status csc_host_free(csc_ctx* ctx, void* p) {
  if (!p) return CSC_OK;
  if (!ctx) {
    // the context that handed the block out is gone (csc_destroy leaves blocks that are still live alone: they back
    // result arrays the caller may still read): the block goes straight back to the driver
    return cudaFreeHost(p) == cudaSuccess ? CSC_OK : csc_fail(nullptr, CSC_ERR_CUDA, "csc_host_free: cudaFreeHost failed");
  }
  CSC_GUARD_BEGIN
  auto it = ctx->host_live.find(p);
  CSC_REQUIRE(ctx, it != ctx->host_live.end(), "csc_host_free: pointer was not allocated by csc_host_alloc of this context");
  ctx->host_free.emplace(it->second, p);
  ctx->host_live.erase(it);
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

csc_status csc_create(int32_t device, csc_ctx** out) {
  if (!out) return csc_fail(nullptr, CSC_ERR_ARG, "csc_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return csc_fail(nullptr, CSC_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path",
                    cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return csc_fail(nullptr, CSC_ERR_ARG, "device %d out of range [0,%d)", device, n);
  csc_ctx* ctx = new (std::nothrow) csc_ctx();
  if (!ctx) return csc_fail(nullptr, CSC_ERR_OOM, "host allocation failed");
  ctx->device = device;
  cudaDeviceProp prop;
#define CREATE_CUDA(call)                                                                  \
  do {                                                                                     \
    cudaError_t e2 = (call);                                                               \
    if (e2 != cudaSuccess) {                                                               \
      csc_fail(nullptr, CSC_ERR_CUDA, "csc_create: %s -> %s", #call, cudaGetErrorString(e2)); \
      delete ctx;                                                                          \
      return CSC_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)
  CREATE_CUDA(cudaSetDevice(device));
  CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  if (prop.major != 10) {
    csc_fail(nullptr, CSC_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
             device, prop.major, prop.minor);
    delete ctx;
    return CSC_ERR_UNSUPPORTED;
  }
  CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  {
    // keep freed blocks in the pool instead of returning them to the driver at every synchronisation
    cudaMemPool_t pool;
    CREATE_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    CREATE_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  ctx->alive = std::make_shared<bool>(true);
  if (const char* lv = getenv("CSC_LOG_LEVEL")) ctx->trace = atoi(lv) >= 2;
  CREATE_CUDA(cudaEventCreate(&ctx->ev_timer0));
  CREATE_CUDA(cudaEventCreate(&ctx->ev_timer1));
  CREATE_CUDA(cudaEventCreate(&ctx->ev_stage0));
  CREATE_CUDA(cudaEventCreate(&ctx->ev_stage1));
  CREATE_CUDA(cudaEventCreate(&ctx->ev_k0));
  CREATE_CUDA(cudaEventCreate(&ctx->ev_k1));
#undef CREATE_CUDA
  *out = ctx;
  return CSC_OK;
}

csc_status csc_destroy(csc_ctx* ctx) {
  if (!ctx) return CSC_OK;
  cudaSetDevice(ctx->device);
  if (ctx->alive) *ctx->alive = false;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
  ctx->gemm_ws.reset();
  ctx->cache.release_all();
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  for (auto& kv : ctx->host_free) cudaFreeHost(kv.second);
  // blocks still live belong to result arrays of the caller: they stay valid and are released by
  // csc_host_free(NULL, p)
  if (ctx->up_stage) cudaFreeHost(ctx->up_stage);
  for (int b = 0; b < 2; ++b) {
    if (ctx->bounce[b]) cudaFreeHost(ctx->bounce[b]);
    if (ctx->ev_bounce[b]) cudaEventDestroy(ctx->ev_bounce[b]);
  }
  if (ctx->ev_timer0) cudaEventDestroy(ctx->ev_timer0);
  if (ctx->ev_timer1) cudaEventDestroy(ctx->ev_timer1);
  if (ctx->ev_stage0) cudaEventDestroy(ctx->ev_stage0);
  if (ctx->ev_stage1) cudaEventDestroy(ctx->ev_stage1);
  if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
  if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return CSC_OK;
}

const char* csc_last_error(const csc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

csc_status csc_synchronize(csc_ctx* ctx) {
  if (!ctx) return CSC_ERR_ARG;
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CSC_OK;
}

csc_status csc_device_info(csc_ctx* ctx, int32_t* sm_count, int64_t* free_bytes, int64_t* total_bytes,
                           int32_t* cc_major, int32_t* cc_minor) {
  if (!ctx) return CSC_ERR_ARG;
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t f = 0, t = 0;
  CSC_CUDA(ctx, cudaMemGetInfo(&f, &t));
  if (sm_count) *sm_count = ctx->sm_count;
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  return CSC_OK;
}

csc_status csc_timer_start(csc_ctx* ctx) {
  if (!ctx) return CSC_ERR_ARG;
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  CSC_CUDA(ctx, cudaEventRecord(ctx->ev_timer0, ctx->stream));
  return CSC_OK;
}

csc_status csc_timer_stop(csc_ctx* ctx, double* elapsed_ms) {
  if (!ctx || !elapsed_ms) return CSC_ERR_ARG;
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  CSC_CUDA(ctx, cudaEventRecord(ctx->ev_timer1, ctx->stream));
  CSC_CUDA(ctx, cudaEventSynchronize(ctx->ev_timer1));
  float ms = 0.f;
  CSC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_timer0, ctx->ev_timer1));
  *elapsed_ms = ms;
  return CSC_OK;
}

csc_status csc_stage_times(csc_ctx* ctx, double* ms, int64_t* kernel_launches, int32_t reset) {
  if (!ctx) return CSC_ERR_ARG;
  if (ms)
    for (int i = 0; i < CSC_T_COUNT; ++i) ms[i] = ctx->stage_ms[i];
  if (kernel_launches) *kernel_launches = ctx->launches;
  if (reset) {
    for (int i = 0; i < CSC_T_COUNT; ++i) ctx->stage_ms[i] = 0.0;
    ctx->launches = 0;
  }
  return CSC_OK;
}

csc_status csc_flush_l2(csc_ctx* ctx) {
  if (!ctx) return CSC_ERR_ARG;
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->flush_buf) {
    ctx->flush_bytes = (size_t)256 << 20;  // 256 MiB > 126 MB L2
    cudaError_t e = cudaMalloc(&ctx->flush_buf, ctx->flush_bytes);
    if (e != cudaSuccess) {
      ctx->flush_buf = nullptr;
      return csc_fail(ctx, CSC_ERR_OOM, "flush buffer: %s", cudaGetErrorString(e));
    }
  }
  CSC_CUDA(ctx, cudaMemsetAsync(ctx->flush_buf, 1, ctx->flush_bytes, ctx->stream));
  return CSC_OK;
}

// ---- vectors ----------------------------------------------------------------------------------
static size_t dtype_size(int32_t dt) {
  switch (dt) {
    case CSC_DT_F32: return 4;
    case CSC_DT_F64: return 8;
    case CSC_DT_I32: return 4;
    case CSC_DT_I64: return 8;
  }
  return 0;
}

csc_status csc_vec_alloc(csc_ctx* ctx, int64_t n, int32_t dtype, csc_vec** out) {
  if (!ctx || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, n >= 0 && dtype_size(dtype) > 0, "csc_vec_alloc: bad n=%lld or dtype=%d", (long long)n, dtype);
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  csc_vec* v = new csc_vec();
  v->ctx = ctx;
  v->n = n;
  v->dtype = dtype;
  csc_status s = dev_alloc(ctx, (size_t)n * dtype_size(dtype), &v->mem, true);
  if (s != CSC_OK) {
    delete v;
    return s;
  }
  *out = v;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

csc_status csc_vec_wrap(csc_ctx* ctx, void* device_ptr, int64_t n, int32_t dtype, csc_vec** out) {
  if (!ctx || !out || !device_ptr) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, n >= 0 && dtype_size(dtype) > 0, "csc_vec_wrap: bad n or dtype");
  csc_vec* v = new csc_vec();
  v->ctx = ctx;
  v->n = n;
  v->dtype = dtype;
  v->mem = std::make_shared<DevMem>();
  v->mem->p = device_ptr;
  v->mem->bytes = (size_t)n * dtype_size(dtype);
  v->mem->device = ctx->device;
  v->mem->owned = false;
  *out = v;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

csc_status csc_vec_free(csc_vec* v) {
  delete v;
  return CSC_OK;
}

csc_status csc_vec_info(const csc_vec* v, int64_t* n, int32_t* dtype, void** device_ptr) {
  if (!v) return CSC_ERR_ARG;
  if (n) *n = v->n;
  if (dtype) *dtype = v->dtype;
  if (device_ptr) *device_ptr = v->mem->p;
  return CSC_OK;
}

csc_status csc_vec_upload(csc_vec* v, const void* host, int64_t n) {
  if (!v || !host) return CSC_ERR_ARG;
  csc_ctx* ctx = v->ctx;
  CSC_REQUIRE(ctx, n >= 0 && n <= v->n, "csc_vec_upload: n=%lld exceeds %lld", (long long)n, (long long)v->n);
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  CSC_CUDA(ctx, cudaMemcpyAsync(v->mem->p, host, (size_t)n * dtype_size(v->dtype), cudaMemcpyHostToDevice, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CSC_OK;
}

csc_status csc_vec_download(const csc_vec* v, void* host, int64_t n) {
  if (!v || !host) return CSC_ERR_ARG;
  csc_ctx* ctx = v->ctx;
  CSC_REQUIRE(ctx, n >= 0 && n <= v->n, "csc_vec_download: n=%lld exceeds %lld", (long long)n, (long long)v->n);
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  return d2h_copy(ctx, host, v->mem->p, (size_t)n * dtype_size(v->dtype));
}

// ---- dense --------------------------------------------------------------------------------------
__global__ void k_f64_to_f32(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (float)in[i];
}
__global__ void k_f32_to_f64(const float* __restrict__ in, double* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (double)in[i];
}

__global__ void k_i32_add(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, int32_t add) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i] + add;
}
__global__ void k_i32_to_i64_add(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n, int64_t add) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (int64_t)in[i] + add;
}
csc_status csc_vec_download_as(const csc_vec* v, void* host, int64_t n, int32_t out_dtype, int64_t add) {
  if (!v || !host) return CSC_ERR_ARG;
  csc_ctx* ctx = v->ctx;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, n >= 0 && n <= v->n, "csc_vec_download_as: n=%lld exceeds %lld", (long long)n, (long long)v->n);
  if (out_dtype == v->dtype && add == 0) return csc_vec_download(v, host, n);
  const bool ok = (v->dtype == CSC_DT_I32 && (out_dtype == CSC_DT_I32 || out_dtype == CSC_DT_I64)) ||
                  (v->dtype == CSC_DT_F32 && out_dtype == CSC_DT_F64 && add == 0);
  CSC_REQUIRE(ctx, ok, "csc_vec_download_as: unsupported conversion %d -> %d (add %lld)", v->dtype, out_dtype, (long long)add);
  if (n == 0) return CSC_OK;
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  DevPtr tmp;
  CSC_TRY(dev_alloc(ctx, (size_t)n * dtype_size(out_dtype), &tmp));
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(n, 256), (int64_t)ctx->sm_count * 16);
  if (out_dtype == CSC_DT_I32) k_i32_add<<<grid, 256, 0, ctx->stream>>>(v->as<int32_t>(), (int32_t*)tmp->p, n, (int32_t)add);
  else if (out_dtype == CSC_DT_I64) k_i32_to_i64_add<<<grid, 256, 0, ctx->stream>>>(v->as<int32_t>(), (int64_t*)tmp->p, n, add);
  else k_f32_to_f64<<<grid, 256, 0, ctx->stream>>>(v->as<float>(), (double*)tmp->p, n);
  CSC_LAUNCHED(ctx);
  return d2h_copy(ctx, host, tmp->p, (size_t)n * dtype_size(out_dtype));
  CSC_GUARD_END(ctx)
}

csc_status csc_dense_alloc(csc_ctx* ctx, int64_t n_rows, int64_t n_cols, csc_dense** out) {
  if (!ctx || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, n_rows >= 0 && n_cols >= 0, "csc_dense_alloc: negative shape");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  csc_dense* d = new csc_dense();
  d->ctx = ctx;
  d->n_rows = n_rows;
  d->n_cols = n_cols;
  csc_status s = dev_alloc(ctx, (size_t)n_rows * n_cols * sizeof(float), &d->mem, false);
  if (s != CSC_OK) {
    delete d;
    return s;
  }
  *out = d;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

csc_status csc_dense_upload(csc_ctx* ctx, const void* host, int32_t dtype, int64_t n_rows, int64_t n_cols,
                            csc_dense** out) {
  if (!ctx || !out || !host) return CSC_ERR_ARG;
  CSC_REQUIRE(ctx, dtype == CSC_DT_F32 || dtype == CSC_DT_F64, "csc_dense_upload: dtype must be F32 or F64");
  csc_dense* d = nullptr;
  CSC_TRY(csc_dense_alloc(ctx, n_rows, n_cols, &d));
  int64_t n = n_rows * n_cols;
  cudaError_t e = cudaSuccess;
  if (dtype == CSC_DT_F32) {
    e = cudaMemcpyAsync(d->p(), host, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
  } else {
    DevPtr tmp;
    csc_status s = dev_alloc(ctx, (size_t)n * 8, &tmp);
    if (s != CSC_OK) {
      delete d;
      return s;
    }
    e = cudaMemcpyAsync(tmp->p, host, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && n > 0) {
      k_f64_to_f32<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 148 * 16), 256, 0, ctx->stream>>>(
          (const double*)tmp->p, d->p(), n);
      ctx->launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    delete d;
    return csc_fail(ctx, CSC_ERR_CUDA, "csc_dense_upload: %s", cudaGetErrorString(e));
  }
  *out = d;
  return CSC_OK;
}

csc_status csc_dense_download(const csc_dense* d, void* host, int32_t dtype, int64_t row_lo, int64_t row_hi) {
  if (!d || !host) return CSC_ERR_ARG;
  csc_ctx* ctx = d->ctx;
  CSC_REQUIRE(ctx, dtype == CSC_DT_F32 || dtype == CSC_DT_F64, "csc_dense_download: dtype must be F32 or F64");
  CSC_REQUIRE(ctx, 0 <= row_lo && row_lo <= row_hi && row_hi <= d->n_rows, "csc_dense_download: bad row range");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  int64_t n = (row_hi - row_lo) * d->n_cols;
  const float* src = d->p() + row_lo * d->n_cols;
  if (n == 0) return CSC_OK;
  if (dtype == CSC_DT_F32) {
    CSC_TRY(d2h_copy(ctx, host, src, (size_t)n * 4));
  } else {
    // convert on the device in bounded chunks so that the staging buffer stays small
    const int64_t chunk = (int64_t)32 << 20;  // elements
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, (size_t)std::min(n, chunk) * 8, &tmp));
    for (int64_t o = 0; o < n; o += chunk) {
      int64_t m = std::min(chunk, n - o);
      k_f32_to_f64<<<(unsigned)std::min<int64_t>(ceil_div64(m, 256), 148 * 16), 256, 0, ctx->stream>>>(
          src + o, (double*)tmp->p, m);
      CSC_LAUNCHED(ctx);
      CSC_TRY(d2h_copy(ctx, (double*)host + o, tmp->p, (size_t)m * 8));
    }
  }
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CSC_OK;
}

csc_status csc_dense_info(const csc_dense* d, int64_t* n_rows, int64_t* n_cols, void** device_ptr) {
  if (!d) return CSC_ERR_ARG;
  if (n_rows) *n_rows = d->n_rows;
  if (n_cols) *n_cols = d->n_cols;
  if (device_ptr) *device_ptr = d->mem->p;
  return CSC_OK;
}

csc_status csc_dense_wrap(csc_ctx* ctx, void* device_ptr, int64_t n_rows, int64_t n_cols, csc_dense** out) {
  if (!ctx || !out || !device_ptr) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  csc_dense* d = new csc_dense();
  d->ctx = ctx;
  d->n_rows = n_rows;
  d->n_cols = n_cols;
  d->mem = std::make_shared<DevMem>();
  d->mem->p = device_ptr;
  d->mem->bytes = (size_t)n_rows * n_cols * 4;
  d->mem->device = ctx->device;
  d->mem->owned = false;
  *out = d;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

csc_status csc_dense_free(csc_dense* d) {
  delete d;
  return CSC_OK;
}

}  // extern "C"
