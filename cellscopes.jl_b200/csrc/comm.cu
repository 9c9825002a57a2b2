// Multi-GPU exchange steps of the path INSIDE the library (north_star): one process per GPU, one csc_ctx
// each, NCCL over NVLink for the natural collectives --
//   all-reduce (sum)  per-gene moment tables, the H x H Gram and the column sums        csc_comm_allreduce_sum
//   all-gather        the 50-d embedding (kNN keys) and the k x N neighbour table        csc_comm_allgather_rows / _i32
// Everything is enqueued on the CONTEXT's stream, i.e. ordered with the kernels that produce and consume the buffers:
// no host synchronisation around a collective (round 1 ran them through torch.distributed on torch's stream, with a
// device-wide synchronisation after each one).  The ragged all-gathers need no padding: every rank broadcasts its
// slice into place inside one NCCL group.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2): the library has no link-time dependency on it, a single-GPU
// caller never touches it, and a process that already carries an NCCL (PyTorch's) shares that copy.  The rendezvous
// is the caller's: rank 0 makes a unique id (csc_comm_unique_id), the host moves the 128 bytes to the other ranks by
// whatever means it has (MPI.jl / Distributed.jl / a file / torch.distributed), every rank calls csc_comm_init.
#include <dlfcn.h>

#include "common.cuh"

namespace {

// the slice of nccl.h this file uses (NCCL 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;     // 0 = ncclSuccess
enum { ncclInt32 = 2, ncclInt64 = 4, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string error;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* names[] = {getenv("CSC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
    const char* why = dlerror();                          // one call: dlerror() clears the message it returns
    api.error = why ? why : "dlopen failed";
  }
  if (!api.handle) return &api;
#define NCCL_SYM(field, name)                                        \
  do {                                                               \
    *(void**)(&api.field) = dlsym(api.handle, name);                 \
    if (!api.field) {                                                \
      api.error = std::string("missing symbol ") + name;             \
      api.handle = nullptr;                                          \
      return &api;                                                   \
    }                                                                \
  } while (0)
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  NCCL_SYM(CommInitRank, "ncclCommInitRank");
  NCCL_SYM(CommDestroy, "ncclCommDestroy");
  NCCL_SYM(AllReduce, "ncclAllReduce");
  NCCL_SYM(Broadcast, "ncclBroadcast");
  NCCL_SYM(AllGather, "ncclAllGather");
  NCCL_SYM(GroupStart, "ncclGroupStart");
  NCCL_SYM(GroupEnd, "ncclGroupEnd");
  NCCL_SYM(GetErrorString, "ncclGetErrorString");
  NCCL_SYM(GetVersion, "ncclGetVersion");
#undef NCCL_SYM
  return &api;
}

#define CSC_NCCL(ctx, call)                                                                                          \
  do {                                                                                                               \
    ncclResult_t r__ = (call);                                                                                       \
    if (r__ != 0) return csc_fail((ctx), CSC_ERR_CUDA, "%s:%d %s -> NCCL: %s", __FILE__, __LINE__, #call,             \
                                  nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r__) : "?");               \
  } while (0)

int nccl_dtype(int32_t dt) {
  switch (dt) {
    case CSC_DT_F32: return ncclFloat32;
    case CSC_DT_F64: return ncclFloat64;
    case CSC_DT_I32: return ncclInt32;
    default: return ncclInt64;
  }
}

csc_status require_comm(csc_ctx* ctx) {
  if (!ctx->comm) return csc_fail(ctx, CSC_ERR_ARG, "no communicator on this context: call csc_comm_init first");
  return CSC_OK;
}

// out[offsets[r] * width ... offsets[r+1] * width) <- rank r's slice, every rank; elem = bytes per element
csc_status allgather_ragged(csc_ctx* ctx, const void* local, void* out, const int64_t* offsets, int64_t width, size_t elem,
                            int dtype) {
  NcclApi* n = nccl_api();
  const int world = ctx->comm_world;
  bool equal = true;
  for (int r = 0; r < world; ++r) equal = equal && (offsets[r + 1] - offsets[r] == offsets[1] - offsets[0]);
  if (equal) {
    CSC_NCCL(ctx, n->AllGather(local, out, (size_t)((offsets[1] - offsets[0]) * width), dtype, (ncclComm_t)ctx->comm, ctx->stream));
    return CSC_OK;
  }
  CSC_NCCL(ctx, n->GroupStart());
  for (int r = 0; r < world; ++r) {
    const size_t cnt = (size_t)((offsets[r + 1] - offsets[r]) * width);
    char* dst = (char*)out + (size_t)offsets[r] * width * elem;
    if (cnt == 0) continue;
    const ncclResult_t rc = n->Broadcast(r == ctx->comm_rank ? local : dst, dst, cnt, dtype, r, (ncclComm_t)ctx->comm, ctx->stream);
    if (rc != 0) {
      n->GroupEnd();
      return csc_fail(ctx, CSC_ERR_CUDA, "ncclBroadcast (ragged all-gather) -> NCCL: %s", n->GetErrorString(rc));
    }
  }
  CSC_NCCL(ctx, n->GroupEnd());
  return CSC_OK;
}

csc_status check_offsets(csc_ctx* ctx, const int64_t* offsets, int64_t local_rows, const char* who) {
  CSC_REQUIRE(ctx, offsets != nullptr && offsets[0] == 0, "%s: row_offsets must start at 0", who);
  for (int r = 0; r < ctx->comm_world; ++r)
    CSC_REQUIRE(ctx, offsets[r + 1] >= offsets[r], "%s: row_offsets must not decrease", who);
  CSC_REQUIRE(ctx, offsets[ctx->comm_rank + 1] - offsets[ctx->comm_rank] == local_rows,
              "%s: this rank holds %lld rows, row_offsets say %lld", who, (long long)local_rows,
              (long long)(offsets[ctx->comm_rank + 1] - offsets[ctx->comm_rank]));
  return CSC_OK;
}

}  // namespace

// host fp64 vector summed over the ranks (small tables: the Loess vertex fits dealt out by csc_hvg_finish)
csc_status comm_allreduce_f64(csc_ctx* ctx, double* host_inout, size_t n) {
  if (!ctx->comm || ctx->comm_world <= 1 || n == 0) return CSC_OK;
  DevPtr d;
  CSC_TRY(dev_alloc(ctx, n * 8, &d));
  CSC_CUDA(ctx, cudaMemcpyAsync(d->p, host_inout, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  CSC_NCCL(ctx, nccl_api()->AllReduce(d->p, d->p, n, ncclFloat64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  CSC_CUDA(ctx, cudaMemcpyAsync(host_inout, d->p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CSC_OK;
}

void comm_release(csc_ctx* ctx) {
  if (ctx->comm && nccl_api()->CommDestroy) nccl_api()->CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
}

extern "C" {

csc_status csc_comm_unique_id(void* id_out, int64_t id_bytes) {
  if (!id_out || id_bytes < (int64_t)sizeof(ncclUniqueId)) return csc_fail(nullptr, CSC_ERR_ARG, "csc_comm_unique_id: need 128 bytes");
  NcclApi* n = nccl_api();
  if (!n->handle) return csc_fail(nullptr, CSC_ERR_UNSUPPORTED, "NCCL could not be loaded (%s)", n->error.c_str());
  ncclUniqueId id;
  const ncclResult_t rc = n->GetUniqueId(&id);
  if (rc != 0) return csc_fail(nullptr, CSC_ERR_CUDA, "ncclGetUniqueId -> NCCL: %s", n->GetErrorString(rc));
  memcpy(id_out, &id, sizeof(id));
  return CSC_OK;
}

csc_status csc_comm_init(csc_ctx* ctx, const void* unique_id, int64_t id_bytes, int32_t rank, int32_t world) {
  if (!ctx || !unique_id) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, id_bytes >= (int64_t)sizeof(ncclUniqueId), "csc_comm_init: the unique id has 128 bytes");
  CSC_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world, "csc_comm_init: rank %d of %d", rank, world);
  CSC_REQUIRE(ctx, ctx->comm == nullptr, "csc_comm_init: this context already has a communicator");
  NcclApi* n = nccl_api();
  if (!n->handle) return csc_fail(ctx, CSC_ERR_UNSUPPORTED, "NCCL could not be loaded (%s)", n->error.c_str());
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  CSC_NCCL(ctx, n->CommInitRank(&comm, world, id, rank));
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

csc_status csc_comm_info(const csc_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version) {
  if (!ctx) return CSC_ERR_ARG;
  if (rank) *rank = ctx->comm_rank;
  if (world) *world = ctx->comm_world;
  if (nccl_version) {
    int v = 0;
    if (nccl_api()->handle) nccl_api()->GetVersion(&v);
    *nccl_version = v;
  }
  return CSC_OK;
}

csc_status csc_comm_destroy(csc_ctx* ctx) {
  if (!ctx) return CSC_ERR_ARG;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
  return CSC_OK;
}

// in place, on the context's stream; returns without waiting (the next kernel on the stream sees the sum)
csc_status csc_comm_allreduce_sum(csc_ctx* ctx, csc_vec* v) {
  if (!ctx || !v) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_TRY(require_comm(ctx));
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (v->n == 0) return CSC_OK;
  CSC_NCCL(ctx, nccl_api()->AllReduce(v->mem->p, v->mem->p, (size_t)v->n, nccl_dtype(v->dtype), ncclSum, (ncclComm_t)ctx->comm,
                                      ctx->stream));
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

csc_status csc_comm_allgather_rows(csc_ctx* ctx, const csc_dense* local, const int64_t* row_offsets, csc_dense** out) {
  if (!ctx || !local || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_TRY(require_comm(ctx));
  CSC_TRY(check_offsets(ctx, row_offsets, local->n_rows, "csc_comm_allgather_rows"));
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  csc_dense* d = nullptr;
  CSC_TRY(csc_dense_alloc(ctx, row_offsets[ctx->comm_world], local->n_cols, &d));
  const csc_status rc = allgather_ragged(ctx, local->p(), d->p(), row_offsets, local->n_cols, 4, ncclFloat32);
  if (rc != CSC_OK) {
    csc_dense_free(d);
    return rc;
  }
  *out = d;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

csc_status csc_comm_allgather_i32(csc_ctx* ctx, const csc_vec* local, const int64_t* row_offsets, int32_t width, csc_vec** out) {
  if (!ctx || !local || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_TRY(require_comm(ctx));
  CSC_REQUIRE(ctx, local->dtype == CSC_DT_I32 && width >= 1 && local->n % width == 0, "csc_comm_allgather_i32: int32 [rows x width]");
  CSC_TRY(check_offsets(ctx, row_offsets, local->n / width, "csc_comm_allgather_i32"));
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  csc_vec* v = nullptr;
  CSC_TRY(csc_vec_alloc(ctx, row_offsets[ctx->comm_world] * width, CSC_DT_I32, &v));
  const csc_status rc = allgather_ragged(ctx, local->mem->p, v->mem->p, row_offsets, width, 4, ncclInt32);
  if (rc != CSC_OK) {
    csc_vec_free(v);
    return rc;
  }
  *out = v;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

}  // extern "C"
