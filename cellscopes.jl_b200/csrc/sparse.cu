// Sparse stages: ingest (narrowing), normalize_object, raw gene moments, scale_object.
// All kernels stream the cell-major CSC once, a warp per cell (column), coalesced along the
// non-zeros of the column, with warp-shuffle reductions for the per-cell sums.
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include <cuda_fp16.h>

#include <cub/cub.cuh>

#include "common.cuh"

// ------------------------------------------------------------------------------------------
// ingest: Int64/Float64 host arrays -> int64 colptr (0-based) / int32 rowval / fp32 values
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_colptr_rebase(const T* __restrict__ in, int64_t* __restrict__ out, int64_t n, int64_t base) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (int64_t)in[i] - base;
}
// 0 <= cp[i] <= cp[i+1] <= nnz for every column (every warp-per-column kernel trusts this)
__global__ void k_colptr_check(const int64_t* __restrict__ cp, int64_t n_cells, int64_t nnz, int* __restrict__ bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n_cells; i += stride) {
    const int64_t b = cp[i], e = cp[i + 1];
    if (b < 0 || e < b || e > nnz) atomicMax(bad, 2);
  }
}
template <typename T>
__global__ void k_narrow_idx(const T* __restrict__ in, int32_t* __restrict__ out, int64_t n, int64_t base,
                             int64_t n_rows, int* __restrict__ bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t r = (int64_t)in[i] - base;
    if (r < 0 || r >= n_rows) atomicMax(bad, 1);
    out[i] = (int32_t)r;
  }
}
template <typename T>
__global__ void k_narrow_val(const T* __restrict__ in, float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (float)in[i];
}
__global__ void k_widen_idx(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n, int64_t base) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (int64_t)in[i] + base;
}
__global__ void k_widen_val(const float* __restrict__ in, double* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (double)in[i];
}

static inline unsigned grid_for(int64_t n, int sm) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n, 256), (int64_t)sm * 16));
}

// CTAs of `kern` that are resident per SM at this block size and dynamic shared memory: the persistent kernels divide
// their tasks statically over the grid, so a grid above the resident count runs a thin second wave
template <typename K>
static int resident_ctas(K kern, int threads, size_t smem, int fallback) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fallback;
  }
  return n;
}

static size_t dt_size(int32_t dt) { return (dt == CSC_DT_F32 || dt == CSC_DT_I32) ? 4 : 8; }

// stage `n` host elements through a bounded device buffer and run `conv(dev_src, offset, count)` on each chunk
template <typename F>
static csc_status staged_upload(csc_ctx* ctx, const void* host, size_t elem, int64_t n, F conv) {
  const int64_t chunk = (int64_t)64 << 20;  // elements per chunk
  if (n == 0) return CSC_OK;
  DevPtr stage[2];
  int64_t cap = std::min(n, chunk);
  CSC_TRY(dev_alloc(ctx, (size_t)cap * elem, &stage[0]));
  if (n > chunk) CSC_TRY(dev_alloc(ctx, (size_t)cap * elem, &stage[1]));
  int b = 0;
  for (int64_t o = 0; o < n; o += chunk, b ^= 1) {
    int64_t m = std::min(chunk, n - o);
    // same stream: the copy into stage[b] is ordered after the kernel that last read it
    CSC_CUDA(ctx, cudaMemcpyAsync(stage[b]->p, (const char*)host + (size_t)o * elem, (size_t)m * elem,
                                  cudaMemcpyHostToDevice, ctx->stream));
    CSC_TRY(conv(stage[b]->p, o, m));
  }
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CSC_OK;
}

// ------------------------------------------------------------------------------------------
// Two-lane upload of a large 64-bit array.  The reference's matrices arrive as Int64 row indices and Int64 / Float64
// values (16 B per stored entry); the device keeps int32 / fp32 (8 B).  Copying the wide arrays and narrowing on the
// device is bound by PCIe at 16 B per entry (26.4 GB = 490 ms of a 692 ms end-to-end pass at C3).  Here the array is
// eaten from both ends at once:
//   * host lane: worker threads claim 2 M-element pieces from the FRONT, narrow them into page-locked slots (row
//     indices to uint16 when n_genes <= 65536, values to uint16 when every value of the piece is an integer in
//     [0, 65535] -- raw counts are -- else to int32 / fp32) and hand them to the bus; a uint16 piece is widened to its
//     final place by a small kernel (2 B instead of 8 B on the bus);
//   * bus lane: whenever no narrowed piece is waiting and fewer than two copies are queued, the main thread claims a
//     16 M-element piece from the BACK and copies it wide, narrowed by the device kernels as before.
// The split point is wherever the two lanes meet, so the host cores take exactly the share they can convert while the
// bus is busy, on any mix of core count, memory bandwidth and link speed.  Everything is enqueued on the context's stream.
// ------------------------------------------------------------------------------------------
enum UpKind { UP_IDX_I64 = 0, UP_VAL_F64 = 1, UP_VAL_I64 = 2 };
static const int64_t UP_SLOT_MAX = (int64_t)2 << 20;   // elements per host-lane piece (bus-lane pieces: 8 of these)

__global__ void k_widen16_idx(const uint16_t* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (int32_t)in[i];
}
__global__ void k_widen16_val(const uint16_t* __restrict__ in, float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (float)in[i];
}

// host_narrow.cpp (g++: AVX-512 behind a run-time CPU check): narrow `cnt` elements at `src` into `stage`; returns the
// element size written (2 or 4); row indices outside [0, n_rows) set *bad
int host_narrow_piece(int kind, const void* src, int64_t cnt, void* stage, bool idx16, int64_t base, int64_t n_rows, int* bad);

static csc_status two_lane_upload(csc_ctx* ctx, int kind, const void* host, void* dev, int64_t n, int64_t base,
                                  int64_t n_rows, int* d_bad, int* host_bad) {
  if (n == 0) return CSC_OK;
  int64_t UP_SLOT = UP_SLOT_MAX;               // CSC_UPLOAD_PIECE: smaller pieces, so that tests reach every branch on small inputs
  if (const char* e = getenv("CSC_UPLOAD_PIECE")) UP_SLOT = std::max<int64_t>(16, std::min<int64_t>(UP_SLOT_MAX, atoll(e)));
  const int64_t UP_WIDE = 8 * UP_SLOT;
  // the host cores are shared by the ranks of the box (one process per GPU): each takes its share, minus the bus thread
  unsigned hw = std::thread::hardware_concurrency() / (unsigned)std::max(1, ctx->comm_world);
  int n_workers = (int)std::max(1u, std::min(31u, hw > 1 ? hw - 1 : 1u));
  if (const char* e = getenv("CSC_UPLOAD_THREADS")) n_workers = std::max(1, std::min(64, atoi(e)));
  const int n_slots = 2 * n_workers;
  const size_t slot_bytes = (size_t)UP_SLOT * 4;
  if (ctx->up_stage_bytes < slot_bytes * n_slots) {
    if (ctx->up_stage) cudaFreeHost(ctx->up_stage);
    ctx->up_stage = nullptr;
    ctx->up_stage_bytes = 0;
    CSC_CUDA(ctx, cudaHostAlloc(&ctx->up_stage, slot_bytes * n_slots, cudaHostAllocDefault));
    ctx->up_stage_bytes = slot_bytes * n_slots;
  }
  const bool idx16 = kind == UP_IDX_I64 && n_rows <= 65536;
  DevPtr d16, dwide[2];
  CSC_TRY(dev_alloc(ctx, (size_t)UP_SLOT * 2 * n_slots, &d16));
  for (int b = 0; b < 2; ++b) CSC_TRY(dev_alloc(ctx, (size_t)std::min(n, UP_WIDE) * 8, &dwide[b]));
  struct Events {
    std::vector<cudaEvent_t> ev;
    ~Events() {
      for (auto e : ev) cudaEventDestroy(e);
    }
  } events;
  events.ev.resize(n_slots + 2);
  for (auto& e : events.ev) {
    e = nullptr;
    CSC_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }

  struct Piece {
    int slot;
    int64_t off, cnt;
    int elem;
  };
  std::mutex mu;
  std::condition_variable cv_workers, cv_main;
  int64_t front = 0, back = n;                 // [front, back) is unclaimed
  std::deque<Piece> ready;
  std::vector<char> slot_free(n_slots, 1);
  int active = n_workers;
  bool abort = false;
  int bad_local = 0;

  auto worker = [&](int w) {
    for (;;) {
      Piece pc;
      {
        std::unique_lock<std::mutex> lk(mu);
        if (abort || front >= back) break;
        pc.off = front;
        pc.cnt = std::min(UP_SLOT, back - front);
        front += pc.cnt;
        cv_workers.wait(lk, [&] { return abort || slot_free[2 * w] || slot_free[2 * w + 1]; });
        if (abort) break;
        pc.slot = slot_free[2 * w] ? 2 * w : 2 * w + 1;
        slot_free[pc.slot] = 0;
      }
      int bad = 0;
      pc.elem = host_narrow_piece(kind, (const char*)host + (size_t)pc.off * 8, pc.cnt,
                                (char*)ctx->up_stage + slot_bytes * pc.slot, idx16, base, n_rows, &bad);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (bad) bad_local = 1;
        ready.push_back(pc);
      }
      cv_main.notify_one();
    }
    {
      std::lock_guard<std::mutex> lk(mu);
      --active;
    }
    cv_main.notify_one();
  };
  std::vector<std::thread> pool;
  struct Joiner {                              // on every way out: stop the workers and wait for them
    std::vector<std::thread>& pool;
    std::mutex& mu;
    std::condition_variable& cv;
    bool& abort;
    ~Joiner() {
      {
        std::lock_guard<std::mutex> lk(mu);
        abort = true;
      }
      cv.notify_all();
      for (auto& t : pool) t.join();
    }
  } joiner{pool, mu, cv_workers, abort};
  for (int w = 0; w < n_workers; ++w) {
    try {
      pool.emplace_back(worker, w);
    } catch (...) {
      std::lock_guard<std::mutex> lk(mu);
      --active;                                // the bus lane (and the workers that did start) take its share
    }
  }

  struct Flight {
    cudaEvent_t ev;
    int slot;                                  // host slot to release on completion, -1 for a bus-lane piece
  };
  std::deque<Flight> flight;
  const int sm = ctx->sm_count;
  int wide_b = 0;
  auto retire = [&](bool block) -> cudaError_t {
    while (!flight.empty()) {
      cudaError_t e = block ? cudaEventSynchronize(flight.front().ev) : cudaEventQuery(flight.front().ev);
      if (e == cudaErrorNotReady) return cudaSuccess;
      if (e != cudaSuccess) return e;
      if (flight.front().slot >= 0) {
        {
          std::lock_guard<std::mutex> lk(mu);
          slot_free[flight.front().slot] = 1;
        }
        cv_workers.notify_all();
      }
      flight.pop_front();
      block = false;
    }
    return cudaSuccess;
  };
  for (;;) {
    CSC_CUDA(ctx, retire(false));
    Piece pc{-1, 0, 0, 0};
    bool wide = false, done = false;
    {
      std::unique_lock<std::mutex> lk(mu);
      if (!ready.empty()) {
        pc = ready.front();
        ready.pop_front();
      } else if (front < back && flight.size() < 2) {
        pc.cnt = std::min(UP_WIDE, back - front);
        back -= pc.cnt;
        pc.off = back;
        wide = true;
      } else if (active == 0 && front >= back) {
        done = true;
      } else if (flight.empty()) {
        cv_main.wait_for(lk, std::chrono::microseconds(200));
        continue;
      }
    }
    if (done) break;
    if (pc.cnt == 0) {                         // nothing to issue: wait for the oldest queued copy
      CSC_CUDA(ctx, retire(true));
      continue;
    }
    if (wide) {
      void* st = dwide[wide_b]->p;
      CSC_CUDA(ctx, cudaMemcpyAsync(st, (const char*)host + (size_t)pc.off * 8, (size_t)pc.cnt * 8, cudaMemcpyHostToDevice,
                                    ctx->stream));
      if (kind == UP_IDX_I64)
        k_narrow_idx<int64_t><<<grid_for(pc.cnt, sm), 256, 0, ctx->stream>>>((const int64_t*)st, (int32_t*)dev + pc.off, pc.cnt,
                                                                           base, n_rows, d_bad);
      else if (kind == UP_VAL_F64)
        k_narrow_val<double><<<grid_for(pc.cnt, sm), 256, 0, ctx->stream>>>((const double*)st, (float*)dev + pc.off, pc.cnt);
      else
        k_narrow_val<int64_t><<<grid_for(pc.cnt, sm), 256, 0, ctx->stream>>>((const int64_t*)st, (float*)dev + pc.off, pc.cnt);
      CSC_LAUNCHED(ctx);
      cudaEvent_t ev = events.ev[n_slots + wide_b];
      CSC_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
      flight.push_back({ev, -1});
      wide_b ^= 1;
    } else {
      const char* st = (const char*)ctx->up_stage + slot_bytes * pc.slot;
      if (pc.elem == 4) {
        CSC_CUDA(ctx, cudaMemcpyAsync((char*)dev + (size_t)pc.off * 4, st, (size_t)pc.cnt * 4, cudaMemcpyHostToDevice,
                                      ctx->stream));
      } else {
        uint16_t* d = (uint16_t*)d16->p + (size_t)UP_SLOT * pc.slot;
        CSC_CUDA(ctx, cudaMemcpyAsync(d, st, (size_t)pc.cnt * 2, cudaMemcpyHostToDevice, ctx->stream));
        if (kind == UP_IDX_I64)
          k_widen16_idx<<<grid_for(pc.cnt, sm), 256, 0, ctx->stream>>>(d, (int32_t*)dev + pc.off, pc.cnt);
        else
          k_widen16_val<<<grid_for(pc.cnt, sm), 256, 0, ctx->stream>>>(d, (float*)dev + pc.off, pc.cnt);
        CSC_LAUNCHED(ctx);
      }
      cudaEvent_t ev = events.ev[pc.slot];
      CSC_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
      flight.push_back({ev, pc.slot});
    }
  }
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad_local) *host_bad = 1;
  return CSC_OK;
}

extern "C" csc_status csc_sparse_upload(csc_ctx* ctx, int64_t n_genes, int64_t n_cells, const void* colptr,
                                        const void* rowval, int32_t idx_dtype, const void* nzval, int32_t val_dtype,
                                        int32_t index_base, csc_sparse** out) {
  if (!ctx || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, colptr != nullptr, "csc_sparse_upload: colptr is NULL");
  CSC_REQUIRE(ctx, n_genes > 0 && n_cells >= 0, "csc_sparse_upload: bad shape %lld x %lld", (long long)n_genes,
              (long long)n_cells);
  CSC_REQUIRE(ctx, n_genes < ((int64_t)1 << 31), "csc_sparse_upload: n_genes must fit int32");
  CSC_REQUIRE(ctx, idx_dtype == CSC_DT_I64 || idx_dtype == CSC_DT_I32, "csc_sparse_upload: idx_dtype must be I64 or I32");
  CSC_REQUIRE(ctx, val_dtype >= CSC_DT_F32 && val_dtype <= CSC_DT_I64, "csc_sparse_upload: bad val_dtype");
  CSC_REQUIRE(ctx, index_base == 0 || index_base == 1, "csc_sparse_upload: index_base must be 0 or 1");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_UPLOAD);
  int64_t first, last;
  if (idx_dtype == CSC_DT_I64) {
    first = ((const int64_t*)colptr)[0];
    last = ((const int64_t*)colptr)[n_cells];
  } else {
    first = ((const int32_t*)colptr)[0];
    last = ((const int32_t*)colptr)[n_cells];
  }
  CSC_REQUIRE(ctx, first == index_base, "csc_sparse_upload: colptr[0]=%lld != index_base=%d", (long long)first, index_base);
  int64_t nnz = last - first;
  CSC_REQUIRE(ctx, nnz >= 0, "csc_sparse_upload: negative nnz");
  CSC_REQUIRE(ctx, nnz == 0 || (rowval && nzval), "csc_sparse_upload: rowval/nzval NULL with nnz>0");
  std::unique_ptr<csc_sparse> m(new csc_sparse());
  m->ctx = ctx;
  m->n_genes = n_genes;
  m->n_cells = n_cells;
  m->nnz = nnz;
  CSC_TRY(dev_alloc(ctx, (size_t)(n_cells + 1) * 8, &m->colptr));
  CSC_TRY(dev_alloc(ctx, (size_t)nnz * 4, &m->rowval));
  CSC_TRY(dev_alloc(ctx, (size_t)nnz * 4, &m->val));
  DevPtr bad;
  CSC_TRY(dev_alloc(ctx, 4, &bad, true));
  int sm = ctx->sm_count;
  int64_t* d_cp = (int64_t*)m->colptr->p;
  int32_t* d_rv = (int32_t*)m->rowval->p;
  float* d_v = (float*)m->val->p;
  int* d_bad = (int*)bad->p;
  CSC_TRY(staged_upload(ctx, colptr, dt_size(idx_dtype), n_cells + 1, [&](void* src, int64_t o, int64_t cnt) {
    if (idx_dtype == CSC_DT_I64)
      k_colptr_rebase<int64_t><<<grid_for(cnt, sm), 256, 0, ctx->stream>>>((const int64_t*)src, d_cp + o, cnt, index_base);
    else
      k_colptr_rebase<int32_t><<<grid_for(cnt, sm), 256, 0, ctx->stream>>>((const int32_t*)src, d_cp + o, cnt, index_base);
    CSC_LAUNCHED(ctx);
    return (csc_status)CSC_OK;
  }));
  if (n_cells > 0) {
    k_colptr_check<<<grid_for(n_cells, sm), 256, 0, ctx->stream>>>(d_cp, n_cells, nnz, d_bad);
    CSC_LAUNCHED(ctx);
  }
  // 64-bit inputs of at least 32 M entries take the two-lane upload (see two_lane_upload); CSC_UPLOAD_HOST_NARROW=0 keeps
  // the plain wide copy + device narrowing, =2 takes the two lanes at any size
  const char* hn = getenv("CSC_UPLOAD_HOST_NARROW");
  const bool two_lane = hn ? atoi(hn) == 2 || (atoi(hn) != 0 && nnz >= ((int64_t)32 << 20)) : nnz >= ((int64_t)32 << 20);
  int host_bad = 0;
  if (two_lane && idx_dtype == CSC_DT_I64) {
    CSC_TRY(two_lane_upload(ctx, UP_IDX_I64, rowval, d_rv, nnz, index_base, n_genes, d_bad, &host_bad));
  } else
  CSC_TRY(staged_upload(ctx, rowval, dt_size(idx_dtype), nnz, [&](void* src, int64_t o, int64_t cnt) {
    if (idx_dtype == CSC_DT_I64)
      k_narrow_idx<int64_t><<<grid_for(cnt, sm), 256, 0, ctx->stream>>>((const int64_t*)src, d_rv + o, cnt, index_base, n_genes, d_bad);
    else
      k_narrow_idx<int32_t><<<grid_for(cnt, sm), 256, 0, ctx->stream>>>((const int32_t*)src, d_rv + o, cnt, index_base, n_genes, d_bad);
    CSC_LAUNCHED(ctx);
    return (csc_status)CSC_OK;
  }));
  if (two_lane && (val_dtype == CSC_DT_F64 || val_dtype == CSC_DT_I64)) {
    CSC_TRY(two_lane_upload(ctx, val_dtype == CSC_DT_F64 ? UP_VAL_F64 : UP_VAL_I64, nzval, d_v, nnz, 0, 0, d_bad, &host_bad));
  } else
  CSC_TRY(staged_upload(ctx, nzval, dt_size(val_dtype), nnz, [&](void* src, int64_t o, int64_t cnt) {
    switch (val_dtype) {
      case CSC_DT_F32: k_narrow_val<float><<<grid_for(cnt, sm), 256, 0, ctx->stream>>>((const float*)src, d_v + o, cnt); break;
      case CSC_DT_F64: k_narrow_val<double><<<grid_for(cnt, sm), 256, 0, ctx->stream>>>((const double*)src, d_v + o, cnt); break;
      case CSC_DT_I32: k_narrow_val<int32_t><<<grid_for(cnt, sm), 256, 0, ctx->stream>>>((const int32_t*)src, d_v + o, cnt); break;
      default: k_narrow_val<int64_t><<<grid_for(cnt, sm), 256, 0, ctx->stream>>>((const int64_t*)src, d_v + o, cnt); break;
    }
    CSC_LAUNCHED(ctx);
    return (csc_status)CSC_OK;
  }));
  int h_bad = 0;
  CSC_CUDA(ctx, cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CSC_REQUIRE(ctx, h_bad != 2, "csc_sparse_upload: colptr is not monotone within [index_base, index_base + nnz]");
  CSC_REQUIRE(ctx, h_bad == 0 && host_bad == 0, "csc_sparse_upload: a row index is outside [0, n_genes)");
  *out = m.release();
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_sparse_info(const csc_sparse* m, int64_t* n_genes, int64_t* n_cells, int64_t* nnz) {
  if (!m) return CSC_ERR_ARG;
  if (n_genes) *n_genes = m->n_genes;
  if (n_cells) *n_cells = m->n_cells;
  if (nnz) *nnz = m->nnz;
  return CSC_OK;
}

extern "C" csc_status csc_sparse_free(csc_sparse* m) {
  delete m;
  return CSC_OK;
}

static csc_status range_ptrs(const csc_sparse* m, int64_t lo, int64_t hi, int64_t* b, int64_t* e) {
  csc_ctx* ctx = m->ctx;
  CSC_REQUIRE(ctx, 0 <= lo && lo <= hi && hi <= m->n_cells, "cell range [%lld,%lld) outside [0,%lld)", (long long)lo,
              (long long)hi, (long long)m->n_cells);
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  CSC_CUDA(ctx, cudaMemcpyAsync(b, m->cp() + lo, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaMemcpyAsync(e, m->cp() + hi, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CSC_OK;
}

extern "C" csc_status csc_sparse_range_nnz(const csc_sparse* m, int64_t cell_lo, int64_t cell_hi, int64_t* nnz) {
  if (!m || !nnz) return CSC_ERR_ARG;
  int64_t b, e;
  CSC_TRY(range_ptrs(m, cell_lo, cell_hi, &b, &e));
  *nnz = e - b;
  return CSC_OK;
}

extern "C" csc_status csc_sparse_download(const csc_sparse* m, int64_t cell_lo, int64_t cell_hi, int64_t* colptr_out,
                                          int64_t* rowval_out, double* nzval_out, int32_t index_base) {
  if (!m) return CSC_ERR_ARG;
  csc_ctx* ctx = m->ctx;
  CSC_GUARD_BEGIN
  int64_t b, e;
  CSC_TRY(range_ptrs(m, cell_lo, cell_hi, &b, &e));
  int64_t nc = cell_hi - cell_lo, nnz = e - b;
  int sm = ctx->sm_count;
  if (colptr_out) {
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, (size_t)(nc + 1) * 8, &tmp));
    k_colptr_rebase<int64_t><<<grid_for(nc + 1, sm), 256, 0, ctx->stream>>>(m->cp() + cell_lo, (int64_t*)tmp->p, nc + 1,
                                                                             b - index_base);
    CSC_LAUNCHED(ctx);
    CSC_TRY(d2h_copy(ctx, colptr_out, tmp->p, (size_t)(nc + 1) * 8));
  }
  const int64_t chunk = (int64_t)32 << 20;
  if ((rowval_out || nzval_out) && nnz > 0) {
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, (size_t)std::min(nnz, chunk) * 8, &tmp));
    for (int64_t o = 0; o < nnz && rowval_out; o += chunk) {
      int64_t cnt = std::min(chunk, nnz - o);
      k_widen_idx<<<grid_for(cnt, sm), 256, 0, ctx->stream>>>(m->rv() + b + o, (int64_t*)tmp->p, cnt, index_base);
      CSC_LAUNCHED(ctx);
      CSC_TRY(d2h_copy(ctx, rowval_out + o, tmp->p, (size_t)cnt * 8));
    }
    for (int64_t o = 0; o < nnz && nzval_out; o += chunk) {
      int64_t cnt = std::min(chunk, nnz - o);
      k_widen_val<<<grid_for(cnt, sm), 256, 0, ctx->stream>>>(m->v() + b + o, (double*)tmp->p, cnt);
      CSC_LAUNCHED(ctx);
      CSC_TRY(d2h_copy(ctx, nzval_out + o, tmp->p, (size_t)cnt * 8));
    }
  }
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------
// K1/K2  normalize_object (processing.jl:15-31) fused with the per-gene moments of the RAW counts
//   (sum x, sum x^2)  -> find_variable_genes (processing.jl:138-139)
// (The moments of the normalised values that scale_object needs are NOT taken here any more: round 1 accumulated
// them in 32-bit fixed point scaled by the theoretical range of y, which resolved y^2 to ~3e-5 only and to ~40 for
// norm_method = "none".  k_scale_stats_fx takes them for the selected rows with 64-bit accumulators.)
//
// A CTA owns a contiguous range of cells; a warp owns one cell at a time.  Pass 1 over the column: fp64 column
// sum (warp-shuffle reduction).  Pass 2 (the column is still in L1): y = log((x * inv) * sf + pc), written with
// the same pattern, and the moment updates.  All moment tables live in shared memory as 32-bit integers, the
// only type shared memory adds natively (ATOMS.ADD; 64-bit and float adds are CAS loops), and are flushed to the
// fp64 tables in global memory once per CTA:
//   S1 = sum x            u32, exact (x integer <= 65535; anything else goes straight to the fp64 table)
//   E  = sum (x^2 - x)    u32, exact (x <= 255)            -> sum x^2 = S1 + E
// Integer sums are order-independent, so the result is deterministic; the raw moments are exact.
// HBM traffic: 4 B value read + 4 B value written + 4 B row index per non-zero.
// What bounds the pass (tools/micro/norm_bench.cu, C2 shape): without the table updates it streams at 0.76 of the
// HBM rate; the 3.4 shared-memory atomics per entry run at ~7 lanes/clk/SM (with or without bank conflicts),
// 0.22 ms on their own, and go through the same load/store unit as the global accesses.  The compiler does not move
// a global load across an atomic, so a loop of load -> update -> load exposes one DRAM latency per entry group: the
// loads of UN entries per lane are therefore issued first (0.74 -> 0.52 ms), the next cell's lines are prefetched
// into L2, and log1p of arguments >= 1 uses the hardware logarithm (-> 0.46 ms).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// lanes 0-15: the 128-byte lines of val[b, e); lanes 16-31: those of rowval[b, e).  Only the HEAD of the column (the
// first PF_CAP entries) is prefetched: it hides the DRAM latency of a warp's first loads after a column switch, the
// rest of a long column streams behind it anyway.  Prefetching whole columns of a whole-transcriptome matrix (1650
// entries, 13 KB per column and array pair, x 3500-6000 resident warps) pushed the columns being worked on out of L2:
// ncu at C3 showed k_normalize reading 23.5 GB for 13.2 GB of input and k_select_fill 24.5 GB.
constexpr int64_t PF_CAP = 512;
__device__ __forceinline__ void prefetch_column_l2(const int32_t* rowval, const float* val, int64_t b, int64_t e, int lane) {
  const char* base = lane < 16 ? reinterpret_cast<const char*>(val) : reinterpret_cast<const char*>(rowval);
  if (base == nullptr) return;
  const int64_t hi = min(e, b + PF_CAP) * 4;
  for (int64_t off = ((b * 4) & ~(int64_t)127) + 128 * (lane & 15); off < hi; off += 128 * 16) prefetch_l2(base + off);
}

struct NormArgs {
  const int64_t* colptr;
  const int32_t* rowval;
  const float* raw;
  float* out;
  int64_t n_cells;
  int32_t n_genes;
  double sf, pc;
  int method, nan_to_zero;
  double* gstats;     // fp64 [2G] raw moments (or null)
  int64_t cells_per_cta;
  int32_t n_e;        // RAW = 1: genes [0, n_e) also have an E slot in shared memory (what is left of it beside the S1 table)
};

// y of one stored entry (processing.jl:24-27): t = x * scale, log(t + pc) or t
__device__ __forceinline__ float norm_value(float x, float scale, int method, bool pc_is_one, float pcf, int nan_to_zero) {
  const float t = x * scale;
  float y;
  if (method == CSC_NORM_LOGARITHM) {
    // log(1 + t): from t = 1 on, 1 + t is at least 2, where __logf is good to 3 ulp and the rounding of the
    // sum costs 2^-24 / log 2 relative (4.5e-7 together, against the 1e-5 the path promises); below, and
    // for any other pseudocount, the library functions
    y = pc_is_one ? (t >= 1.f ? __logf(1.f + t) : log1pf(t)) : logf(t + pcf);
  } else {
    y = t;
  }
  if (nan_to_zero && y != y) y = 0.f;
  return y;
}

// raw moments of one stored entry (x, gene g) into the CTA's tables / the fp64 table in global memory (RAW: see k_normalize).
// (A formulation with two predicated shared-memory adds and one rare branch was not faster: 6.0 vs 5.9 ms at C3.)
template <int RAW>
__device__ __forceinline__ void raw_moment_update(float x, int g, unsigned int* s1tab, unsigned int* etab, int n_e, double* gstats,
                                                  int G) {
  const float xr = rintf(x);
  if (RAW != 3 && xr == x && x >= 0.f && x <= 65535.f) {
    const unsigned int xi = (unsigned int)x;
    atomicAdd(&s1tab[g], xi);
    if ((RAW == 2 || (RAW == 1 && g < n_e)) && xi <= 255u) {
      if (xi >= 2u) atomicAdd(&etab[g], xi * xi - xi);
    } else if (xi >= 2u) {
      const double xd = (double)x;
      atomicAdd(&gstats[(int64_t)G + g], xd * xd - xd);
    }
  } else {
    const double xd = (double)x;
    atomicAdd(&gstats[g], xd);
    atomicAdd(&gstats[(int64_t)G + g], xd * xd);
  }
}

// RAW: 0 = no raw moments, 1 = S1 table + E slots for the first n_e genes (E of the others goes to global memory:
//      13.6 % of the C3 entries are counts >= 2, 225 M fp64 atomics per pass when no gene had a slot), 2 = S1 and E tables,
//      3 = no table fits: every update goes to the fp64 table in global memory
// (Measured dead end, round 2: a group of 128 or 256 threads holding a whole 1650-entry column in registers, one read of
// the column instead of two.  DRAM reads fell to the algorithmic 13.5 GB, but what a warp pays once per column -- bounds,
// fp64 reduction, named barrier, reciprocal -- is then paid by 4-8 warps per column and the few resident warps cannot
// hide the latencies: 8.2 ms with 256-thread groups, 15.7 ms with 128-thread groups, against 5.6 ms here.)
// BIG: the tables leave room for one CTA per SM only (whole-transcriptome matrices): 1024 threads and 8 entries per
// lane in flight instead of 768 x 2 CTAs and 4 -- with 24 warps per SM and 1 KB of loads per warp in flight the pass sat
// at 0.37 of the HBM rate with no unit busier than 48 % (ncu): latency, not throughput
template <bool WRITE_NORM, int RAW, bool BIG>
__global__ void __launch_bounds__(BIG ? 1024 : 768, BIG ? 1 : 2)
k_normalize(const NormArgs a) {
  extern __shared__ unsigned int stab[];
  const int G = a.n_genes;
  unsigned int* s1tab = stab;
  constexpr int NRAW = RAW == 2 ? 2 : RAW == 1 ? 1 : 0;
  unsigned int* etab = stab + (NRAW >= 1 ? G : 0);
  constexpr int NTAB = NRAW;
  const int n_e = RAW == 2 ? G : RAW == 1 ? a.n_e : 0;
  if (NTAB > 0) {
    for (int g = threadIdx.x; g < G + n_e; g += blockDim.x) stab[g] = 0u;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarp = blockDim.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * a.cells_per_cta;
  const int64_t c1 = min(c0 + a.cells_per_cta, a.n_cells);
  const bool pc_is_one = (a.pc == 1.0);
  const float pcf = (float)a.pc;
  constexpr int UN = BIG ? 8 : 4;   // entries per lane whose loads are issued before the first update (see the header comment)
  int64_t b = 0, e = 0, nb = 0, ne = 0;
  if (c0 + warp < c1) { b = a.colptr[c0 + warp]; e = a.colptr[c0 + warp + 1]; }
  if (c0 + warp + nwarp < c1) { nb = a.colptr[c0 + warp + nwarp]; ne = a.colptr[c0 + warp + nwarp + 1]; }
  for (int64_t c = c0 + warp; c < c1; c += nwarp) {
    // column bounds two cells ahead, the next cell's lines on their way into L2 while this one is processed
    int64_t nnb = 0, nne = 0;
    if (c + 2 * nwarp < c1) { nnb = a.colptr[c + 2 * nwarp]; nne = a.colptr[c + 2 * nwarp + 1]; }
    if (nb < ne) prefetch_column_l2(RAW > 0 ? a.rowval : nullptr, a.raw, nb, ne, lane);
    float scale = 0.f;
    if (WRITE_NORM) {
      double s = 0.0;
#pragma unroll (BIG ? 8 : 4)
      for (int64_t i = b + lane; i < e; i += 32) s += (double)a.raw[i];
      s = warp_sum(s);
      // (x * (1/s)) * sf of processing.jl:24 with the two cell constants folded in fp64 and applied in fp32: the
      // stored value is fp32 anyway, the result stays within 2 ulp of it.  1/0 = Inf -> 0 * Inf = NaN: the
      // reference's NaN column (processing.jl:18).
      scale = (float)((1.0 / s) * a.sf);
    }
    for (int64_t i0 = b + lane; i0 < e; i0 += 32 * UN) {
      float xs[UN];
      int gs[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int64_t i = i0 + 32 * u;
        xs[u] = i < e ? a.raw[i] : 0.f;
        gs[u] = (RAW > 0 && i < e) ? __ldcs(a.rowval + i) : 0;
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int64_t i = i0 + 32 * u;
        if (i >= e) break;
        const float x = xs[u];
        if (WRITE_NORM) __stcs(a.out + i, norm_value(x, scale, a.method, pc_is_one, pcf, a.nan_to_zero));
        if (RAW > 0) raw_moment_update<RAW>(x, gs[u], s1tab, etab, n_e, a.gstats, G);
      }
    }
    b = nb; e = ne; nb = nnb; ne = nne;
  }
  if (NTAB > 0) {
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      if (NRAW > 0) {
        const unsigned int v = s1tab[g];
        const unsigned int ex = g < n_e ? etab[g] : 0u;
        if (v) atomicAdd(&a.gstats[g], (double)v);
        if (v || ex) atomicAdd(&a.gstats[(int64_t)G + g], (double)v + (double)ex);  // sum x^2 = sum x + sum (x^2 - x)
      }
    }
  }
}

static csc_status launch_normalize(csc_ctx* ctx, const csc_sparse* raw, float* out, double sf, double pc, int method,
                                   int nan_to_zero, double* gstats) {
  if (raw->n_cells == 0) return CSC_OK;
  const size_t G = (size_t)raw->n_genes;
  const size_t budget = (size_t)200 * 1024;
  // which tables fit: raw S1 (+E)
  int raw_mode = 0;
  if (gstats) raw_mode = (2 * G * 4 <= budget) ? 2 : (G * 4 <= budget ? 1 : 3);
  const size_t n_e = raw_mode == 2 ? G : raw_mode == 1 ? std::min(G, (budget - G * 4) / 4) : 0;
  const size_t smem = raw_mode == 1 || raw_mode == 2 ? (G + n_e) * 4 : 0;
  const int ctas_per_sm = smem ? (int)std::max<size_t>(1, std::min<size_t>(2, ((size_t)220 * 1024) / (smem + 1024))) : 2;
  const bool big = ctas_per_sm == 1;
  const int threads = big ? 1024 : 768;
  int64_t grid = (int64_t)ctx->sm_count * ctas_per_sm;
  // a CTA's u32 tables must not overflow: <= 65535 (S1) and <= 65025 (E) per entry * 65536 cells
  int64_t cells_per_cta = std::max<int64_t>(16, ceil_div64(raw->n_cells, grid));
  cells_per_cta = std::min<int64_t>(cells_per_cta, 65536);
  grid = ceil_div64(raw->n_cells, cells_per_cta);
  NormArgs a;
  a.colptr = raw->cp();
  a.rowval = raw->rv();
  a.raw = raw->v();
  a.out = out;
  a.n_cells = raw->n_cells;
  a.n_genes = (int32_t)raw->n_genes;
  a.sf = sf;
  a.pc = pc;
  a.method = method;
  a.nan_to_zero = nan_to_zero;
  a.gstats = gstats;
  a.cells_per_cta = cells_per_cta;
  a.n_e = (int32_t)n_e;
  auto launch = [&](auto kern) -> csc_status {
    if (smem > 48 * 1024)
      CSC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, threads, smem, ctx->stream>>>(a);
    CSC_LAUNCHED(ctx);
    return CSC_OK;
  };
  const bool wn = out != nullptr;
  csc_status st = CSC_OK;
#define NORM_CASE(W, R)                                                      \
  if (wn == W && raw_mode == R) st = big ? launch(k_normalize<W, R, true>) : launch(k_normalize<W, R, false>);
  NORM_CASE(true, 0) NORM_CASE(true, 1) NORM_CASE(true, 2) NORM_CASE(true, 3)
  NORM_CASE(false, 1) NORM_CASE(false, 2) NORM_CASE(false, 3)
#undef NORM_CASE
  return st;
}

extern "C" csc_status csc_normalize(csc_ctx* ctx, const csc_sparse* raw, double scale_factor, int32_t norm_method,
                                    double pseudocount, int32_t nan_to_zero, csc_vec* gene_stats, csc_sparse** out) {
  if (!ctx || !raw || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  // processing.jl:28: throw(ArgumentError("Unknown normalization method"))
  CSC_REQUIRE(ctx, norm_method == CSC_NORM_LOGARITHM || norm_method == CSC_NORM_NONE,
              "Unknown normalization method: %d", norm_method);
  CSC_REQUIRE(ctx, !gene_stats || (gene_stats->dtype == CSC_DT_F64 && gene_stats->n >= 2 * raw->n_genes),
              "csc_normalize: gene_stats must be fp64[2*n_genes]");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_NORMALIZE);
  std::unique_ptr<csc_sparse> m(new csc_sparse());
  m->ctx = ctx;
  m->n_genes = raw->n_genes;
  m->n_cells = raw->n_cells;
  m->nnz = raw->nnz;
  m->colptr = raw->colptr;  // shared pattern: the output keeps the stored structure (processing.jl:24)
  m->rowval = raw->rowval;
  CSC_TRY(dev_alloc(ctx, (size_t)raw->nnz * 4, &m->val));
  CSC_TRY(launch_normalize(ctx, raw, m->v(), scale_factor, pseudocount, norm_method, nan_to_zero,
                           gene_stats ? gene_stats->as<double>() : nullptr));
  // bound on |y| (k_scale_stats_fx scales its fixed-point accumulators by it): t = x / colsum * sf lies in [0, sf] for
  // non-negative counts; entries outside the bound take the fp64 route there, so the bound only has to be usual
  m->vmax_bound = norm_method == CSC_NORM_LOGARITHM
                      ? std::max(std::fabs(std::log(pseudocount > 0 ? pseudocount : 1e-300)), std::fabs(std::log(scale_factor + pseudocount)))
                      : std::fabs(scale_factor);
  *out = m.release();
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_gene_stats_partial(csc_ctx* ctx, const csc_sparse* raw, csc_vec* gene_stats) {
  if (!ctx || !raw || !gene_stats) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, gene_stats->dtype == CSC_DT_F64 && gene_stats->n >= 2 * raw->n_genes,
              "csc_gene_stats_partial: gene_stats must be fp64[2*n_genes]");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_GENE_STATS);
  CSC_TRY(launch_normalize(ctx, raw, nullptr, 1.0, 1.0, 0, 0, gene_stats->as<double>()));
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_hvg_finish(csc_ctx* ctx, const csc_vec* gene_stats, int64_t n_genes, int64_t n_cells_total,
                                     int64_t n_features, double span, double* mean, double* variance,
                                     double* variance_expected, double* variance_standardized, int64_t* order) {
  if (!ctx || !gene_stats) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, gene_stats->dtype == CSC_DT_F64 && gene_stats->n >= 2 * n_genes && n_genes > 0,
              "csc_hvg_finish: gene_stats must be fp64[2*n_genes]");
  CSC_REQUIRE(ctx, n_cells_total >= 2, "csc_hvg_finish: need at least 2 cells");
  CSC_REQUIRE(ctx, span > 0.0 && span <= 1.0, "span must be in the range (0, 1].");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_HVG);
  if (hvg_device_applies(n_genes, span))
    return launch_hvg_device(ctx, gene_stats->as<double>(), n_genes, n_cells_total, span, mean, variance, variance_expected,
                             variance_standardized, order);
  std::vector<double> h((size_t)2 * n_genes);
  CSC_CUDA(ctx, cudaMemcpyAsync(h.data(), gene_stats->mem->p, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return launch_hvg_host(ctx, h.data(), n_genes, n_cells_total, n_features, span, mean, variance, variance_expected,
                         variance_standardized, order, false);
  CSC_GUARD_END(ctx)
}

// the same, called by EVERY rank of the context's communicator with the all-reduced table: the Loess vertex fits are
// dealt out over the ranks and their results all-reduced (identical outputs on every rank)
extern "C" csc_status csc_hvg_finish_collective(csc_ctx* ctx, const csc_vec* gene_stats, int64_t n_genes, int64_t n_cells_total,
                                                int64_t n_features, double span, double* mean, double* variance,
                                                double* variance_expected, double* variance_standardized, int64_t* order) {
  if (!ctx || !gene_stats) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, gene_stats->dtype == CSC_DT_F64 && gene_stats->n >= 2 * n_genes && n_genes > 0,
              "csc_hvg_finish_collective: gene_stats must be fp64[2*n_genes]");
  CSC_REQUIRE(ctx, n_cells_total >= 2, "csc_hvg_finish_collective: need at least 2 cells");
  CSC_REQUIRE(ctx, span > 0.0 && span <= 1.0, "span must be in the range (0, 1].");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_HVG);
  // on the device the whole finish is ~0.3 ms and deterministic: every rank runs it on its copy of the all-reduced
  // table, nothing is exchanged (the host version below deals its vertex fits out over the ranks)
  if (hvg_device_applies(n_genes, span))
    return launch_hvg_device(ctx, gene_stats->as<double>(), n_genes, n_cells_total, span, mean, variance, variance_expected,
                             variance_standardized, order);
  std::vector<double> h((size_t)2 * n_genes);
  CSC_CUDA(ctx, cudaMemcpyAsync(h.data(), gene_stats->mem->p, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return launch_hvg_host(ctx, h.data(), n_genes, n_cells_total, n_features, span, mean, variance, variance_expected,
                         variance_standardized, order, true);
  CSC_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------
// K4  scale_object (processing.jl:65-93)
// ------------------------------------------------------------------------------------------
// per-gene sum y / sum y^2 of the normalised values of the selected rows (processing.jl:66-67 takes mean and var of
// them).  Two kernels:
//
// k_scale_stats_fx -- the usual one.  64-bit fixed-point accumulators in SHARED memory, each as a (lo, hi) pair of u32
// words because shared memory adds 32-bit integers natively (a 64-bit or fp64 shared add is a CAS loop): the low word
// is added with atomicAdd, whose return value tells the adding thread whether ITS add wrapped the word, and only then
// (or when the addend itself is >= 2^32) the high word is touched -- two shared atomics per sum at the scales in use.
// Scales: the largest powers of two that keep a CTA's sums inside 61 bits (see csc_scale_stats_partial): 2^44 for y and
// 2^42 for y^2 with the logarithm method and the default scale factor, i.e. y^2 is resolved to 2e-13 -- fp64-grade, and
// 10^8 times finer than the 32-bit tables of round 1 (which ADVICE r1 found to round y^2 = 4 to 0 for norm_method
// "none").  Integer sums are order-independent -> deterministic.  Values outside the
// bound, NaN and Inf go straight to the fp64 table in global memory.  A CTA flushes its tables once, as fp64 adds.
//
// k_scale_stats -- exact fp64 sums in per-CTA private tables in global memory (atomics of different CTAs never meet
// at an address; k_scale_stats_reduce adds the tables up).  Used when no bound on |y| is known (an uploaded
// normalised matrix), when the bound leaves fewer than 20 fractional bits (norm_method "none" with a large scale
// factor), or when the tables of H rows do not fit in shared memory.
// Loads of four entries per lane are issued before the look-ups, the next column is prefetched into L2.
struct StatsFx {
  float bound;        // |y| <= bound -> fixed-point route
  float s1, s2;       // 2^e1, 2^e2
  double inv_s1, inv_s2;
};

__device__ __forceinline__ void fx_add(unsigned int* lo, unsigned int* hi, long long v) {
  const unsigned int l = (unsigned int)(unsigned long long)v;
  unsigned int h = (unsigned int)((unsigned long long)v >> 32);
  const unsigned int old = atomicAdd(lo, l);
  h += (old + l < old) ? 1u : 0u;     // the carry of THIS add (64-bit two's complement arithmetic, word by word)
  if (h) atomicAdd(hi, h);
}

template <bool IDENT>   // IDENT: the matrix is already the row subset (gene2h is the identity)
__global__ void __launch_bounds__(512, 2)
k_scale_stats_fx(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const float* __restrict__ val,
                 const int32_t* __restrict__ gene2h, int64_t n_cells, int32_t n_feat, const StatsFx fx,
                 double* __restrict__ part) {
  extern __shared__ unsigned int fxtab[];   // [4][n_feat]: y lo, y hi, y^2 lo, y^2 hi
  unsigned int* y1lo = fxtab;
  unsigned int* y1hi = fxtab + n_feat;
  unsigned int* y2lo = fxtab + 2 * n_feat;
  unsigned int* y2hi = fxtab + 3 * n_feat;
  for (int i = threadIdx.x; i < 4 * n_feat; i += blockDim.x) fxtab[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int UN = 4;
  int64_t b = 0, e = 0, nb = 0, ne = 0;
  if (warp < n_cells) { b = colptr[warp]; e = colptr[warp + 1]; }
  if (warp + nwarp < n_cells) { nb = colptr[warp + nwarp]; ne = colptr[warp + nwarp + 1]; }
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    int64_t nnb = 0, nne = 0;
    if (c + 2 * nwarp < n_cells) { nnb = colptr[c + 2 * nwarp]; nne = colptr[c + 2 * nwarp + 1]; }
    if (nb < ne) prefetch_column_l2(rowval, val, nb, ne, lane);
    for (int64_t i0 = b + lane; i0 < e; i0 += 32 * UN) {
      int g[UN];
      float x[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int64_t i = i0 + 32 * u;
        g[u] = i < e ? __ldcs(rowval + i) : -1;
        x[u] = i < e ? __ldcs(val + i) : 0.f;
      }
      int h[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) h[u] = IDENT ? g[u] : (g[u] >= 0 ? __ldg(gene2h + g[u]) : -1);
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (h[u] >= 0) {
          const float y = x[u];
          if (fabsf(y) <= fx.bound) {     // false for NaN / Inf / out-of-range values
            fx_add(&y1lo[h[u]], &y1hi[h[u]], __float2ll_rn(y * fx.s1));        // y * 2^e1 is exact in fp32
            fx_add(&y2lo[h[u]], &y2hi[h[u]], __double2ll_rn((double)y * (double)y * (double)fx.s2));
          } else {
            const double yd = (double)y;
            atomicAdd(&part[h[u]], yd);
            atomicAdd(&part[n_feat + h[u]], yd * yd);
          }
        }
      }
    }
    b = nb; e = ne; nb = nnb; ne = nne;
  }
  __syncthreads();
  for (int h = threadIdx.x; h < n_feat; h += blockDim.x) {
    const long long v1 = (long long)(((unsigned long long)y1hi[h] << 32) | y1lo[h]);
    const long long v2 = (long long)(((unsigned long long)y2hi[h] << 32) | y2lo[h]);
    if (v1) atomicAdd(&part[h], (double)v1 * fx.inv_s1);
    if (v2) atomicAdd(&part[n_feat + h], (double)v2 * fx.inv_s2);
  }
}

__global__ void __launch_bounds__(256)
k_scale_stats(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const float* __restrict__ val,
              const int32_t* __restrict__ gene2h, int64_t n_cells, int32_t n_feat, double* __restrict__ priv) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double* part = priv + (int64_t)blockIdx.x * 2 * n_feat;
  constexpr int UN = 4;
  int64_t b = 0, e = 0, nb = 0, ne = 0;
  if (warp < n_cells) { b = colptr[warp]; e = colptr[warp + 1]; }
  if (warp + nwarp < n_cells) { nb = colptr[warp + nwarp]; ne = colptr[warp + nwarp + 1]; }
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    int64_t nnb = 0, nne = 0;
    if (c + 2 * nwarp < n_cells) { nnb = colptr[c + 2 * nwarp]; nne = colptr[c + 2 * nwarp + 1]; }
    if (nb < ne) prefetch_column_l2(rowval, val, nb, ne, lane);
    for (int64_t i0 = b + lane; i0 < e; i0 += 32 * UN) {
      int g[UN];
      float x[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int64_t i = i0 + 32 * u;
        g[u] = i < e ? __ldcs(rowval + i) : -1;
        x[u] = i < e ? __ldcs(val + i) : 0.f;
      }
      int h[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) h[u] = g[u] >= 0 ? __ldg(gene2h + g[u]) : -1;
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (h[u] >= 0) {
          const double y = (double)x[u];
          atomicAdd(&part[h[u]], y);
          atomicAdd(&part[n_feat + h[u]], y * y);
        }
      }
    }
    b = nb; e = ne; nb = nnb; ne = nne;
  }
}

__global__ void k_scale_stats_reduce(const double* __restrict__ priv, int32_t n_tab, int32_t n_entries, double* __restrict__ part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  double s = 0.0;
  for (int t = 0; t < n_tab; ++t) s += priv[(int64_t)t * n_entries + i];
  part[i] += s;
}

// mean / 1/sd / fill value per selected gene from the (already all-reduced) sums
__global__ void k_scale_params(const double* __restrict__ part, int32_t n_feat, double n_cells_total, float scale_max,
                               int do_scale, int do_center, float* __restrict__ mu, float* __restrict__ inv_sd,
                               float* __restrict__ z0) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_feat) return;
  const double s1 = part[h], s2 = part[n_feat + h];
  const double mean = s1 / n_cells_total;
  const double var = (s2 - s1 * s1 / n_cells_total) / (n_cells_total - 1.0);
  const double sd = sqrt(var);  // sd == 0 -> division by zero below, NaN/Inf like the reference
  const float m = do_center ? (float)mean : 0.f;
  const float is = do_scale ? (float)(1.0 / sd) : 1.f;
  float z = (0.f - m) * is;
  z = (z != z) ? z : fminf(z, scale_max);
  mu[h] = m;
  inv_sd[h] = is;
  z0[h] = z;
}

// a warp builds one slice of a cell's dense row in shared memory (fill value, then scatter of the non-zeros) and
// writes it out with coalesced 16-byte stores: 4*H bytes written per cell, each once.  The row is cut into `nsplit`
// slices handled by different warp-tasks: the shared-memory footprint per warp shrinks accordingly, which is what
// limits the number of rows in flight per SM (8 KB per full row of 2000 genes); the cell's stored entries are read
// once per slice, out of L1/L2.
// The same row builder for the PCA Gram without a dense fp32 matrix in HBM: the finished slice leaves as the two fp16
// halves of the split-precision operand (hi = fp16(z), lo = fp16(z - hi); gram_tc.cu) with the row pitch the Gram
// kernel's tensor maps expect.  Column n_feat of every row is set to 1 and the rest of the padding to 0: the Gram of
// the extended matrix then carries the column sums Z'1 (which the centring needs) in its last column -- the tensor
// core sums them on the way, with the same split-precision accuracy as every other entry.
// Both are built around the memory latency that bounded their first version (ncu: long-scoreboard stalls 27 per
// issued instruction; 1.78 / 1.91 ms at C2): (1) the column bounds of a warp's next two tasks are loaded ahead, (2) the
// lines of the NEXT task's column (values and row indices) are prefetched into L2 while the current row is built,
// (3) the scatter loads a group of UN entries (index, value, then the gene2h look-ups) before the first dependent
// use, (4) the fill value goes to shared memory with 16-byte stores, (5) the grid is the number of CTAs that are
// actually resident (the tasks are divided statically over the grid): 1.31 / 1.44 ms, bit-identical results.
struct ScaleSrc {
  const int64_t* colptr;
  const int32_t* rowval;
  const float* val;
  const int32_t* gene2h;
  const float* mu;
  const float* inv_sd;
  const float* z0;
};

template <int UN>
__device__ __forceinline__ void scatter_scaled(const ScaleSrc& s, float* row, int h0, int h1, int64_t b, int64_t e, int lane,
                                               float scale_max, bool last_use) {
  for (int64_t i0 = b + lane; i0 < e; i0 += 32 * UN) {
    int g[UN];
    float x[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t i = i0 + 32 * u;
      const bool ok = i < e;
      // the column is read once per slice by the same warp: cached loads until the last slice, streaming there
      g[u] = ok ? (last_use ? __ldcs(s.rowval + i) : __ldg(s.rowval + i)) : -1;
      x[u] = ok ? (last_use ? __ldcs(s.val + i) : __ldg(s.val + i)) : 0.f;
    }
    int h[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) h[u] = g[u] >= 0 ? __ldg(s.gene2h + g[u]) : -1;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (h[u] >= h0 && h[u] < h1) {
        float z = (x[u] - __ldg(s.mu + h[u])) * __ldg(s.inv_sd + h[u]);
        z = (z != z) ? z : fminf(z, scale_max);  // min.(x, scale_max): upper clip only, NaN propagates
        row[h[u] - h0] = z;
      }
    }
  }
}

__global__ void __launch_bounds__(256, 4)
k_scale_fill(const ScaleSrc s, int64_t n_cells, int32_t n_feat, int32_t nsplit, int32_t slice, float scale_max,
              float* __restrict__ out) {
  extern __shared__ float srow[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float* row = srow + (size_t)wib * slice;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  // a warp owns a cell and builds its slices one after the other (the column comes from HBM once, then from L1/L2)
  int64_t b = 0, e = 0, nb = 0, ne = 0;
  if (warp < n_cells) { b = s.colptr[warp]; e = s.colptr[warp + 1]; }
  if (warp + nwarp < n_cells) { nb = s.colptr[warp + nwarp]; ne = s.colptr[warp + nwarp + 1]; }
  const bool vec_ok = (n_feat & 3) == 0 && (slice & 3) == 0;
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    int64_t nnb = 0, nne = 0;
    if (c + 2 * nwarp < n_cells) { nnb = s.colptr[c + 2 * nwarp]; nne = s.colptr[c + 2 * nwarp + 1]; }
    if (nb < ne) prefetch_column_l2(s.rowval, s.val, nb, ne, lane);
    for (int sl = 0; sl < nsplit; ++sl) {
      const int h0 = sl * slice;
      const int h1 = min(h0 + slice, n_feat);
      const int len = h1 - h0;
      if (vec_ok && (len & 3) == 0) {
        const float4* z4 = reinterpret_cast<const float4*>(s.z0 + h0);
        float4* r4 = reinterpret_cast<float4*>(row);
        for (int q = lane; q < (len >> 2); q += 32) r4[q] = __ldg(z4 + q);
      } else {
        for (int h = h0 + lane; h < h1; h += 32) row[h - h0] = __ldg(s.z0 + h);
      }
      __syncwarp();
      scatter_scaled<4>(s, row, h0, h1, b, e, lane, scale_max, sl == nsplit - 1);
      __syncwarp();
      float* dst = out + c * (int64_t)n_feat + h0;
      if (vec_ok && (len & 3) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int h = lane; h < (len >> 2); h += 32) __stcs(d4 + h, r4[h]);
      } else {
        for (int h = lane; h < len; h += 32) __stcs(dst + h, row[h]);
      }
      __syncwarp();
    }
    b = nb; e = ne; nb = nnb; ne = nne;
  }
}

template <int NV>   // groups of 8 genes per lane and slice: slice <= 256 * NV genes
__global__ void __launch_bounds__(256, 4)
k_scale_split(const ScaleSrc s, int64_t n_cells, int32_t n_feat, int32_t nsplit, int32_t slice, int32_t row_len,
               int32_t pitch, float scale_max, __half* __restrict__ hi16, __half* __restrict__ lo16) {
  extern __shared__ float srow[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  float* row = srow + (size_t)wib * row_len;            // row_len >= slice: the last slice carries the padding
  const int64_t warp = (int64_t)blockIdx.x * nw + wib;
  const int64_t nwarp = (int64_t)gridDim.x * nw;
  // a warp owns a cell and builds its slices one after the other (the column comes from HBM once, then from L1/L2)
  int64_t b = 0, e = 0, nb = 0, ne = 0;
  if (warp < n_cells) { b = s.colptr[warp]; e = s.colptr[warp + 1]; }
  if (warp + nwarp < n_cells) { nb = s.colptr[warp + nwarp]; ne = s.colptr[warp + nwarp + 1]; }
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    int64_t nnb = 0, nne = 0;
    if (c + 2 * nwarp < n_cells) { nnb = s.colptr[c + 2 * nwarp]; nne = s.colptr[c + 2 * nwarp + 1]; }
    if (nb < ne) prefetch_column_l2(s.rowval, s.val, nb, ne, lane);
    for (int sl = 0; sl < nsplit; ++sl) {
      const int h0 = sl * slice;                          // slice is a multiple of 8
      const int h1 = min(h0 + slice, n_feat);
      const int hw = (sl == nsplit - 1) ? pitch : h1;     // the last slice also writes the padding up to the row pitch
      const int len8 = (hw - h0) >> 3;
      {
        // n_feat is a multiple of 4 here (the caller falls back otherwise): whole float4 groups of z0, then the tail
        const int nfull = (h1 - h0) >> 2;
        const float4* z4 = reinterpret_cast<const float4*>(s.z0 + h0);
        float4* r4 = reinterpret_cast<float4*>(row);
        for (int q = lane; q < nfull; q += 32) r4[q] = __ldg(z4 + q);
        for (int h = h0 + 4 * nfull + lane; h < hw; h += 32) row[h - h0] = h < n_feat ? __ldg(s.z0 + h) : (h == n_feat ? 1.f : 0.f);
      }
      __syncwarp();
      scatter_scaled<4>(s, row, h0, h1, b, e, lane, scale_max, sl == nsplit - 1);
      __syncwarp();
      const float4* r4 = reinterpret_cast<const float4*>(row);
      uint4* dh = reinterpret_cast<uint4*>(hi16 + c * (int64_t)pitch + h0);
      uint4* dl = reinterpret_cast<uint4*>(lo16 + c * (int64_t)pitch + h0);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int q = j * 32 + lane;
        if (q < len8) {
          const float4 v0 = r4[2 * q], v1 = r4[2 * q + 1];
          const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const __half a0 = __float2half_rn(v[2 * u]), a1 = __float2half_rn(v[2 * u + 1]);
            __half2 p = __halves2half2(a0, a1);
            ph[u] = *reinterpret_cast<uint32_t*>(&p);
            p = __floats2half2_rn(v[2 * u] - __half2float(a0), v[2 * u + 1] - __half2float(a1));
            pl[u] = *reinterpret_cast<uint32_t*>(&p);
          }
          __stcs(dh + q, make_uint4(ph[0], ph[1], ph[2], ph[3]));
          __stcs(dl + q, make_uint4(pl[0], pl[1], pl[2], pl[3]));
        }
      }
      __syncwarp();
    }
    b = nb; e = ne; nb = nnb; ne = nne;
  }
}

// features (any order, 0-based) -> gene2h lookup in ORIGINAL gene order (utils.jl:72-76)
static csc_status build_gene2h(csc_ctx* ctx, int64_t n_genes, const int64_t* features, int64_t n_feat, DevPtr* d_map,
                               int64_t* n_rows_out) {
  std::vector<int32_t> map((size_t)n_genes, -1);
  int64_t H = 0;
  if (!features) {
    for (int64_t g = 0; g < n_genes; ++g) map[g] = (int32_t)g;
    H = n_genes;
  } else {
    std::vector<char> mask((size_t)n_genes, 0);
    for (int64_t i = 0; i < n_feat; ++i) {
      CSC_REQUIRE(ctx, features[i] >= 0 && features[i] < n_genes, "feature id %lld outside [0,%lld)",
                  (long long)features[i], (long long)n_genes);
      mask[features[i]] = 1;
    }
    for (int64_t g = 0; g < n_genes; ++g)
      if (mask[g]) map[g] = (int32_t)H++;
  }
  CSC_REQUIRE(ctx, H > 0, "no features selected");
  CSC_TRY(dev_alloc(ctx, (size_t)n_genes * 4, d_map));
  CSC_CUDA(ctx, cudaMemcpyAsync((*d_map)->p, map.data(), (size_t)n_genes * 4, cudaMemcpyHostToDevice, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *n_rows_out = H;
  return CSC_OK;
}

// ------------------------------------------------------------------------------------------
// subset_count (src/scrna/utils.jl:59-87) on the device: the rows `features` of a CSC matrix, in ORIGINAL gene
// order, renumbered 0..H-1.  A whole-transcriptome matrix keeps ~6 % of its rows (2000 of 33 k genes); the scale,
// fill and transform passes then read the subset instead of looking every stored entry up again.
// A warp owns a cell: count pass, exclusive scan, then an order-preserving compaction by warp votes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_select_count(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const int32_t* __restrict__ gene2h,
               int64_t n_cells, int64_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    const int64_t b = colptr[c], e = colptr[c + 1];
    int n = 0;
#pragma unroll 8
    for (int64_t i = b + lane; i < e; i += 32) n += __ldg(gene2h + __ldcs(rowval + i)) >= 0 ? 1 : 0;
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) counts[c] = n;
  }
}

__global__ void __launch_bounds__(256)
k_select_fill(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const float* __restrict__ val,
              const int32_t* __restrict__ gene2h, int64_t n_cells, const int64_t* __restrict__ out_colptr,
              int32_t* __restrict__ out_rowval, float* __restrict__ out_val) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int UN = 4;   // groups of 32 entries whose loads (index, value, look-up) are in flight together
  int64_t b = 0, e = 0, nb = 0, ne = 0;
  if (warp < n_cells) { b = colptr[warp]; e = colptr[warp + 1]; }
  if (warp + nwarp < n_cells) { nb = colptr[warp + nwarp]; ne = colptr[warp + nwarp + 1]; }
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    int64_t nnb = 0, nne = 0;
    if (c + 2 * nwarp < n_cells) { nnb = colptr[c + 2 * nwarp]; nne = colptr[c + 2 * nwarp + 1]; }
    if (nb < ne) prefetch_column_l2(rowval, nullptr, nb, ne, lane);
    int64_t o = out_colptr[c];
    for (int64_t i0 = b; i0 < e; i0 += 32 * UN) {
      int g[UN], h[UN];
      float x[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int64_t i = i0 + 32 * u + lane;
        g[u] = i < e ? __ldcs(rowval + i) : -1;
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) h[u] = g[u] >= 0 ? __ldg(gene2h + g[u]) : -1;
      // the value is fetched for kept entries only: with 6 % of the rows kept ~40 % of the 32-byte sectors of val are
      // touched instead of all of them
#pragma unroll
      for (int u = 0; u < UN; ++u) x[u] = h[u] >= 0 ? __ldcs(val + i0 + 32 * u + lane) : 0.f;
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const unsigned m = __ballot_sync(FULL, h[u] >= 0);
        if (h[u] >= 0) {
          const int64_t p = o + __popc(m & ((1u << lane) - 1u));
          out_rowval[p] = h[u];
          out_val[p] = x[u];
        }
        o += __popc(m);
      }
    }
    b = nb; e = ne; nb = nnb; ne = nne;
  }
}

extern "C" csc_status csc_sparse_select_rows(csc_ctx* ctx, const csc_sparse* m, const int64_t* features, int64_t n_feat,
                                             csc_sparse** out) {
  if (!ctx || !m || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_SCALE_STATS);
  DevPtr map, counts;
  int64_t H = 0;
  CSC_TRY(build_gene2h(ctx, m->n_genes, features, n_feat, &map, &H));
  std::unique_ptr<csc_sparse> r(new csc_sparse());
  r->ctx = ctx;
  r->n_genes = H;
  r->n_cells = m->n_cells;
  r->nnz = 0;
  r->vmax_bound = m->vmax_bound;
  CSC_TRY(dev_alloc(ctx, (size_t)(m->n_cells + 1) * 8, &counts, true));
  CSC_TRY(dev_alloc(ctx, (size_t)(m->n_cells + 1) * 8, &r->colptr));
  if (m->n_cells > 0) {
    k_select_count<<<ctx->sm_count * resident_ctas(k_select_count, 256, 0, 8), 256, 0, ctx->stream>>>(m->cp(), m->rv(), (const int32_t*)map->p, m->n_cells,
                                                                (int64_t*)counts->p);
    CSC_LAUNCHED(ctx);
  }
  {
    size_t tb = 0;
    CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int64_t*)counts->p, (int64_t*)r->colptr->p,
                                                (int)(m->n_cells + 1), ctx->stream));
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, tb, &tmp));
    CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp->p, tb, (const int64_t*)counts->p, (int64_t*)r->colptr->p,
                                                (int)(m->n_cells + 1), ctx->stream));
    ctx->launches += 2;
    int64_t nnz = 0;
    CSC_CUDA(ctx, cudaMemcpyAsync(&nnz, (const int64_t*)r->colptr->p + m->n_cells, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    r->nnz = nnz;
  }
  CSC_TRY(dev_alloc(ctx, (size_t)r->nnz * 4, &r->rowval));
  CSC_TRY(dev_alloc(ctx, (size_t)r->nnz * 4, &r->val));
  if (r->nnz > 0) {
    k_select_fill<<<ctx->sm_count * resident_ctas(k_select_fill, 256, 0, 8), 256, 0, ctx->stream>>>(m->cp(), m->rv(), m->v(), (const int32_t*)map->p, m->n_cells,
                                                               r->cp(), (int32_t*)r->rowval->p, r->v());
    CSC_LAUNCHED(ctx);
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  *out = r.release();
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_scale_stats_partial(csc_ctx* ctx, const csc_sparse* norm, const int64_t* features,
                                              int64_t n_feat, csc_vec* partial) {
  if (!ctx || !norm || !partial) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_SCALE_STATS);
  DevPtr map;
  int64_t H = 0;
  CSC_TRY(build_gene2h(ctx, norm->n_genes, features, n_feat, &map, &H));
  CSC_REQUIRE(ctx, partial->dtype == CSC_DT_F64 && partial->n >= 2 * H, "csc_scale_stats_partial: partial must be fp64[2*H], H=%lld",
              (long long)H);
  if (norm->n_cells > 0 && H > 0) {
    // fixed-point route: a bound on |y| is known, four u32 tables of H rows fit, and the 64-bit accumulators leave
    // enough fractional bits: the scales are the largest powers of two for which a CTA's sums stay below 2^61
    // (|y| 2^e1 cells < 2^61, y^2 2^e2 cells < 2^61), capped at 2^-44 / 2^-52 resolution.  Logarithm method, default
    // scale factor, 3400 cells per CTA: e1 = 44, e2 = 42 (y^2 resolved to 2e-13); "none" with scale factor 1e4 leaves
    // e2 = 22 only and takes the fp64 route.
    const double bound = norm->vmax_bound * 1.0001;
    const bool bounded = std::isfinite(bound) && bound > 0.0;
    const size_t smem = (size_t)4 * H * 4;
    const char* impl = getenv("CSC_SCALE_STATS_IMPL");
    const bool force_exact = impl && strcmp(impl, "fp64") == 0;
    const bool ident = features == nullptr;
    auto kern = ident ? k_scale_stats_fx<true> : k_scale_stats_fx<false>;
    int grid_fx = 1, e1 = -1, e2 = -1;
    if (bounded && smem <= (size_t)96 * 1024 && !force_exact) {
      if (smem > 48 * 1024) CSC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      grid_fx = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * resident_ctas(kern, 512, smem, 2),
                                                           ceil_div64(norm->n_cells, 16)));
      const int cell_bits = (int)std::ceil(std::log2((double)ceil_div64(norm->n_cells, grid_fx) + 16.0));
      e1 = std::min(44, 61 - (int)std::ceil(std::log2(bound)) - cell_bits);
      e2 = std::min(52, 61 - (int)std::ceil(std::log2(bound * bound)) - cell_bits);
    }
    if (e1 >= 28 && e2 >= 32) {
      StatsFx fx;
      fx.bound = (float)bound;
      fx.s1 = (float)std::ldexp(1.0, e1);
      fx.s2 = (float)std::ldexp(1.0, e2);
      fx.inv_s1 = std::ldexp(1.0, -e1);
      fx.inv_s2 = std::ldexp(1.0, -e2);
      const int grid = grid_fx;
      kern<<<grid, 512, smem, ctx->stream>>>(norm->cp(), norm->rv(), norm->v(), (const int32_t*)map->p, norm->n_cells, (int32_t)H,
                                             fx, partial->as<double>());
      CSC_LAUNCHED(ctx);
    } else {
      int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * resident_ctas(k_scale_stats, 256, 0, 4),
                                        ceil_div64(norm->n_cells, 8));
      // the private tables stay below 256 MB (a whole-transcriptome call with every gene selected)
      grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid, ((int64_t)256 << 20) / (2 * H * 8)));
      DevPtr priv;
      CSC_TRY(dev_alloc(ctx, (size_t)grid * 2 * H * 8, &priv));
      CSC_CUDA(ctx, cudaMemsetAsync(priv->p, 0, (size_t)grid * 2 * H * 8, ctx->stream));
      k_scale_stats<<<grid, 256, 0, ctx->stream>>>(norm->cp(), norm->rv(), norm->v(), (const int32_t*)map->p, norm->n_cells,
                                                   (int32_t)H, (double*)priv->p);
      CSC_LAUNCHED(ctx);
      k_scale_stats_reduce<<<(unsigned)ceil_div64(2 * H, 256), 256, 0, ctx->stream>>>((const double*)priv->p, grid,
                                                                                     (int32_t)(2 * H), partial->as<double>());
      CSC_LAUNCHED(ctx);
    }
  }
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

extern "C" csc_status csc_scale_fill(csc_ctx* ctx, const csc_sparse* norm, const int64_t* features, int64_t n_feat,
                                     const csc_vec* partial, int64_t n_cells_total, double scale_max, int32_t do_scale,
                                     int32_t do_center, csc_dense** out) {
  if (!ctx || !norm || !partial || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_SCALE_FILL);
  DevPtr map;
  int64_t H = 0;
  CSC_TRY(build_gene2h(ctx, norm->n_genes, features, n_feat, &map, &H));
  CSC_REQUIRE(ctx, partial->dtype == CSC_DT_F64 && partial->n >= 2 * H, "csc_scale_fill: partial must be fp64[2*H]");
  CSC_REQUIRE(ctx, n_cells_total >= 2, "csc_scale_fill: need at least 2 cells");
  DevPtr params;
  CSC_TRY(dev_alloc(ctx, (size_t)H * 4 * 3, &params));
  float* mu = (float*)params->p;
  float* inv_sd = mu + H;
  float* z0 = inv_sd + H;
  k_scale_params<<<(unsigned)ceil_div64(H, 256), 256, 0, ctx->stream>>>(partial->as<double>(), (int32_t)H,
                                                                        (double)n_cells_total, (float)scale_max, do_scale,
                                                                        do_center, mu, inv_sd, z0);
  CSC_LAUNCHED(ctx);
  csc_dense* d = nullptr;
  CSC_TRY(csc_dense_alloc(ctx, norm->n_cells, H, &d));
  if (norm->n_cells > 0) {
    // slices of at most 1024 genes (multiple of 4): 4 KB per warp, 8 warps per CTA, up to 7 CTAs per SM
    int nsplit = (int)ceil_div64(H, 1024);
    int slice = (int)(ceil_div64(ceil_div64(H, nsplit), 4) * 4);
    nsplit = (int)ceil_div64(H, slice);
    const int warps = 8;
    size_t smem = (size_t)warps * slice * 4;
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_scale_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      const int ctas_per_sm = resident_ctas(k_scale_fill, warps * 32, smem,
                                            (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)220 * 1024) / (smem + 1024))));
      const int64_t grid = std::min<int64_t>((int64_t)ctx->sm_count * ctas_per_sm, ceil_div64(norm->n_cells, warps));
      const ScaleSrc src{norm->cp(), norm->rv(), norm->v(), (const int32_t*)map->p, mu, inv_sd, z0};
      k_scale_fill<<<(unsigned)grid, warps * 32, smem, ctx->stream>>>(src, norm->n_cells, (int32_t)H, nsplit, slice,
                                                                      (float)scale_max, d->p());
      ctx->launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      csc_dense_free(d);
      return csc_fail(ctx, CSC_ERR_CUDA, "csc_scale_fill: %s", cudaGetErrorString(e));
    }
  }
  *out = d;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

int gram_split_pitch(int H);
csc_status gram_tc_split(csc_ctx* ctx, const __half* hi16, const __half* lo16, int64_t n_rows, int H, int pitch, double* gram,
                         double* colsum_from_ones);

// a3 + a4 fused for callers that only need the PCA of the scaled matrix: colsum[H] += column sums of Z, gram[H*H] += Z'Z
// without a dense Z in HBM (csc_scale_fill followed by csc_pca_partial writes 4 H bytes per cell, reads them back and
// writes the split operands; here the split operands are written once).  Falls back to those two calls when H is not
// a multiple of 4 or exceeds what one CTA's column table holds.
extern "C" csc_status csc_scale_gram_partial(csc_ctx* ctx, const csc_sparse* norm, const int64_t* features, int64_t n_feat,
                                             const csc_vec* partial, int64_t n_cells_total, double scale_max, int32_t do_scale,
                                             int32_t do_center, csc_vec* colsum, csc_vec* gram, csc_split** operands) {
  if (!ctx || !norm || !partial || !colsum || !gram) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  if (operands) *operands = nullptr;
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  DevPtr map;
  int64_t H = 0;
  CSC_TRY(build_gene2h(ctx, norm->n_genes, features, n_feat, &map, &H));
  CSC_REQUIRE(ctx, partial->dtype == CSC_DT_F64 && partial->n >= 2 * H, "csc_scale_gram_partial: partial must be fp64[2*H]");
  CSC_REQUIRE(ctx, n_cells_total >= 2, "csc_scale_gram_partial: need at least 2 cells");
  CSC_REQUIRE(ctx, colsum->dtype == CSC_DT_F64 && colsum->n >= H, "csc_scale_gram_partial: colsum must be fp64[H]");
  CSC_REQUIRE(ctx, gram->dtype == CSC_DT_F64 && gram->n >= H * H, "csc_scale_gram_partial: gram must be fp64[H*H]");
  const char* impl = getenv("CSC_GRAM_IMPL");
  if ((H & 3) != 0 || H > 8192 || (impl && strcmp(impl, "simt") == 0)) {
    csc_dense* z = nullptr;
    CSC_TRY(csc_scale_fill(ctx, norm, features, n_feat, partial, n_cells_total, scale_max, do_scale, do_center, &z));
    const csc_status rc = csc_pca_partial(ctx, z, colsum, gram);
    csc_dense_free(z);
    return rc;
  }
  if (norm->n_cells == 0) return CSC_OK;
  DevPtr params, hi16, lo16;
  const int pitch = gram_split_pitch((int)H + 1);      // room for the column of ones
  {
    StageTimer st(ctx, CSC_T_SCALE_FILL);
    CSC_TRY(dev_alloc(ctx, (size_t)H * 4 * 3, &params));
    float* mu = (float*)params->p;
    float* inv_sd = mu + H;
    float* z0 = inv_sd + H;
    k_scale_params<<<(unsigned)ceil_div64(H, 256), 256, 0, ctx->stream>>>(partial->as<double>(), (int32_t)H,
                                                                          (double)n_cells_total, (float)scale_max, do_scale,
                                                                          do_center, mu, inv_sd, z0);
    CSC_LAUNCHED(ctx);
    CSC_TRY(dev_alloc(ctx, (size_t)norm->n_cells * pitch * 2, &hi16));
    CSC_TRY(dev_alloc(ctx, (size_t)norm->n_cells * pitch * 2, &lo16));
    int nsplit = (int)ceil_div64(H, 1024);
    int slice = (int)(ceil_div64(ceil_div64(H, nsplit), 8) * 8);
    nsplit = (int)ceil_div64(H, slice);
    const int last = pitch - (nsplit - 1) * slice;      // the last slice runs to the row pitch
    const int srow_len = std::max(slice, last);
    const int warps = 8;
    const size_t smem = (size_t)warps * srow_len * 4;
    const int ctas_guess = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)220 * 1024) / (smem + 1024)));
    const int64_t n_wtasks = ceil_div64(norm->n_cells, warps);
    const ScaleSrc src{norm->cp(), norm->rv(), norm->v(), (const int32_t*)map->p, mu, inv_sd, z0};
#define CSC_SPLIT_LAUNCH(NV)                                                                                             \
  do {                                                                                                                   \
    CSC_CUDA(ctx, cudaFuncSetAttribute(k_scale_split<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    const int64_t grid = std::min<int64_t>(                                                                              \
        (int64_t)ctx->sm_count * resident_ctas(k_scale_split<NV>, warps * 32, smem, ctas_guess), n_wtasks);               \
    k_scale_split<NV><<<(unsigned)grid, warps * 32, smem, ctx->stream>>>(src, norm->n_cells, (int32_t)H, nsplit, slice,   \
                                                                         srow_len, pitch, (float)scale_max,              \
                                                                         (__half*)hi16->p, (__half*)lo16->p);             \
  } while (0)
    if (srow_len <= 256) CSC_SPLIT_LAUNCH(1);
    else if (srow_len <= 512) CSC_SPLIT_LAUNCH(2);
    else if (srow_len <= 1024) CSC_SPLIT_LAUNCH(4);
    else CSC_SPLIT_LAUNCH(5);
#undef CSC_SPLIT_LAUNCH
    CSC_LAUNCHED(ctx);
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  StageTimer st2(ctx, CSC_T_PCA_GRAM);
  CSC_TRY(gram_tc_split(ctx, (const __half*)hi16->p, (const __half*)lo16->p, norm->n_cells, (int)H, pitch, gram->as<double>(),
                        colsum->as<double>()));
  if (operands) {
    // the caller keeps the split operands: csc_pca_transform_split takes the embedding from them as a tensor-core GEMM
    std::unique_ptr<csc_split> sp(new csc_split());
    sp->ctx = ctx;
    sp->n_rows = norm->n_cells;
    sp->n_feat = H;
    sp->pitch = pitch;
    sp->hi = hi16;
    sp->lo = lo16;
    *operands = sp.release();
  }
  return CSC_OK;
  CSC_GUARD_END(ctx)
}

// ------------------------------------------------------------------------------------------
// PCA embedding straight from the sparse normalised matrix (run_pca's transform, processing.jl:178-179).
// The scaled matrix is sparse-plus-rank-one, z = s + 1 z0' with s supported on the stored entries of the
// selected rows, so
//     (Z - 1 m') U = S U + 1 ((z0 - m)' U):
// ~100 multiply-adds per component and cell instead of H = 2000, and 8 B per stored entry read instead of 4 H
// bytes per cell.  A warp owns a cell; lanes own components (lane, lane + 32, ...); the stored entries are
// loaded 32 at a time (coalesced) and broadcast by shuffle; the loading rows come out of L1/L2.
// (Measured dead end: a 1024-thread CTA per SM with a 1000-row slice of U and the scaling parameters in shared
// memory, selected entries compacted into a per-warp staging area, loads pipelined two batches ahead and the next
// chunk prefetched into L2 -- 1.45 to 1.85 ms against 1.36 ms for this kernel: with 8 warps per scheduler the
// per-batch ballot / compaction / dependent shared-memory chain costs more than the L2 gathers it removes.  Nor is
// this kernel latency-bound: prefetching the next column, pipelining the loads and four loading rows in flight per
// round left it at 1.6 ms.  It moves 50 M selected entries x 256 B of U = 12.8 GB out of L1/L2 in 1.36 ms, i.e. it
// sits on the L2 gather bandwidth.)
// ------------------------------------------------------------------------------------------
template <int NJ>   // kpad = 32 * NJ components
__global__ void __launch_bounds__(256)
k_transform_sparse(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, const float* __restrict__ val,
                   const int32_t* __restrict__ gene2h, const float* __restrict__ mu, const float* __restrict__ inv_sd,
                   const float* __restrict__ z0, const float* __restrict__ U /*[H x kpad]*/, const float* __restrict__ base,
                   int64_t n_cells, int32_t kpad, int32_t k, float scale_max, float* __restrict__ out) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp; c < n_cells; c += nwarp) {
    float acc[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[j] = 0.f;
    const int64_t b = colptr[c], e = colptr[c + 1];
    for (int64_t i0 = b; i0 < e; i0 += 32) {
      const int64_t i = i0 + lane;
      int h = -1;
      float sv = 0.f;
      if (i < e) {
        h = __ldg(gene2h + __ldcs(rowval + i));
        if (h >= 0) {
          float z = (__ldcs(val + i) - __ldg(mu + h)) * __ldg(inv_sd + h);
          z = (z != z) ? z : fminf(z, scale_max);     // min.(x, scale_max): upper clip only, NaN propagates
          sv = z - __ldg(z0 + h);
        }
      }
      unsigned m = __ballot_sync(FULL, h >= 0);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int hj = __shfl_sync(FULL, h, src);
        const float sj = __shfl_sync(FULL, sv, src);
        const float* ur = U + (int64_t)hj * kpad + lane;
#pragma unroll
        for (int j = 0; j < NJ; ++j)      // lanes past k skip the load: the kernel is bound by these L2 sectors
          if (lane + 32 * j < k) acc[j] = fmaf(sj, __ldg(ur + 32 * j), acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int comp = lane + 32 * j;
      if (comp < k) out[c * k + comp] = acc[j] + base[comp];
    }
  }
}

// base[c] = sum_h z0[h] U[h][c] - (m'U)[c]   (one warp per component)
__global__ void k_transform_base(const float* __restrict__ z0, const float* __restrict__ U, const float* __restrict__ pm, int32_t H,
                                 int32_t kpad, float* __restrict__ base) {
  const int c = blockIdx.x;
  double s = 0.0;
  for (int h = threadIdx.x; h < H; h += 32) s += (double)z0[h] * (double)U[(int64_t)h * kpad + c];
  s = warp_sum(s);
  if (threadIdx.x == 0) base[c] = (float)(s - (double)pm[c]);
}

extern "C" csc_status csc_pca_transform_sparse(csc_ctx* ctx, const csc_sparse* norm, const int64_t* features, int64_t n_feat,
                                               const csc_vec* partial, int64_t n_cells_total, double scale_max, int32_t do_scale,
                                               int32_t do_center, const csc_pca* p, csc_dense** out) {
  if (!ctx || !norm || !partial || !p || !out) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_PCA_TRANSFORM);
  DevPtr map;
  int64_t H = 0;
  CSC_TRY(build_gene2h(ctx, norm->n_genes, features, n_feat, &map, &H));
  CSC_REQUIRE(ctx, H == p->n_feat, "csc_pca_transform_sparse: %lld selected rows, model has %lld features", (long long)H,
              (long long)p->n_feat);
  CSC_REQUIRE(ctx, partial->dtype == CSC_DT_F64 && partial->n >= 2 * H, "csc_pca_transform_sparse: partial must be fp64[2*H]");
  CSC_REQUIRE(ctx, p->kpad <= 128, "csc_pca_transform_sparse: more than 128 components is not supported");
  CSC_REQUIRE(ctx, n_cells_total >= 2, "csc_pca_transform_sparse: need at least 2 cells");
  DevPtr params, base;
  CSC_TRY(dev_alloc(ctx, (size_t)H * 4 * 3, &params));
  float* mu = (float*)params->p;
  float* inv_sd = mu + H;
  float* z0 = inv_sd + H;
  k_scale_params<<<(unsigned)ceil_div64(H, 256), 256, 0, ctx->stream>>>(partial->as<double>(), (int32_t)H, (double)n_cells_total,
                                                                        (float)scale_max, do_scale, do_center, mu, inv_sd, z0);
  CSC_LAUNCHED(ctx);
  // the loadings are stored with leading dimension kpad = k rounded up to 16; the kernel wants whole warps of components
  const int kp32 = ((p->k + 31) / 32) * 32;
  DevPtr u32;
  const float* U = (const float*)p->d_loadings_f32->p;
  int ldu = p->kpad;
  if (kp32 != p->kpad) {
    CSC_TRY(dev_alloc(ctx, (size_t)H * kp32 * 4, &u32));
    CSC_CUDA(ctx, cudaMemsetAsync(u32->p, 0, (size_t)H * kp32 * 4, ctx->stream));
    CSC_CUDA(ctx, cudaMemcpy2DAsync(u32->p, (size_t)kp32 * 4, p->d_loadings_f32->p, (size_t)p->kpad * 4, (size_t)p->kpad * 4, (size_t)H,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    U = (const float*)u32->p;
    ldu = kp32;
  }
  CSC_TRY(dev_alloc(ctx, (size_t)kp32 * 4, &base));
  CSC_CUDA(ctx, cudaMemsetAsync(base->p, 0, (size_t)kp32 * 4, ctx->stream));
  k_transform_base<<<p->k, 32, 0, ctx->stream>>>(z0, U, (const float*)p->d_proj_mean->p, (int32_t)H, ldu, (float*)base->p);
  CSC_LAUNCHED(ctx);
  csc_dense* e = nullptr;
  CSC_TRY(csc_dense_alloc(ctx, norm->n_cells, p->k, &e));
  if (norm->n_cells > 0) {
    const unsigned grid = (unsigned)std::min<int64_t>((int64_t)ctx->sm_count * 8, ceil_div64(norm->n_cells, 8));
#define TS_CASE(NJ)                                                                                                        \
  case NJ:                                                                                                                 \
    k_transform_sparse<NJ><<<grid, 256, 0, ctx->stream>>>(norm->cp(), norm->rv(), norm->v(), (const int32_t*)map->p, mu, inv_sd, \
                                                          z0, U, (const float*)base->p, norm->n_cells, ldu, p->k,          \
                                                          (float)scale_max, e->p());                                      \
    break;
    switch (kp32 / 32) {
      TS_CASE(1) TS_CASE(2) TS_CASE(3) TS_CASE(4)
      default: break;
    }
#undef TS_CASE
    ctx->launches++;
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    if (err != cudaSuccess) {
      csc_dense_free(e);
      return csc_fail(ctx, CSC_ERR_CUDA, "csc_pca_transform_sparse: %s", cudaGetErrorString(err));
    }
  }
  *out = e;
  return CSC_OK;
  CSC_GUARD_END(ctx)
}
