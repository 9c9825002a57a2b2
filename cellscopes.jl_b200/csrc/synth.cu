// Synthetic module-structured Poisson counts generated directly on the device: the
// benchmark matrices (500k x 5k, 1M x 33k ...) are too large to build on the host within the
// benchmark's time budget.  Same model and the same counter-based hash as
// cellscopes.jl_b200/synth.py (sample_counts is the NumPy twin):
//   log lambda_gc = a_g + l_c + s_g h_{c,m(g)} - s_g^2/2 ,   x_gc ~ Poisson(lambda_gc)
// Two passes over the (gene, cell) grid, a warp per cell: count the non-zeros, scan, then
// regenerate and write them in gene order (the hash makes the second pass identical to the first).
#include <cub/cub.cuh>

#include "common.cuh"

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t synth_hash3(uint64_t seed, uint64_t a, uint64_t b) {
  uint64_t h = splitmix64(seed ^ 0xD1B54A32D192ED03ull);
  h = splitmix64(h ^ a);
  h = splitmix64(h ^ (b * 0x2545F4914F6CDD1Dull));
  return h;
}
__device__ __forceinline__ double synth_u01(uint64_t h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ float synth_gauss(uint64_t h) {
  const double u1 = ((double)(h >> 32) + 0.5) * (1.0 / 4294967296.0);
  const double u2 = ((double)(h & 0xFFFFFFFFull) + 0.5) * (1.0 / 4294967296.0);
  return (float)(sqrt(-2.0 * log(u1)) * cospi(2.0 * u2));
}

// Poisson by CDF inversion (counts x = #{k : u > CDF(k)}), fp64 so that tiny rates keep their resolution
__device__ __forceinline__ int synth_poisson(double lam, double u) {
  double p = exp(-lam);
  if (u <= p) return 0;
  double cdf = p;
  int k = 0;
  while (u > cdf && k < 100000) {
    ++k;
    p = p * lam / (double)k;
    cdf += p;
    if (p <= 0.0) break;
  }
  return k;
}

struct SynthArgs {
  int64_t n_genes, cell_lo, cell_hi;
  uint64_t seed;
  const float* gene_base;
  const int32_t* gene_module;
  const float* gene_loading;
  const float* type_mean;
  int32_t n_types, n_modules;
  float lib_sd, type_w, noise_w;
};

template <bool WRITE>
__global__ void __launch_bounds__(256)
k_synth(SynthArgs a, int64_t* __restrict__ counts, const int64_t* __restrict__ colptr, int32_t* __restrict__ rowval,
        float* __restrict__ val) {
  extern __shared__ float s_h[];  // [warps][n_modules]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* h = s_h + (size_t)wib * a.n_modules;
  const int64_t n_loc = a.cell_hi - a.cell_lo;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t lc = warp; lc < n_loc; lc += nwarp) {
    const uint64_t cell = (uint64_t)(a.cell_lo + lc);
    const float lib = a.lib_sd * synth_gauss(synth_hash3(a.seed, cell, 0xFFFFFFF1ull));
    const int typ = (int)(synth_hash3(a.seed, cell, 0xFFFFFFF2ull) % (uint64_t)a.n_types);
    __syncwarp();
    for (int m = lane; m < a.n_modules; m += 32)
      h[m] = a.type_w * a.type_mean[(int64_t)typ * a.n_modules + m] +
             a.noise_w * synth_gauss(synth_hash3(a.seed, cell, (uint64_t)m + 0x40000000ull));
    __syncwarp();
    int64_t base = WRITE ? colptr[lc] : 0;
    int64_t total = 0;
    for (int64_t g0 = 0; g0 < a.n_genes; g0 += 32) {
      const int64_t g = g0 + lane;
      int x = 0;
      if (g < a.n_genes) {
        const int mod = a.gene_module[g];
        const float s = a.gene_loading[g];
        const float hm = mod >= 0 ? h[mod] : 0.f;
        const double loglam = (double)a.gene_base[g] + (double)lib + (double)s * (double)hm - 0.5 * (double)s * (double)s;
        x = synth_poisson(exp(loglam), synth_u01(synth_hash3(a.seed, (uint64_t)g, cell)));
      }
      const unsigned mask = __ballot_sync(0xffffffffu, x > 0);
      if (WRITE && x > 0) {
        const int64_t p = base + __popc(mask & ((1u << lane) - 1u));
        rowval[p] = (int32_t)g;
        val[p] = (float)x;
      }
      base += __popc(mask);
      total += __popc(mask);
    }
    if (total == 0) {  // no empty cells: gene (cell mod G) gets one count (documented fix-up)
      if (WRITE && lane == 0) {
        rowval[colptr[lc]] = (int32_t)(cell % (uint64_t)a.n_genes);
        val[colptr[lc]] = 1.f;
      }
      total = 1;
    }
    if (!WRITE && lane == 0) counts[lc] = total;
  }
}

extern "C" csc_status csc_generate_counts(csc_ctx* ctx, int64_t n_genes, int64_t cell_lo, int64_t cell_hi, uint64_t seed,
                                          const float* gene_base, const int32_t* gene_module, const float* gene_loading,
                                          const float* type_mean, int32_t n_types, int32_t n_modules, float lib_sd,
                                          float type_w, float noise_w, csc_sparse** out) {
  if (!ctx || !out || !gene_base || !gene_module || !gene_loading || !type_mean) return CSC_ERR_ARG;
  CSC_GUARD_BEGIN
  CSC_REQUIRE(ctx, n_genes > 0 && n_genes < ((int64_t)1 << 31) && cell_lo >= 0 && cell_hi >= cell_lo,
              "csc_generate_counts: bad shape");
  CSC_REQUIRE(ctx, n_types > 0 && n_modules > 0 && n_modules <= 1024, "csc_generate_counts: bad n_types/n_modules");
  CSC_CUDA(ctx, cudaSetDevice(ctx->device));
  StageTimer st(ctx, CSC_T_GENERATE);
  const int64_t n_loc = cell_hi - cell_lo;
  DevPtr d_base, d_mod, d_load, d_tm;
  CSC_TRY(dev_alloc(ctx, (size_t)n_genes * 4, &d_base));
  CSC_TRY(dev_alloc(ctx, (size_t)n_genes * 4, &d_mod));
  CSC_TRY(dev_alloc(ctx, (size_t)n_genes * 4, &d_load));
  CSC_TRY(dev_alloc(ctx, (size_t)n_types * n_modules * 4, &d_tm));
  CSC_CUDA(ctx, cudaMemcpyAsync(d_base->p, gene_base, (size_t)n_genes * 4, cudaMemcpyHostToDevice, ctx->stream));
  CSC_CUDA(ctx, cudaMemcpyAsync(d_mod->p, gene_module, (size_t)n_genes * 4, cudaMemcpyHostToDevice, ctx->stream));
  CSC_CUDA(ctx, cudaMemcpyAsync(d_load->p, gene_loading, (size_t)n_genes * 4, cudaMemcpyHostToDevice, ctx->stream));
  CSC_CUDA(ctx, cudaMemcpyAsync(d_tm->p, type_mean, (size_t)n_types * n_modules * 4, cudaMemcpyHostToDevice, ctx->stream));
  SynthArgs a;
  a.n_genes = n_genes;
  a.cell_lo = cell_lo;
  a.cell_hi = cell_hi;
  a.seed = seed;
  a.gene_base = (const float*)d_base->p;
  a.gene_module = (const int32_t*)d_mod->p;
  a.gene_loading = (const float*)d_load->p;
  a.type_mean = (const float*)d_tm->p;
  a.n_types = n_types;
  a.n_modules = n_modules;
  a.lib_sd = lib_sd;
  a.type_w = type_w;
  a.noise_w = noise_w;
  std::unique_ptr<csc_sparse> m(new csc_sparse());
  m->ctx = ctx;
  m->n_genes = n_genes;
  m->n_cells = n_loc;
  m->nnz = 0;
  CSC_TRY(dev_alloc(ctx, (size_t)(n_loc + 1) * 8, &m->colptr, true));
  if (n_loc == 0) {
    CSC_TRY(dev_alloc(ctx, 0, &m->rowval));
    CSC_TRY(dev_alloc(ctx, 0, &m->val));
    *out = m.release();
    return CSC_OK;
  }
  DevPtr counts;
  CSC_TRY(dev_alloc(ctx, (size_t)(n_loc + 1) * 8, &counts, true));
  const int threads = 256;
  const size_t smem = (size_t)(threads / 32) * n_modules * 4;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(n_loc, threads / 32), (int64_t)ctx->sm_count * 8);
  k_synth<false><<<grid, threads, smem, ctx->stream>>>(a, (int64_t*)counts->p, nullptr, nullptr, nullptr);
  CSC_LAUNCHED(ctx);
  {
    size_t tmp_bytes = 0;
    CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const int64_t*)counts->p, (int64_t*)m->colptr->p,
                                                n_loc + 1, ctx->stream));
    DevPtr tmp;
    CSC_TRY(dev_alloc(ctx, tmp_bytes, &tmp));
    CSC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp->p, tmp_bytes, (const int64_t*)counts->p, (int64_t*)m->colptr->p,
                                                n_loc + 1, ctx->stream));
    ctx->launches += 2;
    CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  int64_t nnz = 0;
  CSC_CUDA(ctx, cudaMemcpyAsync(&nnz, (const int64_t*)m->colptr->p + n_loc, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  m->nnz = nnz;
  CSC_TRY(dev_alloc(ctx, (size_t)nnz * 4, &m->rowval));
  CSC_TRY(dev_alloc(ctx, (size_t)nnz * 4, &m->val));
  k_synth<true><<<grid, threads, smem, ctx->stream>>>(a, nullptr, (const int64_t*)m->colptr->p, (int32_t*)m->rowval->p,
                                                       (float*)m->val->p);
  CSC_LAUNCHED(ctx);
  CSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = m.release();
  return CSC_OK;
  CSC_GUARD_END(ctx)
}
