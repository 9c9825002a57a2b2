// Host lane of the two-lane upload (sparse.cu: two_lane_upload): narrow one piece of a 64-bit host array into a
// page-locked slot.  Plain C++ compiled by g++ (not nvcc) so that the AVX-512 variant can use intrinsics behind a
// run-time CPU check; the generic loops are the same arithmetic.
//   kind 0: Int64 row indices  -> uint16 (idx16) or int32, minus `base`; an index outside [0, n_rows) sets *bad
//   kind 1: Float64 values     -> uint16 when every value of the piece is an integer in [0, 65535] with a clear sign bit
//                                 (raw counts are), else fp32 (round to nearest even, as the device kernel does)
//   kind 2: Int64 values       -> uint16 when all are in [0, 65535], else fp32
// Returns the element size written (2 or 4).
#include <immintrin.h>
#include <stdint.h>

#define HIDDEN __attribute__((visibility("hidden")))

namespace {

int narrow_generic(int kind, const void* src, int64_t cnt, void* stage, bool idx16, int64_t base, int64_t n_rows, int* bad) {
  if (kind == 0) {
    const int64_t* in = (const int64_t*)src;
    uint64_t viol = 0;
    if (idx16) {
      uint16_t* o = (uint16_t*)stage;
      for (int64_t i = 0; i < cnt; ++i) {
        const int64_t r = in[i] - base;
        viol |= (uint64_t)((uint64_t)r >= (uint64_t)n_rows);
        o[i] = (uint16_t)r;
      }
    } else {
      int32_t* o = (int32_t*)stage;
      for (int64_t i = 0; i < cnt; ++i) {
        const int64_t r = in[i] - base;
        viol |= (uint64_t)((uint64_t)r >= (uint64_t)n_rows);
        o[i] = (int32_t)r;
      }
    }
    if (viol) *bad = 1;
    return idx16 ? 2 : 4;
  }
  if (kind == 1) {
    const double* in = (const double*)src;
    const int64_t* bits = (const int64_t*)src;
    uint16_t* o = (uint16_t*)stage;
    uint32_t miss = 0;
    for (int64_t i = 0; i < cnt; ++i) {
      const double c = in[i];
      const bool in_range = (c <= 65535.0) & (bits[i] >= 0);   // sign bit clear keeps -0.0 out, NaN fails the compare
      const int32_t q = in_range ? (int32_t)c : 0;
      miss |= (uint32_t)(!in_range | ((double)q != c));
      o[i] = (uint16_t)q;
    }
    if (!miss) return 2;
    float* f = (float*)stage;
    for (int64_t i = 0; i < cnt; ++i) f[i] = (float)in[i];
    return 4;
  }
  const int64_t* in = (const int64_t*)src;
  uint16_t* o = (uint16_t*)stage;
  uint64_t miss = 0;
  for (int64_t i = 0; i < cnt; ++i) {
    miss |= (uint64_t)in[i] >> 16;
    o[i] = (uint16_t)in[i];
  }
  if (!miss) return 2;
  float* f = (float*)stage;
  for (int64_t i = 0; i < cnt; ++i) f[i] = (float)in[i];
  return 4;
}

// AVX-512 variant.  A piece is read as NS = 4 interleaved sub-streams (quarter q of the piece at in + q * len): one
// sequential stream per thread keeps too few cache-line fills in flight (the hardware prefetcher follows streams, and a
// core has a dozen fill buffers), eight threads reading one stream each got 25 GB/s out of a host that gives 50 GB/s to
// eight threads reading four streams each.
constexpr int NS = 4;
__attribute__((target("avx512f,avx512vl,avx512bw,avx512dq")))
int narrow_avx512(int kind, const void* src, int64_t cnt, void* stage, bool idx16, int64_t base, int64_t n_rows, int* bad) {
  const int64_t len = (cnt / (8 * NS)) * 8;       // elements per sub-stream, a multiple of the vector width
  const int64_t body = len * NS;
  const int64_t tail = cnt - body;
  if (kind == 0) {
    const int64_t* in = (const int64_t*)src;
    const __m512i vb = _mm512_set1_epi64(base), vn = _mm512_set1_epi64(n_rows);
    __mmask8 viol = 0;
    if (idx16) {
      uint16_t* o = (uint16_t*)stage;
      for (int64_t i = 0; i < len; i += 8) {
#pragma GCC unroll 4
        for (int q = 0; q < NS; ++q) {
          const __m512i a = _mm512_sub_epi64(_mm512_loadu_si512(in + q * len + i), vb);
          viol |= _mm512_cmpge_epu64_mask(a, vn);
          _mm_storeu_si128((__m128i*)(o + q * len + i), _mm512_cvtepi64_epi16(a));
        }
      }
    } else {
      int32_t* o = (int32_t*)stage;
      for (int64_t i = 0; i < len; i += 8) {
#pragma GCC unroll 4
        for (int q = 0; q < NS; ++q) {
          const __m512i a = _mm512_sub_epi64(_mm512_loadu_si512(in + q * len + i), vb);
          viol |= _mm512_cmpge_epu64_mask(a, vn);
          _mm256_storeu_si256((__m256i*)(o + q * len + i), _mm512_cvtepi64_epi32(a));
        }
      }
    }
    if (viol) *bad = 1;
    if (tail) narrow_generic(kind, in + body, tail, (char*)stage + body * (idx16 ? 2 : 4), idx16, base, n_rows, bad);
    return idx16 ? 2 : 4;
  }
  if (kind == 1) {
    const double* in = (const double*)src;
    uint16_t* o = (uint16_t*)stage;
    __m256i hi = _mm256_setzero_si256();
    __mmask8 all_eq = 0xff;
    for (int64_t i = 0; i < len; i += 8) {
#pragma GCC unroll 4
      for (int q = 0; q < NS; ++q) {
        const __m512d a = _mm512_loadu_pd(in + q * len + i);
        const __m256i qa = _mm512_cvttpd_epi32(a);                               // out of range / NaN -> 0x80000000
        // bitwise equality of the round trip: an integer value with a clear sign bit (-0.0 and NaN differ)
        all_eq &= _mm512_cmpeq_epi64_mask(_mm512_castpd_si512(_mm512_cvtepi32_pd(qa)), _mm512_castpd_si512(a));
        hi = _mm256_or_si256(hi, qa);                                            // bits 16..31 set: negative or > 65535
        _mm_storeu_si128((__m128i*)(o + q * len + i), _mm256_cvtepi32_epi16(qa));
      }
    }
    bool ok = all_eq == 0xff && _mm256_testz_si256(hi, _mm256_set1_epi32((int)0xffff0000u));
    if (ok && tail) ok = narrow_generic(kind, in + body, tail, o + body, false, 0, 0, bad) == 2;
    if (ok) return 2;
    float* f = (float*)stage;
    for (int64_t i = 0; i < len; i += 8) {
#pragma GCC unroll 4
      for (int q = 0; q < NS; ++q) _mm256_storeu_ps(f + q * len + i, _mm512_cvtpd_ps(_mm512_loadu_pd(in + q * len + i)));
    }
    for (int64_t i = body; i < cnt; ++i) f[i] = (float)in[i];
    return 4;
  }
  const int64_t* in = (const int64_t*)src;
  uint16_t* o = (uint16_t*)stage;
  __m512i acc = _mm512_setzero_si512();
  for (int64_t i = 0; i < len; i += 8) {
#pragma GCC unroll 4
    for (int q = 0; q < NS; ++q) {
      const __m512i a = _mm512_loadu_si512(in + q * len + i);
      acc = _mm512_or_si512(acc, a);
      _mm_storeu_si128((__m128i*)(o + q * len + i), _mm512_cvtepi64_epi16(a));
    }
  }
  bool ok = _mm512_test_epi64_mask(acc, _mm512_set1_epi64(~(int64_t)0xffff)) == 0;
  if (ok && tail) ok = narrow_generic(kind, in + body, tail, o + body, false, 0, 0, bad) == 2;
  if (ok) return 2;
  float* f = (float*)stage;
  for (int64_t i = 0; i < cnt; ++i) f[i] = (float)in[i];
  return 4;
}

}  // namespace

HIDDEN int host_narrow_piece(int kind, const void* src, int64_t cnt, void* stage, bool idx16, int64_t base, int64_t n_rows,
                             int* bad) {
  static const bool wide = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") &&
                           __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq");
  return wide ? narrow_avx512(kind, src, cnt, stage, idx16, base, n_rows, bad)
              : narrow_generic(kind, src, cnt, stage, idx16, base, n_rows, bad);
}

// test hook (tests/test_host_logic.py drives both variants through ctypes on the CPU): variant 0 generic, 1 AVX-512 when
// the CPU has it (else generic); returns the element size
extern "C" int csc_test_host_narrow(int variant, int kind, const void* src, int64_t cnt, void* stage, int idx16,
                                    int64_t base, int64_t n_rows, int* bad) {
  if (variant == 0) return narrow_generic(kind, src, cnt, stage, idx16 != 0, base, n_rows, bad);
  return host_narrow_piece(kind, src, cnt, stage, idx16 != 0, base, n_rows, bad);
}
