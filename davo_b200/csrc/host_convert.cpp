// CPU side of the host-buffer entry point: the optical flow crosses PCIe as IEEE binary16.
//
// The conv stack reads its input with 11 significant bits (TF32 operands), and the library defines
// the flow input of BOTH entry points as rounded to binary16 (frontend.cuh: flow_q), which has the
// same 11 bits: so the host entry point may round on the CPU and send half the bytes.  The result
// is bit-identical to the device entry point as long as every value converts to a finite half;
// a chunk holding |x| >= 65520 or a NaN is reported and sent as float32 instead (flow_q leaves
// such values alone).  Plain C++ (no CUDA) so that the compiler's per-function target attribute and
// run-time CPU dispatch are available.
#include "host_convert.h"

#include <cmath>
#include <cstring>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define DAVO_X86 1
#endif

namespace davo_host {

// float -> binary16 bits, round to nearest even, subnormals kept, overflow -> infinity
static inline uint16_t half_bits(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((x > 0x7f800000u) ? 0x0200u : 0u));   // inf / NaN
  if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);             // >= 65520 rounds to infinity
  if (x < 0x38800000u) {                                               // below 2^-14: subnormal half
    if (x < 0x33000000u) return (uint16_t)sign;                        // < 2^-25 rounds to zero
    const int e = (int)(x >> 23);                                      // biased exponent, 102..112
    const uint32_t m = (x & 0x007fffffu) | 0x00800000u;
    const int shift = 126 - e;                                         // 14..24 bits dropped
    uint32_t h = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((x - 0x38000000u) >> 13);                              // rebias 127 -> 15, drop 13 bits
  const uint32_t rem = x & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;              // carries into the exponent correctly
  return (uint16_t)(sign | h);
}

static bool to_half_scalar(const float* src, uint16_t* dst, size_t n) {
  bool ok = true;
  for (size_t i = 0; i < n; ++i) {
    const float v = src[i];
    ok = ok && (std::fabs(v) < 65520.0f);                              // false for NaN as well
    dst[i] = half_bits(v);
  }
  return ok;
}

#ifdef DAVO_X86
__attribute__((target("avx,f16c"))) static bool to_half_f16c(const float* src, uint16_t* dst, size_t n) {
  const __m256 lim = _mm256_set1_ps(65520.0f);
  const __m256 absmask = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
  __m256 bad = _mm256_setzero_ps();
  size_t i = 0;
  for (; i + 16 <= n; i += 16) {
    const __m256 a = _mm256_loadu_ps(src + i), b = _mm256_loadu_ps(src + i + 8);
    bad = _mm256_or_ps(bad, _mm256_cmp_ps(_mm256_and_ps(a, absmask), lim, _CMP_NLT_UQ));   // >= limit or NaN
    bad = _mm256_or_ps(bad, _mm256_cmp_ps(_mm256_and_ps(b, absmask), lim, _CMP_NLT_UQ));
    // plain stores: non-temporal ones were measured slower here (the copy engine reads the lines right away)
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), _mm256_cvtps_ph(a, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 8), _mm256_cvtps_ph(b, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
  }
  bool ok = _mm256_movemask_ps(bad) == 0;
  if (i < n) ok = to_half_scalar(src + i, dst + i, n - i) && ok;
  return ok;
}
#endif

bool flows_to_half(const float* src, uint16_t* dst, size_t n) {
#ifdef DAVO_X86
  static const bool f16c = __builtin_cpu_supports("avx") && __builtin_cpu_supports("f16c");
  if (f16c) return to_half_f16c(src, dst, n);
#endif
  return to_half_scalar(src, dst, n);
}

bool flows_to_half_portable(const float* src, uint16_t* dst, size_t n) { return to_half_scalar(src, dst, n); }

}  // namespace davo_host
