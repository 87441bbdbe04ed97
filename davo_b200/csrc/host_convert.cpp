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

// ---- segmentation labels: float32 -> one byte -------------------------------------------------------------------
static void labels_scalar(const float* src, uint8_t* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    const float v = src[i];
    dst[i] = (v > -1.0f && v < 19.0f) ? (uint8_t)(int)v : (uint8_t)255;      // NaN fails both comparisons: no class
  }
}

#ifdef DAVO_X86
// 16 labels per step: in range -> the truncated integer, everything else (NaN included) -> 255
static void labels_sse2(const float* src, uint8_t* dst, size_t n) {
  const __m128 lo = _mm_set1_ps(-1.0f), hi = _mm_set1_ps(19.0f);
  const __m128i inval = _mm_set1_epi32(255);
  size_t i = 0;
  for (; i + 16 <= n; i += 16) {
    __m128i r[4];
    for (int k = 0; k < 4; ++k) {
      const __m128 v = _mm_loadu_ps(src + i + 4 * k);
      const __m128i ok = _mm_castps_si128(_mm_and_ps(_mm_cmpgt_ps(v, lo), _mm_cmplt_ps(v, hi)));   // false for NaN
      const __m128i iv = _mm_cvttps_epi32(v);
      r[k] = _mm_or_si128(_mm_and_si128(ok, iv), _mm_andnot_si128(ok, inval));
    }
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i),
                     _mm_packus_epi16(_mm_packs_epi32(r[0], r[1]), _mm_packs_epi32(r[2], r[3])));
  }
  labels_scalar(src + i, dst + i, n - i);
}

// the same on 32 labels per step; the 256-bit packs work per 128-bit lane, so one dword permute puts the bytes in order
__attribute__((target("avx2"))) static void labels_avx2(const float* src, uint8_t* dst, size_t n) {
  const __m256 lo = _mm256_set1_ps(-1.0f), hi = _mm256_set1_ps(19.0f);
  const __m256i inval = _mm256_set1_epi32(255);
  const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    __m256i r[4];
    for (int k = 0; k < 4; ++k) {
      const __m256 v = _mm256_loadu_ps(src + i + 8 * k);
      const __m256i ok = _mm256_castps_si256(_mm256_and_ps(_mm256_cmp_ps(v, lo, _CMP_GT_OQ), _mm256_cmp_ps(v, hi, _CMP_LT_OQ)));   // false for NaN
      const __m256i iv = _mm256_cvttps_epi32(v);
      r[k] = _mm256_or_si256(_mm256_and_si256(ok, iv), _mm256_andnot_si256(ok, inval));
    }
    const __m256i b = _mm256_packus_epi16(_mm256_packs_epi32(r[0], r[1]), _mm256_packs_epi32(r[2], r[3]));
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_permutevar8x32_epi32(b, order));
  }
  labels_sse2(src + i, dst + i, n - i);
}
#endif

void labels_to_bytes(const float* src, uint8_t* dst, size_t n) {
#ifdef DAVO_X86
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return labels_avx2(src, dst, n);
  return labels_sse2(src, dst, n);
#else
  labels_scalar(src, dst, n);
#endif
}

void labels_to_bytes_portable(const float* src, uint8_t* dst, size_t n) { labels_scalar(src, dst, n); }

}  // namespace davo_host
