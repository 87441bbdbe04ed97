// CPU helpers of the host-buffer entry point (host_convert.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace davo_host {
// src[n] float32 -> dst[n] IEEE binary16 bits, round to nearest even (F16C when the CPU has it).
// Returns false when some value does not convert to a finite half (|x| >= 65520 or NaN): the caller
// then sends the chunk as float32.
bool flows_to_half(const float* src, uint16_t* dst, size_t n);
bool flows_to_half_portable(const float* src, uint16_t* dst, size_t n);   // the scalar path, for tests
// tf.cast(label, int32) (truncation toward zero, davo.py:1115) -> byte; anything outside 0..18 becomes 255 (an
// all-zero one_hot row), NaN included: on the CPU tf.cast gives INT_MIN for it.  AVX2 when the CPU has it.
void labels_to_bytes(const float* src, uint8_t* dst, size_t n);
void labels_to_bytes_portable(const float* src, uint8_t* dst, size_t n);   // the scalar path, for tests
}  // namespace davo_host
