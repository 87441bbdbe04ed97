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
}  // namespace davo_host
