// Shared definitions of the two implicit-GEMM convolution kernels (conv_pm.cuh: pixels on
// the MMA M axis; conv_cm.cuh: output channels on the M axis).
#pragma once
#include "ptx.cuh"

namespace davo {

constexpr int kMaxPatches = 32;
constexpr int kMaxTaps = 144;
constexpr int kTileW = 8;                           // tile columns: one 8-row UMMA group
constexpr int kSlabBytes = 128;                     // 32 tf32
constexpr int kConvThreads = 224;                   // 7 warps
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 226 * 1024;             // of the 227 KB a CTA may own
constexpr int kBarrierBytes = 512;

// A PATCH is the input halo of one tile for one 32-channel slab: one 5-D TMA box
// {32 ch, Wp, 1, Hp, 1} placed with signed coordinates (out-of-bounds = TF 'SAME' zeros).
struct PatchDesc {
  int16_t c;        // inner (channel-axis) start coordinate, before the group offset
  int8_t dw;        // patch origin relative to the tile's first output column
  int8_t par;       // coordinate on the row-parity axis (0 for stride-1 layers)
  int8_t dh;        // patch origin relative to the tile's first output row
  uint8_t ntaps;
  uint16_t tap0;    // first entry in taps[]
};
// A TAP is one 32-float slab of the reduction: a window of the patch x one weight slab.
struct TapDesc {
  uint16_t a_off;   // window origin inside the patch, in 128-B rows: row * Wp + col
  uint16_t b_idx;   // weight slab index
  // Column-widened layers only (conv_pm.cuh, WIDE): the MMA covers a window of the accumulator.
  uint8_t n16;      // MMA N / 16
  uint8_t dcol16;   // first accumulator column / 16
  uint8_t brow8;    // first row of the weight slab / 8
  uint8_t fresh;    // 1: first MMA to touch its accumulator columns in the tile (overwrite)
};

enum { EPI_STORE_RELU = 0, EPI_SUM_RELU = 1 };

// What a store epilogue writes for the accumulator value x (+ bias): relu, rounded to TF32 (ties away, the integer
// form of cvt.rna: add half an ulp to the magnitude, clear 13 bits) for the next layer's tensor-core read -- or,
// ConvParams::raw (the -batch_norm variants), the plain fp32 convolution sum, which the batch-norm kernels then
// normalise, activate and round (bn.cuh).  The mode lives in three per-thread constants, so the hot path pays
// nothing for it: max(x, 0 | -inf), + 0x1000 | 0, & 0xFFFFE000 | 0xFFFFFFFF.
struct EpiAct { float lo; uint32_t add, mask; };
__device__ __forceinline__ EpiAct epi_act(int raw) {
  EpiAct a;
  a.lo = raw ? __uint_as_float(0xFF800000u) : 0.f;
  a.add = raw ? 0u : 0x1000u;
  a.mask = raw ? 0xFFFFFFFFu : 0xFFFFE000u;
  return a;
}
__device__ __forceinline__ float epi_out(float x, const EpiAct& a) {
  return __uint_as_float((__float_as_uint(fmaxf(x, a.lo)) + a.add) & a.mask);
}

// Ordering of the tensor-core reads after a TMA-fed mbarrier wait.  The wait itself orders the
// async-proxy writes of TMA before what the waiting thread issues next; -DDAVO_MMA_FENCES=1
// adds the explicit tcgen05.fence for A/B comparison.
#if defined(DAVO_MMA_FENCES) && DAVO_MMA_FENCES
#define DAVO_MMA_FENCE() tc_fence_after()
#else
#define DAVO_MMA_FENCE() do { } while (0)
#endif

// Shared-memory matrix descriptor (K-major, SWIZZLE_128B): rows 128 B apart, 8-row groups
// `sbo_bytes` apart.  A window of a patch is start = patch + (row*Wp + col)*128, SBO = Wp*128:
// the hardware swizzle is a function of the absolute shared-memory address, so a row-shifted
// window reads back exactly what TMA wrote (measured: tools/experiments/desc_shift.cu).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

// Optional stall accounting (-DDAVO_TIMING, tools/timing_run.py): cycles per CTA:
// [0] patch producer waiting, [1] weight producer waiting, [2] MMA on patch, [3] MMA on
// weights, [4] MMA on acc_empty, [5] epilogue warp 2 on acc_full, [6] epilogue warp 2 busy,
// [7] epilogue warp 2 lifetime.
#ifdef DAVO_TIMING
__device__ long long g_conv_timing[148 * 8];
#define TWAIT(slot, stmt) do { long long t_ = clock64(); stmt; tacc[slot] += clock64() - t_; } while (0)
#else
#define TWAIT(slot, stmt) stmt
#endif

}  // namespace davo
