// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05, TF32 in /
// fp32 accumulate in TMEM), operands fed by TMA.  One kernel serves every conv
// of the PoseNN stack (reference nets/posenn.py:211-215, 238-240):
//
//   out[pixel, n] = act( bias[n] + sum_{kstep} A_kstep[pixel, 0:32] . B_kstep[n, 0:32] )
//
// * M tile = 128 output pixels = a 16 (rows) x 8 (cols) patch of one frame pair.
// * A K-step is one 32-float (128 B) slab of the reduction axis: a filter tap
//   (or, for the thin strided layers, two horizontally adjacent taps) x 32
//   input channels.  Its A operand is ONE 5-D TMA box {32 ch, 8 w, 1, 16 h, 1}
//   of the NHWC activation, placed by signed coordinates, so TF-'SAME' padding
//   (asymmetric included) is TMA's out-of-bounds zero fill and dilation is just
//   the tap offset -- no im2col buffer, no space-to-batch.
// * Stride-2 layers view the input as [N][H/2][2][W/2][2*C]: a tap becomes a
//   unit-stride box at one (row parity, column parity) of that view.
// * The box lands as 128 rows x 128 B with SWIZZLE_128B = the K-major UMMA
//   operand layout; B (weights, pre-packed [kstep][Cout][32], TF32-rounded) is
//   a 2-D TMA box of the same form.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA
// issuer, warps 2-5 epilogue (TMEM -> registers -> bias/ReLU -> global, or the
// spatial-sum epilogue of cnv7).  Accumulators are double-buffered in TMEM so
// the epilogue of tile i overlaps the main loop of tile i+1.  Persistent CTAs
// walk tiles round-robin.
#pragma once
#include "ptx.cuh"

namespace davo {

constexpr int kMaxKSteps = 72;
constexpr int kTileM = 128;
constexpr int kTileH = 16;
constexpr int kTileW = 8;
constexpr int kSlabBytes = 128;                     // 32 tf32
constexpr int kABytes = kTileM * kSlabBytes;        // 16 KB
constexpr int kConvThreads = 192;

struct KStep {
  int16_t c;     // inner (channel-axis) start coordinate, before the group offset
  int8_t dw;     // column offset added to the tile's first output column
  int8_t par;    // coordinate on the row-parity axis (0 for stride-1 layers)
  int8_t dh;     // row offset added to the tile's first output row
  int8_t pad_[3];
};

enum { EPI_STORE_RELU = 0, EPI_SUM_RELU = 1 };

struct ConvParams {
  int num_tiles;        // pairs * groups * tiles_h * tiles_w
  int tiles_w, tiles_h, groups;
  int Hout, Wout;
  int out_stride;       // floats per output pixel (all groups)
  int cin_group_off;    // inner-coordinate offset of group g = g * cin_group_off
  int n_ksteps;
  float* out;           // EPI_STORE: [pairs][Hout][Wout][out_stride]
  const float* bias;    // [groups * BN]
  float* sum_out;       // EPI_SUM:   [pairs][groups][tiles_h*tiles_w][4][BN]
  KStep ks[kMaxKSteps];
};

template <int BN>
struct ConvCfg {
  static constexpr int kBBytes = BN * kSlabBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kAccStride = BN < 32 ? 32 : BN;             // TMEM columns per accumulator
  static constexpr int kTmemCols = 2 * kAccStride;                 // power of two for BN in {16..256}
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-B aligned stage bases.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* full = bars;                 // [S]  TMA -> MMA
  uint64_t* empty = bars + S;            // [S]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * S;     // [2]  MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * S + 2;  // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int tiles_per_pair = tiles_per_img * p.groups;

  if (warp == 0) {
    // ------------------------------------------------------- TMA producer --
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_pair;
        int r = tile - n * tiles_per_pair;
        const int g = r / tiles_per_img;
        r -= g * tiles_per_img;
        const int h0 = (r / p.tiles_w) * kTileH;
        const int w0 = (r % p.tiles_w) * kTileW;
        const int cg = g * p.cin_group_off;
        const int brow0 = g * p.n_ksteps * BN;
        for (int k = 0; k < p.n_ksteps; ++k) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full[stage], Cfg::kStageBytes);
          const KStep s = p.ks[k];
          tma_load_5d(sa, &tmA, &full[stage], cg + s.c, w0 + s.dw, s.par, h0 + s.dh, n);
          tma_load_2d(sa + kABytes, &tmB, &full[stage], 0, brow0 + k * BN);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------- MMA issuer --
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(kTileM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * Cfg::kAccStride;
        for (int k = 0; k < p.n_ksteps; ++k) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + kABytes);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // 4 x (K = 8 tf32 = 32 B): +2 in the >>4 address field
            tc_mma_tf32(d, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0);
          tc_commit(&empty[stage]);         // frees the smem slot when these MMAs retire
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        tc_commit(&acc_full[acc]);          // accumulator complete -> epilogue
      }
    }
  } else {
    // ----------------------------------------------------------- epilogue --
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;            // accumulator row = pixel within the tile
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n = tile / tiles_per_pair;
      int r = tile - n * tiles_per_pair;
      const int g = r / tiles_per_img;
      r -= g * tiles_per_img;
      const int h = (r / p.tiles_w) * kTileH + (m >> 3);
      const int w = (r % p.tiles_w) * kTileW + (m & 7);
      const bool valid = (h < p.Hout) && (w < p.Wout);
      const float* bias = p.bias + g * BN;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t0 = tmem_base + acc * Cfg::kAccStride + (uint32_t(q * 32) << 16);
      if constexpr (EPI == EPI_STORE_RELU) {
        float* dst = p.out + ((size_t)(n * p.Hout + h) * p.Wout + w) * p.out_stride + g * BN;
        if constexpr (BN >= 32) {
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(t0 + c0, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o;
                o.x = round_tf32(fmaxf(__uint_as_float(v[j + 0]) + __ldg(bias + c0 + j + 0), 0.f));
                o.y = round_tf32(fmaxf(__uint_as_float(v[j + 1]) + __ldg(bias + c0 + j + 1), 0.f));
                o.z = round_tf32(fmaxf(__uint_as_float(v[j + 2]) + __ldg(bias + c0 + j + 2), 0.f));
                o.w = round_tf32(fmaxf(__uint_as_float(v[j + 3]) + __ldg(bias + c0 + j + 3), 0.f));
                *reinterpret_cast<float4*>(dst + c0 + j) = o;
              }
            }
          }
        } else {
          uint32_t v[16];
          tmem_ld_32x16(t0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 o;
              o.x = round_tf32(fmaxf(__uint_as_float(v[j + 0]) + __ldg(bias + j + 0), 0.f));
              o.y = round_tf32(fmaxf(__uint_as_float(v[j + 1]) + __ldg(bias + j + 1), 0.f));
              o.z = round_tf32(fmaxf(__uint_as_float(v[j + 2]) + __ldg(bias + j + 2), 0.f));
              o.w = round_tf32(fmaxf(__uint_as_float(v[j + 3]) + __ldg(bias + j + 3), 0.f));
              *reinterpret_cast<float4*>(dst + j) = o;
            }
          }
        }
      } else {
        // Spatial-sum epilogue (cnv7 -> pred -> mean, reference nets/posenn.py:239-241:
        // pred is linear, so only sum_pixels relu(cnv7) is needed).  Each warp
        // reduces its 32 pixels per column with a transpose-reduce butterfly and
        // writes one deterministic partial row; the head kernel adds them in order.
        float* dst = p.sum_out + ((size_t)(n * p.groups + g) * tiles_per_img + r) * 4 * BN + q * BN;
        static_assert(EPI != EPI_SUM_RELU || BN >= 32, "sum epilogue wants BN >= 32");
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t0 + c0, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            f[j] = valid ? fmaxf(__uint_as_float(v[j]) + __ldg(bias + c0 + j), 0.f) : 0.f;
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool hi = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; ++j) {
              const float send = hi ? f[j] : f[j + off];
              const float keep = hi ? f[j + off] : f[j];
              f[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          dst[c0 + lane] = f[0];            // lane L now holds column c0 + L
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace davo
