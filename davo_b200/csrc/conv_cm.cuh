// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05, TF32 in / fp32
// accumulate in TMEM), operands fed by TMA.  One kernel serves every conv of the
// PoseNN stack (reference nets/posenn.py:211-215, 238-240).
//
// Orientation: the GEMM is computed TRANSPOSED, D^T[cout, pixel] = W[cout, k] . X[pixel, k]^T:
//   * M (TMEM lanes)   = 128 output channels (one "m-block"; layers with fewer are zero-padded,
//                        layers with 256 have two m-blocks, each its own tile)
//   * N (TMEM columns) = NPIX output pixels = a (NPIX/8 rows) x 8 (cols) block of one frame pair
//   * K                = filter taps x 32-channel slabs, 4 MMAs of K=8 per tap
// Measured on B200 (tools/experiments/mma_floor.cu): one tcgen05.mma M=128,K=8 costs
// max(53.8, N/2) cycles, so N >= 128 runs the tensor pipe at full rate.
// Putting PIXELS on N makes every layer an N=256 problem, whatever its channel count; and a
// thread of the epilogue then owns one output CHANNEL, so each store instruction of a warp
// writes 32 consecutive channels of one pixel (a whole 128-B line of the NHWC tensor) straight
// from registers -- no shared-memory transpose competing with the tensor core for smem
// bandwidth (the measured bottleneck of the pixels-on-M version).
//
// Pixel operand: never gathered per tap.  A PATCH -- the tile's input halo, {32 ch, Wp cols,
// 1, Hp rows, 1} of the NHWC activation -- is loaded ONCE by a 5-D TMA box placed with signed
// coordinates, so TF-'SAME' padding (asymmetric included) is TMA's out-of-bounds zero fill.
// It lands as Hp*Wp rows of 128 B with SWIZZLE_128B.  Every tap of the patch is then just a
// different UMMA descriptor: start = patch + (row*Wp + col)*128 B, 8-row groups SBO = Wp*128 B
// apart.  (The hardware swizzle is a function of the absolute shared-memory address, so a
// row-shifted window needs no re-layout -- measured, tools/experiments/desc_shift.cu.)
// Dilation is only a tap offset: no im2col buffer, no space-to-batch, each input byte crosses
// L2->SM once per tile.  Stride-2 layers view the input as [N][H/2][2][W/2][2*C]: a tap is a
// unit-stride window at one (row parity, column parity) of that view; one patch per parity.
//
// Weight operand: pre-packed [group][m-block][tap][128][32], TF32-rounded, streamed through a
// ring, one 16-KB 2-D TMA box per tap.
//
// Warp roles (224 threads): warp 0 patch producer, warp 6 weight producer, warp 1 TMEM owner +
// single-thread MMA issuer, warps 2-5 epilogue (TMEM -> registers -> bias/ReLU/round ->
// coalesced global stores, or the per-channel spatial sum of cnv7).  Accumulators are
// double-buffered in TMEM (2 x NPIX <= 512 columns) so the epilogue of tile i overlaps the main
// loop of tile i+1.  Persistent CTAs walk tiles round-robin.
#pragma once
#include "conv_common.cuh"

namespace davo {
namespace cm {

constexpr int kBlockM = 128;                        // output channels per tile
// 11 warps: 0 patch producer, 1 MMA, 6 weight producer, 2-5 and 7-10 epilogue.  A warp may only
// read the TMEM lane quadrant (warp % 4), so the two warps of a quadrant take alternate 32-pixel
// column blocks: the epilogue is bound by one warp's instruction issue, and cnv4's main loop
// (18 taps) is no longer than its epilogue.
constexpr int kThreads = 352;
constexpr int kEpiWarps = 8;
constexpr int kWBytes = kBlockM * kSlabBytes;       // one weight stage: 16 KB

struct ConvParams {
  int num_tiles;        // pairs * groups * tiles_h * tiles_w * m_blocks; CLUSTER: work items (see decode_item_cluster)
  int num_pixel_tiles;  // CLUSTER: pairs * tiles_h * tiles_w
  int tiles_w, tiles_h, groups, m_blocks;
  int Hout, Wout;
  int out_stride;       // floats per output pixel (all groups)
  int cout_g;           // real output channels per group
  int cin_group_off;    // inner-coordinate offset of group g = g * cin_group_off
  int n_patches, n_taps;          // per tile
  int patch_w;                    // Wp
  int patch_bytes;                // Hp * Wp * 128 (what TMA delivers)
  int patch_stage_bytes;          // rounded up to 1024
  int p_stages, w_stages;         // ring depths
  int raw;                        // store epilogue: 1 = plain fp32 sums, no relu, no rounding (-batch_norm: bn.cuh finishes the layer)
  float* out;           // EPI_STORE: [pairs][Hout][Wout][out_stride]
  const float* bias;    // [groups * cout_g]
  float* sum_out;       // EPI_SUM:   [pairs][groups][tiles_h*tiles_w][cout_g]
  PatchDesc patches[kMaxPatches];
  TapDesc taps[kMaxTaps];
};

struct TileCoord { int n, g, mb, h0, w0, t; };

// CLUSTER (2 CTAs): the two CTAs of a cluster walk the same (group, m-block) sequence on two
// different pixel tiles, so each weight slab is fetched from L2 once: CTA r loads rows
// [64r, 64r+64) of the slab and multicasts them into both CTAs' rings.  A stage is refilled only
// after BOTH MMA warps released it (w_empty counts 2; the commit is multicast).  Work item `item`
// = (pixel-tile pair, group, m-block); with an odd number of pixel tiles the last pair's second
// CTA repeats the last tile (identical stores).
template <int NPIX>
__device__ __forceinline__ TileCoord decode_item_cluster(const ConvParams& p, int item, int rank) {
  TileCoord c;
  c.mb = item % p.m_blocks;
  int r = item / p.m_blocks;
  c.g = r % p.groups;
  const int pp = r / p.groups;
  const int tpi = p.tiles_h * p.tiles_w;
  const int pix = min(2 * pp + rank, p.num_pixel_tiles - 1);
  c.t = pix % tpi;
  c.n = pix / tpi;
  c.h0 = (c.t / p.tiles_w) * (NPIX / kTileW);
  c.w0 = (c.t % p.tiles_w) * kTileW;
  return c;
}

template <int NPIX>
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
  TileCoord c;
  c.mb = tile % p.m_blocks;
  int r = tile / p.m_blocks;
  const int tpi = p.tiles_h * p.tiles_w;
  c.t = r % tpi;
  r /= tpi;
  c.g = r % p.groups;
  c.n = r / p.groups;
  c.h0 = (c.t / p.tiles_w) * (NPIX / kTileW);
  c.w0 = (c.t % p.tiles_w) * kTileW;
  return c;
}

// STAGED: the store epilogue goes through shared memory and TMA tile stores (one 4-KB buffer per
// epilogue warp behind the barriers) instead of 32 scalar stores per 32x32 block.
template <int NPIX, int EPI, bool STAGED = false, bool CLUSTER = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands: keep every stage base 1024-B aligned.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const int PS = p.p_stages, WS = p.w_stages;
  uint8_t* smem_p = smem;
  uint8_t* smem_w = smem + PS * p.patch_stage_bytes;
  uint8_t* epi_stage = smem_w + WS * kWBytes;                      // 1024-B aligned; STAGED only
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + (STAGED ? kEpiWarps * 4096 : 0));
  uint64_t* p_full = bars;                        // [kMaxStages] TMA -> MMA
  uint64_t* p_empty = bars + kMaxStages;          // [kMaxStages] MMA -> TMA
  uint64_t* w_full = bars + 2 * kMaxStages;
  uint64_t* w_empty = bars + 3 * kMaxStages;
  uint64_t* acc_full = bars + 4 * kMaxStages;     // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;             // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  constexpr int kTmemCols = 2 * NPIX;

  pdl_launch_dependents();
  const EpiAct ea = epi_act(p.raw);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // known warp-uniform to the compiler
  const int lane = threadIdx.x & 31;
  const int rank = CLUSTER ? (int)cluster_ctarank() : 0;
  const int first = CLUSTER ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int step = CLUSTER ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto decode = [&](int tile) { return CLUSTER ? decode_item_cluster<NPIX>(p, tile, rank) : decode_tile<NPIX>(p, tile); };
#ifdef DAVO_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_cta0 = clock64();
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    if constexpr (STAGED) tma_prefetch_desc(&tmO);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&p_full[i], 1);
      mbar_init(&p_empty[i], 1);
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], CLUSTER ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], EPI == EPI_STORE_RELU ? kEpiWarps : 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if constexpr (CLUSTER) cluster_sync();          // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------- patch producer --
    if (lane == 0) {
      pdl_wait();                 // the activations this layer reads are the previous kernel's output
      int ps = 0;
      uint32_t pphase = 0;
      for (int tile = first; tile < p.num_tiles; tile += step) {
        const TileCoord tc = decode(tile);
        for (int pi = 0; pi < p.n_patches; ++pi) {
          const PatchDesc d = p.patches[pi];
          TWAIT(0, mbar_wait(&p_empty[ps], pphase ^ 1));
          mbar_expect_tx(&p_full[ps], p.patch_bytes);
          tma_load_5d(smem_p + ps * p.patch_stage_bytes, &tmX, &p_full[ps],
                      tc.g * p.cin_group_off + d.c, tc.w0 + d.dw, d.par, tc.h0 + d.dh, tc.n);
          if (++ps == PS) { ps = 0; pphase ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    // --------------------------------------------------- weight producer --
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int tile = first; tile < p.num_tiles; tile += step) {
        const TileCoord tc = decode(tile);
        const int row0 = (tc.g * p.m_blocks + tc.mb) * p.n_taps;
        for (int t = 0; t < p.n_taps; ++t) {            // taps[] is in issue order
          TWAIT(1, mbar_wait(&w_empty[ws], wphase ^ 1));
          mbar_expect_tx(&w_full[ws], kWBytes);
          if constexpr (CLUSTER)      // my half of the slab, into both CTAs (tmW's box is 64 rows here)
            tma_load_2d_multicast(smem_w + ws * kWBytes + rank * (kWBytes / 2), &tmW, &w_full[ws], 0,
                                  (row0 + p.taps[t].b_idx) * kBlockM + rank * (kBlockM / 2), (uint16_t)3);
          else
            tma_load_2d(smem_w + ws * kWBytes, &tmW, &w_full[ws], 0, (row0 + p.taps[t].b_idx) * kBlockM);
          if (++ws == WS) { ws = 0; wphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------- MMA issuer --
    // All 32 lanes run the loops (operands stay on the uniform datapath); one elected lane issues.
    {
      constexpr uint32_t idesc = umma_idesc_tf32(kBlockM, NPIX);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t w_hi = umma_desc_hi(1024), x_hi = umma_desc_hi((uint32_t)p.patch_w * kSlabBytes);
      const uint32_t p_lo0 = umma_desc_lo(smem_u32(smem_p)), p_lo_step = (uint32_t)p.patch_stage_bytes >> 4;
      const uint32_t w_lo0 = umma_desc_lo(smem_u32(smem_w));
      int ps = 0, ws = 0;
      uint32_t pphase = 0, wphase = 0;
      int it = 0;
      for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        TWAIT(4, mbar_wait(&acc_empty[acc], acc_phase ^ 1));
        tc_fence_after();
        const uint32_t d = tmem_u + acc * NPIX;
        int ti = 0;
        for (int pi = 0; pi < p.n_patches; ++pi) {
          const int nt = p.patches[pi].ntaps;
          TWAIT(2, mbar_wait(&p_full[ps], pphase));
          DAVO_MMA_FENCE();
          const uint32_t pa = p_lo0 + (uint32_t)ps * p_lo_step;
          for (int t = 0; t < nt; ++t, ++ti) {
            const uint32_t x_lo = pa + (uint32_t)p.taps[ti].a_off * (kSlabBytes / 16);
            TWAIT(3, mbar_wait(&w_full[ws], wphase));
            DAVO_MMA_FENCE();
            if (elect_one()) {
              tc_mma_tf32_slab(d, w_lo0 + (uint32_t)ws * (kWBytes / 16), w_hi, x_lo, x_hi, idesc, ti != 0);
              if constexpr (CLUSTER) tc_commit_multicast(&w_empty[ws], (uint16_t)3);   // both producers wait for both MMA warps
              else tc_commit(&w_empty[ws]);   // frees the weight slot when these MMAs retire
            }
            if (++ws == WS) { ws = 0; wphase ^= 1; }
          }
          if (elect_one()) tc_commit(&p_empty[ps]);       // frees the patch slot
          if (++ps == PS) { ps = 0; pphase ^= 1; }
        }
        if (elect_one()) tc_commit(&acc_full[acc]);       // accumulator complete -> epilogue
        __syncwarp();
      }
    }
  } else if (EPI == EPI_STORE_RELU || warp < 6) {
    // ----------------------------------------------------------- epilogue --
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = warp > 6 ? 1 : 0;      // which of the quadrant's two warps (store epilogue)
    int it = 0;
    for (int tile = first; tile < p.num_tiles; tile += step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const TileCoord tc = decode(tile);
      const int co = tc.mb * kBlockM + q * 32 + lane;          // this thread's output channel
      const bool co_ok = co < p.cout_g;
      const float bias = co_ok ? __ldg(p.bias + tc.g * p.cout_g + co) : 0.f;
      TWAIT(5, mbar_wait(&acc_full[acc], acc_phase));
      tc_fence_after();
#ifdef DAVO_TIMING
      const long long t_busy0 = clock64();
#endif
      const uint32_t t0 = tmem_base + acc * NPIX + (uint32_t(q * 32) << 16);
      const bool warp_has_work = tc.mb * kBlockM + q * 32 < p.cout_g;   // warp-uniform
      if constexpr (EPI == EPI_STORE_RELU) {
        // lane = channel: one store instruction writes 32 consecutive channels of one pixel
        const size_t pix_stride = p.out_stride;
        float* const obase = p.out + ((size_t)tc.n * p.Hout * p.Wout) * pix_stride + tc.g * p.cout_g + co;
        if constexpr (STAGED) {
          // lane = channel, registers = 32 pixels (4 tile rows x 8 cols): transposed into this
          // warp's buffer as [pixel][32 channels] (128-B rows, 16-B chunks XOR-swizzled by
          // pixel & 7 = the tensor map's swizzle; a warp's 32 scalar stores hit 32 banks), then
          // one TMA store of the box {32 channels, 8 cols, 4 rows}; TMA clips outside the image.
          const uint32_t stg = smem_u32(epi_stage) + (half * 4 + q) * 4096;
          uint32_t lane_off[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) lane_off[k] = ((((uint32_t)lane >> 2) ^ k) << 4) | ((lane & 3) << 2);
          if (warp_has_work) {
#pragma unroll 1
            for (int n0 = half * 32; n0 < NPIX; n0 += 64) {
              uint32_t v[32];
              tmem_ld_32x32(t0 + n0, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j)
                v[j] = __float_as_uint(epi_out(__uint_as_float(v[j]) + bias, ea));
              if (lane == 0) tma_store_wait_read<0>();    // the previous store has read the buffer
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 32; ++j)
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg + j * 128 + lane_off[j & 7]), "r"(v[j]) : "memory");
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&tmO, reinterpret_cast<const void*>(epi_stage + (half * 4 + q) * 4096),
                             tc.g * p.cout_g + tc.mb * kBlockM + q * 32, tc.w0, tc.h0 + (n0 >> 3), tc.n);
                tma_store_commit();
              }
            }
          }
        } else if (warp_has_work) {
#pragma unroll 1
          for (int n0 = half * 32; n0 < NPIX; n0 += 64) {     // 4 tile rows x 8 cols per step
            uint32_t v[32];
            tmem_ld_32x32(t0 + n0, v);
            tmem_ld_wait();
            const int hb = tc.h0 + (n0 >> 3);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int hh = hb + (j >> 3), ww = tc.w0 + (j & 7);
              if (co_ok && hh < p.Hout && ww < p.Wout)
                obase[((size_t)hh * p.Wout + ww) * pix_stride] =
                    epi_out(__uint_as_float(v[j]) + bias, ea);
            }
          }
        }
      } else {
        // Spatial-sum epilogue (cnv7 -> pred -> mean, reference nets/posenn.py:239-241: pred
        // is linear, so only sum_pixels relu(cnv7) is needed).  A thread owns a channel, so
        // the sum over the tile's pixels is thread-local; one deterministic partial per tile,
        // added in fixed order by the head kernel.  No atomics.
        float s = 0.f;
        if (warp_has_work) {
#pragma unroll 1
          for (int n0 = 0; n0 < NPIX; n0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(t0 + n0, v);
            tmem_ld_wait();
            const int hb = tc.h0 + (n0 >> 3);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int hh = hb + (j >> 3), ww = tc.w0 + (j & 7);
              if (hh < p.Hout && ww < p.Wout) s += fmaxf(__uint_as_float(v[j]) + bias, 0.f);
            }
          }
        }
        if (co_ok)
          p.sum_out[((size_t)(tc.n * p.groups + tc.g) * (p.tiles_h * p.tiles_w) + tc.t) * p.cout_g + co] = s;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
#ifdef DAVO_TIMING
      tacc[6] += clock64() - t_busy0;
#endif
    }
    if constexpr (STAGED) {
      if (lane == 0) tma_store_wait_read<0>();            // shared memory must outlive the last stores
    }
  }
#ifdef DAVO_TIMING
  if (lane == 0 && (warp <= 2 || warp == 6) && blockIdx.x < 148) {
    long long* o = g_conv_timing + blockIdx.x * 8;
    if (warp == 0) o[0] = tacc[0];
    if (warp == 6) o[1] = tacc[1];
    if (warp == 1) { o[2] = tacc[2]; o[3] = tacc[3]; o[4] = tacc[4]; }
    if (warp == 2) { o[5] = tacc[5]; o[6] = tacc[6]; o[7] = clock64() - t_cta0; }
  }
#endif

  tc_fence_before();
  __syncthreads();
  if constexpr (CLUSTER) cluster_sync();          // the peer may still be writing my ring / arriving on my barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace cm
}  // namespace davo
