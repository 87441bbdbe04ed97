// "-batch_norm" (reference nets/posenn.py:206): normalizer_fn=slim.batch_norm on every conv of the PoseNN except
// pred.  The reference passes no normalizer_params, so slim.batch_norm runs with its default is_training=True AT
// TEST TIME: every layer is normalised with the mean and (biased) variance of the batch of the CALL it belongs to
// (the shared nets call the PoseNN twice: all tgt->src0 pairs, then all tgt->src1 pairs, davo.py:1456-1457), and
// test_kitti_pose.py:129 restores trainable variables only, so the moving averages never matter.  No biases exist;
// the only extra trainable variable is BatchNorm/beta (scale=False).  epsilon = 0.001.
//
//   y = relu((conv - mean_g[c]) * rsqrt(var_g[c] + 1e-3) + beta[c]),   g = call group of the unit
//
// The conv kernels write plain fp32 sums (ConvParams::raw); three small HBM-bound kernels finish the layer in place:
//   bn_stats_kernel     per-channel sum and sum of squares, double accumulators, fixed partial layout (deterministic)
//   bn_finalize_kernel  partials -> mean, rstd per (group, channel)
//   bn_apply_kernel     normalise, +beta, relu, round to TF32 (the next layer's operand format); the last layer
//                       (cnv7) reduces over the map instead of storing (pred is linear: mean(pred(x)) = pred(mean(x)))
// An ablation path: correctness first, one extra read + write of every activation.
#pragma once
#include "frontend.cuh"

namespace davo {

constexpr int kBnSplits = 64;            // blocks per call group in bn_stats_kernel
constexpr int kBnMaxC = 512;             // cnv7 of the decouple nets: rotation | translation

struct BnParams {
  int units;               // units (frame pairs / samples) of this pass
  int pair0, pair_mode;    // pair_of_slot: a pair unit's call group is its k (source index); sample units: one group
  int ngroups;             // 2: shared nets computing every pair; 1 otherwise
  int H, W, Hp, Wp, C;     // map, its pitch, channels per pixel
  float* x;                // [units][Hp][Wp][C], normalised in place
  const float* beta;       // [C]
  double* part;            // [ngroups][kBnSplits][2][C]
  float* mean;             // [ngroups][C]
  float* rstd;             // [ngroups][C]
  float* sum_out;          // last layer: [units][C] spatial sums of the activated map (head_kernel, nparts = 1)
};

__device__ __forceinline__ int bn_group(const BnParams& p, int u) {
  if (p.ngroups == 1) return 0;
  int b, k;
  pair_of_slot(p.pair_mode, p.pair0 + u, &b, &k);
  return k;
}

// grid (kBnSplits, ngroups), 256 threads.  A thread owns channel(s) c = t % min(C,256) (+256) and every
// (256 / min(C,256))-th pixel of the split's share.
__global__ void __launch_bounds__(256) bn_stats_kernel(const BnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int g = blockIdx.y;
  const int cw = p.C < 256 ? p.C : 256;                 // channels covered by one row of threads
  const int rows = 256 / cw;                            // pixel rows of threads
  const int c = threadIdx.x % cw, r = threadIdx.x / cw;
  const long long npix = (long long)p.units * p.H * p.W;
  const long long per = (npix + kBnSplits - 1) / kBnSplits;
  const long long beg = (long long)blockIdx.x * per, end = min(beg + per, npix);
  double s[2] = {0.0, 0.0}, q[2] = {0.0, 0.0};
  for (long long i = beg + r; i < end; i += rows) {
    const int u = (int)(i / (p.H * p.W));
    if (bn_group(p, u) != g) continue;
    const int rem = (int)(i - (long long)u * p.H * p.W), h = rem / p.W, w = rem - h * p.W;
    const float* px = p.x + (((size_t)u * p.Hp + h) * p.Wp + w) * p.C;
    for (int k = 0, cc = c; cc < p.C; cc += 256, ++k) {
      const double v = (double)px[cc];
      s[k] += v; q[k] += v * v;
    }
  }
  __shared__ double sh[2][256];
  for (int k = 0, cc = c; cc < p.C; cc += 256, ++k) {
    sh[0][threadIdx.x] = s[k]; sh[1][threadIdx.x] = q[k];
    __syncthreads();
    if (r == 0) {
      double a = 0.0, b = 0.0;
      for (int j = 0; j < rows; ++j) { a += sh[0][j * cw + c]; b += sh[1][j * cw + c]; }   // fixed order
      double* o = p.part + (((size_t)g * kBnSplits + blockIdx.x) * 2) * p.C;
      o[cc] = a; o[p.C + cc] = b;
    }
    __syncthreads();
  }
}

// grid (ngroups), 256 threads
__global__ void __launch_bounds__(256) bn_finalize_kernel(const BnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int g = blockIdx.x;
  // pixels of this group: every unit of the pass whose call group is g
  int n_units = 0;
  for (int u = 0; u < p.units; ++u) n_units += bn_group(p, u) == g ? 1 : 0;
  const double n = (double)n_units * p.H * p.W;
  for (int c = threadIdx.x; c < p.C; c += 256) {
    double a = 0.0, b = 0.0;
    for (int sp = 0; sp < kBnSplits; ++sp) {
      const double* o = p.part + (((size_t)g * kBnSplits + sp) * 2) * p.C;
      a += o[c]; b += o[p.C + c];
    }
    const double mean = n > 0 ? a / n : 0.0;
    const double var = n > 0 ? fmax(b / n - mean * mean, 0.0) : 0.0;      // biased variance (fused_batch_norm normalises with it)
    p.mean[g * p.C + c] = (float)mean;
    p.rstd[g * p.C + c] = (float)(1.0 / sqrt(var + 1e-3));                // slim.batch_norm epsilon
  }
}

// one thread per (pixel, 4 channels), in place
__global__ void __launch_bounds__(256) bn_apply_kernel(const BnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = p.C / 4;
  const long long total = (long long)p.units * p.H * p.W * c4n;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % c4n);
  const long long pix = idx / c4n;
  const int u = (int)(pix / (p.H * p.W));
  const int rem = (int)(pix - (long long)u * p.H * p.W), h = rem / p.W, w = rem - h * p.W;
  const int g = bn_group(p, u);
  float4* px = reinterpret_cast<float4*>(p.x + (((size_t)u * p.Hp + h) * p.Wp + w) * p.C) + c4;
  const float4 m = *reinterpret_cast<const float4*>(p.mean + g * p.C + 4 * c4);
  const float4 rs = *reinterpret_cast<const float4*>(p.rstd + g * p.C + 4 * c4);
  const float4 be = *reinterpret_cast<const float4*>(p.beta + 4 * c4);
  const float4 v = *px;
  *px = make_float4(round_tf32(fmaxf((v.x - m.x) * rs.x + be.x, 0.f)), round_tf32(fmaxf((v.y - m.y) * rs.y + be.y, 0.f)),
                    round_tf32(fmaxf((v.z - m.z) * rs.z + be.z, 0.f)), round_tf32(fmaxf((v.w - m.w) * rs.w + be.w, 0.f)));
}

// last layer: grid (units), 256 threads: sum over the map of the activated values, per channel (fixed order)
__global__ void __launch_bounds__(256) bn_apply_sum_kernel(const BnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int u = blockIdx.x, g = bn_group(p, u);
  for (int c = threadIdx.x; c < p.C; c += 256) {
    const float m = p.mean[g * p.C + c], rs = p.rstd[g * p.C + c], be = p.beta[c];
    float a = 0.f;
    for (int h = 0; h < p.H; ++h)
      for (int w = 0; w < p.W; ++w)
        a += fmaxf((p.x[(((size_t)u * p.Hp + h) * p.Wp + w) * p.C + c] - m) * rs + be, 0.f);
    p.sum_out[(size_t)u * p.C + c] = a;
  }
}

}  // namespace davo
