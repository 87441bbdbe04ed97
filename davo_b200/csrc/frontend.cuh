// HBM-bound front end and head of the pose path:
//   se_pool_kernel  : sum over the frame of (|.|, normalised) optical flow      (attention_module.py:66)
//   pack_kernel     : SE fully-connected layers -> 19 class weights, per-pixel
//                     class-weight gather, u8 -> [-1,1] frames, masking, and the
//                     packed PoseNN input                                       (attention_module.py:89-101,
//                                                                                davo.py:1115,1178,1404-1442,1519-1522)
//   head_kernel     : spatial mean + 1x1 pred + 0.01 scale                      (posenn.py:240-250)
// Frame pair index p = 2 * sample + source (0: tgt->src0, 1: tgt->src1), the
// row order of pred_poses (davo.py:1456-1458).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ptx.cuh"

namespace davo {

constexpr int kPoolSplits = 16;
constexpr int kPackedC = 16;     // packed PoseNN input channels: 10 real + 6 zero
constexpr int kNumClasses = 19;

struct FrontParams {
  int H, W;
  int pair0;             // first global frame pair of this micro-batch
  int npairs;
  int in_mode;           // 1: flows are concatenated (v1)
  int att_src;           // 0 none, 1 se_flow, 2 static
  int att_tgt_ones;
  int mask_rgb, mask_flow;
  int se_act;            // 0 relu, 1 tanh, 2 lrelu
  int flow_abs;          // 0 none, 1 both, 2 h, 3 v
  int flow_norm;
  const uint8_t* img;    // [B][H][3W][3]
  const float* flow;     // [B][4][H][W][2]
  const float* seg;      // [B][3][H][W][1]
  const float* se_w;     // W1[2][8] b1[8] W2[8][19] b2[19]  (195 floats)
  const float* static_w; // sigmoid(seg_channel_weight)[19]
  float* pool_part;      // [mb][kPoolSplits][2]
  float* att_w;          // [mb][19]
  float* packed;         // [mb][H][W][16]
};

__device__ __forceinline__ float se_in_x(float v, const FrontParams& p) {
  if (p.flow_norm) v = (v - 0.32140523f) / 15.384229f;
  if (p.flow_abs == 1 || p.flow_abs == 2) v = fabsf(v);
  return v;
}
__device__ __forceinline__ float se_in_y(float v, const FrontParams& p) {
  if (p.flow_norm) v = (v - 0.32140523f) / 15.384229f;
  if (p.flow_abs == 1 || p.flow_abs == 3) v = fabsf(v);
  return v;
}

// grid (kPoolSplits, npairs), 256 threads: deterministic partial sums of the SE input.
__global__ void __launch_bounds__(256) se_pool_kernel(const FrontParams p) {
  const int pl = blockIdx.y;
  const int pg = p.pair0 + pl;
  const int b = pg >> 1, k = pg & 1;
  const size_t hw = (size_t)p.H * p.W;
  const float4* src = reinterpret_cast<const float4*>(p.flow + ((size_t)b * 4 + k) * hw * 2);
  const int n4 = (int)(hw / 2);                      // float4 = 2 pixels
  const int per = (n4 + kPoolSplits - 1) / kPoolSplits;
  const int beg = blockIdx.x * per;
  const int end = min(beg + per, n4);
  float sx = 0.f, sy = 0.f;
  for (int i = beg + threadIdx.x; i < end; i += 256) {
    const float4 v = __ldg(src + i);
    sx += se_in_x(v.x, p) + se_in_x(v.z, p);
    sy += se_in_y(v.y, p) + se_in_y(v.w, p);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  __shared__ float red[8][2];
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5][0] = sx;
    red[threadIdx.x >> 5][1] = sy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ax = 0.f, ay = 0.f;
    for (int i = 0; i < 8; ++i) { ax += red[i][0]; ay += red[i][1]; }
    p.pool_part[((size_t)pl * kPoolSplits + blockIdx.x) * 2 + 0] = ax;
    p.pool_part[((size_t)pl * kPoolSplits + blockIdx.x) * 2 + 1] = ay;
  }
}

__device__ __forceinline__ float se_activation(float v, int act) {
  if (act == 1) return tanhf(v);
  if (act == 2) return v > 0.f ? v : 0.2f * v;
  return fmaxf(v, 0.f);
}

// grid (blocks_per_pair, npairs), 256 threads; 4 threads per pixel, each writes
// one float4 of the 16-channel packed pixel:
//   [tgt r g b | 0 0 | src r g b | src flow x y | 0 x 6]      (davo.py:1439-1442, posenn.py:198)
__global__ void __launch_bounds__(256) pack_kernel(const FrontParams p) {
  __shared__ float s_fc1[8];
  __shared__ float s_w[kNumClasses];
  const int pl = blockIdx.y;
  const int pg = p.pair0 + pl;
  const int b = pg >> 1, k = pg & 1;
  const size_t hw = (size_t)p.H * p.W;

  if (p.att_src == 1) {
    // SE excitation (attention_module.py:89-101) from the pooled partials.
    if (threadIdx.x < 8) {
      float px = 0.f, py = 0.f;
      for (int s = 0; s < kPoolSplits; ++s) {
        px += p.pool_part[((size_t)pl * kPoolSplits + s) * 2 + 0];
        py += p.pool_part[((size_t)pl * kPoolSplits + s) * 2 + 1];
      }
      const float inv = 1.0f / (float)hw;
      px *= inv;
      py *= inv;
      const float* W1 = p.se_w;            // [2][8]
      const float* b1 = p.se_w + 16;       // [8]
      const int j = threadIdx.x;
      s_fc1[j] = se_activation(px * W1[j] + py * W1[8 + j] + b1[j], p.se_act);
    }
    __syncthreads();
    if (threadIdx.x < kNumClasses) {
      const float* W2 = p.se_w + 24;       // [8][19]
      const float* b2 = p.se_w + 24 + 8 * kNumClasses;
      const int c = threadIdx.x;
      float a = b2[c];
#pragma unroll
      for (int j = 0; j < 8; ++j) a += s_fc1[j] * W2[j * kNumClasses + c];
      const float wgt = 1.0f / (1.0f + expf(-a));
      s_w[c] = wgt;
      if (blockIdx.x == 0) p.att_w[(size_t)pl * kNumClasses + c] = wgt;
    }
  } else if (p.att_src == 2) {
    if (threadIdx.x < kNumClasses) {
      s_w[threadIdx.x] = p.static_w[threadIdx.x];
      if (blockIdx.x == 0) p.att_w[(size_t)pl * kNumClasses + threadIdx.x] = s_w[threadIdx.x];
    }
  } else {
    if (threadIdx.x < kNumClasses) {
      s_w[threadIdx.x] = 1.0f;
      if (blockIdx.x == 0) p.att_w[(size_t)pl * kNumClasses + threadIdx.x] = 1.0f;
    }
  }
  __syncthreads();

  const int t = blockIdx.x * 256 + threadIdx.x;
  const int pix = t >> 2, qtr = t & 3;
  if (pix >= (int)hw) return;
  const int h = pix / p.W, w = pix - h * p.W;
  const int W3 = 3 * p.W;

  // class weight of this pixel in the source frame (davo.py:1115, 1178)
  float a_src = 1.0f;
  if (p.att_src != 0) {
    const float lab_f = __ldg(p.seg + (((size_t)b * 3 + (k == 0 ? 0 : 2)) * hw + pix));
    const int lab = (int)lab_f;                       // tf.cast truncates toward zero
    a_src = (lab >= 0 && lab < kNumClasses) ? s_w[lab] : 0.0f;
  }
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (qtr == 0) {
    const uint8_t* px = p.img + ((size_t)(b * p.H + h) * W3 + (p.W + w)) * 3;   // tgt = centre frame
    float a_tgt = 1.0f;
    if (!p.att_tgt_ones && p.att_src != 0) {
      const int lab = (int)__ldg(p.seg + (((size_t)b * 3 + 1) * hw + pix));
      a_tgt = (lab >= 0 && lab < kNumClasses) ? s_w[lab] : 0.0f;
    }
    const float m = p.mask_rgb ? a_tgt : 1.0f;
    o.x = ((float)px[0] * (1.0f / 255.0f) * 2.0f - 1.0f) * m;
    o.y = ((float)px[1] * (1.0f / 255.0f) * 2.0f - 1.0f) * m;
    o.z = ((float)px[2] * (1.0f / 255.0f) * 2.0f - 1.0f) * m;
  } else if (qtr == 1) {
    const uint8_t* px = p.img + ((size_t)(b * p.H + h) * W3 + ((k == 0 ? 0 : 2 * p.W) + w)) * 3;
    const float m = p.mask_rgb ? a_src : 1.0f;
    o.y = ((float)px[0] * (1.0f / 255.0f) * 2.0f - 1.0f) * m;
    o.z = ((float)px[1] * (1.0f / 255.0f) * 2.0f - 1.0f) * m;
    o.w = ((float)px[2] * (1.0f / 255.0f) * 2.0f - 1.0f) * m;
  } else if (qtr == 2) {
    if (p.in_mode == 1) {
      const float2 f = __ldg(reinterpret_cast<const float2*>(p.flow + ((size_t)b * 4 + k) * hw * 2) + pix);
      const float m = p.mask_flow ? a_src : 1.0f;
      o.x = f.x * m;
      o.y = f.y * m;
    }
  }
  o.x = round_tf32(o.x);
  o.y = round_tf32(o.y);
  o.z = round_tf32(o.z);
  o.w = round_tf32(o.w);
  reinterpret_cast<float4*>(p.packed + ((size_t)pl * hw + pix) * kPackedC)[qtr] = o;
}

struct HeadParams {
  int pair0, npairs;
  int nparts;             // partial rows per (pair, branch) = tiles * 4
  float inv_hw;           // 1 / (H7 * W7)
  const float* sums;      // [mb][2][nparts][256]
  const float* wpred;     // [2][256][3]
  const float* bpred;     // [2][3]
  float* pose_out;        // [B][2][6] -> pair-major [2B][6]
};

// grid npairs, 256 threads.  mean_{h,w} pred(cnv7) == pred(mean_{h,w} cnv7): pred is linear
// (posenn.py:240-241); pose = 0.01 * [rot(3), trans(3)] (posenn.py:248-250).
__global__ void __launch_bounds__(256) head_kernel(const HeadParams p) {
  __shared__ float s_mean[2][256];
  const int pl = blockIdx.x;
  const int c = threadIdx.x;
  for (int br = 0; br < 2; ++br) {
    const float* src = p.sums + ((size_t)(pl * 2 + br) * p.nparts) * 256 + c;
    float a = 0.f;
    for (int i = 0; i < p.nparts; ++i) a += src[(size_t)i * 256];
    s_mean[br][c] = a * p.inv_hw;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 6) {
    const int br = warp / 3, j = warp % 3;
    float a = 0.f;
    for (int i = lane; i < 256; i += 32) a += s_mean[br][i] * p.wpred[(br * 256 + i) * 3 + j];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0)
      p.pose_out[(size_t)(p.pair0 + pl) * 6 + br * 3 + j] = 0.01f * (a + p.bpred[br * 3 + j]);
  }
}

// ------------------------------------------------------------------------------
// Plain fp32 direct convolution (CUDA cores).  NOT on the product path: it is the
// on-GPU cross-check the tests use to tell a tcgen05/TMA descriptor error from a
// host-side packing error (davo_debug_set_conv_impl).
struct DirectConvParams {
  int npairs, Hin, Win, Cin_total, cin_off, Cin;   // input view + channels used
  int Hout, Wout, Cout, out_stride, cout_off;
  int kh, kw, stride, dil, pad_t, pad_l;
  int relu, round_out;
  int cmap[16];          // weight input channel -> input tensor channel (cnv1 only), else identity
  int use_cmap;
  const float* in;       // [n][Hin][Win][Cin_total]
  const float* w;        // HWIO [kh][kw][Cin][Cout]
  const float* bias;     // [Cout]
  float* out;            // [n][Hout][Wout][out_stride]
};

__global__ void __launch_bounds__(256) conv_direct_kernel(const DirectConvParams p) {
  const long long total = (long long)p.npairs * p.Hout * p.Wout * p.Cout;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % p.Cout);
  long long r = idx / p.Cout;
  const int ow = (int)(r % p.Wout);
  r /= p.Wout;
  const int oh = (int)(r % p.Hout);
  const int n = (int)(r / p.Hout);
  float acc = p.bias[co];
  for (int ty = 0; ty < p.kh; ++ty) {
    const int ih = oh * p.stride + ty * p.dil - p.pad_t;
    if (ih < 0 || ih >= p.Hin) continue;
    for (int tx = 0; tx < p.kw; ++tx) {
      const int iw = ow * p.stride + tx * p.dil - p.pad_l;
      if (iw < 0 || iw >= p.Win) continue;
      const float* ip = p.in + ((size_t)(n * p.Hin + ih) * p.Win + iw) * p.Cin_total + p.cin_off;
      const float* wp = p.w + ((size_t)(ty * p.kw + tx) * p.Cin) * p.Cout + co;
      for (int ci = 0; ci < p.Cin; ++ci) {
        const int ic = p.use_cmap ? p.cmap[ci] : ci;
        acc = fmaf(ip[ic], wp[(size_t)ci * p.Cout], acc);
      }
    }
  }
  if (p.relu) acc = fmaxf(acc, 0.f);
  if (p.round_out) acc = round_tf32(acc);
  p.out[((size_t)(n * p.Hout + oh) * p.Wout + ow) * p.out_stride + p.cout_off + co] = acc;
}

// Spatial sum of [n][H][W][C] into the head's partial layout with nparts = 1.
__global__ void __launch_bounds__(256) spatial_sum_kernel(const float* in, int hw, int C, int cstride,
                                                          int coff, float* out /*[n][256]*/) {
  const int n = blockIdx.x, c = threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int i = 0; i < hw; ++i) a += in[((size_t)n * hw + i) * cstride + coff + c];
  out[(size_t)n * 256 + c] = a;
}

}  // namespace davo
