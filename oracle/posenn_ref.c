/*
 * posenn_ref.c -- plain-C, double-precision restatement of DAVO's pose forward
 * path for ONE sample (frame triple).  TEST INFRASTRUCTURE ONLY (see
 * oracle/__init__.py); pinned through oracle/davo_oracle.py, which tests hold to fixtures made by the reference's own graph code.
 *
 * It is written independently of oracle/davo_oracle.py (scalar loops, no
 * library convolution) so that the two restatements check each other on the
 * TF semantics that are easy to get wrong: asymmetric 'SAME' padding, dilation,
 * HWIO weight order, frame unpacking order and the one-hot class gather.
 * Citations are to the reference checkout.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libposenn_ref.so oracle/posenn_ref.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* TF padding='SAME': out = ceil(in/stride); total = max((out-1)*stride + (k-1)*rate + 1 - in, 0);
 * pad_before = total / 2 (the extra pixel goes after). */
static void same_pad(int in, int k, int stride, int rate, int* out, int* before) {
  int o = (in + stride - 1) / stride;
  int total = (o - 1) * stride + (k - 1) * rate + 1 - in;
  if (total < 0) total = 0;
  *out = o;
  *before = total / 2;
}

/* slim.conv2d(x, Cout, [k,k], stride, rate) + bias + optional ReLU (nets/posenn.py:205-215).
 * x [H][W][Cin], w [k][k][Cin][Cout] (HWIO), y [Ho][Wo][Cout]. */
static double* conv2d_same(const double* x, int H, int W, int Cin, const float* w, const float* b,
                           int k, int Cout, int stride, int rate, int relu, int* Ho, int* Wo) {
  int pt, pl;
  same_pad(H, k, stride, rate, Ho, &pt);
  same_pad(W, k, stride, rate, Wo, &pl);
  double* y = (double*)malloc(sizeof(double) * (size_t)(*Ho) * (*Wo) * Cout);
  for (int oh = 0; oh < *Ho; ++oh)
    for (int ow = 0; ow < *Wo; ++ow) {
      double* yp = y + ((size_t)oh * (*Wo) + ow) * Cout;
      for (int co = 0; co < Cout; ++co) yp[co] = b[co];
      for (int ty = 0; ty < k; ++ty) {
        int ih = oh * stride + ty * rate - pt;
        if (ih < 0 || ih >= H) continue;
        for (int tx = 0; tx < k; ++tx) {
          int iw = ow * stride + tx * rate - pl;
          if (iw < 0 || iw >= W) continue;
          const double* xp = x + ((size_t)ih * W + iw) * Cin;
          const float* wp = w + ((size_t)(ty * k + tx) * Cin) * Cout;
          for (int ci = 0; ci < Cin; ++ci) {
            double xv = xp[ci];
            const float* wr = wp + (size_t)ci * Cout;
            for (int co = 0; co < Cout; ++co) yp[co] += xv * (double)wr[co];
          }
        }
      }
      if (relu)
        for (int co = 0; co < Cout; ++co) if (yp[co] < 0) yp[co] = 0;
    }
  return y;
}

typedef struct {
  /* pose_exp_net/cnv{1..5}, pose/{rotation,translation}/{cnv6,cnv7,pred}: HWIO + bias */
  const float *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4, *w5, *b5;
  const float *w6[2], *b6[2], *w7[2], *b7[2], *wp[2], *bp[2];
  /* se_flow/{bottleneck_fc,recover_fc}/{kernel,bias}; static seg_channel_weight/weight */
  const float *se_w1, *se_b1, *se_w2, *se_b2, *static_w;
  int cin1;       /* 10 (v1) or 6 (v0) */
  int c6;         /* cnv6 width */
  int att_src;    /* 0 none, 1 se_flow, 2 static */
  int mask_rgb, mask_flow;
  int se_act;     /* 0 relu 1 tanh 2 lrelu */
  int flow_abs;   /* 0 none 1 both 2 h 3 v */
  int flow_norm;
} posenn_weights;

static double act(double v, int kind) {
  if (kind == 1) return tanh(v);
  if (kind == 2) return v > 0 ? v : 0.2 * v;
  return v > 0 ? v : 0;
}

/* One PoseNN evaluation: decouple_sharednet_v0_dilation (nets/posenn.py:189-254). */
static void posenn(const double* in, int H, int W, const posenn_weights* P, double* pose6) {
  int h, w_, h2, w2;
  double* c1 = conv2d_same(in, H, W, P->cin1, P->w1, P->b1, 7, 16, 2, 1, 1, &h, &w_);      /* :211 */
  double* c2 = conv2d_same(c1, h, w_, 16, P->w2, P->b2, 5, 32, 2, 1, 1, &h2, &w2);         /* :212 */
  double* c3 = conv2d_same(c2, h2, w2, 32, P->w3, P->b3, 3, 64, 1, 2, 1, &h, &w_);         /* :213 */
  double* c4 = conv2d_same(c3, h, w_, 64, P->w4, P->b4, 3, 128, 1, 4, 1, &h, &w_);         /* :214 */
  double* c5 = conv2d_same(c4, h, w_, 128, P->w5, P->b5, 3, 256, 1, 8, 1, &h, &w_);        /* :215 */
  for (int br = 0; br < 2; ++br) {                                                         /* :222 */
    int h7, w7, hp, wq;
    double* c6 = conv2d_same(c5, h, w_, 256, P->w6[br], P->b6[br], 3, P->c6, 1, 2, 1, &hp, &wq);   /* :238 */
    double* c7 = conv2d_same(c6, hp, wq, P->c6, P->w7[br], P->b7[br], 3, 256, 2, 1, 1, &h7, &w7);  /* :239 */
    double* pr = conv2d_same(c7, h7, w7, 256, P->wp[br], P->bp[br], 1, 3, 1, 1, 0, &hp, &wq);      /* :240 */
    for (int j = 0; j < 3; ++j) {
      double a = 0;
      for (int i = 0; i < hp * wq; ++i) a += pr[(size_t)i * 3 + j];
      pose6[br * 3 + j] = 0.01 * a / (hp * wq);                                            /* :241, :250 */
    }
    free(c6); free(c7); free(pr);
  }
  free(c1); free(c2); free(c3); free(c4); free(c5);
}

/* Whole graph for one sample (davo.py:955-1494).  img [H][3W][3] u8, flow [4][H][W][2],
 * seg [3][H][W][1]; pose_out [2][6].  The target map is forced to ones (se_flow /
 * static / none variants, davo.py:1393, 1404-1412). */
int posenn_ref_forward(const uint8_t* img, const float* flow, const float* seg, int H, int W,
                       const posenn_weights* P, double* pose_out, double* att_w_out /*[2][19] or NULL*/) {
  const size_t hw = (size_t)H * W;
  const int C = P->cin1;
  double* in = (double*)malloc(sizeof(double) * hw * C);
  for (int k = 0; k < 2; ++k) {
    /* davo.py:978-982: flows [0, flow[0], flow[1]]; :1000-1004: seg [seg[1], seg[0], seg[2]] */
    const float* fl = flow + (size_t)k * hw * 2;
    const float* sg = seg + (size_t)(k == 0 ? 0 : 2) * hw;
    double w19[19];
    for (int c = 0; c < 19; ++c) w19[c] = 1.0;
    if (P->att_src == 1) {
      /* attention_module.py:66: global mean of the (normalised, abs) flow */
      double px = 0, py = 0;
      for (size_t i = 0; i < hw; ++i) {
        double x = fl[2 * i], y = fl[2 * i + 1];
        if (P->flow_norm) { x = (x - 0.32140523) / 15.384229; y = (y - 0.32140523) / 15.384229; }
        if (P->flow_abs == 1 || P->flow_abs == 2) x = fabs(x);
        if (P->flow_abs == 1 || P->flow_abs == 3) y = fabs(y);
        px += x; py += y;
      }
      px /= (double)hw; py /= (double)hw;
      double fc1[8];
      for (int j = 0; j < 8; ++j)                                         /* :89-94 */
        fc1[j] = act(px * P->se_w1[j] + py * P->se_w1[8 + j] + P->se_b1[j], P->se_act);
      for (int c = 0; c < 19; ++c) {                                      /* :96-101 */
        double a = P->se_b2[c];
        for (int j = 0; j < 8; ++j) a += fc1[j] * P->se_w2[j * 19 + c];
        w19[c] = 1.0 / (1.0 + exp(-a));
      }
    } else if (P->att_src == 2) {
      for (int c = 0; c < 19; ++c) w19[c] = 1.0 / (1.0 + exp(-(double)P->static_w[c]));   /* posenn.py:391 */
    }
    if (att_w_out) memcpy(att_w_out + k * 19, w19, sizeof w19);
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        const size_t i = (size_t)h * W + w;
        double a = 1.0;
        if (P->att_src != 0) {
          int lab = (int)sg[i];                                           /* davo.py:1115 */
          a = (lab >= 0 && lab < 19) ? w19[lab] : 0.0;
        }
        /* data_loader.py:537-557: tgt = centre frame, src0 = left, src1 = right */
        const uint8_t* pt = img + ((size_t)h * 3 * W + (W + w)) * 3;
        const uint8_t* ps = img + ((size_t)h * 3 * W + ((k == 0 ? 0 : 2 * W) + w)) * 3;
        double* o = in + i * C;
        int c = 0;
        for (int j = 0; j < 3; ++j) o[c++] = pt[j] * (1.0 / 255.0) * 2.0 - 1.0;           /* davo.py:1519-1522 */
        if (C == 10) { o[c++] = 0; o[c++] = 0; }
        for (int j = 0; j < 3; ++j) o[c++] = (ps[j] * (1.0 / 255.0) * 2.0 - 1.0) * (P->mask_rgb ? a : 1.0);
        if (C == 10) {
          o[c++] = fl[2 * i] * (P->mask_flow ? a : 1.0);
          o[c++] = fl[2 * i + 1] * (P->mask_flow ? a : 1.0);
        }
      }
    posenn(in, H, W, P, pose_out + k * 6);
  }
  free(in);
  return 0;
}
