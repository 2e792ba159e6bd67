/*
 * CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into librcv_b200.so).
 *
 * Plain-C restatement, with double accumulation, of the primitive operations the reference's
 * hot path delegates to PyTorch ATen (the reference pins no torch version, README.md:10-15;
 * definitions follow the published semantics of torch.nn.functional and were checked against
 * torch 2.11 CPU by tests/test_oracle_c.py):
 *   ref_conv2d            F.conv2d           model.py:112,130-133,170,259,411,554
 *   ref_conv_transpose2d  F.conv_transpose2d model.py:186-187 (k3, s2, p1, op1)
 *   ref_bn_train/eval     F.batch_norm       model.py:113,134,171,188
 *   ref_maxpool2x2        F.max_pool2d       model.py:97
 *   ref_weighted_ce       CrossEntropyLoss2d model.py:76-82
 *   ref_argmax_confusion  torch.max + the per-image confusion loop, train.py:128-153
 * Layout: fp32 NCHW contiguous, labels int64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX4(n, c, h, w, C, H, W) ((((int64_t)(n) * (C) + (c)) * (H) + (h)) * (W) + (w))

/* y[N,Co,Ho,Wo] = cross-correlation of x[N,Ci,H,W] with w[Co,Ci,k,k] (+ bias), zero padding. */
void ref_conv2d(const float* x, const float* w, const float* bias, float* y, int N, int Ci, int H, int W,
                int Co, int k, int stride, int pad, int dil) {
  const int Ho = (H + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  const int Wo = (W + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  for (int n = 0; n < N; ++n)
    for (int co = 0; co < Co; ++co)
      for (int oy = 0; oy < Ho; ++oy)
        for (int ox = 0; ox < Wo; ++ox) {
          double acc = bias ? (double)bias[co] : 0.0;
          for (int ci = 0; ci < Ci; ++ci)
            for (int ky = 0; ky < k; ++ky) {
              const int iy = oy * stride - pad + ky * dil;
              if (iy < 0 || iy >= H) continue;
              for (int kx = 0; kx < k; ++kx) {
                const int ix = ox * stride - pad + kx * dil;
                if (ix < 0 || ix >= W) continue;
                acc += (double)x[IDX4(n, ci, iy, ix, Ci, H, W)] *
                       (double)w[(((int64_t)co * Ci + ci) * k + ky) * k + kx];
              }
            }
          y[IDX4(n, co, oy, ox, Co, Ho, Wo)] = (float)acc;
        }
}

/* ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1): scatter form.
 * y[N,Co,2H,2W]; w[Ci,Co,3,3]; out coordinate = 2*in + k - 1. */
void ref_conv_transpose2d(const float* x, const float* w, const float* bias, float* y, int N, int Ci, int H,
                          int W, int Co) {
  const int Ho = 2 * H, Wo = 2 * W;
  double* acc = (double*)calloc((size_t)N * Co * Ho * Wo, sizeof(double));
  for (int n = 0; n < N; ++n)
    for (int ci = 0; ci < Ci; ++ci)
      for (int iy = 0; iy < H; ++iy)
        for (int ix = 0; ix < W; ++ix) {
          const double v = x[IDX4(n, ci, iy, ix, Ci, H, W)];
          for (int co = 0; co < Co; ++co)
            for (int ky = 0; ky < 3; ++ky) {
              const int oy = 2 * iy + ky - 1;
              if (oy < 0 || oy >= Ho) continue;
              for (int kx = 0; kx < 3; ++kx) {
                const int ox = 2 * ix + kx - 1;
                if (ox < 0 || ox >= Wo) continue;
                acc[IDX4(n, co, oy, ox, Co, Ho, Wo)] += v * (double)w[(((int64_t)ci * Co + co) * 3 + ky) * 3 + kx];
              }
            }
        }
  for (int n = 0; n < N; ++n)
    for (int co = 0; co < Co; ++co)
      for (int64_t p = 0; p < (int64_t)Ho * Wo; ++p) {
        const int64_t i = ((int64_t)n * Co + co) * Ho * Wo + p;
        y[i] = (float)(acc[i] + (bias ? (double)bias[co] : 0.0));
      }
  free(acc);
}

/* BatchNorm2d eval: y = (x-mean)/sqrt(var+eps)*gamma+beta. */
void ref_bn_eval(const float* x, float* y, int N, int C, int64_t HW, const float* gamma, const float* beta,
                 const float* mean, const float* var, float eps) {
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      const double inv = 1.0 / sqrt((double)var[c] + (double)eps);
      for (int64_t p = 0; p < HW; ++p) {
        const int64_t i = ((int64_t)n * C + c) * HW + p;
        y[i] = (float)(((double)x[i] - (double)mean[c]) * inv * (double)gamma[c] + (double)beta[c]);
      }
    }
}

/* BatchNorm2d train: batch mean, biased variance for the normalisation; running stats updated with
 * momentum and the unbiased variance. */
void ref_bn_train(const float* x, float* y, int N, int C, int64_t HW, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, float* save_mean,
                  float* save_invstd) {
  const double cnt = (double)N * (double)HW;
  for (int c = 0; c < C; ++c) {
    double s = 0.0, s2 = 0.0;
    for (int n = 0; n < N; ++n)
      for (int64_t p = 0; p < HW; ++p) s += (double)x[((int64_t)n * C + c) * HW + p];
    const double mean = s / cnt;
    for (int n = 0; n < N; ++n)
      for (int64_t p = 0; p < HW; ++p) {
        const double d = (double)x[((int64_t)n * C + c) * HW + p] - mean;
        s2 += d * d;
      }
    const double var = s2 / cnt;
    const double inv = 1.0 / sqrt(var + (double)eps);
    for (int n = 0; n < N; ++n)
      for (int64_t p = 0; p < HW; ++p) {
        const int64_t i = ((int64_t)n * C + c) * HW + p;
        y[i] = (float)(((double)x[i] - mean) * inv * (double)gamma[c] + (double)beta[c]);
      }
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    if (running_var) running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * (s2 / (cnt - 1.0)));
    if (save_mean) save_mean[c] = (float)mean;
    if (save_invstd) save_invstd[c] = (float)inv;
  }
}

/* MaxPool2d(2,2): first maximum in row-major window order, NaN wins (ATen max_pool2d). */
void ref_maxpool2x2(const float* x, float* y, int64_t* idx, int N, int C, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  for (int64_t pl = 0; pl < (int64_t)N * C; ++pl)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox) {
        float best = -INFINITY;
        int64_t bi = (int64_t)(2 * oy) * W + 2 * ox;
        for (int dy = 0; dy < 2; ++dy)
          for (int dx = 0; dx < 2; ++dx) {
            const int64_t i = (int64_t)(2 * oy + dy) * W + 2 * ox + dx;
            const float v = x[pl * H * W + i];
            if (v > best || isnan(v)) { best = v; bi = i; }
          }
        y[(pl * Ho + oy) * Wo + ox] = best;
        if (idx) idx[(pl * Ho + oy) * Wo + ox] = bi;
      }
}

/* loss = sum_p w[y_p] * (-log softmax(z_p)[y_p]) / sum_p w[y_p]; returns the loss and, if
 * dlogits != NULL, its gradient. */
double ref_weighted_ce(const float* logits, const int64_t* target, const float* w, int N, int C, int64_t HW,
                       float* dlogits) {
  double num = 0.0, den = 0.0;
  for (int n = 0; n < N; ++n)
    for (int64_t p = 0; p < HW; ++p) {
      const int64_t y = target[(int64_t)n * HW + p];
      double mx = -INFINITY, se = 0.0;
      for (int c = 0; c < C; ++c) mx = fmax(mx, (double)logits[((int64_t)n * C + c) * HW + p]);
      for (int c = 0; c < C; ++c) se += exp((double)logits[((int64_t)n * C + c) * HW + p] - mx);
      const double wy = w ? (double)w[y] : 1.0;
      num += wy * (mx + log(se) - (double)logits[((int64_t)n * C + y) * HW + p]);
      den += wy;
    }
  if (dlogits)
    for (int n = 0; n < N; ++n)
      for (int64_t p = 0; p < HW; ++p) {
        const int64_t y = target[(int64_t)n * HW + p];
        double mx = -INFINITY, se = 0.0;
        for (int c = 0; c < C; ++c) mx = fmax(mx, (double)logits[((int64_t)n * C + c) * HW + p]);
        for (int c = 0; c < C; ++c) se += exp((double)logits[((int64_t)n * C + c) * HW + p] - mx);
        const double wy = w ? (double)w[y] : 1.0;
        for (int c = 0; c < C; ++c) {
          const double sm = exp((double)logits[((int64_t)n * C + c) * HW + p] - mx) / se;
          dlogits[((int64_t)n * C + c) * HW + p] = (float)(wy * (sm - (c == y ? 1.0 : 0.0)) / den);
        }
      }
  return num / den;
}

/* argmax over classes (lowest index among equal maxima) and per-image confusion
 * conf[n][p][l] = #(argmax == p && target == l); returns the number of correct pixels. */
int64_t ref_argmax_confusion(const float* logits, const int64_t* target, int N, int C, int64_t HW,
                             int64_t* argmax, int64_t* conf) {
  int64_t correct = 0;
  if (conf) memset(conf, 0, sizeof(int64_t) * (size_t)N * C * C);
  for (int n = 0; n < N; ++n)
    for (int64_t p = 0; p < HW; ++p) {
      int best = 0;
      float bv = logits[((int64_t)n * C) * HW + p];
      for (int c = 1; c < C; ++c) {
        const float v = logits[((int64_t)n * C + c) * HW + p];
        if (v > bv) { bv = v; best = c; }
      }
      const int64_t y = target[(int64_t)n * HW + p];
      if (argmax) argmax[(int64_t)n * HW + p] = best;
      if (conf && y >= 0 && y < C) conf[((int64_t)n * C + best) * C + y]++;
      correct += (best == y);
    }
  return correct;
}
