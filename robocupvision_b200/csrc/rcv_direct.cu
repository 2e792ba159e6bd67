// Direct convolution on CUDA cores for the narrow outer layers (<= 16 output channels, short
// reductions: the 3->8 input layer, 8->16 stride-2, 16->8 transposed, the 1x1 class head and the
// input gradients of their neighbours).  These layers are bandwidth-shaped (1.5-12 FLOP/B): the
// implicit-GEMM engines pay more for staging operands than the arithmetic is worth.
//
// One thread = PIX (2 or 4) pixel-grid points x ALL output channels in registers.  Lanes are consecutive
// pixels, so every activation load and every NCHW store is coalesced; the 9x tap re-use of the
// input is served by L1.  Weights of the CTA's parity class live in shared memory as
// [k = (channel, tap)][CBP] and are read as broadcast float4 (4 output channels per LDS.128).
// Same RcvIgemm problem description and the same fused epilogue as the GEMM engines: bias,
// ReLU / folded-BN affine in either order, residual, train-mode BatchNorm sum / sum-of-squares.
#include "rcv_common.cuh"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float apply_epi(float v, int epi, float sc, float sh) {
  switch (epi) {
    case RCV_EPI_RELU: return fmaxf(v, 0.f);
    case RCV_EPI_RELU_AFFINE: return fmaf(sc, fmaxf(v, 0.f), sh);
    case RCV_EPI_AFFINE_RELU: return fmaxf(fmaf(sc, v, sh), 0.f);
    case RCV_EPI_AFFINE: return fmaf(sc, v, sh);
    default: return v;
  }
}

template <int CBP, int PIX>
__global__ void __launch_bounds__(NT) direct_conv_kernel(const RcvIgemm p) {
  rcv_pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int cls = blockIdx.z;
  const int T = p.taps[cls].n;
  const int K = p.CA * T;
  float* ws = reinterpret_cast<float*>(smem_raw);            // [K][CBP]
  int2* ktab = reinterpret_cast<int2*>(ws + (size_t)K * CBP);  // [K] (byte offset, tap)
  float* cst = reinterpret_cast<float*>(ktab + K);           // [3][CBP] bias, scale, shift
  __shared__ double red[2][NT / 32][CBP];
  __shared__ int s_ext[4];  // min dy, max dy, min dx, max dx over the taps of this class

  const int tid = threadIdx.x;
  const int HWin = p.Hin * p.Win;
  const int HWg = p.Hg * p.Wg;
  const int M = p.N * HWg;

  for (int e = tid; e < K * CBP; e += NT) {
    const int k = e / CBP, cb = e - k * CBP;
    const int ca = k / T, t = k - ca * T;
    ws[e] = cb < p.CB ? __ldg(p.w + (size_t)cb * p.wsB + (size_t)ca * p.wsA + p.taps[cls].wi[t]) : 0.f;
  }
  for (int k = tid; k < K; k += NT) {
    const int ca = k / T, t = k - ca * T;
    ktab[k] = make_int2(4 * (ca * HWin + p.taps[cls].dy[t] * p.Win + p.taps[cls].dx[t]), t);
  }
  if (tid == 0) {
    int a = 127, b = -127, c = 127, d = -127;
    for (int t = 0; t < T; ++t) {
      a = min(a, (int)p.taps[cls].dy[t]); b = max(b, (int)p.taps[cls].dy[t]);
      c = min(c, (int)p.taps[cls].dx[t]); d = max(d, (int)p.taps[cls].dx[t]);
    }
    s_ext[0] = a; s_ext[1] = b; s_ext[2] = c; s_ext[3] = d;
  }
  if (tid < CBP) {
    const bool in = tid < p.CB;
    cst[tid] = (in && p.bias) ? __ldg(p.bias + tid) : 0.f;
    cst[CBP + tid] = (in && p.scale) ? __ldg(p.scale + tid) : 1.f;
    cst[2 * CBP + tid] = (in && p.shift) ? __ldg(p.shift + tid) : 0.f;
  }
  __syncthreads();

  const char* inb = reinterpret_cast<const char*>(p.in);
  const int m0 = blockIdx.x * (NT * PIX) + tid;
  uint32_t boff[PIX], tapmask[PIX];
  size_t obase[PIX];
  bool mrow[PIX];
  const int HWo = p.Hout * p.Wout;
#pragma unroll
  for (int q = 0; q < PIX; ++q) {
    const int m = m0 + q * NT;
    mrow[q] = m < M;
    int pn = 0, pi = 0, pj = 0;
    if (mrow[q]) {
      pn = m / HWg;
      const int r = m - pn * HWg;
      pi = r / p.Wg;
      pj = r - pi * p.Wg;
    }
    const int gy0 = pi * p.gs, gx0 = pj * p.gs;
    boff[q] = 4u * (uint32_t)(pn * p.CA * HWin + gy0 * p.Win + gx0);
    uint32_t tm = 0;
    const bool interior = mrow[q] && gy0 + s_ext[0] >= 0 && gy0 + s_ext[1] < p.Hin && gx0 + s_ext[2] >= 0 &&
                          gx0 + s_ext[3] < p.Win;
    if (interior) {
      tm = 0xffffffffu;  // every tap reads inside the image (the common case: no per-tap tests)
    } else if (mrow[q]) {
      for (int t = 0; t < T; ++t) {
        const int iy = gy0 + p.taps[cls].dy[t], ix = gx0 + p.taps[cls].dx[t];
        if ((unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win) tm |= 1u << t;
      }
    }
    tapmask[q] = tm;
    const int oy = pi * p.ostep + (cls >> 1), ox = pj * p.ostep + (cls & 1);
    obase[q] = (size_t)pn * p.CB * HWo + (size_t)oy * p.Wout + ox;
  }

  float acc[PIX][CBP];
#pragma unroll
  for (int q = 0; q < PIX; ++q)
#pragma unroll
    for (int c = 0; c < CBP; ++c) acc[q][c] = 0.f;

  bool all_in = true;
#pragma unroll
  for (int q = 0; q < PIX; ++q) all_in = all_in && tapmask[q] == 0xffffffffu;
  const bool fast = __all_sync(0xffffffffu, all_in);  // warp-uniform: no divergence in the hot loop

  auto fma_row = [&](int k, const float (&x)[PIX]) {
    const float4* wr = reinterpret_cast<const float4*>(ws + (size_t)k * CBP);
#pragma unroll
    for (int c4 = 0; c4 < CBP / 4; ++c4) {
      const float4 w4 = wr[c4];
#pragma unroll
      for (int q = 0; q < PIX; ++q) {
        acc[q][4 * c4 + 0] = fmaf(x[q], w4.x, acc[q][4 * c4 + 0]);
        acc[q][4 * c4 + 1] = fmaf(x[q], w4.y, acc[q][4 * c4 + 1]);
        acc[q][4 * c4 + 2] = fmaf(x[q], w4.z, acc[q][4 * c4 + 2]);
        acc[q][4 * c4 + 3] = fmaf(x[q], w4.w, acc[q][4 * c4 + 3]);
      }
    }
  };
  if (fast) {
    const char* b0[PIX];
#pragma unroll
    for (int q = 0; q < PIX; ++q) b0[q] = inb + boff[q];
#pragma unroll 9
    for (int k = 0; k < K; ++k) {
      const int off = ktab[k].x;
      float x[PIX];
#pragma unroll
      for (int q = 0; q < PIX; ++q) x[q] = __ldg(reinterpret_cast<const float*>(b0[q] + off));
      fma_row(k, x);
    }
  } else {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const int2 e = ktab[k];
      float x[PIX];
#pragma unroll
      for (int q = 0; q < PIX; ++q)
        x[q] = ((tapmask[q] >> e.y) & 1u) ? __ldg(reinterpret_cast<const float*>(inb + (boff[q] + (uint32_t)e.x))) : 0.f;
      fma_row(k, x);
    }
  }

  // ---------------- epilogue ----------------
  const int epi = p.epilogue;
  float s1[CBP], s2[CBP];
#pragma unroll
  for (int c = 0; c < CBP; ++c) s1[c] = s2[c] = 0.f;
  const bool has_res = p.residual != nullptr;
  const int ncb = p.CB;
#pragma unroll
  for (int q = 0; q < PIX; ++q) {
    if (!mrow[q]) continue;
    float* o = p.out + obase[q];
    const float* rs = has_res ? p.residual + obase[q] : nullptr;
#pragma unroll
    for (int c = 0; c < CBP; ++c) {
      if (c < ncb) {
        float v = apply_epi(acc[q][c] + cst[c], epi, cst[CBP + c], cst[2 * CBP + c]);
        if (has_res) v += __ldg(rs);
        *o = v;
        s1[c] += v;
        s2[c] += v * v;
      }
      o += HWo;
      if (has_res) rs += HWo;
    }
  }
  if (p.stats) {
    const int lane = tid & 31, wi = tid >> 5;
#pragma unroll
    for (int c = 0; c < CBP; ++c) {
      float a = s1[c], b = s2[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane == 0) { red[0][wi][c] = (double)a; red[1][wi][c] = (double)b; }
    }
    __syncthreads();
    if (tid < 2 * CBP) {
      const int which = tid / CBP, c = tid - which * CBP;
      if (c < p.CB) {
        double t = 0.0;
        for (int w = 0; w < NT / 32; ++w) t += red[which][w][c];
        atomicAdd(p.stats + which * p.CB + c, t);
      }
    }
  }
}

template <int CBP, int PIX>
int launch_direct(const RcvIgemm& p, cudaStream_t st) {
  int maxT = 0;
  for (int c = 0; c < p.nclass; ++c) maxT = p.taps[c].n > maxT ? p.taps[c].n : maxT;
  const int K = p.CA * maxT;
  const size_t smem = (size_t)K * CBP * 4 + (size_t)K * 8 + 3 * CBP * 4;
  const int64_t M = (int64_t)p.N * p.Hg * p.Wg;
  RCV_REQUIRE(M < (1ll << 31), RCV_ERR_UNSUPPORTED, "direct_conv: problem too large");
  dim3 grid(rcv_cdiv(M, NT * PIX), 1, p.nclass);
  rcv_launch(direct_conv_kernel<CBP, PIX>, dim3(grid), dim3(NT), smem, st, p);
  RCV_CHECK_LAUNCH("direct_conv_kernel");
  return RCV_OK;
}

}  // namespace

bool rcv_direct_supported(const RcvIgemm& p) {
  int maxT = 0;
  for (int c = 0; c < p.nclass; ++c) maxT = p.taps[c].n > maxT ? p.taps[c].n : maxT;
  const int K = p.CA * maxT;
  const int cbp = p.CB <= 8 ? 8 : 16;
  return p.CB <= 16 && (size_t)K * cbp * 4 + (size_t)K * 8 <= 40 * 1024 &&
         (int64_t)p.N * p.CA * p.Hin * p.Win < (1ll << 30);
}

int rcv_launch_direct(const RcvIgemm& p, cudaStream_t st) {
  RCV_REQUIRE(rcv_direct_supported(p), RCV_ERR_UNSUPPORTED, "direct_conv: geometry outside the kernel's limits");
  // pixels per thread amortise the broadcast weight reads: 2 per thread (4 measured no faster on B200)
  return p.CB <= 8 ? launch_direct<8, 2>(p, st) : launch_direct<16, 2>(p, st);
}
