// Weight gradient as a pixel-reduction GEMM:
//   dW[cb][k=(ca,t)] = sum_m row[m][cb] * gathered[m][k]
// Both operands are pixel-contiguous in NCHW, so shared memory keeps the pixel
// index innermost ([channel][16 pixels], pitch 20 floats) and each thread reads
// float4s along the pixel axis for TMW row-channels x TNW taps (interleaved
// channel assignment keeps the float4 reads bank-conflict free).  The pixel
// range is split over grid.z; partial tiles are combined with fp32 RED atomics
// into the zero-initialised gradient.
#include <stdlib.h>

#include "rcv_common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MC = 16;   // pixels per staged chunk
constexpr int MCP = 20;  // smem pitch (floats)

template <int BMW, int BNW, int TMW, int TNW>
__global__ void __launch_bounds__(NT) wgrad_kernel(const RcvWgrad p) {
  rcv_pdl_enter();
  constexpr int TXN = BMW / TMW, TYN = BNW / TNW;
  static_assert(TXN * TYN == NT, "thread grid");
  constexpr int DQ = (BMW * 4 + NT - 1) / NT;  // row-tensor quads per thread
  constexpr int AQ = (BNW * 4 + NT - 1) / NT;  // gathered quads per thread
  __shared__ __align__(16) float Ds[2][BMW][MCP];
  __shared__ __align__(16) float As[2][BNW][MCP];
  __shared__ int2 tabA[BNW];
  __shared__ int tabW[BNW];

  const int tid = threadIdx.x;
  const int T = p.taps.n;
  const int K = p.CA * T;
  const int HWin = p.Hin * p.Win;
  const int HWg = p.Hg * p.Wg;
  const int M = p.N * HWg;
  const int k0 = blockIdx.x * BNW;
  const int cb0 = blockIdx.y * BMW;
  const int mbeg = blockIdx.z * p.slab;
  const int mend = min(M, mbeg + p.slab);
  const bool vec = (HWg & 3) == 0;

  for (int kl = tid; kl < BNW; kl += NT) {
    int k = k0 + kl;
    if (k < K) {
      int ca = k / T, t = k - ca * T;
      int dy = p.taps.dy[t], dx = p.taps.dx[t];
      tabA[kl] = make_int2(ca * HWin + dy * p.Win + dx, ((dy + 16) << 8) | (dx + 16));
      tabW[kl] = ca * p.wsA + p.taps.wi[t];
    } else {
      tabA[kl] = make_int2(0, 0);
      tabW[kl] = 0;
    }
  }
  __syncthreads();

  // This thread always stages pixel quad q of every chunk: pixels mq..mq+3.
  const int q = tid & 3;
  int mq = mbeg + q * 4;
  int pn, pi, pj;  // (image, grid row, grid col) of pixel mq
  {
    int mm = mq < M ? mq : 0;
    pn = mm / HWg;
    int r = mm - pn * HWg;
    pi = r / p.Wg;
    pj = r - pi * p.Wg;
  }

  float4 rd[DQ];
  float4 ra[AQ];

  auto load_chunk = [&]() {
    // coordinates of the four pixels
    int en[4], ei[4], ej[4];
    en[0] = pn; ei[0] = pi; ej[0] = pj;
#pragma unroll
    for (int e = 1; e < 4; ++e) {
      en[e] = en[e - 1]; ei[e] = ei[e - 1]; ej[e] = ej[e - 1] + 1;
      if (ej[e] >= p.Wg) { ej[e] = 0; ei[e]++; if (ei[e] >= p.Hg) { ei[e] = 0; en[e]++; } }
    }
    bool ok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) ok[e] = (mq + e) < mend;
    // dense row tensor
#pragma unroll
    for (int i = 0; i < DQ; ++i) {
      int el = tid + i * NT;
      int cl = el >> 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cl < BMW && cb0 + cl < p.CB) {
        const int cb = cb0 + cl;
        if (vec && ok[3]) {
          v = __ldg(reinterpret_cast<const float4*>(
              p.row + ((size_t)en[0] * p.CB + cb) * HWg + ei[0] * p.Wg + ej[0]));
        } else {
          float t4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            t4[e] = ok[e] ? __ldg(p.row + ((size_t)en[e] * p.CB + cb) * HWg + ei[e] * p.Wg + ej[e]) : 0.f;
          v = make_float4(t4[0], t4[1], t4[2], t4[3]);
        }
      }
      rd[i] = v;
    }
    // gathered tensor
#pragma unroll
    for (int i = 0; i < AQ; ++i) {
      int el = tid + i * NT;
      int kl = el >> 2;
      float t4[4] = {0.f, 0.f, 0.f, 0.f};
      if (kl < BNW && k0 + kl < K) {
        const int2 te = tabA[kl];
        const int dy = (te.y >> 8) - 16, dx = (te.y & 0xff) - 16;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int iy = ei[e] * p.gs + dy, ix = ej[e] * p.gs + dx;
          if (ok[e] && (unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win)
            t4[e] = __ldg(p.src + (size_t)en[e] * p.CA * HWin + te.x + (ei[e] * p.gs) * p.Win +
                          ej[e] * p.gs);
        }
      }
      ra[i] = make_float4(t4[0], t4[1], t4[2], t4[3]);
    }
    // advance to the next chunk
    mq += MC;
    pj += MC;
    while (pj >= p.Wg) { pj -= p.Wg; pi++; }
    while (pi >= p.Hg) { pi -= p.Hg; pn++; }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int i = 0; i < DQ; ++i) {
      int el = tid + i * NT;
      int cl = el >> 2;
      if (cl < BMW) *reinterpret_cast<float4*>(&Ds[buf][cl][q * 4]) = rd[i];
    }
#pragma unroll
    for (int i = 0; i < AQ; ++i) {
      int el = tid + i * NT;
      int kl = el >> 2;
      if (kl < BNW) *reinterpret_cast<float4*>(&As[buf][kl][q * 4]) = ra[i];
    }
  };

  const int tx = tid % TXN;
  const int ty = tid / TXN;
  float acc[TMW][TNW];
#pragma unroll
  for (int i = 0; i < TMW; ++i)
#pragma unroll
    for (int j = 0; j < TNW; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  const bool do_bias = (p.dbias != nullptr) && blockIdx.x == 0 && tid < BMW;

  if (mbeg < mend) {
    load_chunk();
    store_chunk(0);
  }
  __syncthreads();

  for (int mc = mbeg, it = 0; mc < mend; mc += MC, ++it) {
    const int buf = it & 1;
    const bool more = mc + MC < mend;
    if (more) load_chunk();
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      float4 a[TMW], b[TNW];
#pragma unroll
      for (int i = 0; i < TMW; ++i)
        a[i] = *reinterpret_cast<const float4*>(&Ds[buf][tx + TXN * i][qq * 4]);
#pragma unroll
      for (int j = 0; j < TNW; ++j)
        b[j] = *reinterpret_cast<const float4*>(&As[buf][ty + TYN * j][qq * 4]);
#pragma unroll
      for (int i = 0; i < TMW; ++i)
#pragma unroll
        for (int j = 0; j < TNW; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
    if (do_bias) {
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const float4 v = *reinterpret_cast<const float4*>(&Ds[buf][tid][qq * 4]);
        bsum += (v.x + v.y) + (v.z + v.w);
      }
    }
    if (more) store_chunk(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TMW; ++i) {
    const int cb = cb0 + tx + TXN * i;
    if (cb >= p.CB) continue;
#pragma unroll
    for (int j = 0; j < TNW; ++j) {
      const int kl = ty + TYN * j;
      if (k0 + kl < K) atomicAdd(p.dw + (size_t)cb * p.wsB + tabW[kl], acc[i][j]);
    }
  }
  if (do_bias && cb0 + tid < p.CB) atomicAdd(p.dbias + cb0 + tid, bsum);
}


// ---------------------------------------------------------------------------------------------
// Tiny weight gradients (<= 8 dense channels: the 3->8 input layer, the 8->5 class head): the
// whole result is a few dozen numbers reduced over ~10^6 pixels, i.e. a streaming reduction.
// One thread = one pixel at a time (lanes along pixels: coalesced), T x CBP x CAG accumulators in
// registers for CAG gathered channels, then shuffle + shared-memory reduction and one atomic per
// weight per CTA.  blockIdx.y walks the gathered channels in groups of CAG.
template <int T, int CBP, int CAG>
__global__ void __launch_bounds__(NT) small_wgrad_kernel(const RcvWgrad p) {
  rcv_pdl_enter();
  constexpr int NACC = CAG * T * CBP;
  __shared__ float red[NT / 32][NACC + CBP];
  const int tid = threadIdx.x;
  const int HWin = p.Hin * p.Win;
  const int HWg = p.Hg * p.Wg;
  const int M = p.N * HWg;
  const int ca0 = blockIdx.y * CAG;
  const bool do_bias = p.dbias != nullptr && blockIdx.y == 0;
  float acc[CAG][T][CBP];
  float bs[CBP];
#pragma unroll
  for (int g = 0; g < CAG; ++g)
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int c = 0; c < CBP; ++c) acc[g][t][c] = 0.f;
#pragma unroll
  for (int c = 0; c < CBP; ++c) bs[c] = 0.f;

  for (int m = blockIdx.x * NT + tid; m < M; m += gridDim.x * NT) {
    const int n = m / HWg, r = m - n * HWg;
    const int i = r / p.Wg, j = r - i * p.Wg;
    float dyv[CBP];
#pragma unroll
    for (int c = 0; c < CBP; ++c) {
      dyv[c] = c < p.CB ? __ldg(p.row + ((size_t)n * p.CB + c) * HWg + r) : 0.f;
      bs[c] += dyv[c];
    }
    const int gy0 = i * p.gs, gx0 = j * p.gs;
#pragma unroll
    for (int g = 0; g < CAG; ++g) {
      const int ca = ca0 + g;
      if (ca < p.CA) {
        const float* src = p.src + ((size_t)n * p.CA + ca) * HWin;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int iy = gy0 + p.taps.dy[t], ix = gx0 + p.taps.dx[t];
          const float x = ((unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win)
                              ? __ldg(src + iy * p.Win + ix) : 0.f;
#pragma unroll
          for (int c = 0; c < CBP; ++c) acc[g][t][c] = fmaf(x, dyv[c], acc[g][t][c]);
        }
      }
    }
  }

  const int lane = tid & 31, wi = tid >> 5;
#pragma unroll
  for (int g = 0; g < CAG; ++g)
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int c = 0; c < CBP; ++c) {
        float v = acc[g][t][c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[wi][(g * T + t) * CBP + c] = v;
      }
#pragma unroll
  for (int c = 0; c < CBP; ++c) {
    float v = bs[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wi][NACC + c] = v;
  }
  __syncthreads();
  for (int e = tid; e < NACC + CBP; e += NT) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) v += red[w][e];
    if (e < NACC) {
      const int c = e % CBP, t = (e / CBP) % T, g = e / (CBP * T);
      const int ca = ca0 + g;
      if (c < p.CB && ca < p.CA) atomicAdd(p.dw + (size_t)c * p.wsB + (size_t)ca * p.wsA + p.taps.wi[t], v);
    } else if (do_bias) {
      const int c = e - NACC;
      if (c < p.CB) atomicAdd(p.dbias + c, v);
    }
  }
}

template <int T, int CBP, int CAG>
int launch_small(const RcvWgrad& p, cudaStream_t st) {
  const int64_t M = (int64_t)p.N * p.Hg * p.Wg;
  RCV_REQUIRE(M < (1ll << 31), RCV_ERR_UNSUPPORTED, "wgrad: problem too large");
  const int ygroups = rcv_cdiv(p.CA, CAG);
  int xblocks = rcv_cdiv(148 * 4, ygroups);
  const int maxx = rcv_cdiv(M, NT);
  if (xblocks > maxx) xblocks = maxx;
  if (xblocks < 1) xblocks = 1;
  dim3 grid(xblocks, ygroups);
  rcv_launch(small_wgrad_kernel<T, CBP, CAG>, dim3(grid), dim3(NT), 0, st, p);
  RCV_CHECK_LAUNCH("small_wgrad_kernel");
  return RCV_OK;
}

template <int BMW, int BNW, int TMW, int TNW>
int launch_cfg(RcvWgrad p, cudaStream_t st) {
  const int K = p.CA * p.taps.n;
  const int64_t M = (int64_t)p.N * p.Hg * p.Wg;
  RCV_REQUIRE(M < (1ll << 31), RCV_ERR_UNSUPPORTED, "wgrad: problem too large");
  const int tiles = rcv_cdiv(K, BNW) * rcv_cdiv(p.CB, BMW);
  int splits = rcv_cdiv(148 * 4, tiles);
  const int max_splits = rcv_cdiv(M, MC * 8);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int slab = rcv_cdiv(M, splits);
  slab = ((slab + MC - 1) / MC) * MC;
  splits = rcv_cdiv(M, slab);
  p.slab = slab;
  dim3 grid(rcv_cdiv(K, BNW), rcv_cdiv(p.CB, BMW), splits);
  rcv_launch(wgrad_kernel<BMW, BNW, TMW, TNW>, dim3(grid), dim3(NT), 0, st, p);
  RCV_CHECK_LAUNCH("wgrad_kernel");
  return RCV_OK;
}

}  // namespace

// Which engine runs a weight-gradient problem (rcv_engine).  RCV_NARROW_WGRAD=0: A/B runs.
int rcv_pick_wgrad_engine(const RcvWgrad& p) {
  static const int use_narrow = getenv("RCV_NARROW_WGRAD") ? atoi(getenv("RCV_NARROW_WGRAD")) : 1;
  if (p.math == RCV_MATH_TF32X3) return RCV_ENGINE_UMMA;
  // few channels on the dense side: TMA-staged register-accumulating kernel
  if (use_narrow && rcv_narrow_wgrad_supported(p)) return RCV_ENGINE_NARROW;
  if (rcv_math_auto(p.math) && rcv_umma_wgrad_pays(p)) return RCV_ENGINE_UMMA;
  return RCV_ENGINE_SIMT;
}

int rcv_launch_wgrad(const RcvWgrad& p, cudaStream_t st) {
  const int eng = rcv_pick_wgrad_engine(p);
  RCV_REQUIRE(p.in_scale == nullptr || eng == RCV_ENGINE_UMMA, RCV_ERR_UNSUPPORTED,
              "wgrad: normalise-on-load is a feature of the tensor-core quad-gather kernel only");
  if (eng == RCV_ENGINE_NARROW) return rcv_launch_narrow_wgrad(p, st);
  if (eng == RCV_ENGINE_UMMA) return rcv_launch_wgrad_umma(p, st);
  if (p.CB <= 8 && p.taps.n == 9 && p.CA <= 16) return launch_small<9, 8, 1>(p, st);
  if (p.CB <= 8 && p.taps.n == 1 && p.CA <= 64) return launch_small<1, 8, 8>(p, st);
  if (p.CB > 64) return launch_cfg<128, 128, 8, 8>(p, st);
  if (p.CB > 32) return launch_cfg<64, 128, 4, 8>(p, st);
  if (p.CB > 16) return launch_cfg<32, 128, 4, 4>(p, st);
  if (p.CB > 8) return launch_cfg<16, 128, 2, 4>(p, st);
  return launch_cfg<8, 64, 1, 2>(p, st);
}
