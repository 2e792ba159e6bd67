// fp32 implicit-GEMM engine on CUDA cores: convolution forward, input gradient
// and the transposed-convolution family (as four output-parity classes).
//
// GEMM view:  M = pixel-grid points (n,i,j) flattened over the whole batch,
//             N = output channels, K = (input channel, tap).
// A[m][k] is gathered on the fly (im2col never touches HBM): every CTA stages a
// KC x BM slab of gathered activations and a KC x BN slab of weights in shared
// memory (double buffered, global->register->smem so the next slab's loads fly
// under the FMAs), and each thread owns an 8 x TN register tile.  NCHW makes
// both the gather (lanes = consecutive pixels) and the store (float4 of four
// consecutive pixels per channel) coalesced.
//
// Epilogue (fused): + bias, ReLU / folded-BN affine in either order, + residual
// (decoder skip), per-channel sum / sum-of-squares for train-mode BatchNorm.
#include <stdlib.h>

#include "rcv_common.cuh"

namespace {

constexpr int NT = 256;

template <int BM, int BN, int TN, int KC>
struct Cfg {
  static constexpr int TXN = BM / 8;
  static constexpr int TYN = BN / TN;
  static constexpr int BP = BN + 4;
  static constexpr int GP = BM < NT ? BM : NT;            // distinct gather pixels per pass
  static constexpr int PPT = BM / GP;                     // pixels per thread
  static constexpr int KSTR = NT / GP;                    // k interleave between threads
  static constexpr int KPT = KC / KSTR;                   // k's per thread per pixel
  static constexpr int BEL = (KC * BN + NT - 1) / NT;     // weight elements per thread
  static_assert(TXN * TYN == NT, "thread grid");
  static_assert(KPT * KSTR == KC, "k split");
  static size_t smem_bytes(int K) {
    return sizeof(float) * (2 * KC * BM + 2 * KC * BP) + (size_t)K * 12;
  }
};

__device__ __forceinline__ float apply_epi(float v, int epi, float sc, float sh) {
  switch (epi) {
    case RCV_EPI_RELU: return fmaxf(v, 0.f);
    case RCV_EPI_RELU_AFFINE: return fmaf(sc, fmaxf(v, 0.f), sh);
    case RCV_EPI_AFFINE_RELU: return fmaxf(fmaf(sc, v, sh), 0.f);
    case RCV_EPI_AFFINE: return fmaf(sc, v, sh);
    default: return v;
  }
}

template <int BM, int BN, int TN, int KC>
__global__ void __launch_bounds__(NT) igemm_kernel(const RcvIgemm p) {
  rcv_pdl_enter();
  using C = Cfg<BM, BN, TN, KC>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);   // [2][KC][BM]
  float* Bs = As + 2 * KC * BM;                     // [2][KC][BP]
  int2* tabA = reinterpret_cast<int2*>(Bs + 2 * KC * C::BP);

  const int tid = threadIdx.x;
  const int cls = blockIdx.z;
  const int T = p.taps[cls].n;
  const int K = p.CA * T;
  int* tabW = reinterpret_cast<int*>(tabA + K);
  const int HWin = p.Hin * p.Win;
  const int HWg = p.Hg * p.Wg;
  const int M = p.N * HWg;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  for (int k = tid; k < K; k += NT) {
    int ca = k / T, t = k - ca * T;
    int dy = p.taps[cls].dy[t], dx = p.taps[cls].dx[t];
    tabA[k] = make_int2(ca * HWin + dy * p.Win + dx, ((dy + 16) << 8) | (dx + 16));
    tabW[k] = ca * p.wsA + p.taps[cls].wi[t];
  }

  // gather coordinates of this thread's pixels
  const int gp = tid % C::GP;
  const int ksub = tid / C::GP;
  int gy0[C::PPT], gx0[C::PPT];
  const float* gbase[C::PPT];
#pragma unroll
  for (int q = 0; q < C::PPT; ++q) {
    int m = m0 + gp + q * NT;
    if (m < M) {
      int n = m / HWg, r = m - n * HWg;
      int i = r / p.Wg, j = r - i * p.Wg;
      gy0[q] = i * p.gs;
      gx0[q] = j * p.gs;
      gbase[q] = p.in + (size_t)n * p.CA * HWin + gy0[q] * p.Win + gx0[q];
    } else {
      gy0[q] = -1000000;  // every tap fails the bounds test
      gx0[q] = 0;
      gbase[q] = p.in;
    }
  }
  __syncthreads();

  float ra[C::PPT * C::KPT];
  float rb[C::BEL];

  auto load_tiles = [&](int kc0) {
#pragma unroll
    for (int q = 0; q < C::PPT; ++q) {
#pragma unroll
      for (int i = 0; i < C::KPT; ++i) {
        int k = kc0 + ksub + i * C::KSTR;
        float v = 0.f;
        if (k < K) {
          int2 e = tabA[k];
          int iy = gy0[q] + (e.y >> 8) - 16;
          int ix = gx0[q] + (e.y & 0xff) - 16;
          if ((unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win)
            v = __ldg(gbase[q] + e.x);
        }
        ra[q * C::KPT + i] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < C::BEL; ++i) {
      int e = tid + i * NT;
      float v = 0.f;
      if (e < KC * BN) {
        int kk = e % KC, n = e / KC;
        int k = kc0 + kk, co = n0 + n;
        if (k < K && co < p.CB) v = __ldg(p.w + tabW[k] + (size_t)co * p.wsB);
      }
      rb[i] = v;
    }
  };
  auto store_tiles = [&](int buf) {
    float* as = As + buf * KC * BM;
    float* bs = Bs + buf * KC * C::BP;
#pragma unroll
    for (int q = 0; q < C::PPT; ++q)
#pragma unroll
      for (int i = 0; i < C::KPT; ++i)
        as[(ksub + i * C::KSTR) * BM + gp + q * NT] = ra[q * C::KPT + i];
#pragma unroll
    for (int i = 0; i < C::BEL; ++i) {
      int e = tid + i * NT;
      if (e < KC * BN) bs[(e % KC) * C::BP + e / KC] = rb[i];
    }
  };

  const int tx = tid % C::TXN;
  const int ty = tid / C::TXN;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();

  for (int kc = 0, it = 0; kc < K; kc += KC, ++it) {
    const int buf = it & 1;
    const bool more = kc + KC < K;
    if (more) load_tiles(kc + KC);
    const float* as = As + buf * KC * BM;
    const float* bs = Bs + buf * KC * C::BP;
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float a[8], b[TN];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(as + kk * BM + tx * 4);
      *reinterpret_cast<float4*>(&a[4]) =
          *reinterpret_cast<const float4*>(as + kk * BM + BM / 2 + tx * 4);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(bs + kk * C::BP + ty * 4);
      if (TN == 8)
        *reinterpret_cast<float4*>(&b[TN - 4]) =
            *reinterpret_cast<const float4*>(bs + kk * C::BP + BN / 2 + ty * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---------------- epilogue ----------------
  const int epi = p.epilogue;
  const bool vec = (p.ostep == 1) && ((HWg & 3) == 0);
  const int ca_ = cls >> 1, cb_ = cls & 1;
  float ssum[TN], ssq[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) ssum[j] = ssq[j] = 0.f;

#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int mg = m0 + g * (BM / 2) + tx * 4;
    if (mg >= M) continue;
    if (vec) {
      const int n = mg / HWg, r = mg - n * HWg;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int co = n0 + ((TN == 8 && j >= 4) ? BN / 2 - 4 : 0) + ty * 4 + j;
        if (co >= p.CB) continue;
        const float bi = p.bias ? __ldg(p.bias + co) : 0.f;
        const float sc = p.scale ? __ldg(p.scale + co) : 1.f;
        const float sh = p.shift ? __ldg(p.shift + co) : 0.f;
        const size_t off = ((size_t)n * p.CB + co) * HWg + r;
        float4 v;
        v.x = apply_epi(acc[g * 4 + 0][j] + bi, epi, sc, sh);
        v.y = apply_epi(acc[g * 4 + 1][j] + bi, epi, sc, sh);
        v.z = apply_epi(acc[g * 4 + 2][j] + bi, epi, sc, sh);
        v.w = apply_epi(acc[g * 4 + 3][j] + bi, epi, sc, sh);
        if (p.residual) {
          const float4 rr = *reinterpret_cast<const float4*>(p.residual + off);
          v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
        }
        *reinterpret_cast<float4*>(p.out + off) = v;
        ssum[j] += (v.x + v.y) + (v.z + v.w);
        ssq[j] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = mg + i;
        if (m >= M) continue;
        const int n = m / HWg, r = m - n * HWg;
        const int gi = r / p.Wg, gj = r - gi * p.Wg;
        const int oy = gi * p.ostep + ca_, ox = gj * p.ostep + cb_;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int co = n0 + ((TN == 8 && j >= 4) ? BN / 2 - 4 : 0) + ty * 4 + j;
          if (co >= p.CB) continue;
          const float bi = p.bias ? __ldg(p.bias + co) : 0.f;
          const float sc = p.scale ? __ldg(p.scale + co) : 1.f;
          const float sh = p.shift ? __ldg(p.shift + co) : 0.f;
          const size_t off = (((size_t)n * p.CB + co) * p.Hout + oy) * p.Wout + ox;
          float v = apply_epi(acc[g * 4 + i][j] + bi, epi, sc, sh);
          if (p.residual) v += p.residual[off];
          p.out[off] = v;
          ssum[j] += v;
          ssq[j] += v * v;
        }
      }
    }
  }

  if (p.stats) {
    constexpr int LPT = C::TXN < 32 ? C::TXN : 32;  // lanes sharing one ty inside a warp
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float s = ssum[j], s2 = ssq[j];
#pragma unroll
      for (int o = LPT / 2; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const int co = n0 + ((TN == 8 && j >= 4) ? BN / 2 - 4 : 0) + ty * 4 + j;
      if ((tid % LPT) == 0 && co < p.CB) {
        atomicAdd(p.stats + co, (double)s);
        atomicAdd(p.stats + p.CB + co, (double)s2);
      }
    }
  }
}

template <int BM, int BN, int TN, int KC>
int launch_cfg(const RcvIgemm& p, cudaStream_t st) {
  using C = Cfg<BM, BN, TN, KC>;
  int maxT = 0;
  for (int c = 0; c < p.nclass; ++c) maxT = p.taps[c].n > maxT ? p.taps[c].n : maxT;
  const size_t smem = C::smem_bytes(p.CA * maxT);
  static bool attr_done = false;  // benign race: idempotent
  if (!attr_done) {
    cudaFuncSetAttribute(igemm_kernel<BM, BN, TN, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         96 * 1024);
    attr_done = true;
  }
  RCV_REQUIRE(smem <= 96 * 1024, RCV_ERR_UNSUPPORTED, "igemm: K=%d needs %zu B smem", p.CA * maxT,
              smem);
  const int64_t M = (int64_t)p.N * p.Hg * p.Wg;
  RCV_REQUIRE(M < (1ll << 31) && (int64_t)p.N * p.CB * p.Hout * p.Wout < (1ll << 40), RCV_ERR_UNSUPPORTED,
              "igemm: problem too large");
  dim3 grid(rcv_cdiv(M, BM), rcv_cdiv(p.CB, BN), p.nclass);
  rcv_launch(igemm_kernel<BM, BN, TN, KC>, dim3(grid), dim3(NT), smem, st, p);
  RCV_CHECK_LAUNCH("igemm_kernel");
  return RCV_OK;
}

}  // namespace

// AUTO: tensor cores wherever the reduction is long enough to amortise operand staging
// (measured on B200 at batch 64, tools/umma_probe.py: K >= 128 and >= 16 output channels);
// the 3-/8-channel outer layers and the 1x1 class head stay on CUDA cores.
bool rcv_umma_pays(const RcvIgemm& p) {
  int maxT = 0;
  for (int c = 0; c < p.nclass; ++c) maxT = p.taps[c].n > maxT ? p.taps[c].n : maxT;
  return p.CA * maxT >= 128 && p.CB >= 16 && rcv_umma_supported(p);
}

static int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Which engine runs a problem (rcv_engine).  RCV_NARROW=0 / RCV_DIRECT=0 switch the two direct
// engines off for A/B runs.
int rcv_pick_engine(const RcvIgemm& p, bool have_packed) {
  static const int use_narrow = env_flag("RCV_NARROW", 1);
  static const int use_direct = env_flag("RCV_DIRECT", 1);
  if (p.math == RCV_MATH_TF32X3) return RCV_ENGINE_UMMA;
  // 16 -> <= 16 stride-1 3x3: the persistent tensor-core kernel (rcv_umma_halo.cu) beats the FFMA2 kernel
  if (rcv_math_auto(p.math) && have_packed && rcv_umma_c16_ok(p)) return RCV_ENGINE_UMMA;
  // <= 16 output channels: TMA-staged FFMA2 direct convolution (exact fp32), whatever the reduction length
  if (use_narrow && rcv_narrow_supported(p)) return RCV_ENGINE_NARROW;
  if (rcv_math_auto(p.math) && have_packed && rcv_umma_pays(p)) return RCV_ENGINE_UMMA;
  if (use_direct && rcv_direct_supported(p)) return RCV_ENGINE_DIRECT;
  return RCV_ENGINE_SIMT;
}

int rcv_launch_igemm(const RcvIgemm& p, cudaStream_t st) {
  const int eng = rcv_pick_engine(p, p.wpacked != nullptr);
  RCV_REQUIRE(p.in_scale == nullptr || eng == RCV_ENGINE_UMMA, RCV_ERR_UNSUPPORTED,
              "normalise-on-load is a feature of the halo-staged tensor-core kernel only");
  RCV_REQUIRE(p.residual == nullptr || p.res_C == p.CB || eng == RCV_ENGINE_NARROW, RCV_ERR_UNSUPPORTED,
              "a residual with fewer channels than the output (res_channels %d < %d) is taken by the narrow-layer engine only",
              p.res_C, p.CB);
  switch (eng) {
    case RCV_ENGINE_UMMA: return rcv_launch_igemm_umma(p, st);
    case RCV_ENGINE_NARROW: return rcv_launch_narrow(p, st);
    case RCV_ENGINE_DIRECT: return rcv_launch_direct(p, st);
    default: return rcv_launch_igemm_simt(p, st);
  }
}

int rcv_launch_igemm_simt(const RcvIgemm& p, cudaStream_t st) {
  if (p.CB > 64) return launch_cfg<128, 128, 8, 8>(p, st);
  if (p.CB > 32) return launch_cfg<128, 64, 4, 8>(p, st);
  if (p.CB > 16) return launch_cfg<256, 32, 4, 8>(p, st);
  if (p.CB > 8) return launch_cfg<512, 16, 4, 4>(p, st);
  return launch_cfg<1024, 8, 4, 4>(p, st);
}
