// tcgen05 / TMEM implicit-GEMM engine (sm_100a): convolution forward, input gradient and
// the transposed-convolution family on the 5th-generation tensor cores, at fp32-level
// accuracy through a 3-term TF32 split (a*b ~= ah*bh + al*bh + ah*bl, |err| ~ 2^-21).
//
// GEMM view (same RcvIgemm problem as the CUDA-core engine in rcv_igemm.cu):
//   M = pixel-grid points (n,i,j) of the whole batch, N = output channels,
//   K = (tap, input channel)  -- tap-major so that a 32-wide K block touches few taps.
//
// One CTA = one 128 x BN output tile; 128 threads:
//   * all four warps are PRODUCERS: thread t owns tile row t (one pixel).  Per K block of
//     32 it gathers 32 activations straight from NCHW global memory (coalesced across the
//     warp: lanes are consecutive pixels), splits them into tf32 hi / lo parts and writes
//     both as the K-major, 128-byte-swizzled canonical UMMA layout (one 128 B row per
//     pixel; 16 B chunk c of row r lives at chunk c ^ (r & 7)) -- conflict-free STS.128.
//     The weight tile (BN rows) is staged the same way.  Next block's global loads are
//     issued before this block's MMAs so they fly under the tensor-core work.
//   * thread 0 ISSUES: 3 x tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) per 8-wide K step,
//     accumulating in TMEM; tcgen05.commit on an mbarrier frees the smem stage.
//   * all four warps run the EPILOGUE: tcgen05.ld (32 lanes x 16 columns per warp), then
//     bias / ReLU / folded-BN affine / residual, coalesced NCHW stores (lanes = pixels),
//     and the train-mode BatchNorm per-channel sum / sum-of-squares via a shuffle
//     transpose-reduce and one double atomic per channel per warp.
#include "rcv_common.cuh"

namespace {

constexpr int NT = 128;
constexpr int BM = 128;
constexpr int BK = 32;  // fp32 elements per K block: one 128-byte swizzle row
constexpr int STAGES = 2;
constexpr int MAXT = 9;

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread = lane)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major: 1) | [32,46) SBO>>4 (8 rows x
// 128 B = 1024) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, tf32 x tf32, both
// operands K-major, N at [17,23) in units of 8, M at [24,29) in units of 16.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;  // exact in fp32; the tensor core reads its upper 19 bits
}

__device__ __forceinline__ float apply_epi(float v, int epi, float sc, float sh) {
  switch (epi) {
    case RCV_EPI_RELU: return fmaxf(v, 0.f);
    case RCV_EPI_RELU_AFFINE: return fmaf(sc, fmaxf(v, 0.f), sh);
    case RCV_EPI_AFFINE_RELU: return fmaxf(fmaf(sc, v, sh), 0.f);
    case RCV_EPI_AFFINE: return fmaf(sc, v, sh);
    default: return v;
  }
}

// Sum over the 32 lanes of 16 per-lane values: afterwards a[0] on every lane holds the warp
// total of value index (lane >> 1).  16 shuffles instead of 80.
__device__ __forceinline__ void warp_transpose_reduce16(float (&a)[16], int lane) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int half = 8 >> s;      // values kept per lane after this step
    const int mask = 16 >> s;     // partner distance
    const bool up = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? a[i] : a[i + half];
      const float keep = up ? a[i + half] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
  }
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
}

template <int BN>
struct Smem {
  static constexpr int A_BYTES = BM * 128;  // one hi or lo A tile
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = TILE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers etc.*/;
};

template <int BN>
__global__ void __launch_bounds__(NT) umma_igemm_kernel(const RcvIgemm p) {
  using S = Smem<BN>;
  constexpr int BCH = BN * 8 / NT;  // 16-byte weight chunks per thread per K block
  static_assert(BCH >= 1, "BN too small");
  constexpr int TCOLS = BN < 32 ? 32 : BN;

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  unsigned char* gen_tiles = smem_raw + (tiles - raw);
  unsigned char* misc = gen_tiles + S::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);               // [STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 64);
  int* s_toff = reinterpret_cast<int*>(misc + 128);                  // [MAXT] input offset of a tap
  int* s_twi = s_toff + 16;                                          // [MAXT] weight offset of a tap

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int cls = blockIdx.z;
  const int T = p.taps[cls].n;
  const int CA = p.CA;
  const int K = CA * T;
  const int HWin = p.Hin * p.Win;
  const int HWg = p.Hg * p.Wg;
  const int M = p.N * HWg;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int nkb = (K + BK - 1) / BK;

  if (tid < MAXT) {
    const int t = tid < T ? tid : 0;
    s_toff[tid] = p.taps[cls].dy[t] * p.Win + p.taps[cls].dx[t];
    s_twi[tid] = p.taps[cls].wi[t];
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- this thread's pixel (tile row) --------------------------------------------------
  const int m = m0 + tid;
  const bool mrow = m < M;
  int pn = 0, pi = 0, pj = 0;
  if (mrow) {
    pn = m / HWg;
    const int r = m - pn * HWg;
    pi = r / p.Wg;
    pj = r - pi * p.Wg;
  }
  const int gy0 = pi * p.gs, gx0 = pj * p.gs;
  const float* gbase = p.in + (size_t)pn * CA * HWin + gy0 * p.Win + gx0;
  uint32_t tapmask = 0;  // bit t: tap t reads inside the image for this pixel
  if (mrow) {
    for (int t = 0; t < T; ++t) {
      const int iy = gy0 + p.taps[cls].dy[t], ix = gx0 + p.taps[cls].dx[t];
      if ((unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win) tapmask |= 1u << t;
    }
  }

  float va[BK];
  float vb[BCH * 4];

  auto load_regs = [&](int kb) {
    const int k0 = kb * BK;
    {
      int tap = k0 / CA;
      int ca = k0 - tap * CA;
#pragma unroll
      for (int i = 0; i < BK; ++i) {
        float v = 0.f;
        if (tap < T && ((tapmask >> tap) & 1u)) v = __ldg(gbase + (size_t)ca * HWin + s_toff[tap]);
        va[i] = v;
        if (++ca == CA) { ca = 0; ++tap; }
      }
    }
#pragma unroll
    for (int q = 0; q < BCH; ++q) {
      const int ch = tid + q * NT;  // chunk id: row = ch % BN, 16-byte chunk = ch / BN
      const int nrow = ch % BN, c = ch / BN;
      const int co = n0 + nrow;
      int k = k0 + c * 4;
      int tap = k / CA;
      int ca = k - tap * CA;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = 0.f;
        if (tap < T && co < p.CB) v = __ldg(p.w + (size_t)co * p.wsB + (size_t)ca * p.wsA + s_twi[tap]);
        vb[q * 4 + e] = v;
        if (++ca == CA) { ca = 0; ++tap; }
      }
    }
  };

  auto store_smem = [&](int stage) {
    unsigned char* st = gen_tiles + stage * S::STAGE_BYTES;
    unsigned char* a_hi = st;
    unsigned char* a_lo = st + S::A_BYTES;
    unsigned char* b_hi = st + 2 * S::A_BYTES;
    unsigned char* b_lo = b_hi + S::B_BYTES;
    const int row = tid;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 h, l;
      split_tf32(va[4 * c + 0], h.x, l.x);
      split_tf32(va[4 * c + 1], h.y, l.y);
      split_tf32(va[4 * c + 2], h.z, l.z);
      split_tf32(va[4 * c + 3], h.w, l.w);
      const int off = row * 128 + ((c ^ (row & 7)) << 4);
      *reinterpret_cast<float4*>(a_hi + off) = h;
      *reinterpret_cast<float4*>(a_lo + off) = l;
    }
#pragma unroll
    for (int q = 0; q < BCH; ++q) {
      const int ch = tid + q * NT;
      const int nrow = ch % BN, c = ch / BN;
      float4 h, l;
      split_tf32(vb[4 * q + 0], h.x, l.x);
      split_tf32(vb[4 * q + 1], h.y, l.y);
      split_tf32(vb[4 * q + 2], h.z, l.z);
      split_tf32(vb[4 * q + 3], h.w, l.w);
      const int off = nrow * 128 + ((c ^ (nrow & 7)) << 4);
      *reinterpret_cast<float4*>(b_hi + off) = h;
      *reinterpret_cast<float4*>(b_lo + off) = l;
    }
  };

  constexpr uint32_t idesc = make_idesc(BM, BN);

  // ---- main loop ------------------------------------------------------------------------
  load_regs(0);
  for (int kb = 0; kb < nkb; ++kb) {
    const int stage = kb % STAGES;
    const int use = kb / STAGES;  // how many times this stage has been filled before
    if (use > 0) mbar_wait(smem_u32(&bars[stage]), (uint32_t)((use - 1) & 1));
    store_smem(stage);
    if (kb + 1 < nkb) load_regs(kb + 1);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sbase = tiles + stage * S::STAGE_BYTES;
      const uint64_t a_hi = make_desc(sbase);
      const uint64_t a_lo = make_desc(sbase + S::A_BYTES);
      const uint64_t b_hi = make_desc(sbase + 2 * S::A_BYTES);
      const uint64_t b_lo = make_desc(sbase + 2 * S::A_BYTES + S::B_BYTES);
      const int krem = K - kb * BK;
      const int ksteps = krem >= BK ? BK / 8 : (krem + 7) / 8;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 32 >> 4);  // 8 tf32 = 32 bytes along K inside the swizzle row
        umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, (kb | ks) != 0);
        umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
        umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, 1u);
      }
      umma_commit(smem_u32(&bars[stage]));
    }
  }
  {
    const int last = nkb - 1;
    mbar_wait(smem_u32(&bars[last % STAGES]), (uint32_t)((last / STAGES) & 1));
    tc_fence_after();
  }

  // ---- epilogue -------------------------------------------------------------------------
  const int epi = p.epilogue;
  const int HWo = p.Hout * p.Wout;
  size_t obase = 0;
  if (mrow) {
    const int ca_ = cls >> 1, cb_ = cls & 1;
    const int oy = pi * p.ostep + ca_, ox = pj * p.ostep + cb_;
    obase = (size_t)pn * p.CB * HWo + (size_t)oy * p.Wout + ox;
  }
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 16) {
    if (n0 + c0 >= p.CB) break;  // warp-uniform
    float acc[16];
    tmem_ld16(trow + c0, acc);
    float s1[16], s2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int co = n0 + c0 + j;
      float v = 0.f;
      if (mrow && co < p.CB) {
        const float bi = p.bias ? __ldg(p.bias + co) : 0.f;
        const float sc = p.scale ? __ldg(p.scale + co) : 1.f;
        const float sh = p.shift ? __ldg(p.shift + co) : 0.f;
        const size_t off = obase + (size_t)co * HWo;
        v = apply_epi(acc[j] + bi, epi, sc, sh);
        if (p.residual) v += __ldg(p.residual + off);
        p.out[off] = v;
      }
      s1[j] = v;
      s2[j] = v * v;
    }
    if (p.stats) {
      warp_transpose_reduce16(s1, lane);
      warp_transpose_reduce16(s2, lane);
      const int co = n0 + c0 + (lane >> 1);
      if ((lane & 1) == 0 && co < p.CB) {
        atomicAdd(p.stats + co, (double)s1[0]);
        atomicAdd(p.stats + p.CB + co, (double)s2[0]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TCOLS);
}

template <int BN>
int launch_bn(const RcvIgemm& p, cudaStream_t st) {
  using S = Smem<BN>;
  static bool attr_done = false;  // benign race: idempotent
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(umma_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::TOTAL);
    if (e != cudaSuccess) {
      rcv_set_error("umma_igemm: cannot reserve %d B of shared memory: %s", S::TOTAL, cudaGetErrorString(e));
      return RCV_ERR_CUDA;
    }
    attr_done = true;
  }
  const int64_t M = (int64_t)p.N * p.Hg * p.Wg;
  RCV_REQUIRE(M < (1ll << 31) && (int64_t)p.N * p.CB * p.Hout * p.Wout < (1ll << 40), RCV_ERR_UNSUPPORTED,
              "umma_igemm: problem too large");
  dim3 grid(rcv_cdiv(M, BM), rcv_cdiv(p.CB, BN), p.nclass);
  RCV_REQUIRE(grid.y <= 65535, RCV_ERR_UNSUPPORTED, "umma_igemm: too many output-channel tiles");
  umma_igemm_kernel<BN><<<grid, NT, S::TOTAL, st>>>(p);
  RCV_CHECK_LAUNCH("umma_igemm_kernel");
  return RCV_OK;
}

}  // namespace

int rcv_launch_igemm_umma(const RcvIgemm& p, cudaStream_t st) {
  for (int c = 0; c < p.nclass; ++c)
    RCV_REQUIRE(p.taps[c].n >= 1 && p.taps[c].n <= MAXT, RCV_ERR_UNSUPPORTED, "umma_igemm: %d taps",
                p.taps[c].n);
  if (p.CB > 64) return launch_bn<128>(p, st);
  if (p.CB > 32) return launch_bn<64>(p, st);
  if (p.CB > 16) return launch_bn<32>(p, st);
  return launch_bn<16>(p, st);
}
