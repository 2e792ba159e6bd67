// tcgen05 / TMEM implicit-GEMM engine (sm_100a): convolution forward, input gradient and
// the transposed-convolution family on the 5th-generation tensor cores, at fp32-level
// accuracy through a 3-term TF32 split (a*b ~= ah*bh + al*bh + ah*bl, |err| ~ 2^-21).
//
// GEMM view (same RcvIgemm problem as the CUDA-core engine in rcv_igemm.cu):
//   M = pixel-grid points (n,i,j) of the whole batch, N = output channels,
//   K = (tap, input channel)  -- tap-major so that a 32-wide K block touches few taps.
//
// Weights are PRE-PACKED once per forward (rcv_conv_pack): for every (parity class, N tile,
// K block) the exact shared-memory image of the B operand -- tf32 hi part and lo part, K-major
// rows of 128 bytes with the 128-byte swizzle already applied -- so a stage of B is one
// contiguous cp.async.bulk copy completing on an mbarrier.
//
// One CTA = one 128 x BN output tile, G producer groups of 128 threads + one issuer warp:
//   * PRODUCERS: group g fills K blocks g, g+G, g+2G, ... into A-ring stage g; thread t of a
//     group owns tile row t (one pixel).  Per K block of 32 it gathers 32 activations straight
//     from NCHW global memory (coalesced across the warp: lanes are consecutive pixels), splits
//     them into tf32 hi / lo parts and writes both in the K-major 128B-swizzled canonical UMMA
//     layout (16 B chunk c of row r at chunk c ^ (r & 7): conflict-free STS.128), then
//     fence.proxy.async + one mbarrier arrive per warp.  The proxy fence is a full MEMBAR that
//     drains the thread's outstanding loads, so a thread never prefetches across it; the global
//     latency is hidden instead by the G groups working on G different K blocks at once.
//   * last warp, lane 0, ISSUER: keeps SB bulk copies of B in flight, waits for a stage's A and B,
//     issues 3 x tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) per 8-wide K step -- the hi*hi
//     product into one TMEM accumulator, the two correction products into a second one (the
//     tensor core truncates when it accumulates; keeping the small terms apart keeps that error
//     relative to THEIR magnitude) -- and tcgen05.commit's the stage's empty barriers.
//   * all producer warps, EPILOGUE (warp w: TMEM lanes 32*(w%4).., column group w/4):
//     tcgen05.ld (32 lanes x 16 columns) of both accumulators,
//     bias / ReLU / folded-BN affine / residual, coalesced NCHW stores (lanes = pixels), and the
//     train-mode BatchNorm per-channel sum / sum-of-squares via a shuffle transpose-reduce and
//     one double atomic per channel per warp.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "rcv_common.cuh"
#include "rcv_umma.cuh"

namespace {

constexpr int BK = 32;            // fp32 elements per K block: one 128-byte swizzle row
constexpr int GTHREADS = 128;     // threads of one producer group: one per tile row
constexpr int BM = 128;
constexpr int MAXT = 9;
constexpr int RCV_UMMA_MAX_TABLE_K = 2304;  // per-k gather table (channel counts that are not multiples of 32)

using namespace rcv_umma;

// Phase-timing instrumentation (tools/umma_phases.py): compiled in only with -DRCV_PROF=1
#ifndef RCV_PROF
#define RCV_PROF 0
#endif
#if RCV_PROF
#define RCV_PROF_ON(cond) (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (cond))
#define RCV_PROF_I(e) do { if (RCV_PROF_ON(kb < 128)) p.prof[2048 + kb * 4 + (e)] = clock64(); } while (0)
#define RCV_PROF_P(e) do { if (RCV_PROF_ON(row == 0 && kb < 128)) p.prof[kb * 8 + (e)] = clock64(); } while (0)
#define RCV_PROF_E(e) do { if (RCV_PROF_ON(row == 0)) p.prof[4000 + grp * 4 + (e)] = clock64(); } while (0)
#else
#define RCV_PROF_I(e) do { } while (0)
#define RCV_PROF_P(e) do { } while (0)
#define RCV_PROF_E(e) do { } while (0)
#endif

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
    const int half = 8 >> s;   // values kept per lane after this step
    const int mask = 16 >> s;  // partner distance
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

// Tile configuration per output-channel tile width.
// KB_ = fp32 elements per K block = one swizzle row: 32 (128-byte swizzle) or 16 (64-byte swizzle: stages half
// the size, so the BN = 128 configuration fits two CTAs per SM -- 150 tiles on 148 SMs become one wave, and
// a second CTA's MMAs fill the barrier-latency bubbles of the first)
template <int BN_, int G_, int KB_ = 32>
struct Cfg {
  static constexpr int BN = BN_;
  static constexpr int G = G_;                       // producer groups = ring depth
  static constexpr int KB = KB_;
  static constexpr int ROWB = KB_ * 4;               // bytes per operand row
  static constexpr int NPROD = G_ * GTHREADS;
  static constexpr int NT = NPROD + 64;              // + the MMA-issuer warp + the B-loader warp
  static constexpr int A_BYTES = BM * ROWB;          // one hi or lo A tile
  static constexpr int A_STAGE = 2 * A_BYTES;
  static constexpr int B_STAGE = BN_ * ROWB * 2;     // hi rows then lo rows
  static constexpr int RING = G_ * (BM + BN_) * KB_ * 8;
  static constexpr int CTAS = RING <= 50 * 1024 ? 3 : RING <= 100 * 1024 ? 2 : 1;  // CTAs per SM the ring is sized for
  static constexpr int STAGE = A_STAGE + B_STAGE;    // one ring stage: A hi, A lo, B hi, B lo
  static constexpr int TILE_BYTES = G_ * STAGE;
  static constexpr int FIXED = TILE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers etc.*/ + 3 * BN_ * 4;
  // + the per-k gather table (8 B per k) when the channel count is not a multiple of 32
  static constexpr int TCOLS = 2 * BN_ < 32 ? 32 : 2 * BN_;  // main + correction accumulators
};

// byte offset of 16-byte chunk c of operand row `row` in the K-major swizzled canonical layout
template <int KB>
__host__ __device__ __forceinline__ int sw_off(int row, int c) {
  return KB == 32 ? row * 128 + ((c ^ (row & 7)) << 4)          // Swizzle<3,4,3>: chunk ^= address bits [7,10)
                  : row * 64 + ((c ^ ((row >> 1) & 3)) << 4);   // Swizzle<2,4,3>: chunk ^= address bits [7,9)
}
// shared-memory matrix descriptor of a K-major swizzled tile with rows of KB fp32 (see make_desc in rcv_umma.cuh)
template <int KB>
__device__ __forceinline__ uint64_t make_desc_kb(uint32_t saddr) {
  if (KB == 32) return make_desc(saddr);
  // SWIZZLE_64B: stride between 8-row groups = 8 x 64 B, layout type 4
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

static int g_bn_cap = 0, g_force_g = 0, g_kb128 = 0, g_kb64 = 0, g_kb32 = 0;  // experiments: RCV_UMMA_BNCAP / RCV_UMMA_G
__host__ inline int umma_bn(int CB) {
  if (g_bn_cap == 0) {
    const char* e = getenv("RCV_UMMA_BNCAP");
    g_bn_cap = e ? atoi(e) : 128;
    e = getenv("RCV_UMMA_G");
    g_force_g = e ? atoi(e) : -1;
    e = getenv("RCV_UMMA_KB128");  // K block of the BN = 128 configuration: 16 (default) or 32
    g_kb128 = e ? atoi(e) : 16;
    if (g_kb128 != 16 && g_kb128 != 32) g_kb128 = 16;
    e = getenv("RCV_UMMA_KB64");   // K block of the BN = 64 configuration: 32 (two 48 KB stages) or 16 (four 24 KB stages)
    g_kb64 = e ? atoi(e) : 32;
    if (g_kb64 != 16 && g_kb64 != 32) g_kb64 = 32;
    e = getenv("RCV_UMMA_KB32");   // K block of the BN = 32 short-reduction configuration: 32 (2 CTAs/SM) or 16 (3 CTAs/SM)
    g_kb32 = e ? atoi(e) : 32;
    if (g_kb32 != 16 && g_kb32 != 32) g_kb32 = 32;
  }
  const int bn = CB > 64 ? 128 : CB > 32 ? 64 : CB > 16 ? 32 : 16;
  return bn > g_bn_cap ? g_bn_cap : bn;
}
// fp32 elements per K block of a layer's configuration (the packed panel layout depends on it)
__host__ inline int umma_kb(int CB) {
  const int bn = umma_bn(CB);
  return bn == 128 ? g_kb128 : bn == 64 ? g_kb64 : bn == 32 ? g_kb32 : 32;
}
__host__ __device__ inline int max_taps(const RcvIgemm& p) {
  int m = 0;
  for (int c = 0; c < p.nclass; ++c) m = p.taps[c].n > m ? p.taps[c].n : m;
  return m;
}

template <int BN, int G, int KB>
__global__ void __launch_bounds__(Cfg<BN, G, KB>::NT, Cfg<BN, G, KB>::CTAS) umma_igemm_kernel(const RcvIgemm p) {
  rcv_pdl_enter();
  using C = Cfg<BN, G, KB>;
  constexpr int NPROD = C::NPROD, NT = C::NT;
  constexpr int BK = KB;  // shadows the file-level default inside the kernel

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  unsigned char* gen_tiles = smem_raw + (tiles - raw);
  unsigned char* misc = gen_tiles + C::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);  // full[G] empty[G] done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 120);
  int* s_toff = reinterpret_cast<int*>(misc + 128);     // [MAXT] input offset of a tap
  float* s_cst = reinterpret_cast<float*>(misc + 256);  // [3][BN] bias, scale, shift of this N tile
  int2* s_ktab = reinterpret_cast<int2*>(misc + 256 + 3 * BN * 4);  // [nkb*BK] per-k (offset, tap), non-uniform path
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * G, bar_done = bar_empty + 8 * G;
  static_assert(8 * (2 * G + 1) <= 120, "barrier area");

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
  // fast modes (RCV_MATH_TF32 / RCV_MATH_BF16 on a layer this kernel runs): one TF32 MMA per product -- the lo
  // halves are neither staged, copied nor multiplied, and the correction accumulator is never read
  const bool fast = p.math >= RCV_MATH_TF32;

  if (tid < MAXT) {
    const int t = tid < T ? tid : 0;
    s_toff[tid] = p.taps[cls].dy[t] * p.Win + p.taps[cls].dx[t];
  }
  for (int c = tid; c < BN; c += NT) {
    const int co = n0 + c;
    const bool in = co < p.CB;
    s_cst[c] = (in && p.bias) ? __ldg(p.bias + co) : 0.f;
    s_cst[BN + c] = (in && p.scale) ? __ldg(p.scale + co) : 1.f;
    s_cst[2 * BN + c] = (in && p.shift) ? __ldg(p.shift + co) : 0.f;
  }
  if ((CA % BK) != 0) {
    for (int k = tid; k < nkb * BK; k += NT) {
      const int tap = k / CA, ca = k - tap * CA;
      s_ktab[k] = tap < T ? make_int2(4 * (ca * HWin + p.taps[cls].dy[tap] * p.Win + p.taps[cls].dx[tap]), tap)
                          : make_int2(0, 31);
    }
  }
  if (tid == 0) {
    for (int s = 0; s < G; ++s) {
      mbar_init(bar_full + 8 * s, GTHREADS / 32 + 1);  // 4 producer warps + the B loader's expect_tx
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == NPROD / 32) tmem_alloc(smem_u32(tmem_slot), C::TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == NPROD / 32 + 1) {
    // ================================ B LOADER ========================================
    // one bulk copy of a packed weight block per ring stage, as soon as the stage is free
    if (elect_one() && !(p.debug & 16)) {
      const int kbmax = (CA * max_taps(p) + BK - 1) / BK;  // K blocks per (class, N tile) in the pack
      const unsigned char* gB = reinterpret_cast<const unsigned char*>(p.wpacked) +
                                ((size_t)(cls * gridDim.y + blockIdx.y) * kbmax) * C::B_STAGE;
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % G, u = kb / G;
        if (u > 0) mbar_wait(bar_empty + 8 * st, (uint32_t)((u - 1) & 1));
        const uint32_t bbytes = fast ? C::B_STAGE / 2 : C::B_STAGE;  // hi rows come first in a packed block
        mbar_expect_tx(bar_full + 8 * st, bbytes);
        bulk_g2s(tiles + st * C::STAGE + C::A_STAGE, gB + (size_t)kb * C::B_STAGE, bbytes, bar_full + 8 * st);
      }
    } else if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) mbar_arrive(bar_full + 8 * (kb % G));  // timing experiments only
    }
  } else if (warp == NPROD / 32) {
    // ================================ MMA ISSUER ======================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      // parity mode: a_hi * [b_hi | b_lo] in ONE instruction (N = 2*BN: the stage holds the hi rows then the lo rows,
      // the correction accumulator follows the main one), then a_lo * b_hi -- two MMAs per K step instead of three
      constexpr uint32_t idesc2 = make_idesc(BM, BN < 128 ? 2 * BN : BN);
      const uint32_t d_main = tmem_base, d_corr = tmem_base + BN;
      mbar_wait(bar_full, 0);
      for (int kb0 = 0; kb0 < nkb; kb0 += G) {
        const uint32_t par = (uint32_t)((kb0 / G) & 1);
#pragma unroll
        for (int st = 0; st < G; ++st) {  // stage index is compile-time: descriptors fold to constants + base
          const int kb = kb0 + st;
          if (kb < nkb) {
            RCV_PROF_I(0);
            tc_fence_after();
            const uint32_t abase = tiles + st * C::STAGE, bbase = abase + C::A_STAGE;
            const uint64_t a_hi = make_desc_kb<KB>(abase), a_lo = make_desc_kb<KB>(abase + C::A_BYTES);
            const uint64_t b_hi = make_desc_kb<KB>(bbase);
            const int krem = K - kb * BK;
            const int ksteps = krem >= BK ? BK / 8 : (krem + 7) / 8;
            if (!(p.debug & 1)) {
              // 8 tf32 = 32 B = 2 x 16 B along K inside the swizzled row per step
              constexpr bool FUSE = BN < 128;  // BN = 128 tiles are throughput-bound: they keep three instructions
              const uint64_t b_lo = make_desc_kb<KB>(bbase + BN * C::ROWB);
#pragma unroll
              for (int ks = 0; ks < BK / 8; ++ks) {
                if (ks < ksteps) {
                  const uint32_t acc = ks == 0 ? (uint32_t)(kb != 0) : 1u;
                  if (fast) {
                    umma_tf32(d_main, a_hi + 2 * ks, b_hi + 2 * ks, idesc, acc);
                  } else if (FUSE) {
                    umma_tf32(d_main, a_hi + 2 * ks, b_hi + 2 * ks, idesc2, acc);
                    umma_tf32(d_corr, a_lo + 2 * ks, b_hi + 2 * ks, idesc, 1u);
                  } else {
                    umma_tf32(d_corr, a_lo + 2 * ks, b_hi + 2 * ks, idesc, acc);
                    umma_tf32(d_corr, a_hi + 2 * ks, b_lo + 2 * ks, idesc, 1u);
                    umma_tf32(d_main, a_hi + 2 * ks, b_hi + 2 * ks, idesc, acc);
                  }
                }
              }
            }
            RCV_PROF_I(1);
            // poll the next block's operands while this block's MMAs drain from the queue; the
            // commit below frees a stage the next block does not depend on
            if (kb + 1 < nkb)
              mbar_wait(bar_full + 8 * ((st + 1) % G), st + 1 == G ? par ^ 1u : par);
            RCV_PROF_I(2);
            umma_commit(bar_empty + 8 * st);
            if (kb == nkb - 1) umma_commit(bar_done);
            RCV_PROF_I(3);
          }
        }
      }
    }
  } else {
    // ================================ PRODUCERS =======================================
    const int row = tid & (BM - 1);
    const int grp = tid / GTHREADS;  // producer group (warp-uniform)
    const int m = m0 + row;
    const bool mrow = m < M;
    int pn = 0, pi = 0, pj = 0;
    if (mrow) {
      pn = m / HWg;
      const int r = m - pn * HWg;
      pi = r / p.Wg;
      pj = r - pi * p.Wg;
    }
    const int gy0 = pi * p.gs, gx0 = pj * p.gs;
    // 32-bit byte offsets from the (uniform) tensor base keep the gather at ~4 instructions/load
    const char* inb = reinterpret_cast<const char*>(p.in);
    const uint32_t boff0 = 4u * (uint32_t)(pn * CA * HWin + gy0 * p.Win + gx0);
    const uint32_t cstride = 4u * (uint32_t)HWin;
    uint32_t tapmask = 0;  // bit t: tap t reads inside the image for this pixel
    if (mrow) {
      for (int t = 0; t < T; ++t) {
        const int iy = gy0 + p.taps[cls].dy[t], ix = gx0 + p.taps[cls].dx[t];
        if ((unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win) tapmask |= 1u << t;
      }
    }
    const bool uni = (CA % BK) == 0;  // every K block lies inside one tap
    unsigned char* a_hi = gen_tiles + grp * C::STAGE;
    unsigned char* a_lo = a_hi + C::A_BYTES;
    const uint32_t my_full = bar_full + 8 * grp, my_empty = bar_empty + 8 * grp;
    // (tap, channel) of this group's next K block on the uniform path
    int tap = (grp * BK) / CA, ca = (grp * BK) % CA;

    for (int kb = grp, use = 0; kb < nkb; kb += G, ++use) {
      float va[BK];
      RCV_PROF_P(0);
      if (p.debug & 2) {
#pragma unroll
        for (int i = 0; i < BK; ++i) va[i] = 1.f;
      } else if (uni) {
        const bool ok = (tapmask >> tap) & 1u;
        const uint32_t b = boff0 + 4u * (uint32_t)(ca * HWin + s_toff[tap]);
#pragma unroll
        for (int i = 0; i < BK; ++i)
          va[i] = ok ? __ldg(reinterpret_cast<const float*>(inb + (b + (uint32_t)i * cstride))) : 0.f;
        ca += G * BK;
        while (ca >= CA) { ca -= CA; ++tap; }
      } else {
        const int2* tab = s_ktab + kb * BK;
#pragma unroll
        for (int i = 0; i < BK; ++i) {
          const int2 e = tab[i];  // (byte offset, tap or 31 beyond K)
          va[i] = ((tapmask >> e.y) & 1u) ? __ldg(reinterpret_cast<const float*>(inb + (boff0 + (uint32_t)e.x)))
                                          : 0.f;
        }
      }
      RCV_PROF_P(1);
      if (use > 0) mbar_wait(my_empty, (uint32_t)((use - 1) & 1));
      RCV_PROF_P(2);
      if (!(p.debug & 4))
#pragma unroll
      for (int c = 0; c < BK / 4; ++c) {
        float4 h, l;
        split_tf32(va[4 * c + 0], h.x, l.x);
        split_tf32(va[4 * c + 1], h.y, l.y);
        split_tf32(va[4 * c + 2], h.z, l.z);
        split_tf32(va[4 * c + 3], h.w, l.w);
        const int off = sw_off<KB>(row, c);
        *reinterpret_cast<float4*>(a_hi + off) = h;
        if (!fast) *reinterpret_cast<float4*>(a_lo + off) = l;
      }
      RCV_PROF_P(3);
      fence_proxy_async_smem();
      RCV_PROF_P(4);
      __syncwarp();
      if (lane == 0) mbar_arrive(my_full);
      RCV_PROF_P(5);
    }

    // ================================ EPILOGUE ========================================
    RCV_PROF_E(0);
    mbar_wait(bar_done, 0);
    RCV_PROF_E(1);
    tc_fence_after();
    const int epi = p.epilogue;
    const int HWo = p.Hout * p.Wout;
    size_t obase = 0;
    if (mrow) {
      const int ca_ = cls >> 1, cb_ = cls & 1;
      const int oy = pi * p.ostep + ca_, ox = pj * p.ostep + cb_;
      obase = (size_t)pn * p.CB * HWo + (size_t)oy * p.Wout + ox;
    }
    // warps of different groups share TMEM lanes 32*(w%4)..+31 and interleave the 16-column chunks
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const bool has_res = p.residual != nullptr, has_stats = p.stats != nullptr;
#pragma unroll 1
    for (int c0 = grp * 16; c0 < BN; c0 += G * 16) {
      if (n0 + c0 >= p.CB || (p.debug & 32)) break;  // warp-uniform
      uint32_t rm[16], rc[16];
      tmem_ld16_nowait(trow + c0, rm);
      if (!fast) {
        tmem_ld16_nowait(trow + BN + c0, rc);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) rc[j] = 0u;
      }
      const int nvalid = min(16, p.CB - (n0 + c0));  // channels of this chunk that exist
      float* optr = p.out + obase + (size_t)(n0 + c0) * HWo;
      const float* rptr = has_res ? p.residual + obase + (size_t)(n0 + c0) * HWo : nullptr;
      float res[16];
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 16; ++j) res[j] = (mrow && j < nvalid) ? __ldg(rptr + (size_t)j * HWo) : 0.f;
      }
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float acc = __uint_as_float(rm[j]) + __uint_as_float(rc[j]) + s_cst[c0 + j];
        float y = apply_epi(acc, epi, s_cst[BN + c0 + j], s_cst[2 * BN + c0 + j]);
        if (has_res) y += res[j];
        v[j] = (mrow && j < nvalid) ? y : 0.f;
      }
      if (mrow) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nvalid) optr[(size_t)j * HWo] = v[j];
      }
      if (has_stats) {
        float s2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) s2[j] = v[j] * v[j];
        warp_transpose_reduce16(v, lane);
        warp_transpose_reduce16(s2, lane);
        const int co = n0 + c0 + (lane >> 1);
        if ((lane & 1) == 0 && co < p.CB) {
          atomicAdd(p.stats + co, (double)v[0]);
          atomicAdd(p.stats + p.CB + co, (double)s2[0]);
        }
      }
    }
  }

#if RCV_PROF
  if (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid & 127) == 0 && tid < NPROD)
    p.prof[4000 + (tid >> 7) * 4 + 2] = clock64();
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == NPROD / 32) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TCOLS);
  }
}

// One thread per 16-byte chunk of the packed image: 4 consecutive k of one weight row, split
// into hi / lo and written at the swizzled position.
__device__ __forceinline__ void pack_chunk(const RcvIgemm& p, int BN, int ntiles, int kbmax, int KB,
                                           unsigned char* __restrict__ packed, int64_t q) {
  const int cpr = KB / 4;  // 16-byte chunks per operand row
  const int64_t per_block = (int64_t)BN * cpr;
  {
    const int row = (int)(q % BN);
    const int c = (int)((q / BN) % cpr);
    int64_t blk = q / per_block;
    const int kb = (int)(blk % kbmax);
    blk /= kbmax;
    const int tile = (int)(blk % ntiles);
    const int cls = (int)(blk / ntiles);
    const int T = p.taps[cls].n;
    const int co = tile * BN + row;
    int k = kb * KB + c * 4;
    int tap = k / p.CA;
    int ca = k - tap * p.CA;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[e] = 0.f;
      if (tap < T && co < p.CB)
        v[e] = __ldg(p.w + (size_t)co * p.wsB + (size_t)ca * p.wsA + p.taps[cls].wi[tap]);
      if (++ca == p.CA) { ca = 0; ++tap; }
    }
    float4 h, l;
    split_tf32(v[0], h.x, l.x);
    split_tf32(v[1], h.y, l.y);
    split_tf32(v[2], h.z, l.z);
    split_tf32(v[3], h.w, l.w);
    const size_t rowb = (size_t)KB * 4;
    unsigned char* base = packed + ((size_t)(cls * ntiles + tile) * kbmax + kb) * ((size_t)BN * rowb * 2);
    const int off = KB == 32 ? sw_off<32>(row, c) : sw_off<16>(row, c);
    *reinterpret_cast<float4*>(base + off) = h;
    *reinterpret_cast<float4*>(base + (size_t)BN * rowb + off) = l;
  }
}

// bf16 panel (RCV_MATH_BF16, halo-staged kernel): K blocks of 64 channels = rows of 128 bytes, 128-byte swizzle,
// one copy.  One thread per 16-byte chunk = 8 consecutive channels of one weight row.  KB == 64 marks the layout.
__device__ __forceinline__ void pack_chunk_bf16(const RcvIgemm& p, int BN, int ntiles, int kbmax,
                                                unsigned char* __restrict__ packed, int64_t q) {
  const int row = (int)(q % BN);
  const int c = (int)((q / BN) % 8);
  int64_t blk = q / ((int64_t)BN * 8);
  const int kb = (int)(blk % kbmax);
  const int tile = (int)(blk / kbmax);
  const int co = tile * BN + row;
  const int k = kb * 64 + c * 8;
  const int tap = k / p.CA, ca = k - tap * p.CA;  // CA % 64 == 0: a chunk never straddles two taps
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v0 = 0.f, v1 = 0.f;
    if (tap < p.taps[0].n && co < p.CB) {
      const float* src = p.w + (size_t)co * p.wsB + p.taps[0].wi[tap];
      v0 = __ldg(src + (size_t)(ca + 2 * e) * p.wsA);
      v1 = __ldg(src + (size_t)(ca + 2 * e + 1) * p.wsA);
    }
    const __nv_bfloat162 b = __floats2bfloat162_rn(v0, v1);
    w[e] = *reinterpret_cast<const uint32_t*>(&b);
  }
  unsigned char* base = packed + ((size_t)tile * kbmax + kb) * ((size_t)BN * 128);
  *reinterpret_cast<uint4*>(base + sw_off<32>(row, c)) = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256) pack_kernel(const RcvIgemm p, int BN, int ntiles, int kbmax, int KB,
                                                   unsigned char* __restrict__ packed) {
  rcv_pdl_enter();
  const int64_t total = KB == 64 ? (int64_t)ntiles * kbmax * BN * 8 : (int64_t)p.nclass * ntiles * kbmax * BN * (KB / 4);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total;
       q += (int64_t)gridDim.x * blockDim.x) {
    if (KB == 64) pack_chunk_bf16(p, BN, ntiles, kbmax, packed, q);
    else pack_chunk(p, BN, ntiles, kbmax, KB, packed, q);
  }
}

// All layers' panels in one launch: a device-resident job table (built once on the host, the
// weight and panel pointers are stable) with the prefix sum of 16-byte chunks per job.
__global__ void __launch_bounds__(256) pack_multi_kernel(const RcvPackJob* __restrict__ jobs, int njobs) {
  rcv_pdl_enter();
  __shared__ long long s_begin[RCV_PACK_MAX_JOBS + 1];
  for (int j = threadIdx.x; j <= njobs; j += blockDim.x)
    s_begin[j] = j < njobs ? jobs[j].chunk_begin : jobs[njobs - 1].chunk_begin + jobs[njobs - 1].chunks;
  __syncthreads();
  const long long total = s_begin[njobs];
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total;
       q += (long long)gridDim.x * blockDim.x) {
    int j = 0;
    while (q >= s_begin[j + 1]) ++j;
    const RcvPackJob& jb = jobs[j];
    if (jb.kb == 64) pack_chunk_bf16(jb.p, jb.BN, jb.ntiles, jb.kbmax, jb.packed, q - s_begin[j]);
    else pack_chunk(jb.p, jb.BN, jb.ntiles, jb.kbmax, jb.kb, jb.packed, q - s_begin[j]);
  }
}

template <int BN, int G, int KB = 32>
int launch_bn(const RcvIgemm& p, cudaStream_t st) {
  using C = Cfg<BN, G, KB>;
  constexpr int BK = KB;
  constexpr int MAX_SMEM = C::FIXED + RCV_UMMA_MAX_TABLE_K * 8;
  static_assert(MAX_SMEM <= 227 * 1024, "shared memory budget");
  static bool attr_done = false;  // benign race: idempotent
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(umma_igemm_kernel<BN, G, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         MAX_SMEM);
    if (C::CTAS > 1)
      cudaFuncSetAttribute(umma_igemm_kernel<BN, G, KB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      rcv_set_error("umma_igemm: cannot reserve %d B of shared memory: %s", MAX_SMEM, cudaGetErrorString(e));
      return RCV_ERR_CUDA;
    }
    attr_done = true;
  }
  const int kpad = rcv_cdiv((int64_t)p.CA * max_taps(p), BK) * BK;
  const int smem = C::FIXED + ((p.CA % BK) != 0 ? kpad * 8 : 0);
  RCV_REQUIRE((int64_t)p.N * p.CA * p.Hin * p.Win < (1ll << 30), RCV_ERR_UNSUPPORTED,
              "umma_igemm: input tensor too large for 32-bit byte offsets");
  const int64_t M = (int64_t)p.N * p.Hg * p.Wg;
  RCV_REQUIRE(M < (1ll << 31) && (int64_t)p.N * p.CB * p.Hout * p.Wout < (1ll << 40), RCV_ERR_UNSUPPORTED,
              "umma_igemm: problem too large");
  dim3 grid(rcv_cdiv(M, BM), rcv_cdiv(p.CB, BN), p.nclass);
  RCV_REQUIRE(grid.y <= 65535, RCV_ERR_UNSUPPORTED, "umma_igemm: too many output-channel tiles");
  rcv_launch(umma_igemm_kernel<BN, G, KB>, dim3(grid), dim3(C::NT), smem, st, p);
  RCV_CHECK_LAUNCH("umma_igemm_kernel");
  return RCV_OK;
}

int check_taps(const RcvIgemm& p) {
  for (int c = 0; c < p.nclass; ++c)
    RCV_REQUIRE(p.taps[c].n >= 1 && p.taps[c].n <= MAXT, RCV_ERR_UNSUPPORTED, "umma_igemm: %d taps",
                p.taps[c].n);
  RCV_REQUIRE(rcv_umma_supported(p), RCV_ERR_UNSUPPORTED,
              "umma_igemm: %d reduced channels (not a multiple of 32) x %d taps exceeds the gather table",
              p.CA, max_taps(p));
  return RCV_OK;
}

}  // namespace

bool rcv_umma_supported(const RcvIgemm& p) {
  return (p.CA % umma_kb(p.CB)) == 0 || (int64_t)p.CA * max_taps(p) <= RCV_UMMA_MAX_TABLE_K;
}

// panel layout of a layer: K-block width in elements (64 = the bf16 layout of RCV_MATH_BF16 halo layers)
static int panel_kb(const RcvIgemm& p) {
  if (rcv_umma_c16_ok(p)) return 16;  // the persistent 16-channel kernel keeps the whole 9 x [16 rows x 64 B] panel resident
  return rcv_umma_halo_bf16_ok(p, umma_bn(p.CB)) ? 64 : umma_kb(p.CB);
}

size_t rcv_umma_packed_bytes(const RcvIgemm& p) {
  if (panel_kb(p) == 64) {
    const int BN = umma_bn(p.CB);
    return (size_t)rcv_cdiv(p.CB, BN) * (p.CA * 9 / 64) * BN * 128;
  }
  const int BN = umma_bn(p.CB), KB = panel_kb(p);
  const int ntiles = rcv_cdiv(p.CB, BN);
  const int kbmax = rcv_cdiv((int64_t)p.CA * max_taps(p), KB);
  return (size_t)p.nclass * ntiles * kbmax * BN * KB * 8;
}

int rcv_launch_umma_pack(const RcvIgemm& p, void* packed, cudaStream_t st) {
  int rc = check_taps(p);
  if (rc) return rc;
  RCV_REQUIRE(((uintptr_t)packed & 127) == 0, RCV_ERR_BAD_ARG, "conv_pack: packed buffer must be 128-byte aligned");
  const int BN = umma_bn(p.CB), KB = panel_kb(p);
  const int ntiles = rcv_cdiv(p.CB, BN);
  const int kbmax = rcv_cdiv((int64_t)p.CA * max_taps(p), KB);
  const int64_t total = KB == 64 ? (int64_t)ntiles * kbmax * BN * 8 : (int64_t)p.nclass * ntiles * kbmax * BN * (KB / 4);
  int blocks = rcv_cdiv(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  rcv_launch(pack_kernel, dim3(blocks), dim3(256), 0, st, p, BN, ntiles, kbmax, KB,
             reinterpret_cast<unsigned char*>(packed));
  RCV_CHECK_LAUNCH("pack_kernel");
  return RCV_OK;
}

int rcv_umma_pack_job(const RcvIgemm& p, void* packed, long long chunk_begin, RcvPackJob* job) {
  int rc = check_taps(p);
  if (rc) return rc;
  RCV_REQUIRE(((uintptr_t)packed & 127) == 0, RCV_ERR_BAD_ARG, "conv_pack: packed buffer must be 128-byte aligned");
  job->p = p;
  job->BN = umma_bn(p.CB);
  job->ntiles = rcv_cdiv(p.CB, job->BN);
  job->kb = panel_kb(p);
  job->kbmax = rcv_cdiv((int64_t)p.CA * max_taps(p), job->kb);
  job->packed = reinterpret_cast<unsigned char*>(packed);
  job->chunk_begin = chunk_begin;
  job->chunks = job->kb == 64 ? (long long)job->ntiles * job->kbmax * job->BN * 8
                              : (long long)p.nclass * job->ntiles * job->kbmax * job->BN * (job->kb / 4);
  return RCV_OK;
}

int rcv_launch_umma_pack_multi(const RcvPackJob* dev_jobs, int njobs, long long total_chunks, cudaStream_t st) {
  RCV_REQUIRE(njobs >= 1 && njobs <= RCV_PACK_MAX_JOBS, RCV_ERR_BAD_ARG, "pack_multi: %d jobs (1..%d)", njobs,
              RCV_PACK_MAX_JOBS);
  long long blocks = (total_chunks + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  rcv_launch(pack_multi_kernel, dim3((int)blocks), dim3(256), 0, st, dev_jobs, njobs);
  RCV_CHECK_LAUNCH("pack_multi_kernel");
  return RCV_OK;
}

size_t rcv_umma_workspace_bytes(const RcvIgemm& p) {
  const int bn = umma_bn(p.CB);
  if (!(rcv_umma_halo_ok(p, bn, umma_kb(p.CB)) || rcv_umma_halo_bf16_ok(p, bn))) return 0;
  return rcv_umma_halo_workspace_bytes(p, bn);
}

bool rcv_umma_takes_input_transform(const RcvIgemm& p) {
  return rcv_umma_halo_ok(p, umma_bn(p.CB), umma_kb(p.CB)) || rcv_umma_halo_bf16_ok(p, umma_bn(p.CB));
}

int rcv_launch_igemm_umma(const RcvIgemm& p_in, cudaStream_t st) {
  RcvIgemm p = p_in;
  static int dbg = -1;  // RCV_UMMA_DEBUG: timing experiments only (results are wrong when set)
  if (dbg < 0) {
    const char* e = getenv("RCV_UMMA_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  p.debug = dbg;
  p.prof = g_rcv_prof;
  int rc = check_taps(p);
  if (rc) return rc;
  RCV_REQUIRE(p.wpacked != nullptr, RCV_ERR_BAD_ARG,
              "tensor-core conv needs packed weights (rcv_conv_pack); none were given");
  RCV_REQUIRE(((uintptr_t)p.wpacked & 127) == 0, RCV_ERR_BAD_ARG, "packed weights must be 128-byte aligned");
  // Ring depth = producer groups.  Long reductions: 3-4 stages, one CTA per SM.  Short ones (a few
  // K blocks per tile): 2 stages so that 2 CTAs share an SM and overlap prologue / epilogue.
  if (rcv_umma_c16_ok(p)) return rcv_launch_igemm_umma_c16(p, st);
  const int bn = umma_bn(p.CB);
  const bool deep = g_force_g > 0 ? g_force_g >= 3 : (int64_t)p.CA * max_taps(p) > 10 * BK;
  // stride-1 3x3 layers: nine tap-shifted descriptors over one staged patch instead of nine gathers
  if (rcv_umma_halo_ok(p, bn, umma_kb(p.CB)) || rcv_umma_halo_bf16_ok(p, bn))
    return rcv_launch_igemm_umma_halo(p, bn, umma_kb(p.CB), st);
  RCV_REQUIRE(p.in_scale == nullptr, RCV_ERR_UNSUPPORTED,
              "normalise-on-load needs the halo-staged kernel (stride-1 3x3, reduced channels a multiple of 32, short rows)");
  switch (bn) {
    case 128: return umma_kb(p.CB) == 16 ? launch_bn<128, 3, 16>(p, st) : launch_bn<128, 3>(p, st);
    // BN = 64: two CTAs per SM always (a 150-tile layer is then one wave; measured 38 -> 27 us for 64->64 and
    // 60 -> 44 us for 128->64 at batch 64, 89 -> 64 / 143 -> 106 us at batch 256)
    case 64:
      if (umma_kb(p.CB) == 16) return g_force_g == 2 ? launch_bn<64, 2, 16>(p, st) : launch_bn<64, 4, 16>(p, st);
      return launch_bn<64, 2>(p, st);
    case 32:
      if (umma_kb(p.CB) == 16) return launch_bn<32, 2, 16>(p, st);
      return deep ? launch_bn<32, 4>(p, st) : launch_bn<32, 2>(p, st);
    default: return deep ? launch_bn<16, 4>(p, st) : launch_bn<16, 2>(p, st);
  }
}
