// tcgen05 implicit GEMM with a HALO-STAGED A operand: stride-1 3x3 convolutions (dilation 1 or 2) and their
// input gradients, the belly of every net on the path (model.py:105-116, 126-142, 166-176).
//
// The general engine (rcv_umma.cu) re-gathers the im2col operand once per tap: 9 x (128 pixels x 32 channels)
// global loads, tf32 splits and swizzled stores per channel block, and that gather -- not the tensor pipe -- is
// what bounds it.  Here the pixels of the whole batch are laid out as ONE padded,
// flattened image: row r = d + n*(H+d) + i, pitch PW = W + d, so consecutive images share their halo rows and
// consecutive rows share their halo columns, and every tap is the constant offset dy*PW + dx in that space.
// A tile is 128 consecutive positions (halo positions are computed and discarded: 11 % at 15x20, 6 % at
// 30x40).  Per 32-channel block the producers stage the positions the tile and its halo touch ONCE --
// L = 128 + 2*(d*PW + d) rows of 128 bytes, tf32 hi and lo copies, written with the 128-byte swizzle of their
// absolute shared-memory address -- and the MMA issuer reads the nine taps through nine matrix descriptors whose
// start address is shifted by whole rows (tools/umma_shift_probe.cu: the hardware swizzle is a function of the
// absolute address, so any row shift reads the right data).  6.7x less staging work at 15x20.
// The B operand is the same pre-packed panel the general engine uses (rcv_conv_pack), streamed by bulk copies
// through a ring; 3xTF32 with two TMEM accumulators and the fused epilogue are unchanged.
// Roles: 8 producer / epilogue warps, one MMA-issuer warp, one B-loader warp; two CTAs per SM.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "rcv_common.cuh"
#include "rcv_umma.cuh"

namespace {
using namespace rcv_umma;

constexpr int BM = 128;
constexpr int NPROD = 256;
constexpr int NT = NPROD + 64;
constexpr int MAXL = 256;  // staged rows per channel block: one per producer thread

struct HaloGeo {
  int32_t d, PW, HP, S, L, Lpad, nkc, kbmax, nbs;  // nbs: B ring depth (2..4), as many stages as fit beside the patch
  // split reduction (needs RcvIgemm::ws): pixel tiles [0, tfull) run whole; every later tile is handled by nsplit
  // CTAs that each reduce a contiguous share of the channel blocks, leave their partial accumulator in the
  // workspace and count themselves in; the last one to arrive sums the partials in share order and runs the epilogue
  int32_t tfull, nsplit;
  int64_t Mh;
  int32_t tsign;  // tap t reads offset tsign * d * ((t / 3 - 1) * PW + (t % 3 - 1)): +1 forward order, -1 flipped (input gradient)
};

template <int KBB>
__device__ __forceinline__ uint64_t make_desc_b(uint32_t saddr) {
  if (KBB == 32) return make_desc(saddr);
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);  // SWIZZLE_64B rows of 16 fp32
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

__device__ __forceinline__ void warp_transpose_reduce16(float (&a)[16], int lane) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int half = 8 >> s;
    const int mask = 16 >> s;
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

// BF = false: fp32 operands as tf32 hi (+ lo) parts, channel blocks of 32 (one 128-byte row of fp32);
// BF = true (RCV_MATH_BF16): bf16 operands, channel blocks of 64 (one 128-byte row of bf16), kind::f16 MMAs with
// K = 16 per instruction, one accumulator; KBB is ignored (the B rows are 128 bytes as well).
template <int BN, int KBB, bool BF = false>
struct HCfg {
  static constexpr int CBLK = BF ? 64 : 32;           // channels per staged block
  static constexpr int BROWB = BF ? 128 : KBB * 4;    // bytes per B row
  static constexpr int B_STAGE = BF ? BN * 128 : BN * BROWB * 2;  // fp32: hi rows then lo rows
  static constexpr int MAXBS = 12;                    // deepest B ring
  static constexpr int SUB = BF ? 1 : 32 / KBB;       // B K-blocks per (tap, channel block)
  static constexpr int KSTEPS = BF ? 4 : KBB / 8;     // MMAs (32 bytes of K each) per B K-block
  static constexpr int TCOLS = BF ? (BN < 32 ? 32 : BN) : (2 * BN < 32 ? 32 : 2 * BN);
  static constexpr int MISC = 256 + 3 * BN * 4;       // barriers, epilogue constants (+ 2*CA floats of input scale / shift)
};

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// cute::UMMA::InstrDescriptor for kind::f16: fp32 accumulate (c_format 1 at [4,6)), a_format / b_format = BF16 (1) at
// [7,10) / [10,13), both operands K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x (low half-word, lower address) = lo
  return *reinterpret_cast<const uint32_t*>(&v);
}

// position in the padded flattened space -> pixel; false for halo positions
__device__ __forceinline__ bool decode_pos(long long q, const HaloGeo& g, int N, int H, int W, int& n, int& i, int& j) {
  if (q < 0 || q >= g.Mh) return false;
  const int r = (int)(q / g.PW);
  j = (int)(q - (long long)r * g.PW);
  const int rr = r - g.d;
  if (rr < 0 || j >= W) return false;
  n = rr / g.HP;
  i = rr - n * g.HP;
  return n < N && i < H;
}

// Split reduction, epilogue side (kept out of line: its registers must not weigh on the kernel's hot loops).
// Leaves this share's partial accumulator (main + correction) in the workspace as [column][row] (lanes = rows:
// coalesced), counts the share in, and returns true in the ONE CTA that arrived last -- after summing all shares
// in share order (the result does not depend on who arrived last) back over its main accumulator in TMEM.
template <int BN>
__device__ __noinline__ bool split_exchange(void* ws, int slot, int part, int nsplit, uint32_t trow, int ncols, int grp,
                                            int row, int tid, bool fast, uint32_t* s_arrived) {
  uint32_t* counters = reinterpret_cast<uint32_t*>(ws);
  float* base_part = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ws) + 1024) +
                     (size_t)slot * nsplit * (BN * BM);
  float* mine = base_part + (size_t)part * (BN * BM);
#pragma unroll 1
  for (int c0 = grp * 16; c0 < BN; c0 += 32) {
    if (c0 >= ncols) break;
    uint32_t rm[16];
    tmem_ld16_nowait(trow + c0, rm);
    if (!fast) {
      uint32_t rc[16];
      tmem_ld16_nowait(trow + BN + c0, rc);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) rm[j] = __float_as_uint(__uint_as_float(rm[j]) + __uint_as_float(rc[j]));
    } else {
      tmem_ld_wait();
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) __stcg(mine + (size_t)(c0 + j) * BM + row, __uint_as_float(rm[j]));
  }
  __threadfence();
  asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
  if (tid == 0) *s_arrived = atomicAdd(counters + slot, 1u);
  asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
  if (*s_arrived != (uint32_t)(nsplit - 1)) return false;
  __threadfence();
  if (tid == 0) counters[slot] = 0u;  // every share has arrived: leave the counter ready for the next launch
#pragma unroll 1
  for (int c0 = grp * 16; c0 < BN; c0 += 32) {
    if (c0 >= ncols) break;
    // all shares' loads of the chunk in flight at once (up to 64 per thread), then summed in share order
    float v[4][16];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u < nsplit) {
        const float* src = base_part + (size_t)u * (BN * BM) + (size_t)c0 * BM + row;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[u][j] = __ldcg(src + (size_t)j * BM);
      }
    }
    uint32_t a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float t = v[0][j];
#pragma unroll
      for (int u = 1; u < 4; ++u)
        if (u < nsplit) t += v[u][j];
      a[j] = __float_as_uint(t);
    }
    tmem_st16(trow + c0, a);
  }
  tmem_st_wait();
  return true;
}

template <int BN, int KBB, bool BF>
__global__ void __launch_bounds__(NT, 2) umma_halo_kernel(const RcvIgemm p, const HaloGeo g) {
  rcv_pdl_enter();
  using C = HCfg<BN, KBB, BF>;
  constexpr int SUB = C::SUB;
  constexpr int CBLK = C::CBLK;
  const int NBS = g.nbs;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gen = smem_raw + (base - raw);
  const uint32_t patch_bytes = (uint32_t)g.Lpad * 128u;
  constexpr uint32_t NPATCH = BF ? 1u : 2u;  // bf16: one copy of the patch; fp32: tf32 hi and lo copies
  const uint32_t a_hi_s = base, a_lo_s = base + patch_bytes, b_s = base + NPATCH * patch_bytes;
  unsigned char* misc = gen + NPATCH * patch_bytes + NBS * C::B_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);  // patch_full, patch_empty, done, bfull[NBS], bempty[NBS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 240);
  uint32_t* s_arrived = reinterpret_cast<uint32_t*>(misc + 244);  // split reduction: shares that arrived before this CTA
  float* s_cst = reinterpret_cast<float*>(misc + 256);
  float* s_in = s_cst + 3 * BN;  // [2][CA] input scale, shift (normalise-on-load)
  const uint32_t bar_pfull = smem_u32(bars), bar_pempty = bar_pfull + 8, bar_done = bar_pfull + 16;
  const uint32_t bar_bfull = bar_pfull + 24, bar_bempty = bar_bfull + 8 * C::MAXBS;
  static_assert(24 + 16 * C::MAXBS <= 240, "barrier area");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int CA = p.CA, H = p.Hin, W = p.Win, HW = H * W;
  int tile = blockIdx.x, part = -1, cb0 = 0, cb1 = g.nkc;
  if (tile >= g.tfull) {
    const int r = tile - g.tfull;
    tile = g.tfull + r / g.nsplit;
    part = r - (r / g.nsplit) * g.nsplit;
    cb0 = part * g.nkc / g.nsplit;
    cb1 = (part + 1) * g.nkc / g.nsplit;
  }
  const long long q0 = (long long)tile * BM;
  const int n0 = blockIdx.y * BN;
  const bool fast = BF || p.math >= RCV_MATH_TF32;  // one MMA per product: no lo halves, no correction accumulator

  for (int c = tid; c < BN; c += NT) {
    const int co = n0 + c;
    const bool in = co < p.CB;
    s_cst[c] = (in && p.bias) ? __ldg(p.bias + co) : 0.f;
    s_cst[BN + c] = (in && p.scale) ? __ldg(p.scale + co) : 1.f;
    s_cst[2 * BN + c] = (in && p.shift) ? __ldg(p.shift + co) : 0.f;
  }
  const bool nl = p.in_scale != nullptr;
  if (nl) {
    for (int c = tid; c < CA; c += NT) {
      s_in[c] = __ldg(p.in_scale + c);
      s_in[CA + c] = __ldg(p.in_shift + c);
    }
  }
  if (tid == 0) {
    mbar_init(bar_pfull, NPROD / 32);
    mbar_init(bar_pempty, 1);
    mbar_init(bar_done, 1);
    for (int s = 0; s < NBS; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    fence_barrier_init();
  }
  if (warp == NPROD / 32) tmem_alloc(smem_u32(tmem_slot), C::TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == NPROD / 32 + 1) {
    // ================================ B LOADER ========================================
    if (elect_one()) {
      const unsigned char* gB = reinterpret_cast<const unsigned char*>(p.wpacked) +
                                ((size_t)blockIdx.y * g.kbmax) * C::B_STAGE;
      const int kper = BF ? CA / 64 : CA / KBB;  // B K-blocks per tap in the pack (K = tap-major, then channel)
      int it = 0;
      for (int cb = cb0; cb < cb1; ++cb)
        for (int t = 0; t < 9; ++t)
          for (int sub = 0; sub < SUB; ++sub, ++it) {
            const int st = it % NBS, u = it / NBS;
            if (u > 0) mbar_wait(bar_bempty + 8 * st, (uint32_t)((u - 1) & 1));
            const uint32_t bbytes = (fast && !BF) ? C::B_STAGE / 2 : C::B_STAGE;  // fp32 panels: hi rows come first
            mbar_expect_tx(bar_bfull + 8 * st, bbytes);
            const int kbp = t * kper + cb * SUB + sub;
            bulk_g2s(b_s + st * C::B_STAGE, gB + (size_t)kbp * C::B_STAGE, bbytes, bar_bfull + 8 * st);
          }
    }
  } else if (warp == NPROD / 32) {
    // ================================ MMA ISSUER ======================================
    if (elect_one()) {
      constexpr uint32_t idesc = BF ? make_idesc_bf16(BM, BN) : make_idesc(BM, BN);
      constexpr uint32_t idesc2 = make_idesc(BM, BN < 128 ? 2 * BN : BN);  // fp32 parity mode: main | correction in one instruction
      const uint32_t d_main = tmem_base, d_corr = tmem_base + BN;
      int it = 0;
      for (int cb = cb0; cb < cb1; ++cb) {
        mbar_wait(bar_pfull, (uint32_t)((cb - cb0) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          const int ty = t / 3 - 1, tx = t - (t / 3) * 3 - 1;
          const uint32_t roff = (uint32_t)(g.S + g.tsign * g.d * (ty * g.PW + tx)) * 128u;
          const uint64_t a_hi = make_desc(a_hi_s + roff), a_lo = make_desc(a_lo_s + roff);
#pragma unroll
          for (int sub = 0; sub < SUB; ++sub, ++it) {
            const int st = it % NBS;
            mbar_wait(bar_bfull + 8 * st, (uint32_t)((it / NBS) & 1));
            tc_fence_after();
            const uint32_t bb = b_s + st * C::B_STAGE;
            const uint64_t b_hi = BF ? make_desc(bb) : make_desc_b<KBB>(bb);
#pragma unroll
            for (int ks = 0; ks < C::KSTEPS; ++ks) {
              const int ka = sub * C::KSTEPS + ks;  // K step (32 bytes) inside the 128-byte A row
              const uint32_t first = (it == 0 && ks == 0) ? 0u : 1u;
              if (BF) {
                umma_bf16(d_main, a_hi + 2 * ka, b_hi + 2 * ks, idesc, first);
                continue;
              }
              if (!fast) {
                // Two MMAs per K step instead of three: a B stage holds the BN hi rows and then the BN lo rows, so
                // ONE instruction with N = 2*BN computes a_hi * [b_hi | b_lo] into the adjacent main and correction
                // accumulators; the second adds a_lo * b_hi to the correction accumulator.
                // (neutral at BN = 128, whose tiles are throughput-bound: those keep the three-instruction form)
                if (BN < 128) {
                  umma_tf32(d_main, a_hi + 2 * ka, b_hi + 2 * ks, idesc2, first);
                  umma_tf32(d_corr, a_lo + 2 * ka, b_hi + 2 * ks, idesc, 1u);
                } else {
                  const uint64_t b_lo = make_desc_b<KBB>(bb + BN * C::BROWB);
                  umma_tf32(d_corr, a_lo + 2 * ka, b_hi + 2 * ks, idesc, first);
                  umma_tf32(d_corr, a_hi + 2 * ka, b_lo + 2 * ks, idesc, 1u);
                  umma_tf32(d_main, a_hi + 2 * ka, b_hi + 2 * ks, idesc, first);
                }
              } else {
                umma_tf32(d_main, a_hi + 2 * ka, b_hi + 2 * ks, idesc, first);
              }
            }
            umma_commit(bar_bempty + 8 * st);
          }
        }
        umma_commit(bar_pempty);  // the patch may be overwritten once these MMAs have read it
      }
      umma_commit(bar_done);
    }
  } else {
    // ================================ PRODUCERS =======================================
    // thread = one staged position (row of the patch); 32 channels of it per channel block
    const bool has_row = tid < g.L;
    int pn = 0, pi = 0, pj = 0;
    const bool pvalid = has_row && decode_pos(q0 - g.S + tid, g, p.N, H, W, pn, pi, pj);
    const char* inb = reinterpret_cast<const char*>(p.in);
    const uint32_t boff0 = 4u * (uint32_t)((pn * CA) * HW + pi * W + pj);
    const uint32_t cstride = 4u * (uint32_t)HW;
    for (int cb = cb0; cb < cb1; ++cb) {
      if (BF) {
        // 64 channels of this position as bf16: two half blocks of 32 loads each, four 16-byte chunks per half
        uint32_t pk[32];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float va[32];
          const uint32_t b = boff0 + (uint32_t)(cb * 64 + hf * 32) * cstride;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            va[i] = pvalid ? __ldg(reinterpret_cast<const float*>(inb + (b + (uint32_t)i * cstride))) : 0.f;
          if (nl && pvalid) {
            const bool rl = p.in_relu != 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float t = fmaf(s_in[cb * 64 + hf * 32 + i], va[i], s_in[CA + cb * 64 + hf * 32 + i]);
              va[i] = rl ? fmaxf(t, 0.f) : t;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[hf * 16 + i] = pack_bf16x2(va[2 * i], va[2 * i + 1]);
        }
        if (cb > cb0) mbar_wait(bar_pempty, (uint32_t)((cb - cb0 - 1) & 1));
        if (has_row) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int off = tid * 128 + ((c ^ (tid & 7)) << 4);
            *reinterpret_cast<uint4*>(gen + off) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          }
        }
      } else {
      float va[32];
      const uint32_t b = boff0 + (uint32_t)(cb * 32) * cstride;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        va[i] = pvalid ? __ldg(reinterpret_cast<const float*>(inb + (b + (uint32_t)i * cstride))) : 0.f;
      if (nl && pvalid) {  // the producer block's BatchNorm, applied to real pixels only: halo positions stay zero
        const bool rl = p.in_relu != 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float t = fmaf(s_in[cb * 32 + i], va[i], s_in[CA + cb * 32 + i]);
          va[i] = rl ? fmaxf(t, 0.f) : t;
        }
      }
      if (cb > cb0) mbar_wait(bar_pempty, (uint32_t)((cb - cb0 - 1) & 1));
      if (has_row) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 h, l;
          split_tf32(va[4 * c + 0], h.x, l.x);
          split_tf32(va[4 * c + 1], h.y, l.y);
          split_tf32(va[4 * c + 2], h.z, l.z);
          split_tf32(va[4 * c + 3], h.w, l.w);
          const int off = tid * 128 + ((c ^ (tid & 7)) << 4);
          *reinterpret_cast<float4*>(gen + off) = h;
          if (!fast) *reinterpret_cast<float4*>(gen + patch_bytes + off) = l;
        }
      }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pfull);
    }

    // ================================ EPILOGUE ========================================
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int row = tid & (BM - 1);
    const int grp = tid / BM;  // column half
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    // Split reduction: leave this share's partial accumulator (main + correction) in the workspace as
    // [column][row] (lanes = rows: coalesced), count in, and go on only if every other share has arrived; the
    // finishing CTA sums all shares in share order (so the result does not depend on who arrived last), writes
    // the sums back over its main accumulator in TMEM and runs the ordinary epilogue on them.
    bool finish = true, skip_corr = fast;
    if (part >= 0) {
      finish = split_exchange<BN>(p.ws, (tile - g.tfull) * gridDim.y + blockIdx.y, part, g.nsplit, trow, p.CB - n0, grp,
                                  row, tid, fast, s_arrived);
      skip_corr = true;
    }
    int en = 0, ei = 0, ej = 0;
    const bool mrow = decode_pos(q0 + row, g, p.N, H, W, en, ei, ej);
    const int epi = p.epilogue;
    const int HWo = p.Hout * p.Wout;
    const size_t obase = mrow ? (size_t)en * p.CB * HWo + (size_t)ei * p.Wout + ej : 0;
    const bool has_res = p.residual != nullptr, has_stats = p.stats != nullptr;
#pragma unroll 1
    for (int c0 = grp * 16; c0 < BN && finish; c0 += 32) {
      if (n0 + c0 >= p.CB) break;
      uint32_t rm[16], rc[16];
      tmem_ld16_nowait(trow + c0, rm);
      if (!skip_corr) {
        tmem_ld16_nowait(trow + BN + c0, rc);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) rc[j] = 0u;
      }
      const int nvalid = min(16, p.CB - (n0 + c0));
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

  tc_fence_before();
  __syncthreads();
  if (warp == NPROD / 32) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TCOLS);
  }
}

bool geometry(const RcvIgemm& p, HaloGeo* out) {
  if (p.nclass != 1 || p.gs != 1 || p.ostep != 1 || p.taps[0].n != 9) return false;
  if (p.Hout != p.Hin || p.Wout != p.Win || p.Hg != p.Hin || p.Wg != p.Win) return false;
  if ((p.CA % 32) != 0 || p.CA > 1024) return false;
  HaloGeo g;
  memset(&g, 0, sizeof(g));
  int d = 0;
  for (int t = 0; t < 9; ++t) {
    const int a = abs((int)p.taps[0].dy[t]), b = abs((int)p.taps[0].dx[t]);
    d = a > d ? a : d;
    d = b > d ? b : d;
  }
  if (d != 1 && d != 2) return false;
  // the nine taps must be the {-d, 0, d}^2 grid in row-major order, as it is (forward) or point-reflected (input
  // gradient): the issuer derives each tap's row shift from its index
  const int sgn = p.taps[0].dy[0] < 0 ? 1 : -1;
  for (int t = 0; t < 9; ++t) {
    if (p.taps[0].dy[t] != sgn * (t / 3 - 1) * d || p.taps[0].dx[t] != sgn * (t % 3 - 1) * d) return false;
  }
  g.tsign = sgn;
  g.d = d;
  g.PW = p.Win + d;
  g.HP = p.Hin + d;
  g.S = d * g.PW + d;
  g.L = BM + 2 * g.S;
  g.Lpad = (g.L + 7) & ~7;
  g.nkc = p.CA / 32;
  g.Mh = ((int64_t)p.N * g.HP + d) * g.PW;
  if (g.L > MAXL) return false;
  if ((int64_t)p.N * p.CA * p.Hin * p.Win >= (1ll << 30) || g.Mh >= (1ll << 31)) return false;
  *out = g;
  return true;
}

// B ring depth that fits beside the patch in a two-CTAs-per-SM shared-memory budget (0: nothing fits)
size_t smem_budget() {  // RCV_HALO_SMEM_KB (experiments): 113 = two CTAs per SM (default), up to 227 = one
  static const int kb = getenv("RCV_HALO_SMEM_KB") ? atoi(getenv("RCV_HALO_SMEM_KB")) : 113;
  return (size_t)(kb < 48 ? 48 : kb > 227 ? 227 : kb) * 1024;
}

int ring_depth(const HaloGeo& g, int bn, int kbb, int ca, bool bf, size_t* smem) {
  const size_t budget = smem_budget();
  const size_t fixed = 1024 + (bf ? 1 : 2) * (size_t)g.Lpad * 128 + 256 + 3 * (size_t)bn * 4 + 2 * (size_t)ca * 4;
  const size_t stage = bf ? (size_t)bn * 128 : (size_t)bn * kbb * 8;
  for (int nbs = 12; nbs >= 2; --nbs)
    if (fixed + nbs * stage <= budget) {
      if (smem) *smem = fixed + nbs * stage;
      return nbs;
    }
  return 0;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

// Split-reduction plan of a layer (HaloGeo::tfull / nsplit) and the workspace it needs.  nkc = channel blocks of
// the kernel variant that will run.  Two cases pay: (a) a grid slightly larger than a whole number of SM waves --
// at batch 64 the 15x20 layers are 169 tiles on 148 SMs, so 21 SMs carry two tiles while 127 wait: the left-over
// tiles are cut into up to four shares that run as second CTAs beside whole tiles; (b) a grid far smaller than
// the machine (small batches, batch-1 latency): every tile is cut.
size_t split_plan(const RcvIgemm& p, int bn, int nkc, int64_t Mh, int* tfull, int* nsplit) {
  static const int on = getenv("RCV_UMMA_SPLIT") ? atoi(getenv("RCV_UMMA_SPLIT")) : 1;
  const int T = rcv_cdiv(Mh, BM), ny = rcv_cdiv(p.CB, bn), S = sm_count();
  *tfull = T;
  *nsplit = 1;
  const int parts = nkc < 4 ? nkc : 4;
  // the fast modes issue a third (tf32) or a sixth (bf16) of the MMA work: measured, the exchange then costs more
  // than the balance gains (bf16 128->128 @15x20 batch 64: 19.6 us whole, 21.7 us split)
  if (!on || parts < 2 || p.math >= RCV_MATH_TF32) return 0;
  int tf = T;
  if ((int64_t)T * ny * 2 <= S) tf = 0;
  else if (ny == 1 && T > S && (T % S) * 2 <= S) tf = T - T % S;
  if (tf == T || (int64_t)(T - tf) * ny > 256) return 0;  // 256 arrival counters
  *tfull = tf;
  *nsplit = parts;
  return 1024 + (size_t)(T - tf) * ny * parts * ((size_t)bn * BM * 4);
}

template <int BN, int KBB, bool BF = false>
int launch_h(const RcvIgemm& p, HaloGeo g, cudaStream_t st) {
  if (BF) {
    g.nkc = p.CA / 64;
    g.kbmax = (p.CA * 9) / 64;
  } else {
    g.kbmax = (p.CA * 9) / KBB;
  }
  size_t smem = 0;
  g.nbs = ring_depth(g, BN, KBB, p.CA, BF, &smem);
  RCV_REQUIRE(g.nbs >= 2, RCV_ERR_UNSUPPORTED, "umma_halo: the patch leaves no room for the weight ring");
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(umma_halo_kernel<BN, KBB, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_budget());
    if (e != cudaSuccess) {
      rcv_set_error("umma_halo: cannot reserve shared memory: %s", cudaGetErrorString(e));
      return RCV_ERR_CUDA;
    }
    cudaFuncSetAttribute(umma_halo_kernel<BN, KBB, BF>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    attr_done = true;
  }
  const int T = rcv_cdiv(g.Mh, BM);
  const size_t need = split_plan(p, BN, g.nkc, g.Mh, &g.tfull, &g.nsplit);
  if (need == 0 || p.ws == nullptr || p.ws_bytes < need || ((uintptr_t)p.ws & 127) != 0) {
    g.tfull = T;  // no (or too small a) workspace: every tile whole
    g.nsplit = 1;
  }
  dim3 grid(g.tfull + (T - g.tfull) * g.nsplit, rcv_cdiv(p.CB, BN), 1);
  rcv_launch(umma_halo_kernel<BN, KBB, BF>, dim3(grid), dim3(NT), smem, st, p, g);
  RCV_CHECK_LAUNCH("umma_halo_kernel");
  return RCV_OK;
}

}  // namespace

// Stride-1 3x3 (dilation 1 or 2) problems whose reduced channel count is a multiple of 32 and whose rows are
// short enough for a one-row-per-thread patch: bn / kbb are the N tile and B K-block of the layer's packed panel.
bool rcv_umma_halo_ok(const RcvIgemm& p, int bn, int kbb) {
  static const int on = getenv("RCV_UMMA_HALO") ? atoi(getenv("RCV_UMMA_HALO")) : 1;
  HaloGeo g;
  if (!on || !geometry(p, &g)) return false;
  return bn >= 32 && ring_depth(g, bn, kbb, p.CA, false, nullptr) >= 2;
}

// RCV_MATH_BF16: the same layers with a reduced channel count that is a multiple of 64 run with bf16 operands
// (their packed panel then has the bf16 layout: rcv_umma_packed_bytes / pack follow this predicate)
bool rcv_umma_halo_bf16_ok(const RcvIgemm& p, int bn) {
  static const int on = getenv("RCV_UMMA_BF16") ? atoi(getenv("RCV_UMMA_BF16")) : 1;
  static const int halo_on = getenv("RCV_UMMA_HALO") ? atoi(getenv("RCV_UMMA_HALO")) : 1;
  HaloGeo g;
  if (!on || !halo_on || p.math != RCV_MATH_BF16 || (p.CA % 64) != 0 || bn < 32 || !geometry(p, &g)) return false;
  return ring_depth(g, bn, 32, p.CA, true, nullptr) >= 2;
}

size_t rcv_umma_halo_workspace_bytes(const RcvIgemm& p, int bn) {
  HaloGeo g;
  if (!geometry(p, &g)) return 0;
  int tfull, nsplit;
  return split_plan(p, bn, rcv_umma_halo_bf16_ok(p, bn) ? p.CA / 64 : g.nkc, g.Mh, &tfull, &nsplit);
}

int rcv_launch_igemm_umma_halo(const RcvIgemm& p, int bn, int kbb, cudaStream_t st) {
  HaloGeo g;
  RCV_REQUIRE(geometry(p, &g), RCV_ERR_UNSUPPORTED, "umma_halo: geometry outside the kernel's limits");
  if (rcv_umma_halo_bf16_ok(p, bn)) {
    if (bn == 128) return launch_h<128, 32, true>(p, g, st);
    if (bn == 64) return launch_h<64, 32, true>(p, g, st);
    if (bn == 32) return launch_h<32, 32, true>(p, g, st);
  }
  if (bn == 128 && kbb == 16) return launch_h<128, 16>(p, g, st);
  if (bn == 128 && kbb == 32) return launch_h<128, 32>(p, g, st);
  if (bn == 64 && kbb == 32) return launch_h<64, 32>(p, g, st);
  if (bn == 64 && kbb == 16) return launch_h<64, 16>(p, g, st);
  if (bn == 32 && kbb == 32) return launch_h<32, 32>(p, g, st);
  rcv_set_error("umma_halo: no configuration for BN=%d, K block %d", bn, kbb);
  return RCV_ERR_UNSUPPORTED;
}

// =====================================================================================================================
// 16-channel layers (16 -> <= 16, stride-1 3x3, dilation 1: ROBO_UNet Level1.Conv1, the second conv of every --UNet
// level-1 block, and their input gradients) on the tensor cores: a PERSISTENT kernel.
//
// A tile of 128 positions of such a layer is only 54 small MMAs (9 taps x 16 channels = 18 K steps x 3 products,
// N = 16), ~2 k cycles -- less than the fixed cost of a CTA of the kernel above (barrier init, TMEM allocation,
// first load, epilogue), which is why these layers ran as FFMA2 direct convolutions at 50 % of the FP32 pipe
// (narrow_conv_kernel<16,0,1>, 46 us at batch 64).  Here every CTA allocates once and loops over its tiles
// (blockIdx.x, + gridDim.x, ...) with everything double-buffered:
//   * the whole weight panel (9 taps x 16 rows x 64 B, hi + lo = 18 KB) stays in shared memory for the CTA's lifetime;
//   * 8 producer warps stage the patch of tile t+1 (64-byte rows = 16 channels, SWIZZLE_64B, hi / lo copies, tap
//     shifts by descriptor start as above; up to two positions per thread) while
//   * the MMA warp runs the 54 MMAs of tile t into one of two TMEM accumulator pairs and
//   * 4 epilogue warps drain tile t-1 (fused bias / ReLU / affine / residual / BatchNorm statistics, NCHW stores).
namespace {

constexpr int C16_GT = 160;                     // threads of one producer group: group g stages the tiles lt = g (mod 2)
constexpr int C16_NPROD = 2 * C16_GT;           // producer threads (warps 0-9)
constexpr int C16_NEPI = 128;                   // epilogue threads (warps 11-14: TMEM lane quarters 3, 0, 1, 2)
constexpr int C16_NT = C16_NPROD + 32 + C16_NEPI;
constexpr int C16_WB = 9 * 16 * 64 * 2;         // resident weight panel bytes (hi rows then lo rows per tap block)

bool geometry16(const RcvIgemm& p, HaloGeo* out) {
  if (p.nclass != 1 || p.gs != 1 || p.ostep != 1 || p.taps[0].n != 9 || p.in_scale != nullptr) return false;
  if (p.Hout != p.Hin || p.Wout != p.Win || p.Hg != p.Hin || p.Wg != p.Win) return false;
  if (p.CA != 16 || p.CB > 16 || p.CB < 1) return false;
  const int sgn = p.taps[0].dy[0] < 0 ? 1 : -1;
  for (int t = 0; t < 9; ++t)
    if (p.taps[0].dy[t] != sgn * (t / 3 - 1) || p.taps[0].dx[t] != sgn * (t % 3 - 1)) return false;  // dilation 1
  HaloGeo g;
  memset(&g, 0, sizeof(g));
  g.tsign = sgn;
  g.d = 1;
  g.PW = p.Win + 1;
  g.HP = p.Hin + 1;
  g.S = g.PW + 1;
  g.L = BM + 2 * g.S;
  g.Lpad = (g.L + 7) & ~7;
  g.nkc = 1;
  g.Mh = ((int64_t)p.N * g.HP + 1) * g.PW;
  if (g.L > 2 * C16_GT) return false;  // at most two staged positions per producer thread
  if ((int64_t)p.N * p.CA * p.Hin * p.Win >= (1ll << 30) || g.Mh >= (1ll << 31)) return false;
  *out = g;
  return true;
}

size_t c16_smem(const HaloGeo& g) { return 1024 + 4 * (size_t)g.Lpad * 64 + C16_WB + 512; }

__global__ void __launch_bounds__(C16_NT, 2) umma_c16_kernel(const RcvIgemm p, const HaloGeo g, int ntiles) {
  rcv_pdl_enter();
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gen = smem_raw + (base - raw);
  const uint32_t patch_bytes = (uint32_t)g.Lpad * 64u;   // one copy (hi or lo) of one buffer
  const uint32_t buf_bytes = 2u * patch_bytes;
  const uint32_t wb_s = base + 2u * buf_bytes;           // resident weight panel
  unsigned char* misc = gen + 2u * buf_bytes + C16_WB;
  // barriers: pfull[2], pempty[2], tfull[2], tempty[2], wfull
  const uint32_t bars = smem_u32(misc);
  const uint32_t bar_pfull = bars, bar_pempty = bars + 16, bar_tfull = bars + 32, bar_tempty = bars + 48, bar_wfull = bars + 64;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 80);
  float* s_cst = reinterpret_cast<float*>(misc + 128);   // bias, scale, shift [3][16]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.Hin, W = p.Win, HW = H * W;
  const bool fast = p.math >= RCV_MATH_TF32;

  if (tid < 16) {
    const bool in = tid < p.CB;
    s_cst[tid] = (in && p.bias) ? __ldg(p.bias + tid) : 0.f;
    s_cst[16 + tid] = (in && p.scale) ? __ldg(p.scale + tid) : 1.f;
    s_cst[32 + tid] = (in && p.shift) ? __ldg(p.shift + tid) : 0.f;
  }
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_pfull + 8 * b, C16_GT / 32);
      mbar_init(bar_pempty + 8 * b, 1);
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, C16_NEPI / 32);
    }
    mbar_init(bar_wfull, 1);
    fence_barrier_init();
  }
  if (warp == C16_NPROD / 32) tmem_alloc(smem_u32(tmem_slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == C16_NPROD / 32) {
    // ================================ WEIGHTS + MMA ISSUER =============================
    if (elect_one()) {
      mbar_expect_tx(bar_wfull, C16_WB);
      bulk_g2s(wb_s, p.wpacked, C16_WB, bar_wfull);
      mbar_wait(bar_wfull, 0);
      // Two MMAs per K step instead of three: the tap block holds the hi rows and then the lo rows of the weights,
      // so ONE N = 32 instruction computes a_hi * [b_hi | b_lo] into the adjacent main and correction columns, and
      // a second, N = 16, adds a_lo * b_hi to the correction columns.  At N <= 32 an MMA costs its A-operand fetch
      // (~64 cycles measured), so this is a third off the MMA phase.
      constexpr uint32_t idesc = make_idesc(BM, 16), idesc32 = make_idesc(BM, 32);
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int buf = lt & 1;
        const uint32_t ph = (uint32_t)((lt >> 1) & 1);
        const uint32_t d_main = tmem_base + buf * 32, d_corr = d_main + 16;
        if (lt >= 2) mbar_wait(bar_tempty + 8 * buf, ph ^ 1u);  // the epilogue has drained this accumulator pair
        mbar_wait(bar_pfull + 8 * buf, ph);
        tc_fence_after();
        const uint32_t a_hi_s = base + buf * buf_bytes, a_lo_s = a_hi_s + patch_bytes;
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          const int ty = t / 3 - 1, tx = t - (t / 3) * 3 - 1;
          const uint32_t roff = (uint32_t)(g.S + g.tsign * (ty * g.PW + tx)) * 64u;
          const uint64_t a_hi = make_desc_b<16>(a_hi_s + roff), a_lo = make_desc_b<16>(a_lo_s + roff);
          const uint32_t bb = wb_s + t * (16 * 64 * 2);
          const uint64_t b_hi = make_desc_b<16>(bb);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t acc = (t == 0 && ks == 0) ? 0u : 1u;
            if (!fast) {
              umma_tf32(d_main, a_hi + 2 * ks, b_hi + 2 * ks, idesc32, acc);  // main | correction (a_hi * b_lo)
              umma_tf32(d_corr, a_lo + 2 * ks, b_hi + 2 * ks, idesc, 1u);
            } else {
              umma_tf32(d_main, a_hi + 2 * ks, b_hi + 2 * ks, idesc, acc);
            }
          }
        }
        umma_commit(bar_pempty + 8 * buf);  // the patch buffer may be overwritten once these MMAs have read it
        umma_commit(bar_tfull + 8 * buf);   // ... and the accumulators are complete
      }
    }
  } else if (warp < C16_NPROD / 32) {
    // ================================ PRODUCERS =======================================
    // Two groups of 5 warps; group grp stages the tiles lt = grp, grp + 2, ... into patch buffer grp, so one group's
    // load round trip runs under the other group's stores (a single group exposed one round trip per tile).
    const char* inb = reinterpret_cast<const char*>(p.in);
    const uint32_t cstride = 4u * (uint32_t)HW;
    const int grp = tid / C16_GT, gt = tid - grp * C16_GT;
    const bool one = gt < g.L, two = gt + C16_GT < g.L;  // staged positions gt and gt + 160 of the tile's patch
    unsigned char* dst = gen + grp * buf_bytes;
    for (int lt = grp, use = 0; lt < my_tiles; lt += 2, ++use) {
      const long long q0 = (long long)((int)blockIdx.x + lt * (int)gridDim.x) * BM;
      int n0, i0, j0, n1 = 0, i1 = 0, j1 = 0;
      const bool v0 = one && decode_pos(q0 - g.S + gt, g, p.N, H, W, n0, i0, j0);
      const bool v1 = two && decode_pos(q0 - g.S + gt + C16_GT, g, p.N, H, W, n1, i1, j1);
      float va[16], vb[16];
      {
        const uint32_t b0 = v0 ? 4u * (uint32_t)((n0 * 16) * HW + i0 * W + j0) : 0u;
        const uint32_t b1 = v1 ? 4u * (uint32_t)((n1 * 16) * HW + i1 * W + j1) : 0u;
#pragma unroll
        for (int i = 0; i < 16; ++i) va[i] = v0 ? __ldg(reinterpret_cast<const float*>(inb + (b0 + (uint32_t)i * cstride))) : 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) vb[i] = v1 ? __ldg(reinterpret_cast<const float*>(inb + (b1 + (uint32_t)i * cstride))) : 0.f;
      }
      if (use >= 1) mbar_wait(bar_pempty + 8 * grp, (uint32_t)((use - 1) & 1));
      if (one) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 h, l;
          split_tf32(va[4 * c + 0], h.x, l.x);
          split_tf32(va[4 * c + 1], h.y, l.y);
          split_tf32(va[4 * c + 2], h.z, l.z);
          split_tf32(va[4 * c + 3], h.w, l.w);
          const int off = gt * 64 + ((c ^ ((gt >> 1) & 3)) << 4);  // SWIZZLE_64B of the absolute address
          *reinterpret_cast<float4*>(dst + off) = h;
          if (!fast) *reinterpret_cast<float4*>(dst + patch_bytes + off) = l;
        }
      }
      if (two) {
        const int r2 = gt + C16_GT;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 h, l;
          split_tf32(vb[4 * c + 0], h.x, l.x);
          split_tf32(vb[4 * c + 1], h.y, l.y);
          split_tf32(vb[4 * c + 2], h.z, l.z);
          split_tf32(vb[4 * c + 3], h.w, l.w);
          const int off = r2 * 64 + ((c ^ ((r2 >> 1) & 3)) << 4);
          *reinterpret_cast<float4*>(dst + off) = h;
          if (!fast) *reinterpret_cast<float4*>(dst + patch_bytes + off) = l;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pfull + 8 * grp);
    }
  } else {
    // ================================ EPILOGUE ========================================
    // thread = one accumulator row; the four epilogue warps own TMEM lane quarters (warp % 4)
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;
    const int epi = p.epilogue;
    const int HWo = p.Hout * p.Wout;
    const bool has_res = p.residual != nullptr, has_stats = p.stats != nullptr;
    // BatchNorm statistics of this warp's 32 rows, summed over ALL tiles of this persistent CTA and flushed once:
    // one fp64 atomic pair per channel per tile (2 400 tiles x 4 warps onto 32 addresses) serialised in L2
    double st1 = 0.0, st2 = 0.0;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int buf = lt & 1;
      const long long q0 = (long long)((int)blockIdx.x + lt * (int)gridDim.x) * BM;
      int en = 0, ei = 0, ej = 0;
      const bool mrow = decode_pos(q0 + row, g, p.N, H, W, en, ei, ej);
      const size_t obase = mrow ? (size_t)en * p.CB * HWo + (size_t)ei * p.Wout + ej : 0;
      float res[16];
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 16; ++j) res[j] = (mrow && j < p.CB) ? __ldg(p.residual + obase + (size_t)j * HWo) : 0.f;
      }
      mbar_wait(bar_tfull + 8 * buf, (uint32_t)((lt >> 1) & 1));
      tc_fence_after();
      uint32_t rm[16], rc[16];
      tmem_ld16_nowait(tmem_base + tlane + buf * 32, rm);
      if (!fast) {
        tmem_ld16_nowait(tmem_base + tlane + buf * 32 + 16, rc);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) rc[j] = 0u;
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);  // the accumulator pair is free for tile lt + 2
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float acc = __uint_as_float(rm[j]) + __uint_as_float(rc[j]) + s_cst[j];
        float y = apply_epi(acc, epi, s_cst[16 + j], s_cst[32 + j]);
        if (has_res) y += res[j];
        v[j] = (mrow && j < p.CB) ? y : 0.f;
      }
      if (mrow) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < p.CB) p.out[obase + (size_t)j * HWo] = v[j];
      }
      if (has_stats) {
        float s2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) s2[j] = v[j] * v[j];
        warp_transpose_reduce16(v, lane);
        warp_transpose_reduce16(s2, lane);
        st1 += (double)v[0];
        st2 += (double)s2[0];
      }
    }
    if (has_stats) {
      const int co = lane >> 1;
      if ((lane & 1) == 0 && co < p.CB) {
        atomicAdd(p.stats + co, st1);
        atomicAdd(p.stats + p.CB + co, st2);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C16_NPROD / 32) {
    __syncwarp();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace

bool rcv_umma_c16_ok(const RcvIgemm& p) {
  static const int on = getenv("RCV_UMMA_C16") ? atoi(getenv("RCV_UMMA_C16")) : 1;
  HaloGeo g;
  return on && geometry16(p, &g) && c16_smem(g) <= 113 * 1024;
}

int rcv_launch_igemm_umma_c16(const RcvIgemm& p, cudaStream_t st) {
  HaloGeo g;
  RCV_REQUIRE(geometry16(p, &g), RCV_ERR_UNSUPPORTED, "umma_c16: geometry outside the kernel's limits");
  const size_t smem = c16_smem(g);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(umma_c16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e != cudaSuccess) {
      rcv_set_error("umma_c16: cannot reserve shared memory: %s", cudaGetErrorString(e));
      return RCV_ERR_CUDA;
    }
    cudaFuncSetAttribute(umma_c16_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_done = true;
  }
  const int ntiles = rcv_cdiv(g.Mh, BM);
  const int slots = 2 * sm_count();
  const int grid = ntiles < slots ? ntiles : slots;
  rcv_launch(umma_c16_kernel, dim3(grid), dim3(C16_NT), smem, st, p, g, ntiles);
  RCV_CHECK_LAUNCH("umma_c16_kernel");
  return RCV_OK;
}
