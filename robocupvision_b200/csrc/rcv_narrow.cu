// Narrow-layer direct convolution (<= 16 output channels): the HBM-shaped outer layers of every net
// on the path -- ROBO_UNet Level0/1, the two last decoder stages, the 1x1 class head, PB_FCN conv0/1,
// LabelProp pre/down1/down2/upConv2/3 -- and the input gradients of their neighbours
// (model.py:105-116, 166-199, 403-414).  Same RcvIgemm problem description and the same fused epilogue as the
// other engines, so conv, dilated conv, stride-2 conv, the four parity classes of the transposed conv
// and the 1x1 head are one kernel.
//
//   * TMA halo staging: one elected thread issues cp.async.bulk.tensor.4d over a tensor map of the
//     NCHW input; the box is (tile columns + 8) x (tile rows + halo) x CC channels of one image.  Rows
//     and columns outside the image are zero-filled by the copy engine, so the hot loop carries no
//     padding tests.  Channel chunks are double-buffered on two mbarriers: the copy of chunk k+1
//     runs under the arithmetic of chunk k.
//   * One thread = 4 consecutive grid points of a row x 2 output slots x ALL output channels in
//     registers.  The two slots are two output rows (ordinary conv: the register window of input
//     rows is shared between them) or the two column parities of a transposed conv / stride-2 input
//     gradient (the thread then writes 8 consecutive output floats).  The window is read with one
//     aligned LDS.128 (two for grid stride 2) plus the halo scalars the tap set needs.
//   * Packed FFMA2 (fma.rn.f32x2): the x value is the scalar operand, a pair of output channels the
//     vector operand, read as broadcast LDS.128 from the [channel][tap][CBP] weight image.  Exact fp32
//     FMA arithmetic (no split, no tensor core): results match the ATen CPU path to rounding order.
//   * Epilogue: bias, ReLU / folded-BN affine in either order, residual, float4 NCHW stores,
//     train-mode BatchNorm sum / sum-of-squares (warp shuffle -> shared -> one fp64 atomic per channel).
#include "rcv_narrow.cuh"

namespace {
using namespace rcv_umma;
using namespace rcv_narrow;

constexpr int PIX = 4;    // grid points per thread along x
constexpr int HL = 2;     // window slots left of the first centre column (dx >= -2)
constexpr int NTMAX = 256;

// Tap structure of a problem, fixed at compile time so the whole (row, slot, tap) nest unrolls into
// straight-line LDS / FFMA2 code with immediate weight offsets.  Canonical tap position j (the
// index into the staged weight image) and where it sits in the register window:
//   S1D1 / S1D2: j = ry*3+rx, input offset ((ry-1)*D, (rx-1)*D); slot q = output row +q
//   S2         : j = ky*3+kx, input offset (ky-1, kx-1) around grid point * 2; slot q = output row +q
//   PAR        : row parity a = blockIdx.z, column parity = slot; j = ky*3+kx of the 3x3 kernel
//                (fine = 2*coarse + k - 1): a=0 -> (ky=1, dy=0); a=1 -> (ky=0, dy=+1), (ky=2, dy=0)
//   K1         : j = 0; slot q = output row +q
enum NarrowKind { NK_S1D1 = 0, NK_S1D2 = 1, NK_S2 = 2, NK_PAR = 3, NK_K1 = 4 };

template <int KIND> struct KindTraits;
template <> struct KindTraits<NK_S1D1> { static constexpr int GS = 1, NR = 4, RYMIN = -1, PARITY = 0; };
template <> struct KindTraits<NK_S1D2> { static constexpr int GS = 1, NR = 6, RYMIN = -2, PARITY = 0; };
template <> struct KindTraits<NK_S2> { static constexpr int GS = 2, NR = 5, RYMIN = -1, PARITY = 0; };
template <> struct KindTraits<NK_PAR> { static constexpr int GS = 1, NR = 2, RYMIN = 0, PARITY = 1; };
template <> struct KindTraits<NK_K1> { static constexpr int GS = 1, NR = 2, RYMIN = 0, PARITY = 0; };

// canonical tap index used by (slot q, window row ri, window column slot d = dx+2), or -1
template <int KIND, int Z>
__host__ __device__ constexpr int tap_at(int q, int ri, int d) {
  if (KIND == NK_S1D1 || KIND == NK_S1D2) {
    const int D = KIND == NK_S1D1 ? 1 : 2;
    const int r = ri - q;  // = ry*D
    const int c = d - 2 + D;  // = rx*D
    if (r < 0 || r > 2 * D || (r % D) != 0 || c < 0 || c > 2 * D || (c % D) != 0) return -1;
    return (r / D) * 3 + (c / D);
  }
  if (KIND == NK_S2) {
    const int ky = ri - 2 * q, kx = d - 1;
    if (ky < 0 || ky > 2 || kx < 0 || kx > 2) return -1;
    return ky * 3 + kx;
  }
  if (KIND == NK_PAR) {
    // rows: a=0 -> ri 0 uses ky=1; a=1 -> ri 0 uses ky=2, ri 1 uses ky=0
    int ky = -1;
    if (Z == 0) ky = ri == 0 ? 1 : -1;
    else ky = ri == 0 ? 2 : (ri == 1 ? 0 : -1);
    // columns: slot (parity) 0 -> dx 0 uses kx=1; slot 1 -> dx 0 uses kx=2, dx +1 uses kx=0
    int kx = -1;
    if (q == 0) kx = d == 2 ? 1 : -1;
    else kx = d == 2 ? 2 : (d == 3 ? 0 : -1);
    if (ky < 0 || kx < 0) return -1;
    return ky * 3 + kx;
  }
  // NK_K1
  return (ri == q && d == 2) ? 0 : -1;
}
template <int KIND, int Z>
__host__ __device__ constexpr bool win_used(int idx) {  // is register-window slot idx (= px*GS + dx + 2) read by any tap
  for (int x = 0; x < PIX; ++x)
    for (int d = 0; d < 5; ++d) {
      if (x * KindTraits<KIND>::GS + d != idx) continue;
      for (int q = 0; q < 2; ++q)
        for (int ri = 0; ri < 8; ++ri)
          if (tap_at<KIND, Z>(q, ri, d) >= 0) return true;
    }
  return false;
}

// Output slots per thread.  8-channel kernels: 2 (two output rows share the register window of input rows).
// 16-channel kernels hold 16 x 4 accumulator pairs per slot: with two slots they need ~200 registers and only
// 8 warps fit an SM (ncu: issue slots 37% busy, nothing to switch to during LDS / FFMA2 latencies), so they
// run one output row per thread at ~110 registers and twice the warps; the parity kind keeps both column
// parities in one thread (8 consecutive output floats per store).
template <int CBP, int KIND>
struct Slots { static constexpr int N = (CBP > 8 && KIND != NK_PAR) ? 1 : 2; };
template <int KIND, int SLOTS>
__host__ __device__ constexpr int window_rows() {
  return KindTraits<KIND>::PARITY ? KindTraits<KIND>::NR : KindTraits<KIND>::NR - (2 - SLOTS) * KindTraits<KIND>::GS;
}

struct NarrowCfg {
  int32_t TR, SPR, RS, R, pitch, CC, nchunk, TWg, ctiles, KK, tiles_per_img, total_tiles;
  int32_t NS;            // pipeline stages (2..4)
  int32_t NH, cbp;       // channel groups and channels per group (8 or 16): kernel template selectors
  uint32_t stage_bytes;  // stage stride (128-byte multiple)
  uint32_t tx_bytes;     // bytes one box delivers
  uint64_t wmap;         // 4 bits per canonical tap position: index into the 3x3 (or 1x1) weight kernel, 15 = unused
};

// acc.xy += x * w.xy
__device__ __forceinline__ void ffma2(unsigned long long& acc, float x, float wx, float wy) {
#ifndef RCV_NARROW_NO_FFMA2
  unsigned long long w, xx;
  asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(wx), "f"(wy));
  asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(w));
#else
  float a0 = __uint_as_float((uint32_t)acc), a1 = __uint_as_float((uint32_t)(acc >> 32));
  a0 = fmaf(x, wx, a0);
  a1 = fmaf(x, wy, a1);
  acc = (unsigned long long)__float_as_uint(a0) | ((unsigned long long)__float_as_uint(a1) << 32);
#endif
}
__device__ __forceinline__ float lo_f(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_f(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }

__device__ __forceinline__ float apply_epi(float v, int epi, float sc, float sh) {
  switch (epi) {
    case RCV_EPI_RELU: return fmaxf(v, 0.f);
    case RCV_EPI_RELU_AFFINE: return fmaf(sc, fmaxf(v, 0.f), sh);
    case RCV_EPI_AFFINE_RELU: return fmaxf(fmaf(sc, v, sh), 0.f);
    case RCV_EPI_AFFINE: return fmaf(sc, v, sh);
    default: return v;
  }
}

// The arithmetic of one channel chunk for one thread: fully unrolled over (window row, slot, tap).
template <int CBP, int KIND, int Z, int CBT>
__device__ __forceinline__ void chunk_math(unsigned long long (&acc)[Slots<CBP, KIND>::N][PIX][CBP / 2], const float* __restrict__ xs,
                                           uint32_t wrow0, int cn, int plane, int pitch, int wstep) {
  using KT = KindTraits<KIND>;
  constexpr int GS = KT::GS;
  constexpr int CEN = PIX * GS;
  constexpr int WN = (PIX - 1) * GS + 5;  // register window: index = px*GS + (dx + 2)
  constexpr int SLOTS = Slots<CBP, KIND>::N;
#pragma unroll 1
  for (int c = 0; c < cn; ++c) {
    const float* xr0 = xs + c * plane;
    const uint32_t wrow = wrow0 + c * wstep * 4;
    static_for<window_rows<KIND, SLOTS>()>([&](auto ri_c) {
      constexpr int ri = decltype(ri_c)::value;
      constexpr bool row_used = [] {
        for (int q = 0; q < SLOTS; ++q)
          for (int d = 0; d < 5; ++d)
            if (tap_at<KIND, Z>(q, ri, d) >= 0) return true;
        return false;
      }();
      if constexpr (row_used) {
        const float* xr = xr0 + ri * pitch;
        float wv[WN];
        {
          const float4 a = *reinterpret_cast<const float4*>(xr);
          wv[HL + 0] = a.x; wv[HL + 1] = a.y; wv[HL + 2] = a.z; wv[HL + 3] = a.w;
          if constexpr (GS == 2) {
            const float4 b = *reinterpret_cast<const float4*>(xr + 4);
            wv[HL + 4] = b.x; wv[HL + 5] = b.y; wv[HL + 6] = b.z; wv[HL + 7] = b.w;
          }
        }
        // halo columns this tap structure reads
        if constexpr (win_used<KIND, Z>(0)) wv[0] = xr[-2];
        if constexpr (win_used<KIND, Z>(1)) wv[1] = xr[-1];
        if constexpr (win_used<KIND, Z>(HL + CEN)) wv[HL + CEN] = xr[CEN];
        if constexpr (HL + CEN + 1 < WN) {
          if constexpr (win_used<KIND, Z>(HL + CEN + 1)) wv[HL + CEN + 1] = xr[CEN + 1];
        }
        static_for<SLOTS>([&](auto q_c) {
          constexpr int q = decltype(q_c)::value;
          static_for<5>([&](auto d_c) {
            constexpr int d = decltype(d_c)::value;
            constexpr int j = tap_at<KIND, Z>(q, ri, d);
            if constexpr (j >= 0) {
#pragma unroll
              for (int c4 = 0; c4 < CBP / 4; ++c4) {
                // volatile: one load per use -- no common-subexpression reuse of a tap's weights between
                // the two slots, which would keep up to 3 taps x CBP registers alive across window rows
                float4 w;
                if constexpr (CBP <= 8) {
                  // 8-channel kernels have the registers to keep a tap's weights for the second slot: a plain
                  // load lets the compiler share it between the two slots (half the weight LDS traffic)
                  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                      : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w)
                      : "r"(wrow + (j * CBT + 4 * c4) * 4));
                } else {
                  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                               : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w)
                               : "r"(wrow + (j * CBT + 4 * c4) * 4));
                }
#pragma unroll
                for (int x = 0; x < PIX; ++x) {
                  ffma2(acc[q][x][2 * c4], wv[x * GS + d], w.x, w.y);
                  ffma2(acc[q][x][2 * c4 + 1], wv[x * GS + d], w.z, w.w);
                }
              }
            }
          });
        });
        // grid stride 2: 11-wide windows; keep the compiler from hoisting every row's window at once (spills)
        if constexpr (GS == 2) asm volatile("" ::: "memory");
      }
    });
  }
}

// CBP = output channels one thread accumulates, NH = channel groups the problem is split into (each tile
// is visited once per group; the weight image and the epilogue constants cover all CBP*NH channels)
template <int CBP, int KIND, int NH>
__global__ void __launch_bounds__(CBP <= 8 ? 256 : 128) __maxnreg__((CBP <= 8 || KIND != NK_PAR) ? 128 : 224)
    narrow_conv_kernel(const __grid_constant__ CUtensorMap tmap, const RcvIgemm p, const NarrowCfg cfg) {
  rcv_pdl_enter();
  using KT = KindTraits<KIND>;
  constexpr int GS = KT::GS;
  constexpr int SLOTS = Slots<CBP, KIND>::N;
  constexpr int CBT = CBP * NH;
  constexpr bool par = KT::PARITY != 0;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[4];
  __shared__ float red[2][NTMAX / 32][CBT];  // per-warp BatchNorm partial sums over this CTA's tiles

  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  float* ws = reinterpret_cast<float*>(sgen + cfg.NS * (size_t)cfg.stage_bytes);  // [CA][KK][CBT]
  float* cst = ws + (size_t)p.CA * cfg.KK * CBT;                              // [3][CBT]
  const uint32_t bar0 = smem_u32(&bars[0]);

  // Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...; a tile = (row parity z, image n, row
  // tile rt, column tile ct).  Work items = (tile, channel chunk), double-buffered through two stages:
  // the copy of item i+NS is issued when item i's arithmetic is done, so it also runs under the epilogue
  // stores of a tile and across tile boundaries.
  const int nchunk = cfg.nchunk;
  const int ntiles = cfg.total_tiles;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nitems = my_tiles * nchunk;
  auto decode = [&](int tile_h, int& h, int& z, int& n, int& i0, int& j0) {
    const int tile = tile_h / NH;
    h = tile_h - tile * NH;  // channel group: fastest, so the groups of a tile read its input back to back
    const int per_z = cfg.tiles_per_img * p.N;
    z = tile / per_z;
    const int r = tile - z * per_z;
    n = r / cfg.tiles_per_img;
    const int r2 = r - n * cfg.tiles_per_img;
    const int rt = r2 / cfg.ctiles, ct = r2 - rt * cfg.ctiles;
    i0 = rt * cfg.TR * cfg.RS;
    j0 = ct * cfg.TWg;
  };
  auto issue = [&](int item) {  // thread 0 only
    const int t = item / nchunk, k = item - t * nchunk;
    int h, z, n, i0, j0;
    decode((int)blockIdx.x + t * (int)gridDim.x, h, z, n, i0, j0);
    const int sg = item % cfg.NS;
    const uint32_t bar = bar0 + 8 * sg;
    mbar_expect_tx(bar, cfg.tx_bytes);
    tma_load_4d(sbase + sg * cfg.stage_bytes, &tmap, j0 * GS - 4, i0 * GS + KT::RYMIN, k * cfg.CC, n, bar);
  };

  if (tid == 0) {
    for (int i = 0; i < cfg.NS; ++i) mbar_init(bar0 + 8 * i, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
    for (int i = 0; i < cfg.NS && i < nitems; ++i) issue(i);
  }
  // weight image [channel][canonical tap][CBP] and epilogue constants (plain loads; overlaps the first copy)
  {
    const int KK = cfg.KK;
    const int tot = p.CA * KK * CBT;
    for (int e = tid; e < tot; e += blockDim.x) {
      const int cb = e % CBT, k = e / CBT;
      const int ca = k / KK, j = k - ca * KK;
      const int wi = (int)((cfg.wmap >> (4 * j)) & 15u);
      ws[e] = (cb < p.CB && wi < 9) ? __ldg(p.w + (size_t)ca * p.wsA + (size_t)cb * p.wsB + wi) : 0.f;
    }
    if (tid < CBT) {
      const bool in = tid < p.CB;
      cst[tid] = (in && p.bias) ? __ldg(p.bias + tid) : 0.f;
      cst[CBT + tid] = (in && p.scale) ? __ldg(p.scale + tid) : 1.f;
      cst[2 * CBT + tid] = (in && p.shift) ? __ldg(p.shift + tid) : 0.f;
    }
    if (lane < CBT) { red[0][wid][lane] = 0.f; red[1][wid][lane] = 0.f; }
  }
  __syncthreads();

  const int ty = tid / cfg.SPR, s = tid - ty * cfg.SPR;
  const int pitch = cfg.pitch;
  const int plane = cfg.R * pitch;
  const int toff = (ty * cfg.RS * GS) * pitch + 4 + s * PIX * GS;  // thread's window origin in a channel plane
  const int wstep = cfg.KK * CBT;
  const int epi = p.epilogue;
  const bool has_res = p.residual != nullptr;
  const bool do_stats = p.stats != nullptr;
  const size_t HWo = (size_t)p.Hout * p.Wout;

  int item = 0;
  for (int t = 0; t < my_tiles; ++t) {
    int h, z, n, i0, j0;
    decode((int)blockIdx.x + t * (int)gridDim.x, h, z, n, i0, j0);
    const int gi = i0 + ty * cfg.RS;    // grid row of slot 0
    const int gj = j0 + s * PIX;        // first grid column
    const bool active = ty < cfg.TR && gi < p.Hg && gj < p.Wg;

    unsigned long long acc[SLOTS][PIX][CBP / 2];
#pragma unroll
    for (int q = 0; q < SLOTS; ++q)
#pragma unroll
      for (int x = 0; x < PIX; ++x)
#pragma unroll
        for (int c = 0; c < CBP / 2; ++c) acc[q][x][c] = 0ull;

    for (int k = 0; k < nchunk; ++k, ++item) {
      const int st = item % cfg.NS;
      mbar_wait(bar0 + 8 * st, (item / cfg.NS) & 1);
      if (active) {
        const float* xs = reinterpret_cast<const float*>(sgen + (size_t)st * cfg.stage_bytes) + toff;
        const int ca0 = k * cfg.CC;
        const int cn = min(cfg.CC, p.CA - ca0);
        const uint32_t wrow0 = smem_u32(ws) + (uint32_t)(ca0 * wstep + h * CBP) * 4u;
        if (par && z == 1)
          chunk_math<CBP, KIND, 1, CBT>(acc, xs, wrow0, cn, plane, pitch, wstep);
        else
          chunk_math<CBP, KIND, 0, CBT>(acc, xs, wrow0, cn, plane, pitch, wstep);
      }
      __syncthreads();  // every reader is done with this stage
      if (tid == 0 && item + cfg.NS < nitems) issue(item + cfg.NS);
    }

    // ---------------- epilogue of the tile ----------------
    // slot geometry: ordinary conv -> slot q is output row gi+q, 4 consecutive floats; parity classes ->
    // output row 2*gi+z, slot q is column parity: 8 consecutive floats interleaving the two slots
    bool vq[2];
    size_t ob[2];
    // the residual tensor may hold only the first res_C channels (partial skip): its batch stride differs
    const size_t rshift = (size_t)n * (size_t)(p.CB - p.res_C) * HWo;
    if (par) {
      vq[0] = active; vq[1] = false;  // one row, both slots written together
      ob[0] = (size_t)n * p.CB * HWo + (size_t)(2 * gi + z) * p.Wout + 2 * gj;
      ob[1] = 0;
    } else {
      vq[0] = active; vq[1] = SLOTS == 2 && active && gi + 1 < p.Hg;
      ob[0] = (size_t)n * p.CB * HWo + (size_t)gi * p.Wout + gj;
      ob[1] = ob[0] + p.Wout;
    }
#pragma unroll
    for (int c2 = 0; c2 < CBP / 2; ++c2) {
#pragma unroll
      for (int hl = 0; hl < 2; ++hl) {
        const int cb = h * CBP + 2 * c2 + hl;  // output channel
        float s1 = 0.f, s2 = 0.f;
        if (cb < p.CB) {
          const float bi = cst[cb], sc = cst[CBT + cb], sh = cst[2 * CBT + cb];
          float v[2][PIX];
#pragma unroll
          for (int q = 0; q < SLOTS; ++q)
#pragma unroll
            for (int x = 0; x < PIX; ++x)
              v[q][x] = apply_epi((hl ? hi_f(acc[q][x][c2]) : lo_f(acc[q][x][c2])) + bi, epi, sc, sh);
          if constexpr (par) {
            if (vq[0]) {
              float* o = p.out + ob[0] + (size_t)cb * HWo;
              float4 a = make_float4(v[0][0], v[1][0], v[0][1], v[1][1]);
              float4 b = make_float4(v[0][2], v[1][2], v[0][3], v[1][3]);
              if (has_res && cb < p.res_C) {
                const float4 ra = __ldg(reinterpret_cast<const float4*>(p.residual + (ob[0] - rshift) + (size_t)cb * HWo));
                const float4 rb = __ldg(reinterpret_cast<const float4*>(p.residual + (ob[0] - rshift) + (size_t)cb * HWo) + 1);
                a.x += ra.x; a.y += ra.y; a.z += ra.z; a.w += ra.w;
                b.x += rb.x; b.y += rb.y; b.z += rb.z; b.w += rb.w;
              }
              reinterpret_cast<float4*>(o)[0] = a;
              reinterpret_cast<float4*>(o)[1] = b;
              s1 = (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
              s2 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < SLOTS; ++q) {
              if (!vq[q]) continue;
              float4 a = make_float4(v[q][0], v[q][1], v[q][2], v[q][3]);
              if (has_res && cb < p.res_C) {
                const float4 ra = __ldg(reinterpret_cast<const float4*>(p.residual + (ob[q] - rshift) + (size_t)cb * HWo));
                a.x += ra.x; a.y += ra.y; a.z += ra.z; a.w += ra.w;
              }
              *reinterpret_cast<float4*>(p.out + ob[q] + (size_t)cb * HWo) = a;
              s1 += (a.x + a.y) + (a.z + a.w);
              s2 += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
            }
          }
        }
        if (do_stats) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          if (lane == 0) { red[0][wid][cb] += s1; red[1][wid][cb] += s2; }  // this warp's own row: no race
        }
      }
    }
  }
  if (do_stats) {
    __syncthreads();
    if (tid < 2 * CBT) {
      const int which = tid / CBT, c = tid - which * CBT;
      if (c < p.CB) {
        double tsum = 0.0;
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) tsum += (double)red[which][w][c];
        atomicAdd(p.stats + which * p.CB + c, tsum);
      }
    }
  }
}

// ---------------------------------------------------------------------------- host side
// Does tap set `ts` equal the canonical structure of (KIND, Z, slot-0 view)?  Fills wmap.
// A tap (dy, dx, wi) of an ordinary problem sits at window row dy - RYMIN (slot 0) and column dx + 2.
template <int KIND, int Z>
bool match_taps(const RcvTapSet& ts, int q, int8_t* wmap, int* seen) {
  using KT = KindTraits<KIND>;
  for (int t = 0; t < ts.n; ++t) {
    const int d = ts.dx[t] + 2;
    // slot q's own row offset inside the window: ordinary kinds shift by q*GS, parity kinds do not
    const int ri = ts.dy[t] - KT::RYMIN + (KT::PARITY ? 0 : q * KT::GS);
    if (d < 0 || d > 4 || ri < 0 || ri >= KT::NR) return false;
    const int j = tap_at<KIND, Z>(q, ri, d);
    if (j < 0) return false;
    if (wmap[j] >= 0 && wmap[j] != ts.wi[t]) return false;
    wmap[j] = ts.wi[t];
    ++*seen;
  }
  return true;
}
template <int KIND, int Z>
int count_taps(int q) {
  int n = 0;
  for (int ri = 0; ri < KindTraits<KIND>::NR; ++ri)
    for (int d = 0; d < 5; ++d) n += tap_at<KIND, Z>(q, ri, d) >= 0;
  return n;
}

// Classify the problem; -1 when its tap structure is none of the compiled kinds.
int classify(const RcvIgemm& p, int8_t wmap[9]) {
  for (int j = 0; j < 9; ++j) wmap[j] = -1;
  int seen = 0;
  if (p.ostep == 2) {
    if (p.nclass != 4 || p.gs != 1) return -1;
    bool ok = match_taps<NK_PAR, 0>(p.taps[0], 0, wmap, &seen) && match_taps<NK_PAR, 0>(p.taps[1], 1, wmap, &seen) &&
              match_taps<NK_PAR, 1>(p.taps[2], 0, wmap, &seen) && match_taps<NK_PAR, 1>(p.taps[3], 1, wmap, &seen);
    ok = ok && p.taps[0].n == count_taps<NK_PAR, 0>(0) && p.taps[1].n == count_taps<NK_PAR, 0>(1) &&
         p.taps[2].n == count_taps<NK_PAR, 1>(0) && p.taps[3].n == count_taps<NK_PAR, 1>(1);
    return ok ? NK_PAR : -1;
  }
  if (p.ostep != 1 || p.nclass != 1) return -1;
  const RcvTapSet& ts = p.taps[0];
  auto try_kind = [&](auto kind_c) -> bool {
    constexpr int K = decltype(kind_c)::value;
    if (p.gs != KindTraits<K>::GS) return false;
    for (int j = 0; j < 9; ++j) wmap[j] = -1;
    int sn = 0;
    return match_taps<K, 0>(ts, 0, wmap, &sn) && ts.n == count_taps<K, 0>(0);
  };
  if (ts.n == 1 && try_kind(std::integral_constant<int, NK_K1>{})) return NK_K1;
  if (try_kind(std::integral_constant<int, NK_S1D1>{})) return NK_S1D1;
  if (try_kind(std::integral_constant<int, NK_S1D2>{})) return NK_S1D2;
  if (try_kind(std::integral_constant<int, NK_S2>{})) return NK_S2;
  return -1;
}

int kind_nr(int kind) {
  switch (kind) {
    case NK_S1D1: return KindTraits<NK_S1D1>::NR;
    case NK_S1D2: return KindTraits<NK_S1D2>::NR;
    case NK_S2: return KindTraits<NK_S2>::NR;
    case NK_PAR: return KindTraits<NK_PAR>::NR;
    default: return KindTraits<NK_K1>::NR;
  }
}

// Tile plan of a problem; false when the geometry is outside the kernel's limits.
bool plan(const RcvIgemm& p, NarrowCfg* out, int* kind_out, int* nthreads, size_t* smem) {
  NarrowCfg c;
  memset(&c, 0, sizeof(c));
  if (p.CB > 16 || p.CB < 1) return false;
  if (p.gs != 1 && p.gs != 2) return false;
  if ((p.Wg & 3) || (p.Win & 3) || (p.Wout & 3)) return false;
  if (((uintptr_t)p.in | (uintptr_t)p.out | (uintptr_t)p.residual) & 15) return false;
  if (p.N > 65535) return false;
  const bool par = p.ostep == 2;
  if (par) {
    if (p.Hout != 2 * p.Hg || p.Wout != 2 * p.Wg) return false;
  } else {
    if (p.Hout != p.Hg || p.Wout != p.Wg) return false;
  }
  int8_t wmap[9];
  const int kind = classify(p, wmap);
  if (kind < 0) return false;
  for (int j = 0; j < 9; ++j) c.wmap |= (uint64_t)(wmap[j] < 0 ? 15 : wmap[j]) << (4 * j);
  // > 8 output channels.  Ordinary kinds: one 16-channel pass, one output row per thread (measured faster than
  // two 8-channel passes: 46 vs 52 us for 16->16 @60x80 x64).  Parity kind: both column parities stay in one
  // thread, which at 16 channels is ~200 registers and 8 warps per SM; two groups of 8 channels through the
  // 8-channel kernel run 3-4 CTAs per SM instead (36 vs 44 us for convT 32->16 @30x40 x64).
  // RCV_NARROW_SPLIT16 = 0 / 1 forces one policy for every kind (A/B runs).
  static const int split16 = env_int("RCV_NARROW_SPLIT16", -1);
  const bool split = p.CB > 8 && (split16 < 0 ? kind == NK_PAR : split16 != 0);
  c.NH = split ? 2 : 1;
  c.cbp = (p.CB > 8 && !split) ? 16 : 8;
  const int slots = (c.cbp > 8 && kind != NK_PAR) ? 1 : 2;
  c.RS = par ? 1 : slots;
  c.KK = kind == NK_K1 ? 1 : 9;
  const int NR = kind_nr(kind) - (par ? 0 : (2 - slots) * p.gs);
  // tile: columns
  const int maxcols = (256 - 8) / p.gs;
  c.ctiles = rcv_cdiv(p.Wg, maxcols);
  c.TWg = ((rcv_cdiv(p.Wg, c.ctiles) + 3) / 4) * 4;
  c.SPR = c.TWg / PIX;
  c.pitch = c.TWg * p.gs + 8;
  if (c.pitch > 256 || c.SPR > NTMAX) return false;
  // tile: rows
  // 16-channel parity kernels hold 128 accumulator registers: 128-thread CTAs at <= 224 registers, two per SM
  int want = (c.cbp > 8 && kind == NK_PAR) ? 128 : env_int("RCV_NARROW_THREADS", 128);
  want = (c.cbp > 8 && want > 128) ? 128 : want;  // launch bounds of the 16-channel kernels
  const int trmax = rcv_cdiv(p.Hg, c.RS);
  int TR = want / c.SPR;
  TR = TR < 1 ? 1 : TR;
  TR = TR > trmax ? trmax : TR;
  while (TR > 1 && (TR - 1) * c.RS * p.gs + NR > 256) --TR;
  c.TR = TR;
  c.R = (TR - 1) * c.RS * p.gs + NR;
  if (c.R > 256) return false;
  // channels per stage: two stages + weights within a shared-memory budget that leaves room for 3-4 CTAs
  // per SM; chunks of equal size
  const size_t wbytes = ((size_t)p.CA * c.KK + 3) * 16 * 4;
  const size_t per_ch = (size_t)c.R * c.pitch * 4;
  int CC = p.CA < 8 ? p.CA : 8;
  const size_t budget = (size_t)env_int("RCV_NARROW_SMEM_KB", 56) * 1024;
  int NS = env_int("RCV_NARROW_STAGES", 2);
  NS = NS < 2 ? 2 : NS > 4 ? 4 : NS;
  c.NS = NS;
  while (CC > 1 && NS * ((CC * per_ch + 127) & ~(size_t)127) + wbytes > budget) --CC;
  c.nchunk = rcv_cdiv(p.CA, CC);
  CC = rcv_cdiv(p.CA, c.nchunk);
  c.CC = CC;
  c.stage_bytes = (uint32_t)((CC * per_ch + 127) & ~(size_t)127);
  c.tx_bytes = (uint32_t)(CC * per_ch);
  const size_t total = NS * (size_t)c.stage_bytes + wbytes + 256;
  if (total > 200 * 1024) return false;
  c.tiles_per_img = rcv_cdiv(p.Hg, c.TR * c.RS) * c.ctiles;
  const int64_t tt = (int64_t)c.tiles_per_img * p.N * (par ? 2 : 1) * c.NH;
  if (tt >= (1ll << 31)) return false;
  c.total_tiles = (int)tt;
  *out = c;
  *kind_out = kind;
  *nthreads = ((TR * c.SPR + 31) / 32) * 32;
  *smem = total;
  return true;
}

template <int CBP, int KIND, int NH>
int launch(const RcvIgemm& p, const NarrowCfg& cfg, int nthreads, size_t smem, cudaStream_t st) {
  CUtensorMap tmap;
  int rc = make_nchw_map(&tmap, p.in, p.N, p.CA, p.Hin, p.Win, cfg.pitch, cfg.R, cfg.CC, "narrow_conv");
  if (rc) return rc;
  static int ctas_per_sm = 1;  // occupancy of this instantiation at the last (block size, shared memory) it was asked for
  static int last_nt = 0;
  static size_t last_smem = 0;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(narrow_conv_kernel<CBP, KIND, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(narrow_conv_kernel<CBP, KIND, NH>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
  }
  if (last_nt != nthreads || last_smem != smem) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, narrow_conv_kernel<CBP, KIND, NH>, nthreads, smem);
    ctas_per_sm = occ < 1 ? 1 : occ;
    last_nt = nthreads;
    last_smem = smem;
  }
  const int64_t slots = (int64_t)num_sms * ctas_per_sm;
  const int grid = (int)(cfg.total_tiles < slots ? cfg.total_tiles : slots);
  rcv_launch(narrow_conv_kernel<CBP, KIND, NH>, dim3(grid), dim3(nthreads), smem, st, tmap, p, cfg);
  RCV_CHECK_LAUNCH("narrow_conv_kernel");
  return RCV_OK;
}

template <int CBP, int NH>
int launch_kind(int kind, const RcvIgemm& p, const NarrowCfg& c, int nt, size_t sm, cudaStream_t st) {
  switch (kind) {
    case NK_S1D1: return launch<CBP, NK_S1D1, NH>(p, c, nt, sm, st);
    case NK_S1D2: return launch<CBP, NK_S1D2, NH>(p, c, nt, sm, st);
    case NK_S2: return launch<CBP, NK_S2, NH>(p, c, nt, sm, st);
    case NK_PAR: return launch<CBP, NK_PAR, NH>(p, c, nt, sm, st);
    default: return launch<CBP, NK_K1, NH>(p, c, nt, sm, st);
  }
}

}  // namespace

bool rcv_narrow_supported(const RcvIgemm& p) {
  NarrowCfg c;
  int nt, kind;
  size_t sm;
  return plan(p, &c, &kind, &nt, &sm);
}

int rcv_launch_narrow(const RcvIgemm& p, cudaStream_t st) {
  NarrowCfg c;
  int nt, kind;
  size_t sm;
  RCV_REQUIRE(plan(p, &c, &kind, &nt, &sm), RCV_ERR_UNSUPPORTED, "narrow_conv: geometry outside the kernel's limits");
  if (c.cbp == 16) return launch_kind<16, 1>(kind, p, c, nt, sm, st);
  return c.NH == 2 ? launch_kind<8, 2>(kind, p, c, nt, sm, st) : launch_kind<8, 1>(kind, p, c, nt, sm, st);
}
