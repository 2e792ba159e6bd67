// Weight gradient on the tensor cores (tcgen05 3xTF32, TMEM accumulators):
//   dW[cb][k = (ca, tap)] = sum over pixels m of  row[m][cb] * gathered[m][k]
// as a GEMM whose REDUCTION runs over pixels:  D[k][cb] += A[k][32 px] * B[cb][32 px]^T
//   A (MMA M = 128 rows) = (channel, tap) slices of the gathered tensor (x for a conv), shifted by
//       the tap and zero outside the image: 4-byte loads, lanes along pixels;
//   B (MMA N = BN rows)  = channels of the dense tensor (dy for a conv): float4 loads along pixels.
// Both operands are pixel-contiguous in NCHW, i.e. naturally K-major for this GEMM.  Putting the
// (channel, tap) rows on M keeps the tile full for every layer (N = 16..128 follows Cout), and makes
// the accumulator rows consecutive weights: the epilogue's fp32 RED atomics (lanes = TMEM lanes =
// consecutive k) are coalesced without a transposition.
// Stride-1 3x3 layers (the bulk of the nets) take the QUAD gather: a k tile is 42 (channel, tap-row)
// units of 3 rows (tap columns); a thread loads one aligned float4 of 4 consecutive pixels per
// (unit, quad), gets the +-dilation neighbours from the adjacent lanes by shuffle (two scalar edge
// loads per 8 lanes), and emits the three tap-column rows as STS.128: 12x fewer load instructions
// and 4x fewer store instructions than the element-wise gather, which remains for stride-2 /
// transposed / 1x1 geometries.
// A CTA owns one 128-row k tile x one BN-wide channel tile and a contiguous pixel range (split over
// pixels across the grid); G producer groups of 128 threads each stage whole 32-pixel chunks
// (hi / lo tf32 parts, 128B-swizzled K-major tiles), one thread issues the MMAs.  The bias
// gradient (sum of the dense tensor over pixels) is accumulated by the B loaders of k tile 0.
#include <stdlib.h>

#include "rcv_common.cuh"
#include "rcv_umma.cuh"

using namespace rcv_umma;

// Phase-timing instrumentation (tools/umma_phases.py wgrad): compiled in only with -DRCV_PROF=1
#ifndef RCV_PROF
#define RCV_PROF 0
#endif
#if RCV_PROF
#define WPROF_ON(cond) (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (cond))
#define WPROF_I(e) do { if (WPROF_ON(c < 128)) p.prof[2048 + c * 4 + (e)] = clock64(); } while (0)
#define WPROF_P(e) do { if (WPROF_ON(gt == 0 && c < 128)) p.prof[c * 8 + (e)] = clock64(); } while (0)
#define WPROF_E(e) do { if (WPROF_ON(gt == 0)) p.prof[4000 + grp * 4 + (e)] = clock64(); } while (0)
#else
#define WPROF_I(e) do { } while (0)
#define WPROF_P(e) do { } while (0)
#define WPROF_E(e) do { } while (0)
#endif

namespace {

constexpr int BM = 128;        // (channel, tap) rows per tile (MMA M)
constexpr int BK = 32;         // pixels per chunk: one 128-byte swizzle row
constexpr int GTHREADS = 128;
constexpr int MAXT = 9;

template <int BN_>
struct WCfg {
  static constexpr int BN = BN_;                    // dense-tensor channels per tile (MMA N)
  static constexpr int G = BN_ == 128 ? 3 : 4;      // producer groups = ring depth
  static constexpr int NPROD = G * GTHREADS;
  static constexpr int NT = NPROD + 32;             // + the MMA-issuer warp
  static constexpr int A_TILE = BM * 128;           // bytes of the hi or lo gathered tile
  static constexpr int B_TILE = BN_ * 128;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;  // A hi, A lo, B hi, B lo
  static constexpr int TILE_BYTES = G * STAGE;
  static constexpr int SMEM = TILE_BYTES + 1024 + 2048;  // + alignment slack + barriers / tables
  static constexpr int TCOLS = 2 * BN_ < 32 ? 32 : 2 * BN_;
  static constexpr int BROWS = BN_ / 16;            // dense rows per loader thread
};

constexpr int QUNITS = 42;  // most (channel, tap-row) units per k tile in QUAD mode: 126 of the 128 rows;
                            // p.qunits <= 42 balances the units over the k tiles

template <int BN, bool QUAD>
__global__ void __launch_bounds__(WCfg<BN>::NT, 1) umma_wgrad_kernel(const RcvWgrad p) {
  rcv_pdl_enter();
  using C = WCfg<BN>;
  constexpr int G = C::G, NPROD = C::NPROD, NT = C::NT;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  unsigned char* gen_tiles = smem_raw + (tiles - raw);
  unsigned char* misc = gen_tiles + C::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);  // full[G] empty[G] done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 96);
  int2* s_tab = reinterpret_cast<int2*>(misc + 128);    // [BM] (element offset, tap | 31)
  int* s_wo = reinterpret_cast<int*>(misc + 128 + BM * 8);  // [BM] weight offset of row k, -1 beyond K
  int* s_tdy = s_wo + BM;                               // [MAXT] tap offsets (avoids indexing the params)
  int* s_tdx = s_tdy + 16;
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * G, bar_done = bar_empty + 8 * G;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int T = p.taps.n;
  const int K = p.CA * T;
  const int HWin = p.Hin * p.Win;
  const int HWg = p.Hg * p.Wg;
  const int M = p.N * HWg;
  const int k0 = QUAD ? blockIdx.x * (3 * p.qunits) : blockIdx.x * BM;  // QUAD: qunits units x 3 tap columns
  const int cb0 = blockIdx.y * BN;
  const int mbeg = blockIdx.z * p.slab;
  const int mend = min(M, mbeg + p.slab);
  const int nchunks = (mend - mbeg + BK - 1) / BK;
  const bool fast = p.math >= RCV_MATH_TF32;  // one TF32 MMA per product: no lo halves, no correction accumulator

  for (int kl = tid; kl < BM; kl += NT) {
    const int k = k0 + kl;
    if (k < K && (!QUAD || kl < 3 * p.qunits)) {
      const int ca = k / T, t = k - ca * T;
      s_tab[kl] = make_int2(ca * HWin + p.taps.dy[t] * p.Win + p.taps.dx[t], t);
      s_wo[kl] = ca * p.wsA + p.taps.wi[t];
    } else {
      s_tab[kl] = make_int2(0, 31);
      s_wo[kl] = -1;
    }
  }
  if (tid < MAXT) {
    s_tdy[tid] = tid < T ? p.taps.dy[tid] : 0;
    s_tdx[tid] = tid < T ? p.taps.dx[tid] : 0;
  }
  // rows of B beyond the dense tensor's channel count stay zero for the whole kernel
  if (cb0 + BN > p.CB) {
    for (int g = 0; g < G; ++g) {
      float4* b = reinterpret_cast<float4*>(gen_tiles + g * C::STAGE + 2 * C::A_TILE);
      for (int i = tid; i < 2 * C::B_TILE / 16; i += NT) b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (QUAD) {
    for (int g = 0; g < G; ++g) {
      float4* a = reinterpret_cast<float4*>(gen_tiles + g * C::STAGE);
      for (int i = tid; i < 2 * C::A_TILE / 16; i += NT) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (tid == 0) {
    for (int s = 0; s < G; ++s) {
      mbar_init(bar_full + 8 * s, GTHREADS / 32);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == NPROD / 32) tmem_alloc(smem_u32(tmem_slot), C::TCOLS);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == NPROD / 32) {
    // ================================ MMA ISSUER ======================================
    if (elect_one() && nchunks > 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN), idesc2 = make_idesc(BM, BN < 128 ? 2 * BN : BN);
      const uint32_t d_main = tmem_base, d_corr = tmem_base + BN;
      mbar_wait(bar_full, 0);
      for (int c0 = 0; c0 < nchunks; c0 += G) {
        const uint32_t par = (uint32_t)((c0 / G) & 1);
#pragma unroll
        for (int st = 0; st < G; ++st) {
          const int c = c0 + st;
          if (c < nchunks) {
            WPROF_I(0);
            tc_fence_after();
            const uint32_t base = tiles + st * C::STAGE;
            const uint64_t a_hi = make_desc(base), a_lo = make_desc(base + C::A_TILE);
            const uint64_t b_hi = make_desc(base + 2 * C::A_TILE);  // BN hi rows, then the BN lo rows
            // parity mode: a_hi * [b_hi | b_lo] in ONE instruction with N = 2*BN (the correction accumulator follows
            // the main one in TMEM), then a_lo * b_hi: two MMAs per K step instead of three
            // (at BN = 128 the N = 256 form measured 2 % slower -- those tiles are throughput-bound -- and keeps three)
            constexpr bool FUSE = BN < 128;
            const uint64_t b_lo = make_desc(base + 2 * C::A_TILE + C::B_TILE);
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
              const uint32_t acc = (ks == 0) ? (uint32_t)(c != 0) : 1u;
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
            WPROF_I(1);
            if (c + 1 < nchunks) mbar_wait(bar_full + 8 * ((st + 1) % G), st + 1 == G ? par ^ 1u : par);
            WPROF_I(2);
            umma_commit(bar_empty + 8 * st);
            if (c == nchunks - 1) umma_commit(bar_done);
            WPROF_I(3);
          }
        }
      }
    }
  } else {
    // ================================ PRODUCERS =======================================
    const int gt = tid & (GTHREADS - 1);
    const int grp = tid / GTHREADS;
    unsigned char* st = gen_tiles + grp * C::STAGE;
    unsigned char* sb = st + 2 * C::A_TILE;
    const uint32_t my_full = bar_full + 8 * grp, my_empty = bar_empty + 8 * grp;
    // dense (B) loader mapping: 16-byte chunk bc of rows br0 + 16*i (a warp covers 4 rows x 128 B)
    const int bc = gt & 7, br0 = gt >> 3;
    // gathered (A) loader mapping: pixel = lane, rows aw + 4*i
    const int aw = gt >> 5;
    const bool do_bias = p.dbias != nullptr && blockIdx.x == 0;
    float bsum[C::BROWS];
#pragma unroll
    for (int i = 0; i < C::BROWS; ++i) bsum[i] = 0.f;

    for (int c = grp, use = 0; c < nchunks; c += G, ++use) {
      const int mc = mbeg + c * BK;
      WPROF_P(0);
      // ---- B: dense tensor, rows = channels, float4 along pixels ----
      float4 rb[C::BROWS];
      {
        const int m = mc + bc * 4;  // HWg % 4 == 0: a quad never straddles two images
        const bool ok = m < mend;
        int n = 0, r = 0;
        if (ok) { n = m / HWg; r = m - n * HWg; }
        const float* src = p.row + ((size_t)n * p.CB + cb0) * HWg + r;
#pragma unroll
        for (int i = 0; i < C::BROWS; ++i) {
          const int rowc = br0 + 16 * i;
          rb[i] = (ok && cb0 + rowc < p.CB) ? __ldg(reinterpret_cast<const float4*>(src + (size_t)rowc * HWg))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      // ---- A: gathered tensor ----
      float ra[32];    // element-wise gather: rows = (channel, tap), lanes along pixels
      float4 qv[3];    // QUAD gather: one aligned float4 per (unit, quad) item, 3 items per thread
      float ql[3][2], qr[3][2];  // left / right neighbour pixels of each item (dilation 1 uses [1] / [0])
      const int qq = gt & 7;     // QUAD: pixel quad of the chunk handled by this thread
      const int qu0 = gt >> 3;   // QUAD: units qu0 + 16*r
      if (QUAD) {
        const int d = p.taps.dx[2];  // dilation: taps are (-d, 0, +d)
        const bool nl = p.in_scale != nullptr, nl_relu = p.in_relu != 0;  // normalise-on-load (uniform)
        const int m = mc + qq * 4;
        const bool ok = m < M;       // values beyond this CTA's pixel range are still needed as neighbours
        int n = 0, i0 = 0, j0 = 0;
        if (ok) {
          n = m / HWg;
          const int r = m - n * HWg;
          i0 = r / p.Wg;
          j0 = r - i0 * p.Wg;
        }
        const bool has_l = j0 > 0, has_r = j0 + 4 < p.Wg;  // the row continues to the left / right
        // Phase 1: every load of the three items (the aligned quad and, at the edges of the 8-quad chunk, the
        // neighbour scalars) is issued before any of them is used.  With the shuffles inside the same loop the three
        // items' round trips ran one after the other (phase timing: ~5400 cycles to get through the load section of
        // a chunk, 3 x ~1300 + the dense operand's batch, against ~1050 cycles of MMAs per chunk).
        bool uok[3];
        float isc[3], ish[3];
        float el[3][2], er[3][2];  // edge neighbours from memory (left: [0] = -2, [1] = -1; right: [0] = +4, [1] = +5)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int u = qu0 + 16 * r;          // unit inside the tile
          const int U = blockIdx.x * p.qunits + u;  // global unit = ca * 3 + tap row
          const int ca = U / 3, ty = U - ca * 3;
          const int iy = i0 + (ty - 1) * d;
          uok[r] = ok && u < p.qunits && ca < p.CA && (unsigned)iy < (unsigned)p.Hin;
          const float* src = p.src + ((size_t)n * p.CA + ca) * HWin + iy * p.Win + j0;
          qv[r] = uok[r] ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
          el[r][0] = el[r][1] = er[r][0] = er[r][1] = 0.f;
          if (qq == 0 && uok[r] && has_l) {
            el[r][1] = __ldg(src - 1);
            if (d == 2) el[r][0] = __ldg(src - 2);
          }
          if (qq == 7 && uok[r] && has_r) {
            er[r][0] = __ldg(src + 4);
            if (d == 2) er[r][1] = __ldg(src + 5);
          }
          isc[r] = 1.f;
          ish[r] = 0.f;
          if (nl && uok[r]) {
            isc[r] = __ldg(p.in_scale + ca);
            ish[r] = __ldg(p.in_shift + ca);
          }
        }
        // Phase 2: the producer block's BatchNorm (real pixels only: padding stays zero), then the neighbours
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          if (nl && uok[r]) {
            qv[r].x = fmaf(isc[r], qv[r].x, ish[r]);
            qv[r].y = fmaf(isc[r], qv[r].y, ish[r]);
            qv[r].z = fmaf(isc[r], qv[r].z, ish[r]);
            qv[r].w = fmaf(isc[r], qv[r].w, ish[r]);
            if (nl_relu) {
              qv[r].x = fmaxf(qv[r].x, 0.f);
              qv[r].y = fmaxf(qv[r].y, 0.f);
              qv[r].z = fmaxf(qv[r].z, 0.f);
              qv[r].w = fmaxf(qv[r].w, 0.f);
            }
            if (qq == 0 && has_l) {
              el[r][1] = fmaf(isc[r], el[r][1], ish[r]);
              if (d == 2) el[r][0] = fmaf(isc[r], el[r][0], ish[r]);
              if (nl_relu) { el[r][1] = fmaxf(el[r][1], 0.f); el[r][0] = fmaxf(el[r][0], 0.f); }
            }
            if (qq == 7 && has_r) {
              er[r][0] = fmaf(isc[r], er[r][0], ish[r]);
              if (d == 2) er[r][1] = fmaf(isc[r], er[r][1], ish[r]);
              if (nl_relu) { er[r][0] = fmaxf(er[r][0], 0.f); er[r][1] = fmaxf(er[r][1], 0.f); }
            }
          }
          // neighbours from the adjacent lanes (same unit, adjacent quad) ...
          ql[r][0] = __shfl_up_sync(0xffffffffu, qv[r].z, 1);
          ql[r][1] = __shfl_up_sync(0xffffffffu, qv[r].w, 1);
          qr[r][0] = __shfl_down_sync(0xffffffffu, qv[r].x, 1);
          qr[r][1] = __shfl_down_sync(0xffffffffu, qv[r].y, 1);
          // ... except at the edges of the 8-quad chunk, where they came from memory
          if (qq == 0) { ql[r][0] = el[r][0]; ql[r][1] = el[r][1]; }
          if (qq == 7) { qr[r][0] = er[r][0]; qr[r][1] = er[r][1]; }
          if (!has_l) { ql[r][0] = 0.f; ql[r][1] = 0.f; }
          if (!has_r) { qr[r][0] = 0.f; qr[r][1] = 0.f; }
        }
      } else {
      {
        const int m = mc + lane;
        const bool ok = m < mend;
        int n = 0, i0 = 0, j0 = 0;
        if (ok) {
          n = m / HWg;
          const int r = m - n * HWg;
          i0 = r / p.Wg;
          j0 = r - i0 * p.Wg;
        }
        const int gy0 = i0 * p.gs, gx0 = j0 * p.gs;
        uint32_t tapmask = 0;
        if (ok) {
          for (int t = 0; t < T; ++t) {
            const int iy = gy0 + s_tdy[t], ix = gx0 + s_tdx[t];
            if ((unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win) tapmask |= 1u << t;
          }
        }
        const float* src = p.src + (size_t)n * p.CA * HWin + gy0 * p.Win + gx0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int2 e = s_tab[aw + 4 * i];
          ra[i] = ((tapmask >> e.y) & 1u) ? __ldg(src + e.x) : 0.f;
        }
      }
      }
      WPROF_P(1);
      if (use > 0) mbar_wait(my_empty, (uint32_t)((use - 1) & 1));
      WPROF_P(2);
#pragma unroll
      for (int i = 0; i < C::BROWS; ++i) {
        const int rowc = br0 + 16 * i;
        if (cb0 + rowc < p.CB) {
          float4 h, l;
          split_tf32(rb[i].x, h.x, l.x);
          split_tf32(rb[i].y, h.y, l.y);
          split_tf32(rb[i].z, h.z, l.z);
          split_tf32(rb[i].w, h.w, l.w);
          const int off = rowc * 128 + ((bc ^ (rowc & 7)) << 4);
          *reinterpret_cast<float4*>(sb + off) = h;
          if (!fast) *reinterpret_cast<float4*>(sb + C::B_TILE + off) = l;
          bsum[i] += (rb[i].x + rb[i].y) + (rb[i].z + rb[i].w);
        }
      }
      if (QUAD) {
        const int d = p.taps.dx[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int u = qu0 + 16 * r;
          if (u < p.qunits) {
            const float4 v = qv[r];
            float4 row[3];
            row[1] = v;
            if (d == 1) {
              row[0] = make_float4(ql[r][1], v.x, v.y, v.z);
              row[2] = make_float4(v.y, v.z, v.w, qr[r][0]);
            } else {
              row[0] = make_float4(ql[r][0], ql[r][1], v.x, v.y);
              row[2] = make_float4(v.z, v.w, qr[r][0], qr[r][1]);
            }
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
              const int rowk = u * 3 + tx;
              float4 h, l;
              split_tf32(row[tx].x, h.x, l.x);
              split_tf32(row[tx].y, h.y, l.y);
              split_tf32(row[tx].z, h.z, l.z);
              split_tf32(row[tx].w, h.w, l.w);
              const int off = rowk * 128 + ((qq ^ (rowk & 7)) << 4);
              *reinterpret_cast<float4*>(st + off) = h;
              if (!fast) *reinterpret_cast<float4*>(st + C::A_TILE + off) = l;
            }
          }
        }
      } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int rowk = aw + 4 * i;
        float h, l;
        split_tf32(ra[i], h, l);
        const int off = rowk * 128 + (((lane >> 2) ^ (rowk & 7)) << 4) + (lane & 3) * 4;
        *reinterpret_cast<float*>(st + off) = h;
        if (!fast) *reinterpret_cast<float*>(st + C::A_TILE + off) = l;
      }
      }
      WPROF_P(3);
      fence_proxy_async_smem();
      WPROF_P(4);
      __syncwarp();
      if (lane == 0) mbar_arrive(my_full);
      WPROF_P(5);
    }

    if (do_bias) {
      // the 8 lanes that share a row (chunks 0..7) are consecutive lanes
#pragma unroll
      for (int i = 0; i < C::BROWS; ++i) {
        float s = bsum[i];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        const int rowc = br0 + 16 * i;
        if (bc == 0 && cb0 + rowc < p.CB && s != 0.f) atomicAdd(p.dbias + cb0 + rowc, s);
      }
    }

    // ================================ EPILOGUE ========================================
    // thread = accumulator row = one weight offset; consecutive lanes hit consecutive addresses
    if (nchunks > 0) {
      WPROF_E(0);
      mbar_wait(bar_done, 0);
      WPROF_E(1);
      tc_fence_after();
      const int lrow = (warp & 3) * 32 + lane;
      const int wo = s_wo[lrow];
      const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
      for (int c0 = grp * 16; c0 < BN; c0 += G * 16) {
        if (cb0 + c0 >= p.CB) break;  // warp-uniform
        uint32_t rm[16], rc[16];
        tmem_ld16_nowait(trow + c0, rm);
        if (!fast) {
          tmem_ld16_nowait(trow + BN + c0, rc);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) rc[j] = 0u;
        }
        tmem_ld_wait();
        if (wo >= 0) {
          float* dst = p.dw + (size_t)(cb0 + c0) * p.wsB + wo;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cb0 + c0 + j < p.CB) atomicAdd(dst + (size_t)j * p.wsB, __uint_as_float(rm[j]) + __uint_as_float(rc[j]));
        }
      }
      WPROF_E(2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NPROD / 32) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TCOLS);
  }
}

template <int BN, bool QUAD>
int launch_w(RcvWgrad p, cudaStream_t st) {
  using C = WCfg<BN>;
  static_assert(C::SMEM <= 227 * 1024, "shared memory budget");
  const int K = p.CA * p.taps.n;
  const int64_t M = (int64_t)p.N * p.Hg * p.Wg;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e =
        cudaFuncSetAttribute(umma_wgrad_kernel<BN, QUAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) {
      rcv_set_error("umma_wgrad: cannot reserve %d B of shared memory: %s", C::SMEM, cudaGetErrorString(e));
      return RCV_ERR_CUDA;
    }
    attr_done = true;
  }
  int ktiles = rcv_cdiv(K, BM);
  if (QUAD) {
    ktiles = rcv_cdiv(p.CA * 3, QUNITS);
    p.qunits = rcv_cdiv(p.CA * 3, ktiles);  // balanced: every k tile gathers the same number of units
  }
  const int tiles = ktiles * rcv_cdiv(p.CB, BN);
  int splits = 148 / tiles;  // one CTA per SM, never a second partial wave
  const int max_splits = rcv_cdiv(M, BK * 4);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int slab = rcv_cdiv(M, splits);
  slab = ((slab + BK - 1) / BK) * BK;
  splits = rcv_cdiv(M, slab);
  p.slab = slab;
  dim3 grid(ktiles, rcv_cdiv(p.CB, BN), splits);
  rcv_launch(umma_wgrad_kernel<BN, QUAD>, dim3(grid), dim3(C::NT), C::SMEM, st, p);
  RCV_CHECK_LAUNCH("umma_wgrad_kernel");
  return RCV_OK;
}

}  // namespace

bool rcv_umma_wgrad_pays(const RcvWgrad& p) {
  const int K = p.CA * p.taps.n;
  return ((p.Hg * p.Wg) & 3) == 0 && K >= 64 && p.CB >= 8 &&
         (int64_t)p.N * p.CA * p.Hin * p.Win < (1ll << 31);
}

// QUAD gather: stride-1 3x3 taps (-d,0,+d)^2 in row-major order, rows a multiple of 4 pixels wide
static bool quad_gather_ok(const RcvWgrad& p) {
  bool quad = p.taps.n == 9 && p.gs == 1 && p.Win == p.Wg && p.Hin == p.Hg && (p.Wg & 3) == 0;
  const int d = p.taps.dx[2];
  quad = quad && (d == 1 || d == 2);
  for (int t = 0; quad && t < 9; ++t)
    quad = p.taps.dy[t] == (t / 3 - 1) * d && p.taps.dx[t] == (t % 3 - 1) * d;
  static int no_quad = -1;  // RCV_WGRAD_QUAD=0: element-wise gather everywhere (A/B runs)
  if (no_quad < 0) {
    const char* e = getenv("RCV_WGRAD_QUAD");
    no_quad = (e && atoi(e) == 0) ? 1 : 0;
  }
  return quad && !no_quad;
}

bool rcv_umma_wgrad_takes_input_transform(const RcvWgrad& p) {
  const int64_t M = (int64_t)p.N * p.Hg * p.Wg;
  return M < (1ll << 31) && ((p.Hg * p.Wg) & 3) == 0 && quad_gather_ok(p);
}

int rcv_launch_wgrad_umma(RcvWgrad p, cudaStream_t st) {
  p.prof = g_rcv_prof;
  const int64_t M = (int64_t)p.N * p.Hg * p.Wg;
  RCV_REQUIRE(M < (1ll << 31), RCV_ERR_UNSUPPORTED, "wgrad: problem too large");
  RCV_REQUIRE(p.taps.n >= 1 && p.taps.n <= MAXT, RCV_ERR_UNSUPPORTED, "wgrad: %d taps", p.taps.n);
  RCV_REQUIRE(((p.Hg * p.Wg) & 3) == 0, RCV_ERR_UNSUPPORTED,
              "tensor-core wgrad needs a pixel count per image that is a multiple of 4 (got %d)", p.Hg * p.Wg);
  const bool quad = quad_gather_ok(p);
  RCV_REQUIRE(p.in_scale == nullptr || quad, RCV_ERR_UNSUPPORTED,
              "wgrad: normalise-on-load needs the quad gather (stride-1 3x3, rows a multiple of 4 pixels wide)");
  if (quad) {
    if (p.CB > 64) return launch_w<128, true>(p, st);
    if (p.CB > 32) return launch_w<64, true>(p, st);
    if (p.CB > 16) return launch_w<32, true>(p, st);
    return launch_w<16, true>(p, st);
  }
  if (p.CB > 64) return launch_w<128, false>(p, st);
  if (p.CB > 32) return launch_w<64, false>(p, st);
  if (p.CB > 16) return launch_w<32, false>(p, st);
  return launch_w<16, false>(p, st);
}
