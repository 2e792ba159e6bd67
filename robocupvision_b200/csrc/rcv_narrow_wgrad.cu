// Weight gradient of the narrow layers (convolution_backward's dW / db for model.py:105-116,
// 166-199, 403-414 with few channels on the dense side): the pixel reduction
//   dW[cb][ca][tap] = sum_{n,i,j} row[n,cb,i,j] * src[n,ca,i*gs+dy_tap, j*gs+dx_tap]
// has a tiny output (<= 32 x CA x 9) and a very long reduction (N*H*W pixels), which is a poor
// tensor-core shape; it is FMA work done where the pixels are:
//   * TMA stages, per tile of 2*TR grid rows of one image, the halo patch of CC `src` channels and
//     the matching tile of all CB `row` channels (two cp.async.bulk.tensor.4d on one mbarrier,
//     double-buffered across tiles by persistent CTAs).  Out-of-image elements are zero-filled
//     by the copy engine: no bounds tests anywhere in the arithmetic.
//   * a warp owns one role = (src channel, group of 8 row channels) and keeps its 9 x 8 partial
//     sums in registers for the whole launch; lanes are different 4-pixel x 2-row strips of the
//     tile, so every shared-memory read is a conflict-free LDS.128 (plus the halo scalars) and the
//     register window of 4..6 input rows is shared by the two output rows and all 8 channels.
//   * one shuffle reduction + fp32 RED per accumulator per warp at the very end.
// grid = (persistent tile walkers, channel chunks of CC = 8/G src channels).
#include "rcv_narrow.cuh"

namespace {
using namespace rcv_umma;
using namespace rcv_narrow;

constexpr int NTW = 256;  // 8 warps
constexpr int PIX = 4;

enum WgKind { WK_S1D1 = 0, WK_S1D2 = 1, WK_S2 = 2, WK_K1 = 3 };
template <int KIND> struct WT;
// GS grid stride, D tap spacing, KD taps per axis, MIN = offset of tap 0, WC window columns, NR window rows (2 output rows)
template <> struct WT<WK_S1D1> { static constexpr int GS = 1, D = 1, KD = 3, MIN = -1, WC = 6, NR = 4; };
template <> struct WT<WK_S1D2> { static constexpr int GS = 1, D = 2, KD = 3, MIN = -2, WC = 8, NR = 6; };
template <> struct WT<WK_S2> { static constexpr int GS = 2, D = 1, KD = 3, MIN = -1, WC = 9, NR = 5; };
template <> struct WT<WK_K1> { static constexpr int GS = 1, D = 1, KD = 1, MIN = 0, WC = 4, NR = 2; };

struct WgCfg {
  int32_t TR, SPR, R, pitch, CC, G, nroles, nparts, TWg, ctiles, tiles_per_img, total_tiles;
  int32_t rows;          // grid rows per tile = 2*TR
  int32_t NS;            // pipeline stages (2..4)
  uint32_t src_bytes, row_bytes, stage_bytes, row_off;
  uint64_t wmap;         // 4 bits per canonical tap (ky*KD+kx): index into the 3x3 / 1x1 kernel
};

template <int KIND, bool BIAS>
__global__ void __launch_bounds__(NTW, 2)
    narrow_wgrad_kernel(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ CUtensorMap map_row,
                        const RcvWgrad p, const WgCfg cfg) {
  rcv_pdl_enter();
  using T = WT<KIND>;
  constexpr int GS = T::GS, D = T::D, KD = T::KD, WC = T::WC, NR = T::NR;
  constexpr int NTAP = KD * KD;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[4];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar0 = smem_u32(&bars[0]);

  const int ca0 = blockIdx.y * cfg.CC;
  const int ntiles = cfg.total_tiles;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  auto issue = [&](int item) {  // thread 0 only
    const int tile = (int)blockIdx.x + item * (int)gridDim.x;
    const int n = tile / cfg.tiles_per_img;
    const int r2 = tile - n * cfg.tiles_per_img;
    const int rt = r2 / cfg.ctiles, ct = r2 - rt * cfg.ctiles;
    const int i0 = rt * cfg.rows, j0 = ct * cfg.TWg;
    const int sg = item % cfg.NS;
    const uint32_t bar = bar0 + 8 * sg;
    const uint32_t dst = sbase + sg * cfg.stage_bytes;
    mbar_expect_tx(bar, cfg.src_bytes + cfg.row_bytes);
    tma_load_4d(dst, &map_src, j0 * GS - 4, i0 * GS + T::MIN, ca0, n, bar);
    tma_load_4d(dst + cfg.row_off, &map_row, j0, i0, 0, n, bar);
  };
  if (tid == 0) {
    for (int i = 0; i < cfg.NS; ++i) mbar_init(bar0 + 8 * i, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
    for (int i = 0; i < cfg.NS && i < my_tiles; ++i) issue(i);
  }
  __syncthreads();

  // role of this warp
  const int role = wid % cfg.nroles, part = wid / cfg.nroles;
  const int cl = role / cfg.G, cbg = role - cl * cfg.G;  // local src channel, group of 8 row channels
  const int ca = ca0 + cl;
  const int cb0 = cbg * 8;
  const bool has_role = part < cfg.nparts && ca < p.CA && cb0 < p.CB;
  const int ncb = min(8, p.CB - cb0);
  const bool do_bias = BIAS && has_role && ca == 0;  // every role sums (straight-line code), one writes

  float acc[NTAP][8];
  float bsum[8];
#pragma unroll
  for (int t = 0; t < NTAP; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) bsum[c] = 0.f;

  const int pitch = cfg.pitch;
  const int plane = cfg.R * pitch;
  const int rplane = cfg.rows * cfg.TWg;
  const int nstrips = cfg.TR * cfg.SPR;
  const int per = (nstrips + cfg.nparts - 1) / cfg.nparts;
  const int sbeg = part * per, send = min(nstrips, sbeg + per);

  for (int item = 0; item < my_tiles; ++item) {
    const int st = item % cfg.NS;
    mbar_wait(bar0 + 8 * st, (item / cfg.NS) & 1);
    if (has_role) {
      const float* ssrc = reinterpret_cast<const float*>(sgen + (size_t)st * cfg.stage_bytes) + cl * plane;
      const float* srow = reinterpret_cast<const float*>(sgen + (size_t)st * cfg.stage_bytes + cfg.row_off) +
                          cb0 * rplane;
      for (int sidx = sbeg + lane; sidx < send; sidx += 32) {
        const int rp = sidx / cfg.SPR, s = sidx - rp * cfg.SPR;
        const float* xr = ssrc + (rp * 2 * GS) * pitch + 4 + s * PIX * GS;
        const float* dr = srow + (rp * 2) * cfg.TWg + s * PIX;
        // register window: W[r][c - MIN] = src[row r][x0*GS + c].  Dilated and stride-2 taps share (almost)
        // no window rows between the two output rows: those kinds walk the output rows one at a time with a
        // KD-row window (half the registers); the dense 3x3 keeps one 4-row window for both.
        constexpr bool SPLITQ = (KIND == WK_S1D2 || KIND == WK_S2);
        constexpr int WR = SPLITQ ? KD : NR;
        auto load_row = [&](float (&Wr)[WC], const float* x) {
          const float4 a = *reinterpret_cast<const float4*>(x);
          Wr[-T::MIN + 0] = a.x; Wr[-T::MIN + 1] = a.y; Wr[-T::MIN + 2] = a.z; Wr[-T::MIN + 3] = a.w;
          if constexpr (GS == 2) {
            const float4 b = *reinterpret_cast<const float4*>(x + 4);
            Wr[-T::MIN + 4] = b.x; Wr[-T::MIN + 5] = b.y; Wr[-T::MIN + 6] = b.z; Wr[-T::MIN + 7] = b.w;
          }
#pragma unroll
          for (int c = T::MIN; c < 0; ++c) Wr[c - T::MIN] = x[c];
#pragma unroll
          for (int c = PIX * GS; c < WC + T::MIN; ++c) Wr[c - T::MIN] = x[c];
        };
#pragma unroll
        for (int qo = 0; qo < (SPLITQ ? 2 : 1); ++qo) {
          float W[WR][WC];
#pragma unroll
          for (int r = 0; r < WR; ++r) load_row(W[r], xr + (SPLITQ ? qo * GS + r * D : r) * pitch);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            // no branch on the channel count: channels past the last one re-read it and accumulate into
            // registers that are never written out -- the unrolled body stays straight-line, so the
            // compiler can run the next channel's LDS under this channel's FMAs
            const float* dc = dr + min(c, ncb - 1) * rplane;
#pragma unroll
            for (int qi = 0; qi < (SPLITQ ? 1 : 2); ++qi) {
              const int q = SPLITQ ? qo : qi;
              const float4 d4 = *reinterpret_cast<const float4*>(dc + q * cfg.TWg);
              const float dv[PIX] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int ky = 0; ky < KD; ++ky)
#pragma unroll
                for (int kx = 0; kx < KD; ++kx)
#pragma unroll
                  for (int x = 0; x < PIX; ++x)
                    acc[ky * KD + kx][c] =
                        fmaf(W[SPLITQ ? ky : qi * GS + ky * D][x * GS + kx * D], dv[x], acc[ky * KD + kx][c]);
              if constexpr (BIAS) bsum[c] += (dv[0] + dv[1]) + (dv[2] + dv[3]);
            }
          }
        }
      }
    }
    __syncthreads();  // every reader is done with this stage
    if (tid == 0 && item + cfg.NS < my_tiles) issue(item + cfg.NS);
  }

  // ---------------- reduction ----------------
  // lanes -> lane 0 (shuffles) -> the CTA's dw-shaped image in shared memory (stage 0 is free now) -> one
  // pass of global REDs by the whole CTA, 16 bytes at a time where the address allows: the gradient of a
  // narrow layer is a handful of cache lines that every CTA hits, and same-line atomics serialise in L2.
  float* sdw = reinterpret_cast<float*>(sgen);          // [CB][run], run = CC * wsA
  const int run = cfg.CC * p.wsA;
  float* sdb = sdw + p.CB * run;                        // [CB]
  for (int e = tid; e < p.CB * run + p.CB; e += NTW) sdw[e] = 0.f;
  __syncthreads();
  if (has_role) {
#pragma unroll
    for (int t = 0; t < NTAP; ++t) {
      const int wi = (int)((cfg.wmap >> (4 * t)) & 15u);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v = acc[t][c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && c < ncb && wi < 9) atomicAdd(&sdw[(cb0 + c) * run + cl * p.wsA + wi], v);
      }
    }
    if (do_bias) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v = bsum[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && c < ncb) atomicAdd(&sdb[cb0 + c], v);
      }
    }
  }
  __syncthreads();
  {
    const int nca = min(cfg.CC, p.CA - ca0);
    const int len = nca * p.wsA;                         // floats of each row-channel's run that are live
    // quads: element e of row cb sits at global index cb*wsB + ca0*wsA + e; split each run at 16-byte
    // boundaries of the global address
    for (int cb = wid; cb < p.CB; cb += NTW / 32) {
      const size_t g0 = (size_t)cb * p.wsB + (size_t)ca0 * p.wsA;
      const float* sr = sdw + cb * run;
      float* gr = p.dw + g0;
      const int head = (int)((4 - (((uintptr_t)gr >> 2) & 3)) & 3);  // scalars before the first 16-byte aligned quad
      const int nq = len > head ? (len - head) / 4 : 0;
      for (int e = lane; e < head && e < len; e += 32) atomicAdd(gr + e, sr[e]);
      for (int qd = lane; qd < nq; qd += 32) {
        const int e = head + 4 * qd;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gr + e), "f"(sr[e]), "f"(sr[e + 1]),
                     "f"(sr[e + 2]), "f"(sr[e + 3])
                     : "memory");
      }
      for (int e = head + 4 * nq + lane; e < len; e += 32) atomicAdd(gr + e, sr[e]);
    }
    if (BIAS && ca0 == 0 && tid < p.CB) atomicAdd(p.dbias + tid, sdb[tid]);
  }
}

// Tap set -> kind + canonical map; -1 if the structure is none of the compiled kinds.
template <int KIND>
bool match(const RcvWgrad& p, uint64_t* wmap) {
  using T = WT<KIND>;
  if (p.gs != T::GS || p.taps.n != T::KD * T::KD) return false;
  int8_t m[9];
  for (int j = 0; j < 9; ++j) m[j] = -1;
  for (int t = 0; t < p.taps.n; ++t) {
    const int oy = p.taps.dy[t] - T::MIN, ox = p.taps.dx[t] - T::MIN;
    if (oy < 0 || ox < 0 || (oy % T::D) || (ox % T::D)) return false;
    const int ky = oy / T::D, kx = ox / T::D;
    if (ky >= T::KD || kx >= T::KD) return false;
    const int j = ky * T::KD + kx;
    if (m[j] >= 0 || p.taps.wi[t] < 0 || p.taps.wi[t] > 8) return false;
    m[j] = p.taps.wi[t];
  }
  uint64_t w = 0;
  for (int j = 0; j < 9; ++j) w |= (uint64_t)(m[j] < 0 ? 15 : m[j]) << (4 * j);
  *wmap = w;
  return true;
}

bool plan(const RcvWgrad& p, WgCfg* out, int* kind_out, size_t* smem) {
  WgCfg c;
  memset(&c, 0, sizeof(c));
  static const int maxcb = env_int("RCV_NARROW_WGRAD_MAXCB", 16);
  if (p.CB < 1 || p.CB > maxcb || p.CB > 32) return false;
  if ((p.Wg & 3) || (p.Win & 3)) return false;
  if (((uintptr_t)p.src | (uintptr_t)p.row) & 15) return false;
  int kind = -1;
  if (match<WK_K1>(p, &c.wmap)) kind = WK_K1;
  else if (match<WK_S1D1>(p, &c.wmap)) kind = WK_S1D1;
  else if (match<WK_S1D2>(p, &c.wmap)) kind = WK_S1D2;
  else if (match<WK_S2>(p, &c.wmap)) kind = WK_S2;
  if (kind < 0) return false;
  const int NR = kind == WK_S1D1 ? 4 : kind == WK_S1D2 ? 6 : kind == WK_S2 ? 5 : 2;
  c.G = rcv_cdiv(p.CB, 8);
  c.CC = 8 / c.G;
  if (c.CC > p.CA) c.CC = p.CA;
  c.nroles = c.CC * c.G;
  c.nparts = 8 / c.nroles;
  const int maxcols = (256 - 8) / p.gs;
  c.ctiles = rcv_cdiv(p.Wg, maxcols);
  c.TWg = ((rcv_cdiv(p.Wg, c.ctiles) + 3) / 4) * 4;
  c.SPR = c.TWg / PIX;
  c.pitch = c.TWg * p.gs + 8;
  if (c.pitch > 256) return false;
  // rows per tile: the candidate (<= 8 row pairs, two stages within the shared-memory budget) that wastes the
  // fewest lanes: strips per role part against whole warps, image rows against whole tiles
  const size_t budget = (size_t)env_int("RCV_NARROW_WGRAD_SMEM_KB", 100) * 1024;
  int NS = env_int("RCV_NARROW_WGRAD_STAGES", 2);
  NS = NS < 2 ? 2 : NS > 4 ? 4 : NS;
  c.NS = NS;
  const int trmax = rcv_cdiv(p.Hg, 2) < 8 ? rcv_cdiv(p.Hg, 2) : 8;
  double best = -1.0;
  WgCfg bc = c;
  for (int TR = 1; TR <= trmax; ++TR) {
    WgCfg t = c;
    t.TR = TR;
    t.rows = 2 * TR;
    t.R = (TR - 1) * 2 * p.gs + NR;
    t.src_bytes = (uint32_t)((size_t)t.CC * t.R * t.pitch * 4);
    t.row_bytes = (uint32_t)((size_t)p.CB * t.rows * t.TWg * 4);
    t.row_off = (t.src_bytes + 127u) & ~127u;
    t.stage_bytes = (t.row_off + t.row_bytes + 127u) & ~127u;
    if (NS * (size_t)t.stage_bytes > budget && TR > 1) break;
    const int per = rcv_cdiv(TR * t.SPR, t.nparts);
    const double e1 = (double)per / (32.0 * rcv_cdiv(per, 32));
    const double e2 = (double)p.Hg / ((double)rcv_cdiv(p.Hg, t.rows) * t.rows);
    const double halo = (double)(t.rows * p.gs) / t.R;  // staged rows that are not halo
    const double e = e1 * e2 * (0.75 + 0.25 * halo);
    if (e >= best) { best = e; bc = t; }
  }
  c = bc;
  if (c.R > 256 || c.rows > 256 || NS * (size_t)c.stage_bytes + 256 > 200 * 1024) return false;
  if ((size_t)(p.CB * c.CC * p.wsA + p.CB) * 4 > NS * (size_t)c.stage_bytes) return false;  // reduction image reuses the stages
  c.tiles_per_img = rcv_cdiv(p.Hg, c.rows) * c.ctiles;
  const int64_t tt = (int64_t)c.tiles_per_img * p.N;
  if (tt >= (1ll << 31)) return false;
  c.total_tiles = (int)tt;
  *out = c;
  *kind_out = kind;
  *smem = NS * (size_t)c.stage_bytes + 256;
  return true;
}

template <int KIND, bool BIAS>
int launch(const RcvWgrad& p, const WgCfg& cfg, size_t smem, cudaStream_t st) {
  CUtensorMap ms, mr;
  int rc = make_nchw_map(&ms, p.src, p.N, p.CA, p.Hin, p.Win, cfg.pitch, cfg.R, cfg.CC, "narrow_wgrad");
  if (rc) return rc;
  rc = make_nchw_map(&mr, p.row, p.N, p.CB, p.Hg, p.Wg, cfg.TWg, cfg.rows, p.CB, "narrow_wgrad");
  if (rc) return rc;
  static int num_sms = 0;
  static size_t last_smem = 0;
  static int occ = 1;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(narrow_wgrad_kernel<KIND, BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(narrow_wgrad_kernel<KIND, BIAS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
  }
  if (last_smem != smem) {
    int o = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, narrow_wgrad_kernel<KIND, BIAS>, NTW, smem);
    occ = o < 1 ? 1 : o;
    last_smem = smem;
  }
  const int nchunks = rcv_cdiv(p.CA, cfg.CC);
  int walkers = (num_sms * occ) / nchunks;
  walkers = walkers < 1 ? 1 : walkers;
  walkers = walkers > cfg.total_tiles ? cfg.total_tiles : walkers;
  dim3 grid(walkers, nchunks, 1);
  rcv_launch(narrow_wgrad_kernel<KIND, BIAS>, dim3(grid), dim3(NTW), smem, st, ms, mr, p, cfg);
  RCV_CHECK_LAUNCH("narrow_wgrad_kernel");
  return RCV_OK;
}

}  // namespace

bool rcv_narrow_wgrad_supported(const RcvWgrad& p) {
  WgCfg c;
  int kind;
  size_t sm;
  return plan(p, &c, &kind, &sm);
}

int rcv_launch_narrow_wgrad(const RcvWgrad& p, cudaStream_t st) {
  WgCfg c;
  int kind;
  size_t sm;
  RCV_REQUIRE(plan(p, &c, &kind, &sm), RCV_ERR_UNSUPPORTED, "narrow_wgrad: geometry outside the kernel's limits");
  const bool bias = p.dbias != nullptr;
  switch (kind) {
    case WK_S1D1: return bias ? launch<WK_S1D1, true>(p, c, sm, st) : launch<WK_S1D1, false>(p, c, sm, st);
    case WK_S1D2: return bias ? launch<WK_S1D2, true>(p, c, sm, st) : launch<WK_S1D2, false>(p, c, sm, st);
    case WK_S2: return bias ? launch<WK_S2, true>(p, c, sm, st) : launch<WK_S2, false>(p, c, sm, st);
    default: return bias ? launch<WK_K1, true>(p, c, sm, st) : launch<WK_K1, false>(p, c, sm, st);
  }
}
