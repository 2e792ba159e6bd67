// BatchNorm2d backward in ONE cooperative launch for tensors that fit in the register files of the grid.
//
// The two-pass form (rcv_bn.cu: bn_bwd_kernel<0> reduce, <1> apply) reads dy and z twice.  For the layers of the
// nets' lower half (<= 3 M elements at batch 64: 128 x 15x20, 64 x 15x20, 32 x 30x40) both tensors fit in the
// registers of one co-resident grid: every CTA loads its slice of one channel once (up to Q float4 of dy and of z
// per thread), adds its partial sums (sum g, sum g*xhat) to the fp64 accumulators, the grid synchronises
// (cooperative launch: all CTAs are resident), and the apply pass runs from the registers.  Traffic: 2 reads + 1
// write instead of 4 reads + 1 write, one launch instead of two.
//
// EXPERIMENTAL: parity-tested only behind RCV_TEST_EXPERIMENTAL=1; ops.bn_bwd uses it only
// when RCV_B200_BN_BWD_FUSED=1.
#include <cooperative_groups.h>

#include "rcv_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NT = 256;
constexpr int Q = 10;  // float4 of each tensor cached per thread: 80 registers

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = l < (NT / 32) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in thread 0
}

// grid = (C, parts): one channel per blockIdx.x, blockIdx.y splits its N*HW elements into slices of at most
// NT*4*Q elements (HW % 4 == 0: a float4 never straddles two images).
__global__ void __launch_bounds__(NT, 2)
    bn_bwd_fused_kernel(int N, int C, int64_t HW, int order, const float* __restrict__ dy, const float* __restrict__ z,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ save_mean, const float* __restrict__ save_invstd, double* sums,
                        float* __restrict__ dconv, float* dgamma, float* dbeta, float* dbias) {
  rcv_pdl_enter();
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[NT / 32];
  const int c = blockIdx.x;
  const float sc = scale[c], sft = shift[c], mean = save_mean[c], invstd = save_invstd[c];
  const int64_t E = (int64_t)N * HW;
  const int64_t per = (((E + gridDim.y - 1) / gridDim.y) + 3) & ~(int64_t)3;
  const int64_t beg = (int64_t)blockIdx.y * per;
  int64_t end = beg + per;
  if (end > E) end = E;

  float4 g[Q], zz[Q];
  float fs1 = 0.f, fs2 = 0.f;
  // element e of the channel -> offset in the NCHW tensor (32-bit: E and the tensor are < 2^31 elements, checked
  // by the launcher)
  const unsigned hw32 = (unsigned)HW, chw = (unsigned)C * hw32, cbase = (unsigned)c * hw32;
  auto offset_of = [&](unsigned e) {
    const unsigned n = e / hw32;
    return n * chw + cbase + (e - n * hw32);
  };
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int64_t e = beg + ((int64_t)q * NT + threadIdx.x) * 4;
    g[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    zz[q] = make_float4(mean, mean, mean, mean);
    if (e < end) {
      const unsigned off = offset_of((unsigned)e);
      g[q] = __ldg(reinterpret_cast<const float4*>(dy + off));
      zz[q] = __ldg(reinterpret_cast<const float4*>(z + off));
    }
  }
  // pass 0 from the registers: the masked gradient replaces g, xhat replaces nothing (recomputed in pass 1)
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    float* gp = reinterpret_cast<float*>(&g[q]);
    const float* zp = reinterpret_cast<const float*>(&zz[q]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float gv = gp[i];
      if (order == RCV_EPI_AFFINE_RELU) gv = (fmaf(sc, zp[i], sft) > 0.f) ? gv : 0.f;
      gp[i] = gv;
      const float xh = (zp[i] - mean) * invstd;
      fs1 += gv;
      fs2 += gv * xh;
    }
  }
  {
    const double t1 = block_sum_d((double)fs1, sh);
    const double t2 = block_sum_d((double)fs2, sh);
    if (threadIdx.x == 0) {
      atomicAdd(sums + c, t1);
      atomicAdd(sums + C + c, t2);
    }
  }
  __threadfence();
  grid.sync();
  const double cnt = (double)E;
  const double S1 = __ldcg(sums + c), S2 = __ldcg(sums + C + c);
  const float m1 = (float)(S1 / cnt), m2 = (float)(S2 / cnt);
  float fd = 0.f;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int64_t e = beg + ((int64_t)q * NT + threadIdx.x) * 4;
    if (e < end) {
      const float* gp = reinterpret_cast<const float*>(&g[q]);
      const float* zp = reinterpret_cast<const float*>(&zz[q]);
      float4 o;
      float* op = reinterpret_cast<float*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xh = (zp[i] - mean) * invstd;
        float d = sc * (gp[i] - m1 - xh * m2);
        if (order == RCV_EPI_RELU_AFFINE) d = zp[i] > 0.f ? d : 0.f;
        op[i] = d;
        fd += d;
      }
      *reinterpret_cast<float4*>(dconv + offset_of((unsigned)e)) = o;
    }
  }
  const double td = block_sum_d((double)fd, sh);
  if (threadIdx.x == 0) {
    if (dbias) atomicAdd(dbias + c, (float)td);
    if (blockIdx.y == 0) {
      if (dgamma) atomicAdd(dgamma + c, (float)S2);
      if (dbeta) atomicAdd(dbeta + c, (float)S1);
    }
  }
}

// Slices per channel so that every slice fits the per-thread cache and the grid is co-resident; 0: does not fit.
int fused_parts(int C, int64_t E) {
  static int cap = -1;  // co-resident CTAs of the kernel on this device
  if (cap < 0) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_fused_kernel, NT, 0) != cudaSuccess) {
      cudaGetLastError();
      return 0;  // (no device: the host-side query answers "no")
    }
    cap = sms * per_sm;
  }
  const int64_t slice = (int64_t)NT * 4 * Q;
  const int64_t need = (E + slice - 1) / slice;
  if (C <= 0 || need * C > cap || need > 65535) return 0;
  int64_t parts = cap / C;  // as many slices as stay co-resident: shorter per-thread loops
  const int64_t most = (E + NT * 4 - 1) / (NT * 4);  // at least one float4 per thread
  if (parts > most) parts = most;
  if (parts < need) parts = need;
  if (parts > 65535) parts = 65535;
  return (int)parts;
}

}  // namespace

extern "C" int rcv_bn_bwd_fused_supported(int32_t N, int32_t C, int64_t HW) {
  if (N <= 0 || C <= 0 || HW <= 0 || (HW & 3) != 0) return 0;
  return fused_parts(C, (int64_t)N * HW) > 0 ? 1 : 0;
}

extern "C" int rcv_bn_bwd_fused(int32_t N, int32_t C, int64_t HW, int order, const float* dy, const float* z,
                                const float* scale, const float* shift, const float* save_mean,
                                const float* save_invstd, double* sums, float* dconv, float* dgamma, float* dbeta,
                                float* dbias, void* stream) {
  RCV_REQUIRE(N > 0 && C > 0 && HW > 0 && dy && z && scale && shift && save_mean && save_invstd && sums && dconv,
              RCV_ERR_BAD_ARG, "bn_bwd_fused: bad arg");
  RCV_REQUIRE(order == RCV_EPI_RELU_AFFINE || order == RCV_EPI_AFFINE_RELU || order == RCV_EPI_AFFINE,
              RCV_ERR_BAD_ARG, "bn_bwd_fused: bad order %d", order);
  RCV_REQUIRE((HW & 3) == 0 && (((uintptr_t)dy | (uintptr_t)z | (uintptr_t)dconv) & 15) == 0, RCV_ERR_UNSUPPORTED,
              "bn_bwd_fused: needs HW %% 4 == 0 and 16-byte aligned tensors");
  RCV_REQUIRE((int64_t)N * C * HW < (1ll << 31), RCV_ERR_UNSUPPORTED, "bn_bwd_fused: tensor too large");
  const int parts = fused_parts(C, (int64_t)N * HW);
  RCV_REQUIRE(parts > 0, RCV_ERR_UNSUPPORTED,
              "bn_bwd_fused: the tensor does not fit the register files of one co-resident grid "
              "(query rcv_bn_bwd_fused_supported; use rcv_bn_bwd_reduce + rcv_bn_bwd_apply)");
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(C, parts);
  cfg.blockDim = dim3(NT);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr;
  memset(&attr, 0, sizeof(attr));
  attr.id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid-wide barrier cannot deadlock
  attr.val.cooperative = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int64_t hw = HW;
  int n = N, cc = C;
  (void)cudaLaunchKernelEx(&cfg, bn_bwd_fused_kernel, n, cc, hw, order, dy, z, scale, shift, save_mean, save_invstd,
                           sums, dconv, dgamma, dbeta, dbias);
  RCV_CHECK_LAUNCH("bn_bwd_fused");
  return RCV_OK;
}
