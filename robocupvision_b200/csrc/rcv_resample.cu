// Decoder-side resampling extras north_star names beside the transposed convolution: MaxUnpool2d(2,2) driven by the
// indices (or 2-bit window codes) of rcv_maxpool2x2_fwd -- the other half of the pool-index / unpool pair -- and 2x
// bilinear upsampling (align_corners=False), each with its adjoint.  The reference itself has neither
// (SURVEY.md section 0): they are pinned against F.max_unpool2d / F.interpolate in tests/test_gpu_ops.py.
// All four are streaming kernels: HBM-bound, algorithmic bytes = 4 B x (elements read + written) (+ the index bytes).
#include "rcv_common.cuh"

namespace {

constexpr int NT = 256;

int blocks_for(int64_t items, int cap) {
  int64_t b = (items + NT - 1) / NT;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// window position 0..3 of a pooled element: from the uint8 code, else from the int64 plane index
__device__ __forceinline__ int window_pos(const uint8_t* code, const int64_t* idx, int64_t o, int oy, int ox, int W) {
  if (code) return code[o];
  const int64_t f = __ldg(idx + o);
  const int r = (int)(f / W) - 2 * oy, c = (int)(f % W) - 2 * ox;
  return ((r & 1) << 1) | (c & 1);
}

// out[N,C,H,W] = MaxUnpool2d(2,2)(y, idx) (+ skip): every output element is written (zeros off the maxima)
__global__ void __launch_bounds__(NT) unpool_fwd_kernel(int64_t total, int H, int W, const float* __restrict__ y,
                                                         const int64_t* __restrict__ idx,
                                                         const uint8_t* __restrict__ code,
                                                         const float* __restrict__ skip, float* __restrict__ out) {
  rcv_pdl_enter();
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t stride = (int64_t)gridDim.x * NT;
  for (int64_t o = (int64_t)blockIdx.x * NT + threadIdx.x; o < total; o += stride) {
    const int ox = (int)(o % Wo);
    const int64_t t = o / Wo;
    const int oy = (int)(t % Ho);
    const int64_t plane = t / Ho;
    const float g = __ldg(y + o);
    const int bc = window_pos(code, idx, o, oy, ox, W);
    const int64_t off = plane * (int64_t)H * W + (int64_t)(2 * oy) * W + 2 * ox;
    float2 r0 = make_float2(bc == 0 ? g : 0.f, bc == 1 ? g : 0.f);
    float2 r1 = make_float2(bc == 2 ? g : 0.f, bc == 3 ? g : 0.f);
    if (skip) {
      const float2 s0 = __ldg(reinterpret_cast<const float2*>(skip + off));
      const float2 s1 = __ldg(reinterpret_cast<const float2*>(skip + off + W));
      r0.x += s0.x; r0.y += s0.y; r1.x += s1.x; r1.y += s1.y;
    }
    *reinterpret_cast<float2*>(out + off) = r0;
    *reinterpret_cast<float2*>(out + off + W) = r1;
  }
}

// dy[N,C,H/2,W/2] = dout gathered at the stored positions
__global__ void __launch_bounds__(NT) unpool_bwd_kernel(int64_t total, int H, int W, const float* __restrict__ dout,
                                                         const int64_t* __restrict__ idx,
                                                         const uint8_t* __restrict__ code, float* __restrict__ dy) {
  rcv_pdl_enter();
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t stride = (int64_t)gridDim.x * NT;
  for (int64_t o = (int64_t)blockIdx.x * NT + threadIdx.x; o < total; o += stride) {
    const int ox = (int)(o % Wo);
    const int64_t t = o / Wo;
    const int oy = (int)(t % Ho);
    const int64_t plane = t / Ho;
    const int bc = window_pos(code, idx, o, oy, ox, W);
    const int64_t off = plane * (int64_t)H * W + (int64_t)(2 * oy + (bc >> 1)) * W + 2 * ox + (bc & 1);
    dy[o] = __ldg(dout + off);
  }
}

// ---- 2x bilinear, align_corners = False ---------------------------------------------------------------------
// Output o reads source s = (o + 0.5) / 2 - 0.5 clamped at 0: even o = 2k (k >= 1): 0.25 x[k-1] + 0.75 x[k]; o = 0:
// x[0]; odd o = 2k+1: 0.75 x[k] + 0.25 x[min(k+1, n-1)].  Interpolation along W first, then along H (ATen's order).
__device__ __forceinline__ void taps(int o, int n, int& i0, int& i1, float& w0, float& w1) {
  if (o & 1) {
    i0 = o >> 1; i1 = min(i0 + 1, n - 1); w0 = 0.75f; w1 = 0.25f;
  } else if (o == 0) {
    i0 = 0; i1 = min(1, n - 1); w0 = 1.f; w1 = 0.f;
  } else {
    i0 = (o >> 1) - 1; i1 = o >> 1; w0 = 0.25f; w1 = 0.75f;
  }
}

// one thread = two adjacent output columns (2k, 2k+1) of one output row: reads x[k-1], x[k], x[k+1] of two rows
__global__ void __launch_bounds__(NT) bilinear2x_fwd_kernel(int64_t total, int H, int W, const float* __restrict__ x,
                                                             const float* __restrict__ skip, float* __restrict__ out) {
  rcv_pdl_enter();
  const int Ho = 2 * H;
  const int64_t stride = (int64_t)gridDim.x * NT;
  for (int64_t q = (int64_t)blockIdx.x * NT + threadIdx.x; q < total; q += stride) {
    const int k = (int)(q % W);
    const int64_t t = q / W;
    const int oy = (int)(t % Ho);
    const int64_t plane = t / Ho;
    int y0, y1;
    float h0, h1;
    taps(oy, H, y0, y1, h0, h1);
    const float* r0 = x + plane * (int64_t)H * W + (int64_t)y0 * W;
    const float* r1 = x + plane * (int64_t)H * W + (int64_t)y1 * W;
    const int km = max(k - 1, 0), kp = min(k + 1, W - 1);
    const float a0 = __ldg(r0 + km), b0 = __ldg(r0 + k), c0 = __ldg(r0 + kp);
    const float a1 = __ldg(r1 + km), b1 = __ldg(r1 + k), c1 = __ldg(r1 + kp);
    // even column 2k: (k == 0) ? x[0] : 0.25 x[k-1] + 0.75 x[k]; odd column 2k+1: 0.75 x[k] + 0.25 x[min(k+1, W-1)]
    const float we0 = k == 0 ? 0.f : 0.25f, we1 = k == 0 ? 1.f : 0.75f;
    const float e0 = we0 * a0 + we1 * b0, e1 = we0 * a1 + we1 * b1;
    const float o0 = 0.75f * b0 + 0.25f * c0, o1 = 0.75f * b1 + 0.25f * c1;
    float2 v = make_float2(h0 * e0 + h1 * e1, h0 * o0 + h1 * o1);
    const int64_t off = (plane * Ho + oy) * (int64_t)(2 * W) + 2 * k;
    if (skip) {
      const float2 s = __ldg(reinterpret_cast<const float2*>(skip + off));
      v.x += s.x; v.y += s.y;
    }
    *reinterpret_cast<float2*>(out + off) = v;
  }
}

// adjoint weights of input index i along one axis of length n over outputs 2i-1 .. 2i+2
__device__ __forceinline__ void adj(int i, int n, float (&w)[4]) {
  w[0] = i >= 1 ? 0.25f : 0.f;                      // output 2i-1 (odd, k = i-1, second tap)
  w[1] = i == 0 ? 1.f : 0.75f;                      // output 2i
  w[2] = i == n - 1 ? 1.f : 0.75f;                  // output 2i+1 (its second tap clamps onto i at the border)
  w[3] = i + 1 <= n - 1 ? 0.25f : 0.f;              // output 2i+2 (even, k = i+1, first tap)
}

// dx[N,C,H,W] = adjoint of the forward: one thread per input element gathers its <= 4x4 outputs
__global__ void __launch_bounds__(NT) bilinear2x_bwd_kernel(int64_t total, int H, int W, const float* __restrict__ dout,
                                                             float* __restrict__ dx) {
  rcv_pdl_enter();
  const int Wo = 2 * W;
  const int64_t stride = (int64_t)gridDim.x * NT;
  for (int64_t e = (int64_t)blockIdx.x * NT + threadIdx.x; e < total; e += stride) {
    const int ix = (int)(e % W);
    const int64_t t = e / W;
    const int iy = (int)(t % H);
    const int64_t plane = t / H;
    float wy[4], wx[4];
    adj(iy, H, wy);
    adj(ix, W, wx);
    const float* base = dout + plane * (int64_t)(2 * H) * Wo;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wy[a] == 0.f) continue;
      const float* row = base + (int64_t)(2 * iy - 1 + a) * Wo + 2 * ix - 1;
      float r = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (wx[b] != 0.f) r += wx[b] * __ldg(row + b);
      acc += wy[a] * r;
    }
    dx[e] = acc;
  }
}

// dst[n, dst_off + c, :] = src[n, src_off + c, :], c < count: the channel slice / concat copy (VEC = 4: float4 rows)
template <int VEC>
__global__ void __launch_bounds__(NT) channel_copy_kernel(int64_t total, int64_t row, int count, const float* __restrict__ src,
                                                           int64_t src_batch, float* __restrict__ dst, int64_t dst_batch) {
  rcv_pdl_enter();
  const int64_t per = (int64_t)count * row;  // elements (of VEC floats) one image moves
  const int64_t stride = (int64_t)gridDim.x * NT;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += stride) {
    const int64_t n = i / per, r = i - n * per;
    if (VEC == 4)
      reinterpret_cast<float4*>(dst + n * dst_batch)[r] = __ldg(reinterpret_cast<const float4*>(src + n * src_batch) + r);
    else
      dst[n * dst_batch + r] = __ldg(src + n * src_batch + r);
  }
}
}  // namespace

extern "C" int rcv_maxunpool2x2_fwd(int32_t N, int32_t C, int32_t H, int32_t W, const float* y, const int64_t* idx,
                                    const uint8_t* code, const float* skip, float* out, void* stream) {
  RCV_REQUIRE(N > 0 && C > 0 && H > 1 && W > 1 && y && out && (idx || code), RCV_ERR_BAD_ARG, "maxunpool_fwd: bad arg");
  RCV_REQUIRE((H & 1) == 0 && (W & 1) == 0, RCV_ERR_UNSUPPORTED, "maxunpool_fwd: H and W must be even (got %dx%d)", H, W);
  const int64_t total = (int64_t)N * C * (H / 2) * (W / 2);
  rcv_launch(unpool_fwd_kernel, dim3(blocks_for(total, 148 * 16)), dim3(NT), 0, (cudaStream_t)stream, total, H, W, y,
             idx, code, skip, out);
  RCV_CHECK_LAUNCH("maxunpool_fwd");
  return RCV_OK;
}

extern "C" int rcv_maxunpool2x2_bwd(int32_t N, int32_t C, int32_t H, int32_t W, const float* dout, const int64_t* idx,
                                    const uint8_t* code, float* dy, void* stream) {
  RCV_REQUIRE(N > 0 && C > 0 && H > 1 && W > 1 && dout && dy && (idx || code), RCV_ERR_BAD_ARG, "maxunpool_bwd: bad arg");
  RCV_REQUIRE((H & 1) == 0 && (W & 1) == 0, RCV_ERR_UNSUPPORTED, "maxunpool_bwd: H and W must be even");
  const int64_t total = (int64_t)N * C * (H / 2) * (W / 2);
  rcv_launch(unpool_bwd_kernel, dim3(blocks_for(total, 148 * 16)), dim3(NT), 0, (cudaStream_t)stream, total, H, W, dout,
             idx, code, dy);
  RCV_CHECK_LAUNCH("maxunpool_bwd");
  return RCV_OK;
}

extern "C" int rcv_upsample_bilinear2x_fwd(int32_t N, int32_t C, int32_t H, int32_t W, const float* x,
                                           const float* skip, float* out, void* stream) {
  RCV_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && x && out, RCV_ERR_BAD_ARG, "upsample_bilinear2x_fwd: bad arg");
  const int64_t total = (int64_t)N * C * (2 * H) * W;
  rcv_launch(bilinear2x_fwd_kernel, dim3(blocks_for(total, 148 * 16)), dim3(NT), 0, (cudaStream_t)stream, total, H, W, x,
             skip, out);
  RCV_CHECK_LAUNCH("upsample_bilinear2x_fwd");
  return RCV_OK;
}

extern "C" int rcv_upsample_bilinear2x_bwd(int32_t N, int32_t C, int32_t H, int32_t W, const float* dout, float* dx,
                                           void* stream) {
  RCV_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && dout && dx, RCV_ERR_BAD_ARG, "upsample_bilinear2x_bwd: bad arg");
  const int64_t total = (int64_t)N * C * H * W;
  rcv_launch(bilinear2x_bwd_kernel, dim3(blocks_for(total, 148 * 16)), dim3(NT), 0, (cudaStream_t)stream, total, H, W,
             dout, dx);
  RCV_CHECK_LAUNCH("upsample_bilinear2x_bwd");
  return RCV_OK;
}

extern "C" int rcv_channel_copy(int64_t N, int64_t HW, int32_t count, const float* src, int32_t src_channels,
                                int32_t src_offset, float* dst, int32_t dst_channels, int32_t dst_offset, void* stream) {
  RCV_REQUIRE(N > 0 && HW > 0 && count > 0 && src && dst, RCV_ERR_BAD_ARG, "channel_copy: bad arg");
  RCV_REQUIRE(src_offset >= 0 && dst_offset >= 0 && src_offset + count <= src_channels &&
                  dst_offset + count <= dst_channels,
              RCV_ERR_BAD_ARG, "channel_copy: channels [%d,+%d) of %d -> [%d,+%d) of %d", src_offset, count,
              src_channels, dst_offset, count, dst_channels);
  const float* s0 = src + (int64_t)src_offset * HW;
  float* d0 = dst + (int64_t)dst_offset * HW;
  const int64_t sb = (int64_t)src_channels * HW, db = (int64_t)dst_channels * HW;
  const bool v4 = (HW % 4 == 0) && (((uintptr_t)s0 | (uintptr_t)d0) % 16 == 0);
  if (v4) {
    const int64_t total = N * count * (HW / 4);
    rcv_launch(channel_copy_kernel<4>, dim3(blocks_for(total, 148 * 16)), dim3(NT), 0, (cudaStream_t)stream, total,
               HW / 4, (int)count, s0, sb, d0, db);
  } else {
    const int64_t total = N * count * HW;
    rcv_launch(channel_copy_kernel<1>, dim3(blocks_for(total, 148 * 16)), dim3(NT), 0, (cudaStream_t)stream, total, HW,
               (int)count, s0, sb, d0, db);
  }
  RCV_CHECK_LAUNCH("channel_copy");
  return RCV_OK;
}
