// Input-side label ops of the hot path's callers (bandwidth-bound elementwise kernels):
//   rcv_label_lut      transform.py:26-49   maskLabel: class-drop relabel, in place, as a lookup table
//   rcv_label_to_pred  transform.py:172-183 labelToPred: int64 label map -> +-1 one-hot planes
//   rcv_lp_assemble    labelPropTrain.py:178-193: the two mirrored 8-channel LabelProp samples of a
//                      frame pair (Y_a, Y_b, Y_a - Y_b, +-1 one-hot of the OTHER frame's labels)
#include "rcv_common.cuh"

namespace {

constexpr int NT = 256;

__global__ void __launch_bounds__(NT) label_lut_kernel(int64_t n, long long* __restrict__ lab, int nlut,
                                                        const long long* __restrict__ lut) {
  rcv_pdl_enter();
  __shared__ long long s[64];
  if (threadIdx.x < nlut) s[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) {
    const long long v = lab[i];
    if (v >= 0 && v < nlut) lab[i] = s[v];
  }
}

// out[b, c, p] = (label[b, p] == c) ? +1 : -1
__global__ void __launch_bounds__(NT) label_to_pred_kernel(int64_t B, int C, int64_t HW,
                                                            const long long* __restrict__ lab,
                                                            float* __restrict__ out) {
  rcv_pdl_enter();
  const int64_t total = B * HW;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int64_t b = i / HW, p = i - b * HW;
    const long long v = lab[i];
    for (int c = 0; c < C; ++c) out[(b * C + c) * HW + p] = (v == c) ? 1.f : -1.f;
  }
}

__global__ void __launch_bounds__(NT) lp_assemble_kernel(int64_t P, int C, int64_t HW,
                                                          const float* __restrict__ ya,
                                                          const float* __restrict__ yb,
                                                          const long long* __restrict__ la,
                                                          const long long* __restrict__ lb,
                                                          float* __restrict__ inputs,
                                                          long long* __restrict__ targets) {
  rcv_pdl_enter();
  const int64_t total = P * HW;
  const int CH = 3 + C;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int64_t q = i / HW, p = i - q * HW;
    const float a = ya[i], b = yb[i];
    const long long va = la[i], vb = lb[i];
    float* s0 = inputs + (2 * q) * CH * HW + p;      // sample 2q   : (a, b, a-b, pred(lab_b)) -> lab_a
    float* s1 = inputs + (2 * q + 1) * CH * HW + p;  // sample 2q+1 : (b, a, b-a, pred(lab_a)) -> lab_b
    s0[0] = a; s0[HW] = b; s0[2 * HW] = a - b;
    s1[0] = b; s1[HW] = a; s1[2 * HW] = b - a;
    for (int c = 0; c < C; ++c) {
      s0[(3 + c) * HW] = (vb == c) ? 1.f : -1.f;
      s1[(3 + c) * HW] = (va == c) ? 1.f : -1.f;
    }
    targets[(2 * q) * HW + p] = va;
    targets[(2 * q + 1) * HW + p] = vb;
  }
}

int blocks_for(int64_t items) {
  int64_t b = (items + NT - 1) / NT;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int rcv_label_lut(int64_t n, int64_t* labels, int32_t nlut, const int64_t* lut, void* stream) {
  RCV_REQUIRE(n > 0 && labels && lut && nlut >= 1 && nlut <= 64, RCV_ERR_BAD_ARG, "label_lut: bad arg");
  rcv_launch(label_lut_kernel, dim3(blocks_for(n)), dim3(NT), 0, (cudaStream_t)stream, n,
             reinterpret_cast<long long*>(labels), nlut, reinterpret_cast<const long long*>(lut));
  RCV_CHECK_LAUNCH("label_lut");
  return RCV_OK;
}

extern "C" int rcv_label_to_pred(int64_t B, int32_t C, int64_t HW, const int64_t* labels, float* out, void* stream) {
  RCV_REQUIRE(B > 0 && C >= 1 && C <= 64 && HW > 0 && labels && out, RCV_ERR_BAD_ARG, "label_to_pred: bad arg");
  rcv_launch(label_to_pred_kernel, dim3(blocks_for(B * HW)), dim3(NT), 0, (cudaStream_t)stream, B, C, HW,
             reinterpret_cast<const long long*>(labels), out);
  RCV_CHECK_LAUNCH("label_to_pred");
  return RCV_OK;
}

extern "C" int rcv_lp_assemble(int64_t P, int32_t C, int64_t HW, const float* ya, const float* yb,
                               const int64_t* la, const int64_t* lb, float* inputs, int64_t* targets,
                               void* stream) {
  RCV_REQUIRE(P > 0 && C >= 1 && C <= 64 && HW > 0 && ya && yb && la && lb && inputs && targets, RCV_ERR_BAD_ARG,
              "lp_assemble: bad arg");
  rcv_launch(lp_assemble_kernel, dim3(blocks_for(P * HW)), dim3(NT), 0, (cudaStream_t)stream, P, C, HW, ya, yb,
             reinterpret_cast<const long long*>(la), reinterpret_cast<const long long*>(lb), inputs,
             reinterpret_cast<long long*>(targets));
  RCV_CHECK_LAUNCH("lp_assemble");
  return RCV_OK;
}
