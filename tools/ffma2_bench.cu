// Microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) throughput on one B200, with the operand patterns a
// direct convolution has (accumulator += x * w, all three in distinct registers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_bench tools/ffma2_bench.cu && /tmp/ffma2_bench
#include <cuda_runtime.h>
#include <stdio.h>
// MODE 0: FFMA  acc = acc*c0 + c1 (constant operands)        1: FFMA2 same shape
// MODE 2: FFMA  acc[i] += x[j]*w[k] (three distinct registers) 3: FFMA2 acc2[i] += x[j] (scalar) * w2[k]
// MODE 4: FFMA2 acc2[i] += x2[j] * w2[k] (three distinct pairs)
template <int MODE>
__global__ void k(float* out, int iters, float w0, float w1) {
  float a[64], x[4], w[8];
#pragma unroll
  for (int i = 0; i < 64; ++i) a[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = w0 + i * 1e-6f + threadIdx.x * 1e-9f;
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = w1 + i * 1e-6f + threadIdx.x * 1e-9f;
  unsigned long long p[32], w2[4], x2[2];
#pragma unroll
  for (int i = 0; i < 32; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
#pragma unroll
  for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(w2[i]) : "f"(w[2 * i]), "f"(w[2 * i + 1]));
#pragma unroll
  for (int i = 0; i < 2; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(x2[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
  unsigned long long wc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(wc) : "f"(w0), "f"(w1));
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 64; ++i) a[i] = fmaf(a[i], w0, w1);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(wc));
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x[i & 3]), "f"(w[(i >> 2) & 7]));
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        unsigned long long xx;
        asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x[i & 3]));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(xx), "l"(w2[(i >> 2) & 3]));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(x2[i & 1]), "l"(w2[(i >> 1) & 3]));
    }
  }
  float s = 0;
  if (MODE == 0 || MODE == 2) {
    for (int i = 0; i < 64; ++i) s += a[i];
  } else {
    for (int i = 0; i < 32; ++i) { float u, v; asm("mov.b64 {%0, %1}, %2;" : "=f"(u), "=f"(v) : "l"(p[i])); s += u + v; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(float* o, const char* name, int warps_per_sm) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 10000;
  const int threads = 128, blocks = 148 * warps_per_sm / 4;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(o, iters, 1.0001f, 1e-7f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)blocks * threads * 64.0 * iters;
    if (rep) printf("%-44s %2d warps/SM: %.3f ms  %.2f TFMA/s (%.1f TFLOP/s)\n", name, warps_per_sm, ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
  }
}
int main() {
  float* o; cudaMalloc(&o, 148 * 16 * 128 * 4);
  for (int wps : {8, 16, 32}) {
    run<0>(o, "FFMA  acc=acc*c0+c1 (const operands)", wps);
    run<1>(o, "FFMA2 acc=acc*c+c", wps);
    run<2>(o, "FFMA  acc+=x*w (3 distinct regs)", wps);
    run<3>(o, "FFMA2 acc2+=x(scalar)*w2", wps);
    run<4>(o, "FFMA2 acc2+=x2*w2 (3 distinct pairs)", wps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
