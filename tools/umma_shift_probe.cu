// Probe for the halo-staged A operand: does a K-major SWIZZLE_128B shared-memory matrix
// descriptor whose start address is shifted by s rows (s * 128 bytes, not a multiple of the 1024-byte swizzle
// atom) read rows s .. s+127 of a tile that was written with the swizzle of its ABSOLUTE address?  If so, the
// nine taps of a 3x3 convolution are nine descriptors over ONE staged [pixel][channel] patch (no 9x im2col).
// Tries base_offset = 0 and base_offset = (start >> 7) & 7 (descriptor bits [49,52)).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I robocupvision_b200/csrc -I include -o /tmp/umma_shift_probe \
//        tools/umma_shift_probe.cu && timeout 60 /tmp/umma_shift_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "rcv_umma.cuh"

using namespace rcv_umma;

constexpr int ROWS = 160;  // patch rows (pixels)
constexpr int BN = 32;     // B rows (output channels)
constexpr int KB = 32;     // fp32 per row = 128 bytes

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity, int max_iter) {
  for (int i = 0; i < max_iter; ++i) {
    uint32_t ok = 0;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

__global__ void __launch_bounds__(128) probe(int shift, int use_base_offset, float* out, int* status) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gen = smem_raw + (base - raw);
  const int tid = threadIdx.x;
  // A patch: element (r, c) = ((r*7 + c*3) % 32); written at the swizzle of its absolute address
  for (int e = tid; e < ROWS * 8; e += 128) {
    const int r = e / 8, ch = e % 8;
    float4 v;
    v.x = (float)((r * 7 + (4 * ch + 0) * 3) % 32); v.y = (float)((r * 7 + (4 * ch + 1) * 3) % 32);
    v.z = (float)((r * 7 + (4 * ch + 2) * 3) % 32); v.w = (float)((r * 7 + (4 * ch + 3) * 3) % 32);
    *reinterpret_cast<float4*>(gen + r * 128 + ((ch ^ (r & 7)) << 4)) = v;
  }
  unsigned char* genB = gen + ROWS * 128;  // 1024-aligned since ROWS*128 is a multiple of 1024
  for (int e = tid; e < BN * 8; e += 128) {
    const int n = e / 8, ch = e % 8;
    float4 v;
    v.x = (float)((n * 5 + 4 * ch + 0) % 16); v.y = (float)((n * 5 + 4 * ch + 1) % 16);
    v.z = (float)((n * 5 + 4 * ch + 2) % 16); v.w = (float)((n * 5 + 4 * ch + 3) % 16);
    *reinterpret_cast<float4*>(genB + n * 128 + ((ch ^ (n & 7)) << 4)) = v;
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (tid < 32) tmem_alloc(smem_u32(&tmem_slot), 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t astart = base + shift * 128;
    uint64_t adesc = make_desc(astart);
    if (use_base_offset) adesc |= (uint64_t)((astart >> 7) & 7u) << 49;
    const uint64_t bdesc = make_desc(base + ROWS * 128);
    const uint32_t idesc = make_idesc(128, BN);
    for (int ks = 0; ks < KB / 8; ++ks) umma_tf32(tmem, adesc + 2 * ks, bdesc + 2 * ks, idesc, ks != 0);
    umma_commit(smem_u32(&bar));
  }
  const bool done = mbar_wait_bounded(smem_u32(&bar), 0, 1 << 22);
  if (!done) { if (tid == 0) *status = -1; }
  tc_fence_after();
  if (done) {
    // warp w reads TMEM lanes 32w..32w+31, columns 0..31
    const int warp = tid >> 5;
    uint32_t r0[16], r1[16];
    tmem_ld16_nowait(tmem + ((uint32_t)(warp * 32) << 16), r0);
    tmem_ld16_nowait(tmem + ((uint32_t)(warp * 32) << 16) + 16, r1);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) { out[tid * BN + j] = __uint_as_float(r0[j]); out[tid * BN + 16 + j] = __uint_as_float(r1[j]); }
    if (tid == 0) *status = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 32);
}

int main() {
  float* out; int* status;
  cudaMalloc(&out, 128 * BN * 4); cudaMalloc(&status, 4);
  const int smem = ROWS * 128 + BN * 128 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  static float h[128 * BN];
  for (int bo = 0; bo < 2; ++bo)
    for (int shift : {0, 1, 2, 3, 5, 7, 8, 9, 13, 22}) {
      cudaMemset(out, 0, sizeof(h)); cudaMemset(status, 0, 4);
      probe<<<1, 128, smem>>>(shift, bo, out, status);
      cudaError_t e = cudaDeviceSynchronize();
      int st = 0; cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost); cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      int bad = 0; double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < BN; ++n) {
          double ref = 0;
          for (int c = 0; c < KB; ++c) ref += (double)(((m + shift) * 7 + c * 3) % 32) * (double)((n * 5 + c) % 16);
          const double d = fabs(ref - h[m * BN + n]);
          if (d > 1e-3) ++bad;
          if (d > maxerr) maxerr = d;
        }
      printf("base_offset=%d shift=%2d: status %d (%s) mismatches %d / %d, max err %.1f\n", bo, shift, st,
             cudaGetErrorString(e), bad, 128 * BN, maxerr);
      if (e != cudaSuccess) return 1;
    }
  return 0;
}
