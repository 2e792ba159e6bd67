// Companion of umma_shift_probe.cu for 64-BYTE rows (experiment queue: 16-channel layers on the
// tensor cores): does a K-major SWIZZLE_64B shared-memory matrix descriptor whose start address is shifted by s rows
// (s * 64 bytes, not a multiple of the 512-byte swizzle atom) read rows s .. s+127 of a tile of 16-float rows that
// was written with the swizzle of its ABSOLUTE address (16-byte chunk ^= address bits [7,9))?  If so, the
// halo-staged kernel extends to 16 channels per staged position.  Tries base_offset = 0 and (start >> 7) & 7.
// NOT YET RUN (written after the round's GPU budget was spent).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I robocupvision_b200/csrc -I include -o /tmp/umma_shift_probe64 \
//        tools/umma_shift_probe64.cu && timeout 60 /tmp/umma_shift_probe64
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "rcv_umma.cuh"

using namespace rcv_umma;

constexpr int ROWS = 160;  // patch rows (pixels)
constexpr int BN = 32;     // B rows (output channels)
constexpr int KB = 16;     // fp32 per row = 64 bytes
constexpr int ROWB = KB * 4;

// SWIZZLE_64B K-major descriptor: stride between 8-row groups = 8 x 64 B, layout type 4 (rcv_umma.cu: make_desc_kb<16>)
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// byte offset of 16-byte chunk ch of row r at the swizzle of the ABSOLUTE address: chunk ^= address bits [7,9)
__device__ __forceinline__ uint32_t sw64(uint32_t tile_base, int r, int ch) {
  const uint32_t row_addr = tile_base + (uint32_t)r * ROWB;
  return (uint32_t)r * ROWB + ((((uint32_t)ch) ^ ((row_addr >> 7) & 3u)) << 4);
}

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
  for (int e = tid; e < ROWS * 4; e += 128) {
    const int r = e / 4, ch = e % 4;
    float4 v;
    v.x = (float)((r * 7 + (4 * ch + 0) * 3) % 32); v.y = (float)((r * 7 + (4 * ch + 1) * 3) % 32);
    v.z = (float)((r * 7 + (4 * ch + 2) * 3) % 32); v.w = (float)((r * 7 + (4 * ch + 3) * 3) % 32);
    *reinterpret_cast<float4*>(gen + sw64(base, r, ch)) = v;
  }
  unsigned char* genB = gen + ROWS * ROWB;  // 1024-aligned since ROWS * 64 is a multiple of 1024
  for (int e = tid; e < BN * 4; e += 128) {
    const int n = e / 4, ch = e % 4;
    float4 v;
    v.x = (float)((n * 5 + 4 * ch + 0) % 16); v.y = (float)((n * 5 + 4 * ch + 1) % 16);
    v.z = (float)((n * 5 + 4 * ch + 2) % 16); v.w = (float)((n * 5 + 4 * ch + 3) % 16);
    *reinterpret_cast<float4*>(genB + sw64(base + ROWS * ROWB, n, ch)) = v;
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (tid < 32) tmem_alloc(smem_u32(&tmem_slot), 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t astart = base + shift * ROWB;
    uint64_t adesc = make_desc64(astart);
    if (use_base_offset) adesc |= (uint64_t)((astart >> 7) & 7u) << 49;
    const uint64_t bdesc = make_desc64(base + ROWS * ROWB);
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
  const int smem = ROWS * ROWB + BN * ROWB + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  static float h[128 * BN];
  for (int bo = 0; bo < 2; ++bo)
    for (int shift : {0, 1, 2, 3, 4, 5, 7, 8, 9, 13, 16, 22, 23}) {
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
