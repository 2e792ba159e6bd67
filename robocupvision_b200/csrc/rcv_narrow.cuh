// Helpers shared by the narrow-layer engines (rcv_narrow.cu: forward / input gradient,
// rcv_narrow_wgrad.cu: weight gradient): TMA tensor-map tile loads, compile-time loops, and the
// driver entry point that encodes tensor maps (fetched through the runtime: no -lcuda).
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>
#include <utility>

#include "rcv_common.cuh"
#include "rcv_umma.cuh"

namespace rcv_narrow {

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, "
      "%5}], [%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

template <int N, class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl<N>(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}


// 4-D tensor map over an fp32 NCHW tensor [N, C, H, W] with box {bw, bh, bc, 1}; out-of-range elements
// are zero-filled.  W must be a multiple of 4 (16-byte strides) and the base 16-byte aligned.
inline int make_nchw_map(CUtensorMap* map, const float* base, int N, int C, int H, int W, int bw, int bh, int bc,
                         const char* who) {
  EncodeTiledFn enc = encode_fn();
  RCV_REQUIRE(enc != nullptr, RCV_ERR_CUDA, "%s: cuTensorMapEncodeTiled not available from the driver", who);
  const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RCV_REQUIRE(r == CUDA_SUCCESS, RCV_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
  return RCV_OK;
}

}  // namespace rcv_narrow
