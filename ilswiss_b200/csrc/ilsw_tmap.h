// ilswiss_b200 -- host-side construction of the TMA tensor maps consumed by ilsw_tc5.cuh.
// cuTensorMapEncodeTiled is resolved through the runtime (cudaGetDriverEntryPoint): the library links cudart only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ilsw {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tmap_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// fp32 matrix with `inner` contiguous elements per row, `outer` rows of stride `ld` floats; box = {32 inner, box_outer rows},
// 128-byte swizzle (16-byte atoms for K-major panels, 32-byte atoms for MN-major ones: see ilsw_tc5.cuh), out-of-bounds
// elements read as zero.  TMA needs a 16-byte aligned base and row stride.
inline bool tmap_operand_ok(const float* base, int ld) { return (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld & 3) == 0; }
inline int make_tmap_2d(CUtensorMap* out, const float* base, int ld, int inner, int outer, int box_outer, bool mn_major) {
  EncodeTiledFn enc = tmap_encoder();
  if (!enc || !tmap_operand_ok(base, ld)) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace ilsw
