// Host-side TMA descriptor construction (cuTensorMapEncodeTiled through the runtime's driver
// entry-point query, so the library has no link-time dependency on libcuda).
#include "common.cuh"

#include <cudaTypedefs.h>
#include <mutex>

namespace tfimm {
namespace {

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
  });
  return fn;
}

}  // namespace

// Generic tiled map, swizzle_bytes in {128, 64, 0 = none}, zero OOB fill.  dims/box innermost first; strides (bytes) for dims 1..rank-1.
int make_tmap(CUtensorMap* map, const void* ptr, int dtype, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, const char* what, int swizzle_bytes,
              const uint32_t* elem_strides) {
  auto encode = get_encode_fn();
  if (encode == nullptr) {
    set_last_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return kCudaError;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0) {
    set_last_error("tensor map %s: base pointer %p is not 16-byte aligned", what, ptr);
    return kInvalidArgument;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstride[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = elem_strides != nullptr ? elem_strides[i] : 1;  // traversal stride (strided convolutions)
    if (i > 0) {
      if ((strides_bytes[i - 1] & 15u) != 0) {
        set_last_error("tensor map %s: stride %llu of dim %d is not a multiple of 16 bytes", what,
                       (unsigned long long)strides_bytes[i - 1], i);
        return kInvalidArgument;
      }
      gstride[i - 1] = strides_bytes[i - 1];
    }
  }
  CUresult r = encode(map, dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                      (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstride, bx, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE,
                      swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                      : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return kCudaError;
  }
  return kOk;
}

// 2D row-major tensor [rows, cols] with leading dimension ld (elements); box = [box_rows, box_cols].
int make_tmap_2d(CUtensorMap* map, const void* ptr, int dtype, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols, const char* what, int swizzle_bytes) {
  const uint64_t esize = dtype == kBF16 ? 2 : 4;
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld * esize};
  const uint32_t box[2] = {box_cols, box_rows};
  return make_tmap(map, ptr, dtype, 2, dims, strides, box, what, swizzle_bytes);
}

}  // namespace tfimm
