// esb200 — host-side TMA tensor-map construction for the kernels that load / store tiles with cp.async.bulk.tensor.
// cuTensorMapEncodeTiled is resolved through the runtime (cudaGetDriverEntryPoint), so libesb200.so carries no link-time
// dependency on libcuda and still loads (and lists its symbols) on a box without a driver.
#include "tc_common.cuh"

#include <mutex>

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;

void resolve() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode = (EncodeTiledFn)fn;
}
}  // namespace

int esb_tma_encode(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                   const unsigned long long* strides_bytes, const unsigned* box, const unsigned* elem_strides,
                   int swizzle_bytes) {
  std::call_once(g_once, resolve);
  if (g_encode == nullptr) {
    esb_set_error("esb_tma_encode: cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return ESB_ECUDA;
  }
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = elem_strides ? elem_strides[i] : 1;
    if (i + 1 < rank) s[i] = strides_bytes[i];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    esb_set_error("esb_tma_encode: cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu..., box %u %u %u..., "
                  "stride0 %llu, swizzle %d)", (int)r, rank, dims[0], rank > 1 ? dims[1] : 0ull, rank > 2 ? dims[2] : 0ull, box[0],
                  rank > 1 ? box[1] : 0u, rank > 2 ? box[2] : 0u, rank > 1 ? strides_bytes[0] : 0ull, swizzle_bytes);
    return ESB_EINVAL;
  }
  return ESB_OK;
}
