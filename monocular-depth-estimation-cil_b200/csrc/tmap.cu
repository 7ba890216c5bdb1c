// Host-side CUtensorMap encoding (driver entry point fetched at run time; no libcuda link dependency).
#include "common.cuh"
#include "tc.cuh"
#include <mutex>

dp_encode_tiled_fn dp_get_encode_tiled() {
  static dp_encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<dp_encode_tiled_fn>(p);
  });
  return fn;
}

int dp_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const uint32_t* elem_strides, int row_bytes) {
  dp_encode_tiled_fn enc = dp_get_encode_tiled();
  if (!enc) return dp_set_error(DP_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                            : CU_TENSOR_MAP_SWIZZLE_NONE;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return dp_set_error(DP_ERR_INVALID, "tensor map base %p not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    if (gs[i] % 16 != 0) return dp_set_error(DP_ERR_INVALID, "tensor map stride %llu not a multiple of 16 bytes",
                                             (unsigned long long)gs[i]);
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return dp_set_error(DP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, box0 %u, rowbytes %d)",
                        (int)r, rank, box[0], row_bytes);
  return DP_OK;
}
