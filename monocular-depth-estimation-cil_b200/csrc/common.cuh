// Shared helpers for the sm_100a kernels behind the C-ABI in include/depth_b200.h.
// No allocation, no synchronisation and no mutable global state live here (the caller owns every
// buffer and the stream); the only process-wide object is the thread-local last-error string.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#define DP_OK 0
#define DP_ERR_INVALID (-1)
#define DP_ERR_WORKSPACE (-2)
#define DP_ERR_CUDA (-3)
#define DP_ERR_UNSUPPORTED (-4)

int dp_set_error(int code, const char* fmt, ...);
void dp_count_launch(int n);

#define DP_CHECK_ARG(cond, ...)                                          \
  do {                                                                   \
    if (!(cond)) return dp_set_error(DP_ERR_INVALID, __VA_ARGS__);       \
  } while (0)

#define DP_CHECK_LAUNCH(name)                                                              \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    dp_count_launch(1);                                                                    \
    if (e__ != cudaSuccess)                                                                \
      return dp_set_error(DP_ERR_CUDA, "%s launch failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

typedef __nv_bfloat16 bf16;

namespace dp {

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of K doubles per thread; result valid in thread 0.  smem: K * 32 doubles.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) smem[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = lane < nwarp ? smem[k * 32 + lane] : 0.0;
      v[k] = warp_sum(x);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__host__ __device__ static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace dp
