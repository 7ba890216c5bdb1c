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

int dp_pdl_enabled(void);   // dp_core.cu: programmatic dependent launch on (default) / off (DP_PDL=0)
int dp_pdl_take(void);      // ... for the next launch of this thread, given the work hint set by dp::pdl_work()
void dp_pdl_hint(double bytes_equivalent);

namespace dp {

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch ----------------------------------------------------------------------------------
// The train step is ~1500 short launches in one stream.  Launched with the programmatic-stream-serialization attribute a
// kernel's blocks may be scheduled while the previous kernel of the stream is still draining (every block of it has
// executed pdl_trigger(), which all kernels here do first thing), so launch latency, block scheduling and the
// prologue (barrier init, TMEM allocation, descriptor prefetch) overlap the predecessor's tail.  pdl_wait() returns once
// the predecessor grid has COMPLETED and its writes are visible: every kernel launched through dp::launch executes it
// before its first global-memory access, which also makes completion transitive along the stream (a grid cannot
// complete before its own wait returned).  Both instructions are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_trigger(); pdl_wait(); }

// Work hint for the next dp::launch of this thread, in bytes of HBM traffic (flops / 200 for tensor-bound launches).
// Measured: the early launch pays on the short kernels (train step 61.7 -> 60.9 ms) and costs ~2 % on a workload of
// long ones (the 896x1152 DPT decoder), so launches above DP_PDL_MAX_MB (default 256 MB ~ 50 us) stay ordinary.
inline void pdl_work(double bytes_equivalent) { dp_pdl_hint(bytes_equivalent); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = dp_pdl_take();
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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
