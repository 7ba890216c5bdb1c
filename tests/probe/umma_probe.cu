// Diagnostic entry point: one CTA, one UMMA chain with fully host-specified descriptors, full TMEM dump.
// Test scaffolding: built into tests/probe/libdepth_b200_probe.so (NOT part of the product library).
// Used by tests/test_umma_probe_gpu.py to pin the shared-memory descriptor conventions (K-major and
// MN-major, every swizzle mode, sub-tile start offsets) and the TMEM accumulator layouts (M=128, M=64)
// the convolution kernels rely on.  Not on the hot path.
#include "../../monocular-depth-estimation-cil_b200/csrc/common.cuh"
#include "../../monocular-depth-estimation-cil_b200/csrc/tc.cuh"
#include "probe.h"

namespace {

struct ProbeArgs {
  int a_nbox, a_box_bytes, a_box_cols, b_nbox, b_box_bytes, b_box_cols;
  int M, N, nk;
  uint32_t idesc;
  uint32_t a_lbo, a_sbo, a_layout, a_kadv, a_off, a_boff;
  uint32_t b_lbo, b_sbo, b_layout, b_kadv, b_off;
  float* out;  // [128][ncols_dump]
  int ncols_dump;
};

__global__ void __launch_bounds__(128) umma_probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmB, ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((p.a_nbox * p.a_box_bytes + 1023) & ~1023);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar_full, 1);
    tc::mbar_init(&bar_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  // zero the dumped TMEM columns first so untouched lanes read as a recognisable value
  // (tcgen05.st is avoided: we simply rely on accumulate=0 for the first MMA and report raw contents)
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(&bar_full, p.a_nbox * p.a_box_bytes + p.b_nbox * p.b_box_bytes);
    for (int i = 0; i < p.a_nbox; ++i) tc::tma_load_2d(sA + i * p.a_box_bytes, &tmA, &bar_full, i * p.a_box_cols, 0);
    for (int i = 0; i < p.b_nbox; ++i) tc::tma_load_2d(sB + i * p.b_box_bytes, &tmB, &bar_full, i * p.b_box_cols, 0);
    tc::mbar_wait(&bar_full, 0);
    tc::fence_after_sync();
    for (int k = 0; k < p.nk; ++k) {
      uint64_t da = tc::make_smem_desc(tc::smem_u32(sA) + p.a_off + k * p.a_kadv, p.a_lbo, p.a_sbo, p.a_layout, p.a_boff);
      uint64_t db = tc::make_smem_desc(tc::smem_u32(sB) + p.b_off + k * p.b_kadv, p.b_lbo, p.b_sbo, p.b_layout);
      tc::umma_bf16(tmem, da, db, p.idesc, k > 0 ? 1u : 0u);
    }
    tc::umma_commit(&bar_done);
  }
  __syncwarp();
  tc::mbar_wait(&bar_done, 0);
  tc::fence_after_sync();
  const int lane_row = warp * 32 + (threadIdx.x & 31);
  for (int c0 = 0; c0 < p.ncols_dump; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) p.out[(size_t)lane_row * p.ncols_dump + c0 + j] = v[j];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int dp_umma_probe(const void* A, int a_rows, int a_cols, int a_box_rows, int a_box_cols, const void* B,
                             int b_rows, int b_cols, int b_box_rows, int b_box_cols, int M, int N, int nk, int a_mn,
                             int b_mn, const uint32_t* adesc /*host: lbo,sbo,layout,kadv,off,base_offset*/,
                             const uint32_t* bdesc /*host*/, float* out, int ncols_dump, cudaStream_t stream) {
  DP_CHECK_ARG(A && B && out && adesc && bdesc, "dp_umma_probe: null pointer");
  DP_CHECK_ARG(ncols_dump % 16 == 0 && ncols_dump <= 512, "dp_umma_probe: ncols_dump");
  CUtensorMap tmA, tmB;
  uint64_t da[2] = {(uint64_t)a_cols, (uint64_t)a_rows}, sa[1] = {(uint64_t)a_cols * 2};
  uint32_t ba[2] = {(uint32_t)a_box_cols, (uint32_t)a_box_rows};
  int rc = dp_make_tmap_bf16(&tmA, A, 2, da, sa, ba, nullptr, a_box_cols * 2);
  if (rc) return rc;
  uint64_t db[2] = {(uint64_t)b_cols, (uint64_t)b_rows}, sb[1] = {(uint64_t)b_cols * 2};
  uint32_t bb[2] = {(uint32_t)b_box_cols, (uint32_t)b_box_rows};
  rc = dp_make_tmap_bf16(&tmB, B, 2, db, sb, bb, nullptr, b_box_cols * 2);
  if (rc) return rc;
  ProbeArgs p;
  p.a_nbox = a_cols / a_box_cols; p.a_box_bytes = a_box_rows * a_box_cols * 2; p.a_box_cols = a_box_cols;
  p.b_nbox = b_cols / b_box_cols; p.b_box_bytes = b_box_rows * b_box_cols * 2; p.b_box_cols = b_box_cols;
  p.M = M; p.N = N; p.nk = nk;
  p.idesc = tc::make_idesc_bf16(M, N, a_mn, b_mn);
  p.a_lbo = adesc[0]; p.a_sbo = adesc[1]; p.a_layout = adesc[2]; p.a_kadv = adesc[3]; p.a_off = adesc[4]; p.a_boff = adesc[5];
  p.b_lbo = bdesc[0]; p.b_sbo = bdesc[1]; p.b_layout = bdesc[2]; p.b_kadv = bdesc[3]; p.b_off = bdesc[4];
  p.out = out; p.ncols_dump = ncols_dump;
  size_t smem = 1024 + ((p.a_nbox * p.a_box_bytes + 1023) & ~1023) + ((p.b_nbox * p.b_box_bytes + 1023) & ~1023);
  DP_CHECK_ARG(smem <= 200 * 1024, "dp_umma_probe: operands too large for shared memory");
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return dp_set_error(DP_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  umma_probe_kernel<<<1, 128, smem, stream>>>(tmA, tmB, p);
  DP_CHECK_LAUNCH("umma_probe_kernel");
  return DP_OK;
}
