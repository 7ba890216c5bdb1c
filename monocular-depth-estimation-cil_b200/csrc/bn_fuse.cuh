// BatchNorm fused around the tcgen05 convolutions: the in-place shared-memory transform that turns a TMA-landed tile of
// the pre-BatchNorm tensor c into act(c * scale + shift) before the tensor core reads it (conv_tc.cu forward / data
// gradient, wgrad_tc.cu weight gradient).  Replaces the separate bn_apply pass of the reference's
// conv -> BatchNorm2d -> ReLU -> conv chains (midas_semantics.py:132-150, 195-203).
#pragma once
#include <cuda_bf16.h>
#include <cstdint>
#include "tc.cuh"

namespace dpf {

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ---- prologue transform (NT threads): BatchNorm scale/shift + activation applied in place to the A box
// a TMA load just wrote, in its swizzled layout.  The box holds `npx` pixel rows of ROWB = 2*KB bytes; a thread owns one
// 16-byte channel group (8 channels) of every (NT / chunks-per-row)-th pixel, so its scale / shift stay in registers for
// the whole k-chunk.  Pixels outside the source plane were zero-filled by TMA and must stay zero: the reference pads the
// activated tensor (nn.Conv2d padding follows nn.ReLU), and act(0 * scale + shift) != 0.
template <int ROWB, int NT>
__device__ __forceinline__ void transform_box(uint32_t base, int npx, int boxW, int bx0, int by0, int pW, int pH,
                                              const float* __restrict__ s_pre, int pre_pad, int ch0, int act, int tid) {
  constexpr int CPP = ROWB / 16;                  // 16-byte chunks per pixel row
  constexpr uint32_t kSwz = ROWB == 128 ? 7u : (ROWB == 64 ? 3u : 1u);
  constexpr int PSTEP = NT / CPP;
  const int j = tid % CPP;
  float sc[8], sh[8];
  {
    const float4 a0 = *reinterpret_cast<const float4*>(s_pre + ch0 + j * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(s_pre + ch0 + j * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(s_pre + pre_pad + ch0 + j * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(s_pre + pre_pad + ch0 + j * 8 + 4);
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
  }
  const int rows = npx / boxW;
  const bool interior = bx0 >= 0 && by0 >= 0 && bx0 + boxW <= pW && by0 + rows <= pH;   // block-uniform
  const float hi = act == 2 ? 6.f : __int_as_float(0x7f800000);
  constexpr int U = 4;
  for (int pix0 = tid / CPP; pix0 < npx; pix0 += U * PSTEP) {
    uint4 v[U];
    uint32_t addr[U];
    bool live[U], inside[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * PSTEP;
      live[u] = pix < npx;
      const uint32_t off = (uint32_t)pix * ROWB + (uint32_t)j * 16u;
      addr[u] = base + (off ^ (((off >> 7) & kSwz) << 4));
      inside[u] = true;
      if (!interior) {
        const int by = pix / boxW, bx = pix - by * boxW;
        inside[u] = (unsigned)(by0 + by) < (unsigned)pH && (unsigned)(bx0 + bx) < (unsigned)pW;
      }
      if (live[u])
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                     : "r"(addr[u]));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!live[u]) continue;
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      uint32_t q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f0 = fmaf(bf16_lo(w[k]), sc[2 * k], sh[2 * k]);
        float f1 = fmaf(bf16_hi(w[k]), sc[2 * k + 1], sh[2 * k + 1]);
        if (act) { f0 = fminf(fmaxf(f0, 0.f), hi); f1 = fminf(fmaxf(f1, 0.f), hi); }
        q[k] = inside[u] ? pack_bf16x2(f0, f1) : 0u;
      }
      tc::st_shared_v4(addr[u], q[0], q[1], q[2], q[3]);
    }
  }
}

}  // namespace dpf
