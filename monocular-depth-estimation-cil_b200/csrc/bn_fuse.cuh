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

// packs (lo, hi) to bf16x2 with the ReLU folded into the conversion (one F2FP instead of two FMNMX + one F2FP)
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t min_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("min.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// ---- prologue transform (NT threads): BatchNorm scale/shift + activation applied in place to the A box a TMA load just
// wrote, in its swizzled layout.  The box holds `npx` pixel rows of ROWB = 2*KB bytes; a thread owns one 16-byte channel
// group (8 channels) of every (NT / chunks-per-row)-th pixel, so its scale / shift stay in registers for the whole
// k-chunk, and - because NT / chunks-per-row pixel rows are a whole number of 1024-byte swizzle atoms - so does the
// swizzled position of its chunk within a row: the address just advances by a constant.  Pixels outside the source
// plane were zero-filled by TMA and must stay zero: the reference pads the activated tensor (nn.Conv2d padding follows
// nn.ReLU), and act(0 * scale + shift) != 0.  Arithmetic: fp32 fma, one rounding to bf16 - bit-identical to dp_bn_apply
// (min(max(v, 0), 6) commutes with the monotone rounding, and 0 and 6 are exact in bf16).
template <int ROWB, int NT>
__device__ __forceinline__ void transform_box(uint32_t base, int npx, int boxW, int bx0, int by0, int pW, int pH,
                                              const float* __restrict__ s_pre, int pre_pad, int ch0, int act, int tid) {
  constexpr int CPP = ROWB / 16;                  // 16-byte chunks per pixel row
  constexpr uint32_t kSwz = ROWB == 128 ? 7u : (ROWB == 64 ? 3u : 1u);
  constexpr int PSTEP = NT / CPP;
  static_assert((PSTEP * ROWB) % 1024 == 0, "a thread's stride must be a whole number of swizzle atoms");
  const int j = tid % CPP, p0 = tid / CPP;
  float sc[8], sh[8];
  {
    const uint32_t tb = tc::smem_u32(s_pre) + (uint32_t)(ch0 + j * 8) * 4u;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[0]), "=f"(sc[1]), "=f"(sc[2]), "=f"(sc[3]) : "r"(tb));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[4]), "=f"(sc[5]), "=f"(sc[6]), "=f"(sc[7]) : "r"(tb + 16u));
    const uint32_t tb2 = tb + (uint32_t)pre_pad * 4u;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sh[0]), "=f"(sh[1]), "=f"(sh[2]), "=f"(sh[3]) : "r"(tb2));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sh[4]), "=f"(sh[5]), "=f"(sh[6]), "=f"(sh[7]) : "r"(tb2 + 16u));
  }
  const int rows = npx / boxW;
  const bool interior = bx0 >= 0 && by0 >= 0 && bx0 + boxW <= pW && by0 + rows <= pH;   // block-uniform
  const uint32_t off0 = (uint32_t)p0 * ROWB + (uint32_t)j * 16u;
  uint32_t addr = base + (off0 ^ (((off0 >> 7) & kSwz) << 4));
  constexpr uint32_t kStep = (uint32_t)PSTEP * ROWB;
  constexpr uint32_t kSix2 = 0x40C040C0u;         // bf16x2 (6.0, 6.0)
  constexpr int U = 4;
  for (int pix0 = p0; pix0 < npx; pix0 += U * PSTEP, addr += U * kStep) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (pix0 + u * PSTEP < npx)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                     : "r"(addr + u * kStep));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * PSTEP;
      if (pix >= npx) break;
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      uint32_t q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float f0 = fmaf(bf16_lo(w[k]), sc[2 * k], sh[2 * k]);
        const float f1 = fmaf(bf16_hi(w[k]), sc[2 * k + 1], sh[2 * k + 1]);
        q[k] = act ? pack_relu_bf16x2(f0, f1) : pack_bf16x2(f0, f1);
        if (act == 2) q[k] = min_bf16x2(q[k], kSix2);
      }
      if (!interior) {
        const int by = pix / boxW, bx = pix - by * boxW;
        if (!((unsigned)(by0 + by) < (unsigned)pH && (unsigned)(bx0 + bx) < (unsigned)pW)) q[0] = q[1] = q[2] = q[3] = 0u;
      }
      tc::st_shared_v4(addr + u * kStep, q[0], q[1], q[2], q[3]);
    }
  }
}

}  // namespace dpf
