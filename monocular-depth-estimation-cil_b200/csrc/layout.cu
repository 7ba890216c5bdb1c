// Layout / dtype boundaries of the NHWC bf16 activation path and weight packing.
//   * the reference's modules speak NCHW fp32 (src/network/*.py); encoder feature maps enter and the
//     gradients w.r.t. them leave through the two transposing converters below;
//   * nn.Conv2d weights stay fp32 OIHW in the module (state_dict compatible, blocks.py:149-161 etc.) and are
//     re-packed to the K-major bf16 layouts conv_tc.cu consumes.
#include "common.cuh"
#include "../../include/depth_b200.h"

namespace {

// (B,C,H,W) fp32 -> (B,H,W,ld) bf16: 32 pixels x 32 channels tiles through shared memory
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, int HW, bf16* __restrict__ dst, long long ld) {
  dp::pdl_prologue();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + (size_t)b * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? __ldg(s + (size_t)c * HW + p) : 0.f;
  }
  __syncthreads();
  bf16* d = dst + (size_t)b * HW * ld;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) d[(size_t)p * ld + c] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// C <= 8 channels into 8-channel pixels (the RGB input of the stem convolution): one thread per pixel, plane reads
// coalesced across the warp, one 16-byte store per pixel with the pad channels written as zeros
__global__ void __launch_bounds__(256) nchw_to_nhwc8_kernel(const float* __restrict__ src, int C, int HW,
                                                            bf16* __restrict__ dst) {
  dp::pdl_prologue();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float* s = src + (size_t)blockIdx.y * C * HW + p;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = c < C ? __ldg(s + (size_t)c * HW) : 0.f;
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(dst + ((size_t)blockIdx.y * HW + p) * 8) = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ src, long long ld, int C, int HW, float* __restrict__ dst) {
  dp::pdl_prologue();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const bf16* s = src + (size_t)b * HW * ld;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? __bfloat162float(s[(size_t)p * ld + c]) : 0.f;
  }
  __syncthreads();
  float* d = dst + (size_t)b * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) d[(size_t)c * HW + p] = tile[threadIdx.x][i];
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  dp::pdl_prologue();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(__ldg(src + i));
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, size_t n) {
  dp::pdl_prologue();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}

// weight [D0][D1][KH][KW] fp32 -> bf16 [tap'][A][ld], (A, b) = swap ? (D1, d0) : (D0, d1); tap' reversed when flip
__global__ void pack_weight_kernel(const float* __restrict__ w, int D0, int D1, int KH, int KW, int swap, int flip,
                                   bf16* __restrict__ dst, int ld) {
  dp::pdl_prologue();
  const long long total = (long long)D0 * D1 * KH * KW;
  const int taps = KH * KW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int d1 = (int)((i / taps) % D1);
    const int d0 = (int)(i / ((long long)taps * D1));
    const int t = flip ? taps - 1 - tap : tap;
    const int A = swap ? D1 : D0;
    const int ai = swap ? d1 : d0, bi = swap ? d0 : d1;
    dst[((size_t)t * A + ai) * ld + bi] = __float2bfloat16_rn(__ldg(w + i));
  }
}

// All weight packs of a step in ONE launch (a train step needs ~230: forward and data-gradient layout of every
// convolution; one tiny kernel each costs more in launch gaps than in work).  blockIdx.y = descriptor.
__global__ void pack_weights_batched_kernel(const dp_pack_desc_t* __restrict__ descs) {
  dp::pdl_prologue();
  const dp_pack_desc_t d = descs[blockIdx.y];
  const float* __restrict__ w = reinterpret_cast<const float*>(d.src);
  bf16* __restrict__ dst = reinterpret_cast<bf16*>(d.dst);
  const int taps = d.KH * d.KW;
  const long long total = (long long)d.D0 * d.D1 * taps;
  const int A = d.swap ? d.D1 : d.D0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int d1 = (int)((i / taps) % d.D1);
    const int d0 = (int)(i / ((long long)taps * d.D1));
    const int t = d.flip ? taps - 1 - tap : tap;
    const int ai = d.swap ? d1 : d0, bi = d.swap ? d0 : d1;
    dst[((size_t)t * A + ai) * d.ld + bi] = __float2bfloat16_rn(__ldg(w + i));
  }
}

}  // namespace

extern "C" {

int dp_pack_conv_weights_batched(const dp_pack_desc_t* descs_device, int n, int blocks_per_weight, cudaStream_t stream) {
  DP_CHECK_ARG(descs_device && n > 0 && n <= 65535 && blocks_per_weight > 0, "dp_pack_conv_weights_batched: bad arguments");
  dp::launch(pack_weights_batched_kernel, dim3(blocks_per_weight, n), 256, 0, stream, descs_device);
  DP_CHECK_LAUNCH("pack_weights_batched_kernel");
  return DP_OK;
}

int dp_nchw_f32_to_nhwc_bf16(const float* src, int B, int C, int H, int W, void* dst, long long dst_ld,
                             cudaStream_t stream) {
  DP_CHECK_ARG(src && dst && B > 0 && C > 0 && dst_ld >= C, "dp_nchw_f32_to_nhwc_bf16: bad arguments");
  const int HW = H * W;
  if (C <= 8 && dst_ld == 8 && B <= 65535 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    dp::launch(nchw_to_nhwc8_kernel, dim3(dp::ceil_div(HW, 256), B), 256, 0, stream, src, C, HW, reinterpret_cast<bf16*>(dst));
    DP_CHECK_LAUNCH("nchw_to_nhwc8_kernel");
    return DP_OK;
  }
  dim3 grid(dp::ceil_div(HW, 32), dp::ceil_div(C, 32), B), block(32, 8);
  dp::launch(nchw_to_nhwc_kernel, grid, block, 0, stream, src, C, HW, reinterpret_cast<bf16*>(dst), dst_ld);
  DP_CHECK_LAUNCH("nchw_to_nhwc_kernel");
  return DP_OK;
}

int dp_nhwc_bf16_to_nchw_f32(const void* src, long long src_ld, int B, int C, int H, int W, float* dst,
                             cudaStream_t stream) {
  DP_CHECK_ARG(src && dst && B > 0 && C > 0 && src_ld >= C, "dp_nhwc_bf16_to_nchw_f32: bad arguments");
  const int HW = H * W;
  dim3 grid(dp::ceil_div(HW, 32), dp::ceil_div(C, 32), B), block(32, 8);
  dp::launch(nhwc_to_nchw_kernel, grid, block, 0, stream, reinterpret_cast<const bf16*>(src), src_ld, C, HW, dst);
  DP_CHECK_LAUNCH("nhwc_to_nchw_kernel");
  return DP_OK;
}

int dp_cast_f32_to_bf16(const float* src, void* dst, size_t n, cudaStream_t stream) {
  DP_CHECK_ARG(src && dst, "dp_cast_f32_to_bf16: null pointer");
  if (n == 0) return DP_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 8 * dp::kNumSMs) blocks = 8 * dp::kNumSMs;
  dp::launch(cast_f32_bf16_kernel, blocks, 256, 0, stream, src, reinterpret_cast<bf16*>(dst), n);
  DP_CHECK_LAUNCH("cast_f32_bf16_kernel");
  return DP_OK;
}

int dp_cast_bf16_to_f32(const void* src, float* dst, size_t n, cudaStream_t stream) {
  DP_CHECK_ARG(src && dst, "dp_cast_bf16_to_f32: null pointer");
  if (n == 0) return DP_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 8 * dp::kNumSMs) blocks = 8 * dp::kNumSMs;
  dp::launch(cast_bf16_f32_kernel, blocks, 256, 0, stream, reinterpret_cast<const bf16*>(src), dst, n);
  DP_CHECK_LAUNCH("cast_bf16_f32_kernel");
  return DP_OK;
}

int dp_pack_conv_weight(const float* w, int D0, int D1, int KH, int KW, int swap, int flip, void* dst, int ld,
                        cudaStream_t stream) {
  DP_CHECK_ARG(w && dst, "dp_pack_conv_weight: null pointer");
  DP_CHECK_ARG(ld >= (swap ? D0 : D1), "dp_pack_conv_weight: row stride too small");
  const long long total = (long long)D0 * D1 * KH * KW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 8 * dp::kNumSMs) blocks = 8 * dp::kNumSMs;
  dp::launch(pack_weight_kernel, blocks, 256, 0, stream, w, D0, D1, KH, KW, swap, flip, reinterpret_cast<bf16*>(dst), ld);
  DP_CHECK_LAUNCH("pack_weight_kernel");
  return DP_OK;
}

}  // extern "C"
