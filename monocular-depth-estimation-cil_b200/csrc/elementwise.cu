// Bandwidth-bound NHWC bf16 kernels around the tensor-core convolutions: ReLU backward masks, bilinear
// resize (both align_corners conventions) forward/backward, train/eval BatchNorm finalize / apply /
// backward, per-channel sums.  All vectorised 8 channels (16 bytes) per thread, coalesced along channels.
//
// Reference arithmetic: nn.ReLU / F.interpolate(mode="bilinear") (blocks.py:226-238, 432-434; dpt_depth.py:147;
// midas_semantics.py:233,243), nn.BatchNorm2d in train and eval mode (midas_semantics.py:40-61,133-151,196).
#include "common.cuh"
#include "../../include/depth_b200.h"

namespace {

using namespace dp;

struct bf8 { uint4 u; };
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    f[2 * i] = __low2float(h);
    f[2 * i + 1] = __high2float(h);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 ld8(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void st8(bf16* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
// read-only 16-byte load the compiler may not sink below later loads (keeps a batch of loads in flight together)
__device__ __forceinline__ uint4 ld8_issue(const bf16* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

inline int grid_for(size_t items, int tpb = 256, int max_waves = 8) {
  size_t b = (items + tpb - 1) / tpb;
  size_t cap = (size_t)max_waves * kNumSMs;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// out = (ga ? ga : 0) + (gb ? gb * (y > 0) : 0)            (vectors of 8)
__global__ void add_relu_bwd_kernel(const bf16* __restrict__ ga, const bf16* __restrict__ gb, const bf16* __restrict__ y,
                                    bf16* __restrict__ out, size_t n8) {
  dp::pdl_prologue();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float a[8], b[8], yy[8], o[8];
    if (ga) unpack8(ld8(ga + i * 8), a);
    if (gb) { unpack8(ld8(gb + i * 8), b); unpack8(ld8(y + i * 8), yy); }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (ga ? a[j] : 0.f) + (gb ? (yy[j] > 0.f ? b[j] : 0.f) : 0.f);
    st8(out + i * 8, pack8(o));
  }
}

__global__ void relu_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, size_t n8) {
  dp::pdl_prologue();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float v[8];
    unpack8(ld8(x + i * 8), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    st8(out + i * 8, pack8(v));
  }
}

// out = a + b (+ c)
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, const bf16* __restrict__ c,
                           bf16* __restrict__ out, size_t n8) {
  dp::pdl_prologue();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float x[8], y[8], z[8], o[8];
    unpack8(ld8(a + i * 8), x);
    unpack8(ld8(b + i * 8), y);
    if (c) unpack8(ld8(c + i * 8), z);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = x[j] + y[j] + (c ? z[j] : 0.f);
    st8(out + i * 8, pack8(o));
  }
}

// dst[p][0..C) = src[p][0..C) with independent pixel strides (channel concat / slice copies).  IDX = unsigned when the
// vector count fits 32 bits: the per-vector divide is then one multiply-shift sequence instead of a 64-bit division.
template <typename IDX>
__global__ void copy_channels_kernel(const bf16* __restrict__ src, long long src_ld, bf16* __restrict__ dst,
                                     long long dst_ld, size_t npix, int C8) {
  dp::pdl_prologue();
  const IDX total = (IDX)(npix * C8);
  const IDX stride = (IDX)gridDim.x * blockDim.x;
  for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const IDX p = i / (IDX)C8;
    const int c8 = (int)(i - p * (IDX)C8);
    st8(dst + (size_t)p * dst_ld + c8 * 8, ld8(src + (size_t)p * src_ld + c8 * 8));
  }
}

// ---- bilinear resize ------------------------------------------------------------------------------------
// source coordinate of output index o (PyTorch upsample_bilinear2d semantics)
__device__ __forceinline__ float src_coord(int o, float scale, int align) {
  if (align) return scale * o;
  float s = scale * (o + 0.5f) - 0.5f;
  return s < 0.f ? 0.f : s;
}
__host__ __device__ inline float resize_scale(int in, int out, int align) {
  if (align) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  return (float)in / (float)out;
}

// One thread = 8 channels of one output column over kRsRows consecutive output rows.  grid: x = ceil(Wo * C/8 / 256),
// y = ceil(Ho / kRsRows), z = B.  When up-sampling, kRsRows output rows draw on at most four consecutive source rows:
// those are loaded once (8 loads, all issued before the first use), interpolated horizontally once, and every output
// row interpolates vertically between two of the four results (a block-uniform pick) - column arithmetic, address
// arithmetic and half of the interpolation are shared by the rows, the arithmetic per output is unchanged.  Otherwise (down-sampling) each output row is produced on its own.
constexpr int kRsRows = 4;

template <int K0, int K1>
__device__ __forceinline__ void vlerp_store(const float (&h)[4][8], float hy, float ly, bf16* dst) {
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = hy * h[K0][j] + ly * h[K1][j];
  st8(dst, pack8(o));
}

__global__ void __launch_bounds__(256, 2) resize_fwd_kernel(const bf16* __restrict__ src, long long src_ld, int B, int Hi,
                                                            int Wi, int C, bf16* __restrict__ dst, long long dst_ld,
                                                            int Ho, int Wo, int align, float sy, float sx) {
  dp::pdl_prologue();
  const int C8 = C / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Wo * C8) return;
  const int ox = idx / C8, c8 = idx - ox * C8;
  const int oy0 = blockIdx.y * kRsRows, b = blockIdx.z;
  const float fx = src_coord(ox, sx, align);
  const int x0 = (int)fx;
  const int x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
  const float lx = fx - x0, hx = 1.f - lx;
  const bf16* s = src + (size_t)b * Hi * Wi * src_ld + c8 * 8;
  bf16* d = dst + (((size_t)b * Ho + oy0) * Wo + ox) * dst_ld + c8 * 8;
  const int nrows = min(kRsRows, Ho - oy0);

  int y0[kRsRows], y1[kRsRows];
  float ly[kRsRows];
#pragma unroll
  for (int r = 0; r < kRsRows; ++r) {
    const float fy = src_coord(min(oy0 + r, Ho - 1), sy, align);
    y0[r] = (int)fy;
    y1[r] = y0[r] + (y0[r] < Hi - 1 ? 1 : 0);
    ly[r] = fy - y0[r];
  }
  const int yb = y0[0];
  if (y1[kRsRows - 1] - yb <= 3) {
    uint4 ua[4], ub[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t row = (size_t)min(yb + k, Hi - 1) * Wi;
      ua[k] = ld8_issue(s + (row + x0) * src_ld);
      ub[k] = ld8_issue(s + (row + x1) * src_ld);
    }
    float h[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a[8], bq[8];
      unpack8(ua[k], a); unpack8(ub[k], bq);
#pragma unroll
      for (int j = 0; j < 8; ++j) h[k][j] = hx * a[j] + lx * bq[j];
    }
#pragma unroll
    for (int r = 0; r < kRsRows; ++r) {
      if (r < nrows) {
        // a block-uniform branch picks the two source rows, so the arithmetic per output stays exactly the four-corner
        // formula F.interpolate uses: hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11)
        bf16* dr = d + (size_t)r * Wo * dst_ld;
        const float hy = 1.f - ly[r];
        switch ((y0[r] - yb) * 4 + (y1[r] - yb)) {
          case 0: vlerp_store<0, 0>(h, hy, ly[r], dr); break;
          case 1: vlerp_store<0, 1>(h, hy, ly[r], dr); break;
          case 5: vlerp_store<1, 1>(h, hy, ly[r], dr); break;
          case 6: vlerp_store<1, 2>(h, hy, ly[r], dr); break;
          case 10: vlerp_store<2, 2>(h, hy, ly[r], dr); break;
          case 11: vlerp_store<2, 3>(h, hy, ly[r], dr); break;
          default: vlerp_store<3, 3>(h, hy, ly[r], dr); break;
        }
      }
    }
  } else {
    for (int r = 0; r < nrows; ++r) {
      const float hy = 1.f - ly[r];
      float a[8], bq[8], c[8], dd[8], o[8];
      unpack8(ld8(s + ((size_t)y0[r] * Wi + x0) * src_ld), a);
      unpack8(ld8(s + ((size_t)y0[r] * Wi + x1) * src_ld), bq);
      unpack8(ld8(s + ((size_t)y1[r] * Wi + x0) * src_ld), c);
      unpack8(ld8(s + ((size_t)y1[r] * Wi + x1) * src_ld), dd);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = hy * (hx * a[j] + lx * bq[j]) + ly[r] * (hx * c[j] + lx * dd[j]);
      st8(d + (size_t)r * Wo * dst_ld, pack8(o));
    }
  }
}

// fp32 variant for the (B,3,H,W) RGB -> DINOv2 input resize and the (B,1,H,W) prediction resize (NCHW planes)
__global__ void resize_planes_f32_kernel(const float* __restrict__ src, int planes, int Hi, int Wi,
                                         float* __restrict__ dst, int Ho, int Wo, int align) {
  dp::pdl_prologue();
  const size_t total = (size_t)planes * Ho * Wo;
  const float sy = resize_scale(Hi, Ho, align), sx = resize_scale(Wi, Wo, align);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    const int oy = (int)((i / Wo) % Ho);
    const size_t pl = i / ((size_t)Wo * Ho);
    const float fy = src_coord(oy, sy, align), fx = src_coord(ox, sx, align);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* s = src + pl * Hi * Wi;
    dst[i] = hy * (hx * __ldg(s + (size_t)y0 * Wi + x0) + lx * __ldg(s + (size_t)y0 * Wi + x1)) +
             ly * (hx * __ldg(s + (size_t)y1 * Wi + x0) + lx * __ldg(s + (size_t)y1 * Wi + x1));
  }
}

// Backward as a gather: input pixel (iy,ix) collects from every output pixel whose 2x2 footprint touches it; each
// candidate re-derives its own (x0, x1, lx) exactly as the forward pass did - deterministic, no atomics.
// weight of output index o on input index i along one axis (the forward's own (x0, x1, lx), re-derived)
__device__ __forceinline__ float axis_weight(int o, int i, int in, float scale, int align) {
  const float f = src_coord(o, scale, align);
  const int i0 = (int)f;
  const int i1 = i0 + (i0 < in - 1 ? 1 : 0);
  const float l = f - i0;
  float w = 0.f;
  if (i0 == i) w += 1.f - l;
  if (i1 == i) w += l;
  return w;
}

// The run of outputs that contribute to input i along one axis: `first` and the weights of first .. first+RW-1.
// Output o contributes when its source coordinate lies in (i-1, i+1); inverting the coordinate map gives the first
// candidate to within one position, the weights themselves are the forward's own (axis_weight).  Returns false when the
// run is longer than RW (down-sampling): the caller then walks [lo, hi].
template <int RW>
__device__ __forceinline__ bool axis_window(int i, int in, int out, float scale, float inv, int align, int& lo, int& hi,
                                            int& first, float (&w)[RW]) {
  const float a = align ? (i - 1) * inv : ((i - 1) + 0.5f) * inv - 0.5f;
  const float bnd = align ? (i + 1) * inv : ((i + 1) + 0.5f) * inv - 0.5f;
  lo = max(0, (int)floorf(a) - 1);
  hi = min(out - 1, (int)ceilf(bnd) + 1);
  if (scale <= 0.f) { lo = 0; hi = out - 1; }
  int f = lo;
#pragma unroll
  for (int t = 0; t < 3; ++t)
    if (f < hi && axis_weight(f, i, in, scale, align) == 0.f) ++f;
  first = f;
#pragma unroll
  for (int q = 0; q < RW; ++q) w[q] = (f + q <= hi) ? axis_weight(f + q, i, in, scale, align) : 0.f;
  bool ok = w[0] != 0.f || f >= hi;             // the advances reached the run (or there is none)
  for (int o = f + RW; o <= hi && ok; ++o) ok = axis_weight(o, i, in, scale, align) == 0.f;
  return ok;
}

// grid: x = ceil(Wi * C/8 / 256), y = Hi, z = B.  Up-sampling (every use on the hot path) touches at most a 4x4 window
// of output pixels per input pixel: the sixteen 16-byte loads are issued unconditionally (zero weight where a slot
// does not contribute), eight at a time ahead of their use; wider windows (down-sampling) walk the candidate range.
__global__ void __launch_bounds__(256, 4) resize_bwd_kernel(const bf16* __restrict__ gout, long long g_ld, int B, int Hi,
                                                            int Wi, int C, bf16* __restrict__ gin, long long gin_ld,
                                                            int Ho, int Wo, int align, float sy, float sx, float inv_sy,
                                                            float inv_sx) {
  dp::pdl_prologue();
  constexpr int RW = 4;
  const int C8 = C / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Wi * C8) return;
  const int ix = idx / C8, c8 = idx - ix * C8;
  const int iy = blockIdx.y, b = blockIdx.z;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  float2 acc2[4];                               // the 4 x 4 window path accumulates packed pairs (same fma per element)
#pragma unroll
  for (int j = 0; j < 4; ++j) acc2[j] = make_float2(0.f, 0.f);
  const bf16* g = gout + (size_t)b * Ho * Wo * g_ld + c8 * 8;
  float wx[RW], wy[RW];
  int oy_lo, oy_hi, ox_lo, ox_hi, fx0, fy0;
  const bool fx_ok = axis_window<RW>(ix, Wi, Wo, sx, inv_sx, align, ox_lo, ox_hi, fx0, wx);
  const bool fy_ok = axis_window<RW>(iy, Hi, Ho, sy, inv_sy, align, oy_lo, oy_hi, fy0, wy);
  if (fx_ok && fy_ok) {
#pragma unroll
    for (int r0 = 0; r0 < RW; r0 += 2) {      // two rows = eight loads in flight per batch
      uint4 u[2][RW];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int oy = min(fy0 + r0 + r, Ho - 1);
#pragma unroll
        for (int q = 0; q < RW; ++q) {
          const int ox = min(fx0 + q, Wo - 1);
          u[r][q] = ld8_issue(g + ((size_t)oy * Wo + ox) * g_ld);   // volatile asm: issued in program order, before the math
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int q = 0; q < RW; ++q) {
          const uint32_t uw[4] = {u[r][q].x, u[r][q].y, u[r][q].z, u[r][q].w};
          const float w = wy[r0 + r] * wx[q];
          const float2 w2 = make_float2(w, w);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            acc2[j] = __ffma2_rn(w2, make_float2(__uint_as_float(uw[j] << 16), __uint_as_float(uw[j] & 0xffff0000u)), acc2[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[2 * j] = acc2[j].x; acc[2 * j + 1] = acc2[j].y; }
  } else {
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const float wyv = axis_weight(oy, iy, Hi, sy, align);
      if (wyv == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const float wxv = axis_weight(ox, ix, Wi, sx, align);
        if (wxv == 0.f) continue;
        float v[8];
        unpack8(ld8(g + ((size_t)oy * Wo + ox) * g_ld), v);
        const float w = wyv * wxv;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += w * v[j];
      }
    }
  }
  st8(gin + (((size_t)b * Hi + iy) * Wi + ix) * gin_ld + c8 * 8, pack8(acc));
}

// ---- per-channel reductions over pixels --------------------------------------------------------------------
// MODE 0: sum x                              -> out[0][C]
// MODE 1: sum x, sum x^2                     -> out[0..1][C]                 (BN statistics from a stored tensor)
// MODE 2: g = dy * (mask ? 0<mask<mask_hi : 1); sum g, sum g*x   -> out[0..1][C]     (BN backward; mask_hi = inf for
//         ReLU, 6 for ReLU6)
constexpr int RED_TPB = 256;
template <int MODE>
__global__ void __launch_bounds__(RED_TPB) chan_reduce_kernel(const bf16* __restrict__ x, long long x_ld,
                                                              const bf16* __restrict__ dy, long long dy_ld,
                                                              const bf16* __restrict__ mask, long long m_ld,
                                                              size_t npix, int C, float* __restrict__ partial,
                                                              float mask_hi, const float* __restrict__ mask_ss) {
  dp::pdl_prologue();
  extern __shared__ float sred[];  // [2][lanes][C]
  const int C8 = C / 8;
  const int lanes = RED_TPB / C8;  // pixel lanes per block (C8 <= 64 -> lanes >= 4); threads beyond lanes*C8 idle
  const int c8 = threadIdx.x % C8, pl = threadIdx.x / C8;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
  if (pl < lanes) {
    float msc[8], msh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      msc[j] = (MODE == 2 && mask_ss) ? __ldg(mask_ss + c8 * 8 + j) : 0.f;
      msh[j] = (MODE == 2 && mask_ss) ? __ldg(mask_ss + C + c8 * 8 + j) : 0.f;
    }
    const size_t pstep = (size_t)gridDim.x * lanes;
    // four pixels per iteration: all loads are issued before any is consumed (bytes in flight, not occupancy, carry
    // the bandwidth of this reduction; the grid stays at two blocks per SM so the partials stay few)
    constexpr int U = 4;
    for (size_t p = (size_t)blockIdx.x * lanes + pl; p < npix; p += U * pstep) {
      size_t pp[U];
#pragma unroll
      for (int u = 0; u < U; ++u) pp[u] = p + u * pstep;
      uint4 rx[U], rg[U], rm[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = pp[u] < npix;
        rx[u] = (ok && (MODE != 2 || x)) ? ld8(x + pp[u] * x_ld + c8 * 8) : make_uint4(0, 0, 0, 0);
        rg[u] = (ok && MODE == 2) ? ld8(dy + pp[u] * dy_ld + c8 * 8) : make_uint4(0, 0, 0, 0);
        rm[u] = (ok && MODE == 2 && mask) ? ld8(mask + pp[u] * m_ld + c8 * 8) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (pp[u] >= npix) continue;
        float xv[8], g[8];
        unpack8(rx[u], xv);
        if (MODE == 2) {
          unpack8(rg[u], g);
          if (mask_ss) {
            // activation recomputed from the pre-BN tensor (already being read) instead of loading the activated
            // output; with mask_ss set, `mask` (optional) is the residual that was added before the activation
            float radd[8];
            unpack8(rm[u], radd);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // fp32 pre-activation, not its bf16 rounding: values just below 6 that round up to 6.0 keep their
              // gradient, as in the reference's fp32 hardtanh backward
              const float m = xv[j] * msc[j] + msh[j] + radd[j];
              g[j] = (m > 0.f && m < mask_hi) ? g[j] : 0.f;
            }
          } else if (mask) {
            float m[8];
            unpack8(rm[u], m);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = (m[j] > 0.f && m[j] < mask_hi) ? g[j] : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (MODE == 0) s0[j] += xv[j];
          if (MODE == 1) { s0[j] += xv[j]; s1[j] += xv[j] * xv[j]; }
          if (MODE == 2) { s0[j] += g[j]; s1[j] += g[j] * xv[j]; }
        }
      }
    }
  }
  if (pl < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sred[(0 * lanes + pl) * C + c8 * 8 + j] = s0[j];
      if (MODE != 0) sred[(1 * lanes + pl) * C + c8 * 8 + j] = s1[j];
    }
  }
  __syncthreads();
  const int nout = (MODE == 0 ? 1 : 2) * C;
  for (int i = threadIdx.x; i < nout; i += RED_TPB) {
    const int which = i / C, c = i - which * C;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += sred[(which * lanes + l) * C + c];
    partial[((size_t)blockIdx.x * 2 + which) * C + c] = s;
  }
}

// BatchNorm2d finalize (train): partial[nparts][2][C] (sum, sumsq over `count` values per channel) ->
//   scale_shift[0][C] = gamma*invstd, [1][C] = beta - mean*gamma*invstd, save[0][C]=mean, save[1][C]=invstd;
//   running stats updated with momentum (unbiased variance), num_batches_tracked += 1.
constexpr int kPartLanes = 32;   // partial-sum lanes per channel in the finalize / fold kernels (block = 32 x kPartLanes)
__global__ void __launch_bounds__(32 * kPartLanes) bn_finalize_kernel(const float* __restrict__ partial, int nparts, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches, float* __restrict__ scale_shift,
                                   float* __restrict__ save) {
  dp::pdl_prologue();
  // block = 32 channels x kPartLanes partial lanes: coalesced 128-byte rows of the partials, fixed-order fp64 folding
  __shared__ double s_s[kPartLanes][32], s_q[kPartLanes][32];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches) *num_batches += 1;
  double s = 0.0, q = 0.0;
  if (c < C) {
    for (int p = pl; p < nparts; p += kPartLanes) {
      s += (double)partial[((size_t)p * 2 + 0) * C + c];
      q += (double)partial[((size_t)p * 2 + 1) * C + c];
    }
  }
  s_s[pl][cl] = s;
  s_q[pl][cl] = q;
  __syncthreads();
  if (pl != 0 || c >= C) return;
  s = 0.0; q = 0.0;
#pragma unroll
  for (int l = 0; l < kPartLanes; ++l) { s += s_s[l][cl]; q += s_q[l][cl]; }
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale_shift[c] = g * invstd;
  scale_shift[C + c] = b - (float)mean * g * invstd;
  save[c] = (float)mean;
  save[C + c] = invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval mode: scale/shift from running statistics
__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                      float* __restrict__ scale_shift, float* __restrict__ save) {
  dp::pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.f / sqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale_shift[c] = g * invstd;
  scale_shift[C + c] = b - rm[c] * g * invstd;
  save[c] = rm[c];
  save[C + c] = invstd;
}

// y = act( x*scale + shift  [+ x2*scale2 + shift2 | + res] )
// The grid-stride is rounded down to a multiple of C/8, so every thread stays on one 8-channel group and keeps its
// scale / shift vectors in registers for its whole pixel walk.
__global__ void __launch_bounds__(256) bn_apply_kernel(const bf16* __restrict__ x, long long x_ld, const float* __restrict__ ss,
                                const bf16* __restrict__ x2, long long x2_ld, const float* __restrict__ ss2,
                                const bf16* __restrict__ res, long long res_ld, size_t npix, int C, int relu,
                                bf16* __restrict__ y, long long y_ld) {
  dp::pdl_prologue();
  const int C8 = C / 8;
  const size_t total = npix * C8;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  const size_t stride = (nthreads / C8) * C8;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= stride) return;
  const int c8 = (int)(tid % C8);
  float sc[8], sh[8], sc2[8], sh2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(ss + c8 * 8 + j);
    sh[j] = __ldg(ss + C + c8 * 8 + j);
    sc2[j] = x2 ? __ldg(ss2 + c8 * 8 + j) : 0.f;
    sh2[j] = x2 ? __ldg(ss2 + C + c8 * 8 + j) : 0.f;
  }
  const size_t pstep = stride / C8;
  constexpr int U = 1;                  // pixels per iteration (measured at 448x576x64: U = 1 5.1 TB/s, U = 4 4.8 TB/s -
                                        // at 57 registers occupancy already carries the bytes in flight)
  for (size_t p0 = tid / C8; p0 < npix; p0 += U * pstep) {
    uint4 rx[U], r2[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t p = p0 + u * pstep;
      const bool ok = p < npix;
      rx[u] = ok ? ld8(x + p * x_ld + c8 * 8) : make_uint4(0, 0, 0, 0);
      r2[u] = (ok && x2) ? ld8(x2 + p * x2_ld + c8 * 8) : make_uint4(0, 0, 0, 0);
      rr[u] = (ok && res) ? ld8(res + p * res_ld + c8 * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t p = p0 + u * pstep;
      if (p >= npix) continue;
      float v[8], o[8];
      unpack8(rx[u], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[j] * sc[j] + sh[j];
      if (x2) {
        unpack8(r2[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += v[j] * sc2[j] + sh2[j];
      }
      if (res) {
        unpack8(rr[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += v[j];
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
      }
      if (relu == 2) {   // ReLU6
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fminf(o[j], 6.f);
      }
      st8(y + p * y_ld + c8 * 8, pack8(o));
    }
  }
  (void)total;
}

// BN backward apply: g = dy*(mask>0);  train: dx = gamma*invstd*(g - sum_g/N - xhat*sum_gxhat/N);  eval: gamma*invstd*g
// red[0][C] = sum g, red[1][C] = sum g*x (raw x): sum_gxhat = invstd*(red1 - mean*red0).
// Also writes dgamma = sum_gxhat, dbeta = sum g (by block 0), and optionally gmask = g (the masked upstream grad,
// which is also the gradient of an identity shortcut / residual).
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const bf16* __restrict__ dy, long long dy_ld, const bf16* __restrict__ mask,
                                    long long m_ld, const bf16* __restrict__ x, long long x_ld,
                                    const float* __restrict__ red, const float* __restrict__ save,
                                    const float* __restrict__ gamma, double count, int train, size_t npix, int C,
                                    bf16* __restrict__ dx, long long dx_ld, bf16* __restrict__ gmask, long long gm_ld,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                                    float mask_hi, const float* __restrict__ mask_ss) {
  dp::pdl_prologue();
  const int C8 = C / 8;
  if (blockIdx.x == 0 && dgamma) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float mean = save[c], invstd = save[C + c];
      const float dg = invstd * (red[C + c] - mean * red[c]);
      dgamma[c] = accumulate ? dgamma[c] + dg : dg;
      dbeta[c] = accumulate ? dbeta[c] + red[c] : red[c];
    }
  }
  // every thread stays on one 8-channel group (stride rounded to a multiple of C/8) and folds the per-channel
  // statistics into three coefficients once:  dx = ca*g + cb*x + cd
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  const size_t stride = (nthreads / C8) * C8;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= stride) return;
  const int c8 = (int)(tid % C8);
  const float invN = (float)(1.0 / count);
  float ca[8], cb[8], cd[8], msc[8], msh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c8 * 8 + j;
    ca[j] = cb[j] = cd[j] = 0.f;
    if (dx) {
      const float mean = __ldg(save + c), invstd = __ldg(save + C + c);
      const float gm = gamma ? __ldg(gamma + c) : 1.f;
      ca[j] = gm * invstd;
      if (train) {
        const float sg = __ldg(red + c), sgx = __ldg(red + C + c);
        const float sgxh = invstd * (sgx - mean * sg);
        // gm*invstd*(g - sg/N - (x-mean)*invstd*sgxh/N)
        cb[j] = -gm * invstd * invstd * sgxh * invN;
        cd[j] = -gm * invstd * sg * invN - cb[j] * mean;
      }
    }
    msc[j] = mask_ss ? __ldg(mask_ss + c) : 0.f;
    msh[j] = mask_ss ? __ldg(mask_ss + C + c) : 0.f;
  }
  const size_t pstep = stride / C8;
  constexpr int U = 4;                  // pixels per iteration: all their loads first (bytes in flight carry this pass)
  for (size_t p0 = tid / C8; p0 < npix; p0 += U * pstep) {
    size_t pp[U];
#pragma unroll
    for (int u = 0; u < U; ++u) pp[u] = p0 + u * pstep;
    uint4 rg[U], rx[U], rm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = pp[u] < npix;
      rg[u] = ok ? ld8(dy + pp[u] * dy_ld + c8 * 8) : make_uint4(0, 0, 0, 0);
      rx[u] = (ok && (dx || mask_ss)) ? ld8(x + pp[u] * x_ld + c8 * 8) : make_uint4(0, 0, 0, 0);
      rm[u] = (ok && mask) ? ld8(mask + pp[u] * m_ld + c8 * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t p = pp[u];
      if (p >= npix) continue;
      float g[8], xv[8], o[8];
      unpack8(rg[u], g);
      unpack8(rx[u], xv);
      if (mask_ss) {
        float radd[8];
        unpack8(rm[u], radd);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float m = xv[j] * msc[j] + msh[j] + radd[j];
          g[j] = (m > 0.f && m < mask_hi) ? g[j] : 0.f;
        }
      } else if (mask) {
        float m[8];
        unpack8(rm[u], m);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = (m[j] > 0.f && m[j] < mask_hi) ? g[j] : 0.f;
      }
      if (gmask) st8(gmask + p * gm_ld + c8 * 8, pack8(g));
      if (dx) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(ca[j], g[j], fmaf(cb[j], xv[j], cd[j]));
        st8(dx + p * dx_ld + c8 * 8, pack8(o));
      }
    }
  }
}

__global__ void __launch_bounds__(32 * kPartLanes) sum_partials_kernel(const float* __restrict__ partial, int nparts, int rows, int C,
                                    float* __restrict__ out, int accumulate) {
  dp::pdl_prologue();
  // block = 32 columns x kPartLanes partial lanes over the flattened [rows*C] vector; partial stride is 2*C per part
  __shared__ double s_s[kPartLanes][32];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cl;
  double s = 0.0;
  if (i < rows * C) {
    const int which = i / C, c = i - which * C;
    for (int p = pl; p < nparts; p += kPartLanes) s += (double)partial[((size_t)p * 2 + which) * C + c];
  }
  s_s[pl][cl] = s;
  __syncthreads();
  if (pl != 0 || i >= rows * C) return;
  s = 0.0;
#pragma unroll
  for (int l = 0; l < kPartLanes; ++l) s += s_s[l][cl];
  out[i] = accumulate ? out[i] + (float)s : (float)s;
}

constexpr int kRedBlocks = 2 * kNumSMs;

}  // namespace

extern "C" {

int dp_add_relu_bwd(const void* g_raw, const void* g_relu, const void* y, void* out, size_t n, cudaStream_t stream) {
  DP_CHECK_ARG(out && (g_raw || g_relu) && (!g_relu || y) && n % 8 == 0, "dp_add_relu_bwd: bad arguments");
  if (n == 0) return DP_OK;
  dp::pdl_work(8.0 * (double)n);
  dp::launch(add_relu_bwd_kernel, grid_for(n / 8), 256, 0, stream, (const bf16*)g_raw, (const bf16*)g_relu, (const bf16*)y,
                                                           (bf16*)out, n / 8);
  DP_CHECK_LAUNCH("add_relu_bwd_kernel");
  return DP_OK;
}

int dp_relu_bf16(const void* x, void* out, size_t n, cudaStream_t stream) {
  DP_CHECK_ARG(x && out && n % 8 == 0, "dp_relu_bf16: bad arguments");
  if (n == 0) return DP_OK;
  dp::pdl_work(4.0 * (double)n);
  dp::launch(relu_kernel, grid_for(n / 8), 256, 0, stream, (const bf16*)x, (bf16*)out, n / 8);
  DP_CHECK_LAUNCH("relu_kernel");
  return DP_OK;
}

int dp_add_bf16(const void* a, const void* b, const void* c, void* out, size_t n, cudaStream_t stream) {
  DP_CHECK_ARG(a && b && out && n % 8 == 0, "dp_add_bf16: bad arguments");
  if (n == 0) return DP_OK;
  dp::pdl_work(8.0 * (double)n);
  dp::launch(add_kernel, grid_for(n / 8), 256, 0, stream, (const bf16*)a, (const bf16*)b, (const bf16*)c, (bf16*)out, n / 8);
  DP_CHECK_LAUNCH("add_kernel");
  return DP_OK;
}

int dp_copy_channels(const void* src, long long src_ld, void* dst, long long dst_ld, size_t npix, int C,
                     cudaStream_t stream) {
  DP_CHECK_ARG(src && dst && C % 8 == 0 && src_ld % 8 == 0 && dst_ld % 8 == 0, "dp_copy_channels: bad arguments");
  if (npix == 0) return DP_OK;
  const size_t items = npix * (C / 8);
  dp::pdl_work(4.0 * (double)npix * C);
  if (items < (size_t)0x7fffff00u)
    dp::launch(copy_channels_kernel<unsigned>, grid_for(items), 256, 0, stream, (const bf16*)src, src_ld, (bf16*)dst, dst_ld, npix, C / 8);
  else
    dp::launch(copy_channels_kernel<size_t>, grid_for(items), 256, 0, stream, (const bf16*)src, src_ld, (bf16*)dst, dst_ld, npix, C / 8);
  DP_CHECK_LAUNCH("copy_channels_kernel");
  return DP_OK;
}

int dp_resize_bilinear_nhwc(const void* src, long long src_ld, int B, int Hi, int Wi, int C, void* dst,
                            long long dst_ld, int Ho, int Wo, int align_corners, cudaStream_t stream) {
  DP_CHECK_ARG(src && dst && C % 8 == 0 && src_ld % 8 == 0 && dst_ld % 8 == 0, "dp_resize_bilinear_nhwc: bad arguments");
  DP_CHECK_ARG(B > 0 && B <= 65535 && Ho > 0 && Ho <= 65535, "dp_resize_bilinear_nhwc: B / Ho out of the grid range");
  const dim3 grid((unsigned)(((size_t)Wo * (C / 8) + 255) / 256), (unsigned)((Ho + kRsRows - 1) / kRsRows), (unsigned)B);
  dp::pdl_work(2.0 * (double)B * C * ((double)Hi * Wi + (double)Ho * Wo));
  dp::launch(resize_fwd_kernel, grid, 256, 0, stream, (const bf16*)src, src_ld, B, Hi, Wi, C, (bf16*)dst, dst_ld, Ho, Wo,
                                              align_corners, resize_scale(Hi, Ho, align_corners),
                                              resize_scale(Wi, Wo, align_corners));
  DP_CHECK_LAUNCH("resize_fwd_kernel");
  return DP_OK;
}

int dp_resize_bilinear_nhwc_bwd(const void* gout, long long g_ld, int B, int Hi, int Wi, int C, void* gin,
                                long long gin_ld, int Ho, int Wo, int align_corners, cudaStream_t stream) {
  DP_CHECK_ARG(gout && gin && C % 8 == 0 && g_ld % 8 == 0 && gin_ld % 8 == 0, "dp_resize_bilinear_nhwc_bwd: bad arguments");
  DP_CHECK_ARG(B > 0 && B <= 65535 && Hi > 0 && Hi <= 65535, "dp_resize_bilinear_nhwc_bwd: B / Hi out of the grid range");
  const dim3 grid((unsigned)(((size_t)Wi * (C / 8) + 255) / 256), (unsigned)Hi, (unsigned)B);
  const float sy = resize_scale(Hi, Ho, align_corners), sx = resize_scale(Wi, Wo, align_corners);
  dp::pdl_work(2.0 * (double)B * C * ((double)Hi * Wi + (double)Ho * Wo));
  dp::launch(resize_bwd_kernel, grid, 256, 0, stream, (const bf16*)gout, g_ld, B, Hi, Wi, C, (bf16*)gin, gin_ld, Ho, Wo,
                                              align_corners, sy, sx, sy > 0.f ? 1.f / sy : 0.f, sx > 0.f ? 1.f / sx : 0.f);
  DP_CHECK_LAUNCH("resize_bwd_kernel");
  return DP_OK;
}

int dp_resize_bilinear_planes_f32(const float* src, int planes, int Hi, int Wi, float* dst, int Ho, int Wo,
                                  int align_corners, cudaStream_t stream) {
  DP_CHECK_ARG(src && dst && planes > 0, "dp_resize_bilinear_planes_f32: bad arguments");
  const size_t items = (size_t)planes * Ho * Wo;
  dp::launch(resize_planes_f32_kernel, grid_for(items), 256, 0, stream, src, planes, Hi, Wi, dst, Ho, Wo, align_corners);
  DP_CHECK_LAUNCH("resize_planes_f32_kernel");
  return DP_OK;
}

int dp_chan_reduce_blocks(void) { return kRedBlocks; }

/* mode 0: sum x; 1: sum x, sum x^2; 2: sum g, sum g*x with g = dy*(mask>0); 3: as 2 with the ReLU6 mask 0<mask<6.
 * partial: float[dp_chan_reduce_blocks()][2][C] */
int dp_chan_reduce(int mode, const void* x, long long x_ld, const void* dy, long long dy_ld, const void* mask,
                   long long m_ld, const float* mask_ss, size_t npix, int C, float* partial, cudaStream_t stream) {
  DP_CHECK_ARG(partial && C % 8 == 0 && C <= 2048 && mode >= 0 && mode <= 3, "dp_chan_reduce: bad arguments");
  DP_CHECK_ARG(mode >= 2 ? (dy != nullptr && x != nullptr) : (x != nullptr), "dp_chan_reduce: null input");
  const float mask_hi = mode == 3 ? 6.f : INFINITY;
  const int C8 = C / 8;
  const int lanes = RED_TPB / C8;
  const size_t smem = (size_t)2 * lanes * C * sizeof(float);
  dp::pdl_work((mode >= 2 ? 4.0 : 2.0) * (double)npix * C);
  if (mode == 0)
    dp::launch(chan_reduce_kernel<0>, kRedBlocks, RED_TPB, smem, stream, (const bf16*)x, x_ld, nullptr, 0, nullptr, 0, npix, C, partial, mask_hi, nullptr);
  else if (mode == 1)
    dp::launch(chan_reduce_kernel<1>, kRedBlocks, RED_TPB, smem, stream, (const bf16*)x, x_ld, nullptr, 0, nullptr, 0, npix, C, partial, mask_hi, nullptr);
  else
    dp::launch(chan_reduce_kernel<2>, kRedBlocks, RED_TPB, smem, stream, (const bf16*)x, x_ld, (const bf16*)dy, dy_ld,
                                                                 (const bf16*)mask, m_ld, npix, C, partial, mask_hi, mask_ss);
  DP_CHECK_LAUNCH("chan_reduce_kernel");
  return DP_OK;
}

int dp_sum_partials(const float* partial, int nparts, int rows, int C, float* out, int accumulate, cudaStream_t stream) {
  DP_CHECK_ARG(partial && out && rows >= 1 && rows <= 2, "dp_sum_partials: bad arguments");
  dp::launch(sum_partials_kernel, dp::ceil_div(rows * C, 32), 32 * kPartLanes, 0, stream, partial, nparts, rows, C, out, accumulate);
  DP_CHECK_LAUNCH("sum_partials_kernel");
  return DP_OK;
}

int dp_bn_finalize(const float* partial, int nparts, int C, double count, const float* gamma, const float* beta,
                   float eps, float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                   float* scale_shift, float* save_mean_invstd, cudaStream_t stream) {
  DP_CHECK_ARG(partial && scale_shift && save_mean_invstd && C > 0 && count > 0, "dp_bn_finalize: bad arguments");
  dp::launch(bn_finalize_kernel, dp::ceil_div(C, 32), 32 * kPartLanes, 0, stream, partial, nparts, C, count, gamma, beta, eps, momentum,
                                                               running_mean, running_var, num_batches_tracked,
                                                               scale_shift, save_mean_invstd);
  DP_CHECK_LAUNCH("bn_finalize_kernel");
  return DP_OK;
}

int dp_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, int C, float* scale_shift, float* save_mean_invstd, cudaStream_t stream) {
  DP_CHECK_ARG(running_mean && running_var && scale_shift && save_mean_invstd, "dp_bn_eval_coeffs: null pointer");
  dp::launch(bn_eval_coeffs_kernel, dp::ceil_div(C, 128), 128, 0, stream, gamma, beta, running_mean, running_var, eps, C,
                                                                  scale_shift, save_mean_invstd);
  DP_CHECK_LAUNCH("bn_eval_coeffs_kernel");
  return DP_OK;
}

int dp_bn_apply(const void* x, long long x_ld, const float* scale_shift, const void* x2, long long x2_ld,
                const float* scale_shift2, const void* res, long long res_ld, size_t npix, int C, int relu, void* y,
                long long y_ld, cudaStream_t stream) {
  DP_CHECK_ARG(x && scale_shift && y && C % 8 == 0 && (!x2 || scale_shift2), "dp_bn_apply: bad arguments");
  dp::pdl_work((x2 || res ? 6.0 : 4.0) * (double)npix * C);
  dp::launch(bn_apply_kernel, grid_for(npix * (C / 8)), 256, 0, stream, (const bf16*)x, x_ld, scale_shift, (const bf16*)x2, x2_ld,
                                                                scale_shift2, (const bf16*)res, res_ld, npix, C, relu,
                                                                (bf16*)y, y_ld);
  DP_CHECK_LAUNCH("bn_apply_kernel");
  return DP_OK;
}

int dp_bn_bwd_apply(const void* dy, long long dy_ld, const void* mask, long long m_ld, const float* mask_ss,
                    const void* x, long long x_ld, const float* red, const float* save_mean_invstd, const float* gamma, double count, int train,
                    size_t npix, int C, void* dx, long long dx_ld, void* gmask, long long gm_ld, float* dgamma,
                    float* dbeta, int accumulate, cudaStream_t stream) {
  DP_CHECK_ARG(dy && C % 8 == 0 && (dx || gmask) && (!mask_ss || x), "dp_bn_bwd_apply: bad arguments");
  DP_CHECK_ARG(!dx || (x && save_mean_invstd && (!train || red)), "dp_bn_bwd_apply: missing statistics");
  DP_CHECK_ARG(!dgamma || (dbeta && red && save_mean_invstd), "dp_bn_bwd_apply: dgamma needs dbeta, red and save");
  // train: bit 0 = batch statistics (train mode); bit 1 = the mask is a ReLU6 output (gradient passes for 0 < mask < 6)
  dp::pdl_work(6.0 * (double)npix * C);
  dp::launch(bn_bwd_apply_kernel, grid_for(npix * (C / 8)), 256, 0, stream, 
      (const bf16*)dy, dy_ld, (const bf16*)mask, m_ld, (const bf16*)x, x_ld, red, save_mean_invstd, gamma, count, train & 1,
      npix, C, (bf16*)dx, dx_ld, (bf16*)gmask, gm_ld, dgamma, dbeta, accumulate, (train & 2) ? 6.f : INFINITY, mask_ss);
  DP_CHECK_LAUNCH("bn_bwd_apply_kernel");
  return DP_OK;
}

}  // extern "C"
