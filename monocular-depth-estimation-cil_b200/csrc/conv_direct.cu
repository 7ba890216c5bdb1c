// CUDA-core direct convolutions for the few layers the tcgen05 implicit GEMM does not cover:
//   * stride-2 3x3 convs and k4/s2/p1, k4/s4, k2/s2 transposed convs (CrossAttention.spatial_reduction /
//     spatial_upsample, midas_semantics.py:38-61; Dinov2Head.resize_layers, dpt_depth.py:49-69), forward,
//     data gradient and weight gradient;
//   * the 16->1 (or C->1) 3x3 depth head with bias + ReLU (midas_semantics.py:203-204) and the
//     1x1 32->1 head of MidasNet_small / DPT (midas_net_custom.py:110-111, dpt_depth.py:282-283).
// One gather-form kernel serves conv forward, transposed-conv forward and both data gradients:
//   out[b,oy,ox,co] = bias[co] + sum_{ky,kx,ci} in[b,iy,ix,ci] * Wg[ky*KW+kx][co][ci]
//   conv rule:        iy = oy*stride - pad + ky
//   transposed rule:  iy = (oy + pad - ky) / stride   when divisible
#include "common.cuh"
#include "../../include/depth_b200.h"

namespace {

using namespace dp;

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

struct GatherArgs {
  const bf16* in; long long in_ld; int Hi, Wi, Ci;
  const bf16* w;   // [KH*KW][Co][Ci]
  const float* bias;
  bf16* out; long long out_ld; int B, Ho, Wo, Co;
  int KH, KW, stride, pad, transposed, relu;
};

// thread = one output pixel x 8 output channels
__global__ void __launch_bounds__(256) conv_gather_kernel(GatherArgs a) {
  const int G = a.Co / 8;
  const size_t total = (size_t)a.B * a.Ho * a.Wo * G;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    size_t p = i / G;
    const int ox = (int)(p % a.Wo); p /= a.Wo;
    const int oy = (int)(p % a.Ho);
    const int b = (int)(p / a.Ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = a.bias ? __ldg(a.bias + g * 8 + j) : 0.f;
    for (int ky = 0; ky < a.KH; ++ky) {
      int iy;
      if (a.transposed) {
        const int t = oy + a.pad - ky;
        if (t < 0 || (t % a.stride) != 0) continue;
        iy = t / a.stride;
      } else {
        iy = oy * a.stride - a.pad + ky;
      }
      if (iy < 0 || iy >= a.Hi) continue;
      for (int kx = 0; kx < a.KW; ++kx) {
        int ix;
        if (a.transposed) {
          const int t = ox + a.pad - kx;
          if (t < 0 || (t % a.stride) != 0) continue;
          ix = t / a.stride;
        } else {
          ix = ox * a.stride - a.pad + kx;
        }
        if (ix < 0 || ix >= a.Wi) continue;
        const bf16* xin = a.in + (((size_t)b * a.Hi + iy) * a.Wi + ix) * a.in_ld;
        const bf16* wt = a.w + ((size_t)(ky * a.KW + kx) * a.Co + g * 8) * a.Ci;
        for (int c = 0; c < a.Ci; c += 8) {
          float xv[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(xin + c)), xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float wv[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(wt + (size_t)j * a.Ci + c)), wv);
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) s += xv[q] * wv[q];
            acc[j] += s;
          }
        }
      }
    }
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
    *reinterpret_cast<uint4*>(a.out + (((size_t)b * a.Ho + oy) * a.Wo + ox) * a.out_ld + g * 8) = pack8(acc);
  }
}

// out[tap][cp][ct] (fp32 partials per pixel chunk) = sum over plain-grid pixels (b,y,x) of
//     P[b,y,x,cp] * T[b, y*stride-pad+ky, x*stride-pad+kx, ct]
// block: 32x32 tile of (cp, ct) for one tap and one pixel chunk; 256 threads, 4 ct per thread.
struct WgDirectArgs {
  const bf16* P; long long p_ld; int Hp, Wp, Cp;
  const bf16* T; long long t_ld; int Ht, Wt, Ct;
  int B, KH, KW, stride, pad, chunks;
  float* partial;  // [chunks][KH*KW][Cp][Ct]
};

__global__ void __launch_bounds__(256) conv_wgrad_direct_kernel(WgDirectArgs a) {
  __shared__ float sP[32][33];
  __shared__ float sT[32][33];
  const int tiles_t = ceil_div(a.Ct, 32);
  const int cp0 = (blockIdx.y / tiles_t) * 32, ct0 = (blockIdx.y % tiles_t) * 32;
  const int tap = blockIdx.z % (a.KH * a.KW), chunk = blockIdx.z / (a.KH * a.KW);
  const int ky = tap / a.KW, kx = tap % a.KW;
  const size_t npix = (size_t)a.B * a.Hp * a.Wp;
  const size_t per = (npix + a.chunks - 1) / a.chunks;
  const size_t p_begin = (size_t)chunk * per, p_end = min(npix, p_begin + per);
  const int cp = threadIdx.x / 8, ctg = (threadIdx.x % 8) * 4;  // this thread: row cp, columns ctg..ctg+3
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (size_t base = p_begin; base < p_end; base += 32) {
    // stage 32 pixels x 32 channels of each operand
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int pi = e / 32, c = e % 32;
      const size_t p = base + pi;
      float pv = 0.f, tv = 0.f;
      if (p < p_end) {
        const int x = (int)(p % a.Wp);
        const int y = (int)((p / a.Wp) % a.Hp);
        const int b = (int)(p / ((size_t)a.Wp * a.Hp));
        if (cp0 + c < a.Cp) pv = __bfloat162float(a.P[p * a.p_ld + cp0 + c]);
        const int ty = y * a.stride - a.pad + ky, tx = x * a.stride - a.pad + kx;
        if (ty >= 0 && ty < a.Ht && tx >= 0 && tx < a.Wt && ct0 + c < a.Ct)
          tv = __bfloat162float(a.T[(((size_t)b * a.Ht + ty) * a.Wt + tx) * a.t_ld + ct0 + c]);
      }
      sP[pi][c] = pv;
      sT[pi][c] = tv;
    }
    __syncthreads();
#pragma unroll 8
    for (int pi = 0; pi < 32; ++pi) {
      const float pv = sP[pi][cp];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += pv * sT[pi][ctg + j];
    }
    __syncthreads();
  }
  if (cp0 + cp < a.Cp) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (ct0 + ctg + j < a.Ct)
        a.partial[(((size_t)chunk * a.KH * a.KW + tap) * a.Cp + cp0 + cp) * a.Ct + ct0 + ctg + j] = acc[j];
  }
}

// out[i] (+)= sum over chunks; optional permutation to OIHW / IOHW: src index (tap, cp, ct) ->
// dst[(cp*Ct + ct)*taps + tap] when perm == 1, dst[(ct*Cp + cp)*taps + tap] when perm == 2, identity when 0
__global__ void wgrad_direct_reduce_kernel(const float* __restrict__ partial, int chunks, int taps, int Cp, int Ct,
                                           int perm, float* __restrict__ out, int accumulate) {
  const long long total = (long long)taps * Cp * Ct;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * total + i];
    const int ct = (int)(i % Ct);
    const int cp = (int)((i / Ct) % Cp);
    const int tap = (int)(i / ((long long)Ct * Cp));
    size_t o = (size_t)i;
    if (perm == 1) o = ((size_t)cp * Ct + ct) * taps + tap;
    if (perm == 2) o = ((size_t)ct * Cp + cp) * taps + tap;
    out[o] = accumulate ? out[o] + s : s;
  }
}

// ---- C -> 1 head convolution (3x3 pad 1 or 1x1), bias, optional ReLU; fp32 (B,H,W) output --------------------
__global__ void __launch_bounds__(256) head_conv_fwd_kernel(const bf16* __restrict__ x, long long x_ld, int B, int H,
                                                            int W, int C, int KS, const float* __restrict__ w /*[C][KS][KS]*/,
                                                            const float* __restrict__ bias, int relu,
                                                            float* __restrict__ out) {
  extern __shared__ float sw[];  // [KS*KS][C]
  for (int i = threadIdx.x; i < KS * KS * C; i += blockDim.x) {
    const int tap = i / C, c = i % C;
    sw[i] = w[(size_t)c * KS * KS + tap];
  }
  __syncthreads();
  const int pad = KS / 2;
  const size_t total = (size_t)B * H * W;
  const float b0 = bias ? __ldg(bias) : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    const int yy = (int)((i / W) % H);
    const int b = (int)(i / ((size_t)W * H));
    float acc = b0;
    for (int r = 0; r < KS; ++r) {
      const int iy = yy + r - pad;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < KS; ++s) {
        const int ix = xx + s - pad;
        if (ix < 0 || ix >= W) continue;
        const bf16* px = x + (((size_t)b * H + iy) * W + ix) * x_ld;
        const float* wt = sw + (r * KS + s) * C;
        for (int c = 0; c < C; c += 8) {
          float v[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(px + c)), v);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc += v[q] * wt[c + q];
        }
      }
    }
    out[i] = relu ? fmaxf(acc, 0.f) : acc;
  }
}

// data gradient: dx[b,y,x,c] = sum_taps g[b, y-(r-pad), x-(s-pad)] * w[c][r][s],  g = dout * (out > 0 if relu)
__global__ void __launch_bounds__(256) head_conv_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                              int relu, int B, int H, int W, int C, int KS,
                                                              const float* __restrict__ w, bf16* __restrict__ dx,
                                                              long long dx_ld) {
  extern __shared__ float sw[];  // [KS*KS][C]
  for (int i = threadIdx.x; i < KS * KS * C; i += blockDim.x) {
    const int tap = i / C, c = i % C;
    sw[i] = w[(size_t)c * KS * KS + tap];
  }
  __syncthreads();
  const int pad = KS / 2, C8 = C / 8;
  const size_t total = (size_t)B * H * W * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    size_t p = i / C8;
    const int xx = (int)(p % W);
    const int yy = (int)((p / W) % H);
    const int b = (int)(p / ((size_t)W * H));
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < KS; ++r) {
      const int oy = yy - (r - pad);
      if (oy < 0 || oy >= H) continue;
      for (int s = 0; s < KS; ++s) {
        const int ox = xx - (s - pad);
        if (ox < 0 || ox >= W) continue;
        const size_t o = ((size_t)b * H + oy) * W + ox;
        float g = __ldg(dout + o);
        if (relu && !(__ldg(out + o) > 0.f)) g = 0.f;
        const float* wt = sw + (r * KS + s) * C + c8 * 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += g * wt[q];
      }
    }
    *reinterpret_cast<uint4*>(dx + p * dx_ld + c8 * 8) = pack8(acc);
  }
}

// weight + bias gradient partials: part[block][KS*KS*C + 1].
// Thread = (pixel lane, tap, 8-channel group): it walks its lane's pixels, multiplies the (ReLU-masked) upstream
// gradient by the 8 input channels under its tap and keeps 8 fp32 accumulators; the pixel lanes are then folded
// through shared memory.  x is re-read once per tap from L1/L2 (9 x 32 B per pixel for C=16), g is a broadcast load.
__global__ void __launch_bounds__(256) head_conv_wgrad_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                              int relu, const bf16* __restrict__ x, long long x_ld, int B,
                                                              int H, int W, int C, int KS, float* __restrict__ part) {
  extern __shared__ float sacc[];  // [lanes][ncombo*8 + 1]
  const int pad = KS / 2, C8 = C / 8;
  const int ncombo = KS * KS * C8;            // (tap, channel group) pairs
  const int lanes = 256 / ncombo;             // pixel lanes per block (>= 1 since ncombo <= 72 is enforced by the host)
  const int combo = threadIdx.x % ncombo, pl = threadIdx.x / ncombo;
  const int tap = combo / C8, c8 = combo % C8;
  const int dy = tap / KS - pad, dx = tap % KS - pad;
  const int nacc = KS * KS * C + 1;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float bsum = 0.f;
  const size_t total = (size_t)B * H * W;
  if (pl < lanes) {
    for (size_t o = (size_t)blockIdx.x * lanes + pl; o < total; o += (size_t)gridDim.x * lanes) {
      float g = __ldg(dout + o);
      if (relu && !(__ldg(out + o) > 0.f)) g = 0.f;
      if (g == 0.f) continue;
      if (combo == 0) bsum += g;
      const int xx = (int)(o % W);
      const int yy = (int)((o / W) % H);
      const int iy = yy + dy, ix = xx + dx;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const size_t b = o / ((size_t)W * H);
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + ((b * H + iy) * W + ix) * x_ld + c8 * 8)), v);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += g * v[q];
    }
  }
  const int stride = ncombo * 8 + 1;
  if (pl < lanes) {
#pragma unroll
    for (int q = 0; q < 8; ++q) sacc[pl * stride + combo * 8 + q] = acc[q];
    if (combo == 0) sacc[pl * stride + ncombo * 8] = bsum;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
    float s = 0.f;
    if (i == nacc - 1) {
      for (int l = 0; l < lanes; ++l) s += sacc[l * stride + ncombo * 8];
    } else {
      const int t = i / C, c = i % C;                // output index order: tap-major, channel-minor
      const int src = (t * C8 + c / 8) * 8 + (c % 8);
      for (int l = 0; l < lanes; ++l) s += sacc[l * stride + src];
    }
    part[(size_t)blockIdx.x * nacc + i] = s;
  }
}

// Faster form for C in {8,16,32,64}: thread = (input pixel, 8-channel group) with the channel group fixed per thread, so the
// activation vector is loaded once (fully coalesced) and multiplied into all KS*KS taps; the masked upstream gradient of
// the KS*KS neighbours comes from L1.  Persistent blocks walk whole image rows (no per-pixel index divisions) and keep
// KS*KS*8 fp32 accumulators per thread; one fixed-order shuffle + shared-memory reduction per block at the end.
template <int KS>
__global__ void __launch_bounds__(256) head_conv_wgrad_rows_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                                   int relu, const bf16* __restrict__ x, long long x_ld,
                                                                   int B, int H, int W, int C, float* __restrict__ part) {
  constexpr int T = KS * KS, pad = KS / 2;
  __shared__ float s_red[8][8][T * 8 + 1];     // [warp][c8][tap*8 + j | bias]
  const int C8 = C / 8;                        // power of two <= 8 (host checked)
  const int c8 = threadIdx.x % C8;
  const int px0 = threadIdx.x / C8, pxs = 256 / C8;
  float acc[T][8];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  float bsum = 0.f;
  const int rows = B * H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int y = row % H;
    const size_t rbase = (size_t)row * W;
    for (int xx = px0; xx < W; xx += pxs) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (rbase + xx) * x_ld + c8 * 8)), v);
#pragma unroll
      for (int r = 0; r < KS; ++r) {
        const int gy = y - r + pad;
        if (gy < 0 || gy >= H) continue;
#pragma unroll
        for (int s2 = 0; s2 < KS; ++s2) {
          const int gx = xx - s2 + pad;
          if (gx < 0 || gx >= W) continue;
          const size_t o = rbase + (size_t)(gy - y) * W + gx;     // same image: rows of one image are contiguous
          float g = __ldg(dout + o);
          if (relu && !(__ldg(out + o) > 0.f)) g = 0.f;
          if (r == pad && s2 == pad && c8 == 0) bsum += g;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[r * KS + s2][j] = fmaf(g, v[j], acc[r * KS + s2][j]);
        }
      }
    }
  }
  // lanes with equal (lane % C8) hold the same channel group: fold them with xor shuffles down to lanes 0..C8-1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = acc[t][j];
      for (int o = 16; o >= C8; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      acc[t][j] = a;
    }
  for (int o = 16; o >= 1; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
  if (lane < C8) {
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) s_red[warp][lane][t * 8 + j] = acc[t][j];
    if (lane == 0) s_red[warp][0][T * 8] = bsum;
  }
  __syncthreads();
  const int nacc = T * C + 1;
  for (int i = threadIdx.x; i < nacc; i += 256) {
    float s = 0.f;
    if (i == nacc - 1) {
      for (int w = 0; w < 8; ++w) s += s_red[w][0][T * 8];
    } else {
      const int t = i / C, c = i % C;            // output index order: tap-major, channel-minor
      for (int w = 0; w < 8; ++w) s += s_red[w][c / 8][t * 8 + (c % 8)];
    }
    part[(size_t)blockIdx.x * nacc + i] = s;
  }
}

// ---- the depth head proper: 16 -> 1, 3x3, pad 1 (midas_semantics.py:203-204) at full resolution ---------------------
// The generic kernels above spend most of their time on 64-bit index divisions, a 144-long dependent FMA chain (forward)
// and scalar gradient loads behind bounds branches (backward): 370 / 410 / 490 us at 32 x 448 x 576 against HBM floors of
// 45 / 50 / 45 us.  These three keep 32-bit indices, packed f32x2 MACs on independent accumulators and unpack / load
// every operand once per thread.
__device__ __forceinline__ void unpack8f2(const uint4& u, float2 (&f)[4]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
}

// forward: thread = 4 vertically adjacent output pixels of one column.  Adjacent lanes sit on adjacent 32-byte pixels
// (a warp load touches 8 cache lines - a horizontal 4-pixel strip per thread put every lane on its own line and ran at the
// L1 tag rate), and the 6 x 3 input vectors under the strip are loaded and unpacked once for the four outputs (4.5
// vectors per output instead of 9); two independent packed accumulators per output, weights from shared memory.
__global__ void __launch_bounds__(256, 2) head16_fwd_kernel(const bf16* __restrict__ x, long long x_ld, int B, int H, int W,
                                                         const float* __restrict__ w /*[16][3][3]*/,
                                                         const float* __restrict__ bias, int relu, float* __restrict__ out) {
  __shared__ __align__(16) float sw[9 * 16];
  if (threadIdx.x < 144) sw[threadIdx.x] = w[(threadIdx.x % 16) * 9 + threadIdx.x / 16];
  __syncthreads();
  const int HG = (H + 3) >> 2;
  const int total = B * HG * W;
  const float b0 = bias ? __ldg(bias) : 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int xx = i % W, rg = i / W;
    const int yg = rg % HG, b = rg / HG;
    const int y0 = yg * 4;
    const long long img = (long long)b * H;
    float2 acc[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) { acc[t][0] = make_float2(0.f, 0.f); acc[t][1] = make_float2(0.f, 0.f); }
    // the six loads of input row rin + 1 are issued before row rin is consumed (the kernel is bound by load latency)
    auto load_row = [&](int rin, uint4 (&u)[3][2]) {
      const int iy = y0 - 1 + rin;
      const bool rok = iy >= 0 && iy < H;
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int ix = xx + s2 - 1;
        const bool ok = rok && ix >= 0 && ix < W;
        const uint4* pp = reinterpret_cast<const uint4*>(x + ((img + (ok ? iy : y0)) * W + (ok ? ix : xx)) * x_ld);
        u[s2][0] = __ldg(pp);
        u[s2][1] = __ldg(pp + 1);
      }
    };
    uint4 ua[3][2], ub[3][2];
    load_row(0, ua);
#pragma unroll
    for (int rin = 0; rin < 6; ++rin) {
      uint4 (&u)[3][2] = (rin & 1) ? ub : ua;
      if (rin + 1 < 6) load_row(rin + 1, (rin & 1) ? ua : ub);
      const int iy = y0 - 1 + rin;
      const bool rok = iy >= 0 && iy < H;
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int ix = xx + s2 - 1;
        if (!(rok && ix >= 0 && ix < W)) continue;             // zero padding: the tap contributes nothing
        float2 v[8];
        {
          float2 a2[4], b2[4];
          unpack8f2(u[s2][0], a2);
          unpack8f2(u[s2][1], b2);
#pragma unroll
          for (int q = 0; q < 4; ++q) { v[q] = a2[q]; v[4 + q] = b2[q]; }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int r = rin - t;                                // tap row of output t under input row rin (compile-time)
          if (r < 0 || r > 2) continue;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t4 = *reinterpret_cast<const float4*>(&sw[(r * 3 + s2) * 16 + q * 4]);
            acc[t][q & 1] = __ffma2_rn(v[2 * q], make_float2(t4.x, t4.y), acc[t][q & 1]);
            acc[t][q & 1] = __ffma2_rn(v[2 * q + 1], make_float2(t4.z, t4.w), acc[t][q & 1]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (y0 + t >= H) continue;
      const float2 st = __fadd2_rn(acc[t][0], acc[t][1]);
      const float rr = b0 + (st.x + st.y);
      out[((img + y0 + t) * W) + xx] = relu ? fmaxf(rr, 0.f) : rr;
    }
  }
}

// data gradient: thread = one pixel, all 16 channels; the nine masked upstream gradients are loaded first
__global__ void __launch_bounds__(256) head16_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                           int relu, int B, int H, int W, const float* __restrict__ w,
                                                           bf16* __restrict__ dx, long long dx_ld) {
  __shared__ __align__(16) float sw[9 * 16];
  if (threadIdx.x < 144) sw[threadIdx.x] = w[(threadIdx.x % 16) * 9 + threadIdx.x / 16];
  __syncthreads();
  const int total = B * H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int xx = i % W, row = i / W;
    const int yy = row % H;
    float g[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int oy = yy - (r - 1), ox = xx - (s2 - 1);
        const bool ok = oy >= 0 && oy < H && ox >= 0 && ox < W;
        const int o = ok ? i - (r - 1) * W - (s2 - 1) : i;     // clamped: the load is unconditional
        float gv = __ldg(dout + o);
        if (relu && !(__ldg(out + o) > 0.f)) gv = 0.f;
        g[r * 3 + s2] = ok ? gv : 0.f;
      }
    float2 acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float2 gg = make_float2(g[t], g[t]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t4 = *reinterpret_cast<const float4*>(&sw[t * 16 + q * 4]);
        acc[2 * q] = __ffma2_rn(gg, make_float2(t4.x, t4.y), acc[2 * q]);
        acc[2 * q + 1] = __ffma2_rn(gg, make_float2(t4.z, t4.w), acc[2 * q + 1]);
      }
    }
    uint4 o0, o1;
    {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[0].x, acc[0].y), h1 = __floats2bfloat162_rn(acc[1].x, acc[1].y);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2].x, acc[2].y), h3 = __floats2bfloat162_rn(acc[3].x, acc[3].y);
      o0 = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2),
                      *reinterpret_cast<uint32_t*>(&h3));
      __nv_bfloat162 h4 = __floats2bfloat162_rn(acc[4].x, acc[4].y), h5 = __floats2bfloat162_rn(acc[5].x, acc[5].y);
      __nv_bfloat162 h6 = __floats2bfloat162_rn(acc[6].x, acc[6].y), h7 = __floats2bfloat162_rn(acc[7].x, acc[7].y);
      o1 = make_uint4(*reinterpret_cast<uint32_t*>(&h4), *reinterpret_cast<uint32_t*>(&h5), *reinterpret_cast<uint32_t*>(&h6),
                      *reinterpret_cast<uint32_t*>(&h7));
    }
    uint4* dst = reinterpret_cast<uint4*>(dx + (long long)i * dx_ld);
    dst[0] = o0;
    dst[1] = o1;
  }
}

// weight + bias gradient partials: part[block][9*16 + 1].  A block owns a contiguous run of image rows and walks it in
// tiles of up to 8 rows of one image: the masked upstream gradient of the tile's rows and of the row above / below
// (zero outside the image) is staged in shared memory with zero borders, so a thread = (pixel column, 8-channel group)
// loads its activation vector once and reads its nine gradients from shared memory - two barriers per tile, the
// activation loads of a whole tile in flight; 9 x 8 fp32 accumulators as packed pairs; fixed-order shuffle +
// shared-memory fold at the end.
constexpr int H16_MAXW = 1024, H16_R = 8;
__global__ void __launch_bounds__(256, 2) head16_wgrad_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                           int relu, const bf16* __restrict__ x, long long x_ld, int B,
                                                           int H, int W, int rows_per_block, float* __restrict__ part) {
  extern __shared__ __align__(16) float h16_smem[];          // [H16_R + 2][W + 2] gradient rows | reduction scratch
  const int SW = W + 2;
  float* s_g = h16_smem;
  float (*s_red)[2][9 * 8 + 1] = reinterpret_cast<float (*)[2][9 * 8 + 1]>(h16_smem + (H16_R + 2) * SW + 2);
  const int c8 = threadIdx.x & 1, px0 = threadIdx.x >> 1;
  float2 acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[t][q] = make_float2(0.f, 0.f);
  float bsum = 0.f;
  const int rows = B * H;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int row0 = r0; row0 < r1;) {
    const int y0 = row0 % H;
    const int nr = min(min(H16_R, H - y0), r1 - row0);         // rows of this tile: one image, this block's run
    __syncthreads();                                           // the previous tile's reads are done
    for (int e0 = threadIdx.x; e0 < (nr + 2) * SW; e0 += 4 * 256) {   // slot k holds gradient row row0 - 1 + k
      float gd[4], go[4];
      bool live[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {                                   // four elements per thread, all loads first
        const int e = e0 + q * 256;
        const int k = e / SW, xs = e - k * SW;
        const bool valid = e < (nr + 2) * SW && ((k >= 1 && k <= nr) || (k == 0 && y0 > 0) || (k == nr + 1 && y0 + nr < H));
        live[q] = valid && xs >= 1 && xs <= W;
        const long long o = live[q] ? (long long)(row0 - 1 + k) * W + xs - 1 : (long long)row0 * W;
        gd[q] = __ldg(dout + o);
        go[q] = relu ? __ldg(out + o) : 1.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = e0 + q * 256;
        if (e < (nr + 2) * SW) s_g[e] = (live[q] && go[q] > 0.f) ? gd[q] : 0.f;
      }
    }
    __syncthreads();
    // this thread's items of the tile: (row rr, column px0 + 128 k); eight activation loads in flight at a time (the
    // kernel is bound by the latency of these loads, not by their bytes)
    const int nk = px0 < W ? (W - px0 + 127) / 128 : 0;
    const int nitems = nr * nk;
    for (int j0 = 0; j0 < nitems; j0 += 8) {
      uint4 u[8];
      int rrs[8], xxs[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j = min(j0 + q, nitems - 1);
        rrs[q] = j / nk;
        xxs[q] = px0 + (j - rrs[q] * nk) * 128;
        u[q] = __ldg(reinterpret_cast<const uint4*>(x + ((long long)(row0 + rrs[q]) * W + xxs[q]) * x_ld + c8 * 8));
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (j0 + q >= nitems) break;
        float2 v[4];
        unpack8f2(u[q], v);
        const float* gm = s_g + rrs[q] * SW + xxs[q];          // gradient row y - 1 (slot rr), column xx - 1 (+1 border)
        // dw[r][s] += g[y - (r-1)][x - (s-1)] * v : tap (r, s) pairs with gradient row y + 1 - r, column xx + 1 - s
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) {
            const float gv = gm[(2 - r) * SW + 2 - s2];
            if (r == 1 && s2 == 1 && c8 == 0) bsum += gv;
            const float2 gg = make_float2(gv, gv);
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[r * 3 + s2][t] = __ffma2_rn(gg, v[t], acc[r * 3 + s2][t]);
          }
      }
    }
    row0 += nr;
  }
  // lanes with equal parity hold the same channel group: fold them with xor shuffles down to lanes 0 / 1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float a = acc[t][q].x, b2 = acc[t][q].y;
      for (int o = 16; o >= 2; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b2 += __shfl_xor_sync(0xffffffffu, b2, o); }
      if (lane < 2) { s_red[warp][lane][t * 8 + 2 * q] = a; s_red[warp][lane][t * 8 + 2 * q + 1] = b2; }
    }
  for (int o = 16; o >= 1; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
  if (lane == 0) s_red[warp][0][72] = bsum;
  __syncthreads();
  for (int i = threadIdx.x; i < 145; i += 256) {
    float sum = 0.f;
    if (i == 144) {
      for (int wq = 0; wq < 8; ++wq) sum += s_red[wq][0][72];
    } else {
      const int t = i / 16, c = i % 16;                        // output index order: tap-major, channel-minor
      for (int wq = 0; wq < 8; ++wq) sum += s_red[wq][c / 8][t * 8 + (c % 8)];
    }
    part[(size_t)blockIdx.x * 145 + i] = sum;
  }
}

__global__ void head_conv_wgrad_reduce_kernel(const float* __restrict__ part, int nblocks, int C, int KS,
                                              float* __restrict__ dw /*[C][KS][KS]*/, float* __restrict__ db,
                                              int accumulate) {
  // one warp per output: lanes stride the per-block partials (fixed order), fp64 butterfly
  const int nacc = KS * KS * C + 1;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= nacc) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += (double)part[(size_t)b * nacc + i];
  s = dp::warp_sum(s);
  if (lane != 0) return;
  if (i == nacc - 1) {
    if (db) db[0] = accumulate ? db[0] + (float)s : (float)s;
  } else {
    const int tap = i / C, c = i % C;
    const size_t o = (size_t)c * KS * KS + tap;
    dw[o] = accumulate ? dw[o] + (float)s : (float)s;
  }
}

constexpr int kHeadWgBlocks = 2 * kNumSMs;

inline int grid_for(size_t items, int tpb = 256, int waves = 8) {
  size_t b = (items + tpb - 1) / tpb, cap = (size_t)waves * kNumSMs;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

inline int wg_chunks(size_t npix, int tiles, int taps) {
  long long want = (4LL * kNumSMs + (long long)tiles * taps - 1) / ((long long)tiles * taps);
  long long maxc = (long long)((npix + 255) / 256);
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  if (want > 512) want = 512;
  return (int)want;
}

}  // namespace

extern "C" {

int dp_conv_gather(const void* in, long long in_ld, int B, int Hi, int Wi, int Ci, const void* w_packed,
                   const float* bias, void* out, long long out_ld, int Ho, int Wo, int Co, int KH, int KW, int stride,
                   int pad, int transposed, int relu, cudaStream_t stream) {
  DP_CHECK_ARG(in && w_packed && out, "dp_conv_gather: null pointer");
  DP_CHECK_ARG(Ci % 8 == 0 && Co % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0, "dp_conv_gather: channels %% 8");
  DP_CHECK_ARG(stride >= 1 && KH >= 1 && KW >= 1, "dp_conv_gather: bad geometry");
  GatherArgs a;
  a.in = (const bf16*)in; a.in_ld = in_ld; a.Hi = Hi; a.Wi = Wi; a.Ci = Ci;
  a.w = (const bf16*)w_packed; a.bias = bias;
  a.out = (bf16*)out; a.out_ld = out_ld; a.B = B; a.Ho = Ho; a.Wo = Wo; a.Co = Co;
  a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad; a.transposed = transposed; a.relu = relu;
  conv_gather_kernel<<<grid_for((size_t)B * Ho * Wo * (Co / 8), 256, 16), 256, 0, stream>>>(a);
  DP_CHECK_LAUNCH("conv_gather_kernel");
  return DP_OK;
}

size_t dp_conv_wgrad_direct_workspace(int B, int Hp, int Wp, int Cp, int Ct, int KH, int KW) {
  const int tiles = dp::ceil_div(Cp, 32) * dp::ceil_div(Ct, 32);
  const int chunks = wg_chunks((size_t)B * Hp * Wp, tiles, KH * KW);
  return (size_t)chunks * KH * KW * Cp * Ct * sizeof(float);
}

/* out[(tap,cp,ct) permuted] = sum_pixels P[pix][cp] * T[pix*stride - pad + tap][ct].
 * perm 0: [tap][Cp][Ct]; 1: [Cp][Ct][tap] (OIHW when P = dY, T = X); 2: [Ct][Cp][tap] */
int dp_conv_wgrad_direct(const void* P, long long p_ld, int Hp, int Wp, int Cp, const void* T, long long t_ld, int Ht,
                         int Wt, int Ct, int B, int KH, int KW, int stride, int pad, int perm, float* out,
                         int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DP_CHECK_ARG(P && T && out && workspace, "dp_conv_wgrad_direct: null pointer");
  const int tiles = dp::ceil_div(Cp, 32) * dp::ceil_div(Ct, 32);
  WgDirectArgs a;
  a.P = (const bf16*)P; a.p_ld = p_ld; a.Hp = Hp; a.Wp = Wp; a.Cp = Cp;
  a.T = (const bf16*)T; a.t_ld = t_ld; a.Ht = Ht; a.Wt = Wt; a.Ct = Ct;
  a.B = B; a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad;
  a.chunks = wg_chunks((size_t)B * Hp * Wp, tiles, KH * KW);
  const size_t need = (size_t)a.chunks * KH * KW * Cp * Ct * sizeof(float);
  if (workspace_bytes < need) return dp_set_error(DP_ERR_WORKSPACE, "dp_conv_wgrad_direct: workspace %zu < %zu",
                                                  workspace_bytes, need);
  a.partial = (float*)workspace;
  dim3 grid(1, tiles, a.chunks * KH * KW);
  DP_CHECK_ARG(grid.z <= 65535, "dp_conv_wgrad_direct: grid too large");
  conv_wgrad_direct_kernel<<<grid, 256, 0, stream>>>(a);
  DP_CHECK_LAUNCH("conv_wgrad_direct_kernel");
  const long long total = (long long)KH * KW * Cp * Ct;
  wgrad_direct_reduce_kernel<<<grid_for((size_t)total), 256, 0, stream>>>(a.partial, a.chunks, KH * KW, Cp, Ct, perm, out,
                                                                          accumulate);
  DP_CHECK_LAUNCH("wgrad_direct_reduce_kernel");
  return DP_OK;
}

int dp_head_conv_fwd(const void* x, long long x_ld, int B, int H, int W, int C, int KS, const float* w,
                     const float* bias, int relu, float* out, cudaStream_t stream) {
  DP_CHECK_ARG(x && w && out && C % 8 == 0 && (KS == 1 || KS == 3), "dp_head_conv_fwd: bad arguments");
  if (C == 16 && KS == 3 && x_ld % 8 == 0 && (long long)B * H * W < (1LL << 31)) {
    head16_fwd_kernel<<<grid_for((size_t)B * ((H + 3) / 4) * W), 256, 0, stream>>>((const bf16*)x, x_ld, B, H, W, w, bias, relu, out);
    DP_CHECK_LAUNCH("head16_fwd_kernel");
    return DP_OK;
  }
  head_conv_fwd_kernel<<<grid_for((size_t)B * H * W), 256, KS * KS * C * sizeof(float), stream>>>(
      (const bf16*)x, x_ld, B, H, W, C, KS, w, bias, relu, out);
  DP_CHECK_LAUNCH("head_conv_fwd_kernel");
  return DP_OK;
}

size_t dp_head_conv_bwd_workspace(int C, int KS) { return (size_t)kHeadWgBlocks * (KS * KS * C + 1) * sizeof(float); }

int dp_head_conv_bwd(const float* dout, const float* out, int relu, const void* x, long long x_ld, int B, int H, int W,
                     int C, int KS, const float* w, void* dx, long long dx_ld, float* dw, float* db, int accumulate,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DP_CHECK_ARG(dout && x && w && workspace && C % 8 == 0 && (KS == 1 || KS == 3) && (!relu || out),
               "dp_head_conv_bwd: bad arguments");
  DP_CHECK_ARG(KS * KS * (C / 8) <= 72, "dp_head_conv_bwd: C too large for the head kernel");
  if (workspace_bytes < dp_head_conv_bwd_workspace(C, KS))
    return dp_set_error(DP_ERR_WORKSPACE, "dp_head_conv_bwd: workspace too small");
  const bool fast16 = C == 16 && KS == 3 && x_ld % 8 == 0 && dx_ld % 8 == 0 && W <= H16_MAXW &&
                      (long long)B * H * W < (1LL << 31);
  if (dx && fast16) {
    head16_dgrad_kernel<<<grid_for((size_t)B * H * W), 256, 0, stream>>>(dout, out, relu, B, H, W, w, (bf16*)dx, dx_ld);
    DP_CHECK_LAUNCH("head16_dgrad_kernel");
  } else if (dx) {
    head_conv_dgrad_kernel<<<grid_for((size_t)B * H * W * (C / 8)), 256, KS * KS * C * sizeof(float), stream>>>(
        dout, out, relu, B, H, W, C, KS, w, (bf16*)dx, dx_ld);
    DP_CHECK_LAUNCH("head_conv_dgrad_kernel");
  }
  if (dw && fast16) {
    const int rows = B * H;
    const int rpb = (rows + kHeadWgBlocks - 1) / kHeadWgBlocks;
    const int nblk = (rows + rpb - 1) / rpb;                        // <= kHeadWgBlocks: the partials buffer has room
    const size_t smem = ((size_t)(H16_R + 2) * (W + 2) + 2 + 8 * 2 * 73) * sizeof(float);
    head16_wgrad_kernel<<<nblk, 256, smem, stream>>>(dout, out, relu, (const bf16*)x, x_ld, B, H, W, rpb, (float*)workspace);
    DP_CHECK_LAUNCH("head16_wgrad_kernel");
    head_conv_wgrad_reduce_kernel<<<ceil_div((KS * KS * C + 1) * 32, 256), 256, 0, stream>>>((const float*)workspace, nblk, C, KS, dw, db, accumulate);
    DP_CHECK_LAUNCH("head_conv_wgrad_reduce_kernel");
    return DP_OK;
  }
  if (dw) {
    const int nacc = KS * KS * C + 1;
    const int ncombo = KS * KS * (C / 8);
    const size_t wg_smem = (size_t)(256 / ncombo) * (ncombo * 8 + 1) * sizeof(float);
    const int C8 = C / 8;
    if ((C8 == 1 || C8 == 2 || C8 == 4 || C8 == 8) && (KS == 3 ? C8 <= 4 : true)) {
      if (KS == 3)
        head_conv_wgrad_rows_kernel<3><<<kHeadWgBlocks, 256, 0, stream>>>(dout, out, relu, (const bf16*)x, x_ld, B, H, W, C,
                                                                           (float*)workspace);
      else
        head_conv_wgrad_rows_kernel<1><<<kHeadWgBlocks, 256, 0, stream>>>(dout, out, relu, (const bf16*)x, x_ld, B, H, W, C,
                                                                           (float*)workspace);
      DP_CHECK_LAUNCH("head_conv_wgrad_rows_kernel");
    } else {
      head_conv_wgrad_kernel<<<kHeadWgBlocks, 256, wg_smem, stream>>>(dout, out, relu, (const bf16*)x, x_ld, B,
                                                                      H, W, C, KS, (float*)workspace);
      DP_CHECK_LAUNCH("head_conv_wgrad_kernel");
    }
    head_conv_wgrad_reduce_kernel<<<dp::ceil_div(nacc * 32, 256), 256, 0, stream>>>((const float*)workspace, kHeadWgBlocks, C,
                                                                               KS, dw, db, accumulate);
    DP_CHECK_LAUNCH("head_conv_wgrad_reduce_kernel");
  }
  return DP_OK;
}

}  // extern "C"
