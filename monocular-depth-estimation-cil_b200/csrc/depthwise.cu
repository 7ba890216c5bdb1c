// Depthwise convolutions of the EfficientNet-Lite3 encoder trunk (the hub model consumed at
// reference src/network/blocks.py:166-186): k3 / k5, stride 1 / 2, NHWC bf16, fp32 accumulation.
// All three passes are bandwidth-bound (9..25 MAC per element), so the design goal is to touch HBM once per tensor:
//   * forward (and, with flipped taps, the stride-1 data gradient): each thread owns 8 channels (one 16-byte vector) of
//     a strip of 4 output pixels and slides the K x (3*S+K) input window through registers; the BatchNorm batch
//     statistics (sum, sum of squares of the stored bf16 value) of the layer that follows are accumulated on the fly
//     and reduced deterministically per block, so the BN needs no extra pass over the output;
//   * stride-2 data gradient: gather over the taps whose parity matches;
//   * weight gradient: thread = 8 channels x one kernel row, K x 8 fp32 accumulators, fixed-order two-level reduction
//     (no atomics).
#include "common.cuh"
#include "../../include/depth_b200.h"

namespace {

using namespace dp;

constexpr int TPB = 256;
constexpr int TX = 4;  // output pixels per thread along x

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = __uint_as_float(w[k] << 16);
    v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    w[k] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 ld8(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

struct DwArgs {
  const bf16* x; long long x_ld;
  int B, Hi, Wi, C;
  const float* w;        // [K*K][C] fp32, tap = ky*K + kx
  int pad_t, pad_l;
  bf16* out; long long out_ld;
  int Ho, Wo;
  float* stats;          // [gridDim.x][2][C] or null
  int G;                 // channel groups (of 8) per block column: min(C/8, 32)
};

// out[b,oy,ox,c] = sum_{ky,kx} w[ky*K+kx][c] * x[b, oy*S - pad_t + ky, ox*S - pad_l + kx, c]
// grid = (pixel-strip blocks, channel chunks of G groups).  A block keeps its chunk's taps in shared memory; thread =
// (channel group c8l = tid % G, strip slot = tid / G), so a warp touches G*16 contiguous bytes per pixel.
template <int K, int S>
__global__ void __launch_bounds__(TPB, 2) dw_fwd_kernel(DwArgs a) {
  extern __shared__ float smem[];            // [K*K][G*8] taps | [TPB][16] statistics scratch
  const int G = a.G, C8 = a.C / 8;
  float* s_w = smem;
  float* s_red = smem + K * K * G * 8;
  const int c8l = threadIdx.x % G, slot = threadIdx.x / G, nslots = TPB / G;
  const int c8 = blockIdx.y * G + c8l;
  const bool active = slot < nslots && c8 < C8;
  for (int i = threadIdx.x; i < K * K * G * 8; i += TPB) {
    const int tap = i / (G * 8), cc = i - tap * (G * 8);
    const int ch = blockIdx.y * G * 8 + cc;
    s_w[i] = ch < a.C ? __ldg(a.w + (size_t)tap * a.C + ch) : 0.f;
  }
  __syncthreads();
  const int strips = (a.Wo + TX - 1) / TX;
  const int nstrips = a.B * a.Ho * strips;
  float st_s[8], st_q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { st_s[j] = 0.f; st_q[j] = 0.f; }
  if (active) {
    for (int st = blockIdx.x * nslots + slot; st < nstrips; st += gridDim.x * nslots) {
      // rows vary fastest: the strip slots of a block (and the blocks next to it) work on vertically adjacent strips
      // at the same time, so the K-row input windows overlap in L1 / L2 instead of being fetched K times
      const int oy = st % a.Ho;
      const int r = st / a.Ho;
      const int sx = r % strips, b = r / strips;
      const int ox0 = sx * TX;
      float2 acc2[TX][4];            // packed f32x2 accumulators: one FFMA2 does two of the eight channels
#pragma unroll
      for (int t = 0; t < TX; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[t][j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * S - a.pad_t + ky;
        if (iy < 0 || iy >= a.Hi) continue;
        const bf16* rowp = a.x + (((long long)b * a.Hi + iy) * a.Wi) * a.x_ld + c8 * 8;
        constexpr int NCOL = (TX - 1) * S + K;
        uint4 raw[NCOL];
#pragma unroll
        for (int col = 0; col < NCOL; ++col) {       // all loads of the row window first, then the FMAs
          const int ix = ox0 * S - a.pad_l + col;
          raw[col] = (ix >= 0 && ix < a.Wi) ? ld8(rowp + (long long)ix * a.x_ld) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float4 w0 = *reinterpret_cast<const float4*>(s_w + (ky * K + kx) * G * 8 + c8l * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(s_w + (ky * K + kx) * G * 8 + c8l * 8 + 4);
          const float2 wv2[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                 make_float2(w1.z, w1.w)};
#pragma unroll
          for (int t = 0; t < TX; ++t) {
            const uint4 u = raw[kx + t * S];
            const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 v2 = make_float2(__uint_as_float(uw[j] << 16), __uint_as_float(uw[j] & 0xffff0000u));
              acc2[t][j] = __ffma2_rn(v2, wv2[j], acc2[t][j]);
            }
          }
        }
      }
      float acc[TX][8];
#pragma unroll
      for (int t = 0; t < TX; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[t][2 * j] = acc2[t][j].x; acc[t][2 * j + 1] = acc2[t][j].y; }
      bf16* op = a.out + (((long long)b * a.Ho + oy) * a.Wo + ox0) * a.out_ld + c8 * 8;
#pragma unroll
      for (int t = 0; t < TX; ++t) {
        if (ox0 + t < a.Wo) {
          const uint4 q = pack8(acc[t]);
          *reinterpret_cast<uint4*>(op + (long long)t * a.out_ld) = q;
          if (a.stats) {
            float f[8];
            unpack8(q, f);   // statistics of the value as stored
#pragma unroll
            for (int j = 0; j < 8; ++j) { st_s[j] += f[j]; st_q[j] = fmaf(f[j], f[j], st_q[j]); }
          }
        }
      }
    }
  }
  if (a.stats) {
    // deterministic block reduction over the strip slots of each channel group
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_red[threadIdx.x * 16 + j] = st_s[j]; s_red[threadIdx.x * 16 + 8 + j] = st_q[j]; }
    __syncthreads();
    if (threadIdx.x < G && blockIdx.y * G + threadIdx.x < C8) {
      float sacc[8], qacc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { sacc[j] = 0.f; qacc[j] = 0.f; }
      for (int sl = 0; sl < nslots; ++sl) {
        const int t = sl * G + threadIdx.x;
#pragma unroll
        for (int j = 0; j < 8; ++j) { sacc[j] += s_red[t * 16 + j]; qacc[j] += s_red[t * 16 + 8 + j]; }
      }
      float* dst = a.stats + (size_t)blockIdx.x * 2 * a.C + (size_t)(blockIdx.y * G + threadIdx.x) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) { dst[j] = sacc[j]; dst[a.C + j] = qacc[j]; }
    }
  }
}

// stride-2 data gradient: dx[b,iy,ix,c] = sum over taps with (iy + pad_t - ky) even: w[ky*K+kx][c] * dy[b,(iy+pad_t-ky)/2,...]
template <int K>
__global__ void __launch_bounds__(TPB) dw_dgrad_s2_kernel(const bf16* __restrict__ dy, long long dy_ld, int B, int Ho, int Wo,
                                                          int C, const float* __restrict__ w, int pad_t, int pad_l,
                                                          bf16* __restrict__ dx, long long dx_ld, int Hi, int Wi) {
  const int C8 = C / 8;
  const long long items = (long long)B * Hi * Wi * C8;
  for (long long idx = (long long)blockIdx.x * TPB + threadIdx.x; idx < items; idx += (long long)gridDim.x * TPB) {
    const int c8 = (int)(idx % C8);
    long long r = idx / C8;
    const int ix = (int)(r % Wi); r /= Wi;
    const int iy = (int)(r % Hi);
    const int b = (int)(r / Hi);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int ty = iy + pad_t - ky;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int tx = ix + pad_l - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= Wo) continue;
        float g[8];
        unpack8(ld8(dy + (((long long)b * Ho + oy) * Wo + ox) * dy_ld + c8 * 8), g);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(ky * K + kx) * C + c8 * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(ky * K + kx) * C + c8 * 8 + 4));
        acc[0] = fmaf(g[0], w0.x, acc[0]); acc[1] = fmaf(g[1], w0.y, acc[1]);
        acc[2] = fmaf(g[2], w0.z, acc[2]); acc[3] = fmaf(g[3], w0.w, acc[3]);
        acc[4] = fmaf(g[4], w1.x, acc[4]); acc[5] = fmaf(g[5], w1.y, acc[5]);
        acc[6] = fmaf(g[6], w1.z, acc[6]); acc[7] = fmaf(g[7], w1.w, acc[7]);
      }
    }
    *reinterpret_cast<uint4*>(dx + (((long long)b * Hi + iy) * Wi + ix) * dx_ld + c8 * 8) = pack8(acc);
  }
}

// weight gradient partials: partial[chunk][ky*K+kx][c] = sum over the chunk's output pixels of dy * x(tap)
// thread = (channel group, kernel row ky, strip slot); each step takes a strip of TX output pixels of one row: TX dy
// vectors and the (TX-1)*S+K input vectors under kernel row ky are loaded up front (independent loads), then multiplied.
constexpr int WG_C8B = 16;   // channel groups per block
template <int K, int S>
__global__ void __launch_bounds__(TPB) dw_wgrad_kernel(const bf16* __restrict__ x, long long x_ld, int B, int Hi, int Wi, int C,
                                                       const bf16* __restrict__ dy, long long dy_ld, int Ho, int Wo,
                                                       int pad_t, int pad_l, int nchunks, float* __restrict__ partial) {
  extern __shared__ float s_acc[];   // [TPB][K*8]
  const int C8 = C / 8;
  const int c8b = C8 < WG_C8B ? C8 : WG_C8B;
  const int per_slot = c8b * K;
  const int nslots = TPB / per_slot;
  const int c8l = threadIdx.x % c8b;
  const int ky = (threadIdx.x / c8b) % K;
  const int slot = threadIdx.x / per_slot;
  const int c8 = blockIdx.x * c8b + c8l;
  const int strips = (Wo + TX - 1) / TX;
  const int nstrips = B * Ho * strips;
  const int per_chunk = (nstrips + nchunks - 1) / nchunks;
  const int s0 = blockIdx.y * per_chunk;
  const int s1 = s0 + per_chunk < nstrips ? s0 + per_chunk : nstrips;
  float acc[K][8];
#pragma unroll
  for (int kx = 0; kx < K; ++kx)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[kx][j] = 0.f;
  if (slot < nslots && c8 < C8) {
    for (int st = s0 + slot; st < s1; st += nslots) {
      const int sx = st % strips;
      const int r = st / strips;
      const int oy = r % Ho, b = r / Ho;
      const int iy = oy * S - pad_t + ky;
      if (iy < 0 || iy >= Hi) continue;
      const int ox0 = sx * TX;
      constexpr int NCOL = (TX - 1) * S + K;
      uint4 graw[TX], xraw[NCOL];
      const bf16* gp = dy + (((long long)b * Ho + oy) * Wo + ox0) * dy_ld + c8 * 8;
#pragma unroll
      for (int t = 0; t < TX; ++t) graw[t] = (ox0 + t < Wo) ? ld8(gp + (long long)t * dy_ld) : make_uint4(0, 0, 0, 0);
      const bf16* rowp = x + (((long long)b * Hi + iy) * Wi) * x_ld + c8 * 8;
#pragma unroll
      for (int col = 0; col < NCOL; ++col) {
        const int ix = ox0 * S - pad_l + col;
        xraw[col] = (ix >= 0 && ix < Wi) ? ld8(rowp + (long long)ix * x_ld) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int t = 0; t < TX; ++t) {
        float g[8];
        unpack8(graw[t], g);
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          float v[8];
          unpack8(xraw[kx + t * S], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[kx][j] = fmaf(g[j], v[j], acc[kx][j]);
        }
      }
    }
  }
#pragma unroll
  for (int kx = 0; kx < K; ++kx)
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[threadIdx.x * (K * 8) + kx * 8 + j] = acc[kx][j];
  __syncthreads();
  if (slot == 0 && c8 < C8) {
    for (int kx = 0; kx < K; ++kx) {
      float s[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = 0.f;
      for (int sl = 0; sl < nslots; ++sl) {
        const int t = sl * per_slot + threadIdx.x;
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += s_acc[t * (K * 8) + kx * 8 + j];
      }
      float* dst = partial + ((size_t)blockIdx.y * K * K + ky * K + kx) * C + c8 * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = s[j];
    }
  }
}

// grad[c][ky][kx] (OIHW with I = 1) (+)= sum_chunks partial[chunk][tap][c]
// block = 32 consecutive (tap, c) columns x 32 chunk lanes: coalesced rows of the partials, fixed-order fp64 fold
__global__ void __launch_bounds__(1024) dw_wgrad_reduce_kernel(const float* __restrict__ partial, int nchunks, int taps, int C,
                                                               float* __restrict__ grad, int accumulate) {
  __shared__ double s_s[32][32];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cl;
  double s = 0.0;
  if (i < taps * C)
    for (int p = pl; p < nchunks; p += 32) s += (double)partial[(size_t)p * taps * C + i];
  s_s[pl][cl] = s;
  __syncthreads();
  if (pl != 0 || i >= taps * C) return;
  s = 0.0;
#pragma unroll
  for (int l = 0; l < 32; ++l) s += s_s[l][cl];
  const int tap = i / C, c = i - tap * C;
  const size_t o = (size_t)c * taps + tap;
  grad[o] = accumulate ? grad[o] + (float)s : (float)s;
}

inline int dw_grid(long long items) {
  long long b = (items + TPB - 1) / TPB;
  const long long cap = 8LL * kNumSMs;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

inline int wg_chunks(int B, int Ho, int Wo, int C) {
  const int C8 = C / 8;
  const int cblocks = (C8 + WG_C8B - 1) / WG_C8B;
  long long npix = (long long)B * Ho * Wo;
  long long n = (6LL * kNumSMs + cblocks - 1) / cblocks;
  if (n > npix / 128) n = npix / 128;
  if (n < 1) n = 1;
  return (int)n;
}

}  // namespace

extern "C" {

static inline int dw_G(int C) { return C / 8 < 32 ? C / 8 : 32; }

int dp_dwconv_fwd_blocks(int B, int Ho, int Wo, int C) {
  const int G = dw_G(C), ny = (C / 8 + G - 1) / G, nslots = TPB / G;
  const long long nstrips = (long long)B * Ho * ((Wo + TX - 1) / TX);
  long long nx = (nstrips + nslots - 1) / nslots;
  long long cap = (4LL * kNumSMs + ny - 1) / ny;
  if (cap < 1) cap = 1;
  if (nx > cap) nx = cap;
  return (int)(nx < 1 ? 1 : nx);
}

/* Depthwise K x K convolution (K = 3 or 5; stride 1 or 2; explicit top / left padding so both symmetric and TF-"SAME"
 * geometries are expressible), NHWC bf16.  w: fp32 [K*K][C] (tap-major).  stats_partials: null or
 * float[dp_dwconv_fwd_blocks()][2][C] receiving per-block (sum, sum of squares) of the stored output for the BatchNorm
 * that follows.  With flipped taps and pad' = K-1-pad this is also the stride-1 data gradient. */
int dp_dwconv_fwd(const void* x, long long x_ld, int B, int Hi, int Wi, int C, const float* w, int K, int stride,
                  int pad_t, int pad_l, void* out, long long out_ld, int Ho, int Wo, float* stats_partials,
                  cudaStream_t stream) {
  DP_CHECK_ARG(x && w && out, "dp_dwconv_fwd: null pointer");
  DP_CHECK_ARG(C % 8 == 0 && C <= 2048 && x_ld % 8 == 0 && out_ld % 8 == 0, "dp_dwconv_fwd: channels must be a multiple of 8 (<= 2048)");
  DP_CHECK_ARG((K == 3 || K == 5) && (stride == 1 || stride == 2), "dp_dwconv_fwd: K %d stride %d", K, stride);
  DP_CHECK_ARG((long long)B * Ho * ((Wo + TX - 1) / TX) < (1LL << 31), "dp_dwconv_fwd: too many pixel strips");
  const int G = dw_G(C);
  DwArgs a{reinterpret_cast<const bf16*>(x), x_ld, B, Hi, Wi, C, w, pad_t, pad_l, reinterpret_cast<bf16*>(out), out_ld,
           Ho, Wo, stats_partials, G};
  dim3 grid(dp_dwconv_fwd_blocks(B, Ho, Wo, C), (C / 8 + G - 1) / G);
  const size_t smem = ((size_t)K * K * G * 8 + (stats_partials ? (size_t)TPB * 16 : 0)) * sizeof(float);
  if (K == 3 && stride == 1) dw_fwd_kernel<3, 1><<<grid, TPB, smem, stream>>>(a);
  else if (K == 3) dw_fwd_kernel<3, 2><<<grid, TPB, smem, stream>>>(a);
  else if (stride == 1) dw_fwd_kernel<5, 1><<<grid, TPB, smem, stream>>>(a);
  else dw_fwd_kernel<5, 2><<<grid, TPB, smem, stream>>>(a);
  DP_CHECK_LAUNCH("dw_fwd_kernel");
  return DP_OK;
}

/* data gradient of the stride-2 depthwise convolution: dy (B,Ho,Wo,C) -> dx (B,Hi,Wi,C); w as in dp_dwconv_fwd */
int dp_dwconv_dgrad_s2(const void* dy, long long dy_ld, int B, int Ho, int Wo, int C, const float* w, int K, int pad_t,
                       int pad_l, void* dx, long long dx_ld, int Hi, int Wi, cudaStream_t stream) {
  DP_CHECK_ARG(dy && w && dx && C % 8 == 0 && (K == 3 || K == 5), "dp_dwconv_dgrad_s2: bad arguments");
  const int grid = dw_grid((long long)B * Hi * Wi * (C / 8));
  if (K == 3)
    dw_dgrad_s2_kernel<3><<<grid, TPB, 0, stream>>>(reinterpret_cast<const bf16*>(dy), dy_ld, B, Ho, Wo, C, w, pad_t, pad_l,
                                                    reinterpret_cast<bf16*>(dx), dx_ld, Hi, Wi);
  else
    dw_dgrad_s2_kernel<5><<<grid, TPB, 0, stream>>>(reinterpret_cast<const bf16*>(dy), dy_ld, B, Ho, Wo, C, w, pad_t, pad_l,
                                                    reinterpret_cast<bf16*>(dx), dx_ld, Hi, Wi);
  DP_CHECK_LAUNCH("dw_dgrad_s2_kernel");
  return DP_OK;
}

size_t dp_dwconv_wgrad_workspace(int B, int Ho, int Wo, int C, int K) {
  return (size_t)wg_chunks(B, Ho, Wo, C) * K * K * C * sizeof(float);
}

/* weight gradient of the depthwise convolution: grad (fp32, [C][1][K][K]) (+)= sum_p dy[p][c] * x[tap(p)][c] */
int dp_dwconv_wgrad(const void* x, long long x_ld, int B, int Hi, int Wi, int C, const void* dy, long long dy_ld, int Ho,
                    int Wo, int K, int stride, int pad_t, int pad_l, float* grad, int accumulate, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream) {
  DP_CHECK_ARG(x && dy && grad && workspace && C % 8 == 0, "dp_dwconv_wgrad: bad arguments");
  DP_CHECK_ARG((K == 3 || K == 5) && (stride == 1 || stride == 2), "dp_dwconv_wgrad: K %d stride %d", K, stride);
  if (workspace_bytes < dp_dwconv_wgrad_workspace(B, Ho, Wo, C, K))
    return dp_set_error(DP_ERR_WORKSPACE, "dp_dwconv_wgrad: workspace too small");
  const int nchunks = wg_chunks(B, Ho, Wo, C);
  const int C8 = C / 8;
  dim3 grid((C8 + WG_C8B - 1) / WG_C8B, nchunks);
  float* partial = reinterpret_cast<float*>(workspace);
  const size_t smem = (size_t)TPB * K * 8 * sizeof(float);
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  const bf16* gb = reinterpret_cast<const bf16*>(dy);
  if (K == 3 && stride == 1) dw_wgrad_kernel<3, 1><<<grid, TPB, smem, stream>>>(xb, x_ld, B, Hi, Wi, C, gb, dy_ld, Ho, Wo, pad_t, pad_l, nchunks, partial);
  else if (K == 3) dw_wgrad_kernel<3, 2><<<grid, TPB, smem, stream>>>(xb, x_ld, B, Hi, Wi, C, gb, dy_ld, Ho, Wo, pad_t, pad_l, nchunks, partial);
  else if (stride == 1) dw_wgrad_kernel<5, 1><<<grid, TPB, smem, stream>>>(xb, x_ld, B, Hi, Wi, C, gb, dy_ld, Ho, Wo, pad_t, pad_l, nchunks, partial);
  else dw_wgrad_kernel<5, 2><<<grid, TPB, smem, stream>>>(xb, x_ld, B, Hi, Wi, C, gb, dy_ld, Ho, Wo, pad_t, pad_l, nchunks, partial);
  DP_CHECK_LAUNCH("dw_wgrad_kernel");
  dw_wgrad_reduce_kernel<<<dp::ceil_div(K * K * C, 32), 1024, 0, stream>>>(partial, nchunks, K * K, C, grad, accumulate);
  DP_CHECK_LAUNCH("dw_wgrad_reduce_kernel");
  return DP_OK;
}

}  // extern "C"
