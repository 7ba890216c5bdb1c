// Depthwise convolutions of the EfficientNet-Lite3 encoder trunk (the hub model consumed at
// reference src/network/blocks.py:166-186): k3 / k5, stride 1 / 2, NHWC bf16, fp32 accumulation.
// By bytes all passes are bandwidth-bound (9..25 MAC per element); in practice the k5 layers are FMA / issue-bound (100
// packed FFMA2 per 16-byte output vector), so the design touches HBM once per tensor AND unpacks every operand once:
//   * stride-1 forward and (flipped taps) stride-1 data gradient: dw_tile_kernel - persistent blocks, one TMA box per
//     (TH+K-1) x (TW+K-1) input tile into one of two shared-memory buffers, R x XO output patch per thread, packed MACs,
//     BatchNorm batch statistics (sum, sum of squares of the stored bf16 value) accumulated on the fly and reduced
//     deterministically per block, so the BN that follows needs no extra pass over the output;
//   * stride-2 forward: dw_fwd_kernel (register window, 75 % of its HBM floor);
//   * stride-2 data gradient: dw_dgrad_s2_tile_kernel - 2 x 2 output blocks from a TMA tile of dy, tap sets compile-time
//     per parity of the top / left padding;
//   * weight gradient (both strides): dw_wgrad_tile_kernel - dy tile and input tile by TMA, thread = (channel group,
//     kernel row, row slot) with a sliding window, fixed-order two-level reduction (no atomics).
// dw_dgrad_s2_kernel / dw_wgrad_kernel are the round-1 gather forms, kept for operands whose pixel stride is not a
// multiple of 8 channels (TMA needs 16-byte strides).
#include "common.cuh"
#include "tc.cuh"
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
  dp::pdl_prologue();
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

// ---- stride-1 forward / data gradient from a shared-memory tile -----------------------------------------------------
// The register-window kernel above re-reads every input vector K times through L1 / L2 (10 loads per output at K = 5)
// and unpacks it for every tap: ~250 instructions per output vector, 13 - 40 % of the HBM floor.  Here a persistent block
// owns a channel chunk of CG groups (CG * 16 bytes per pixel) and walks TH x TW output tiles.  The (TH+K-1) x (TW+K-1)
// input tile arrives by ONE TMA box load (zero fill outside the image = the conv padding and the tile overhang) into one
// of two buffers, so the next tile is in flight while this one is computed.  A thread owns one channel group of an
// R-row x XO-column output patch: it reads each of the (R+K-1) x (XO+K-1) input vectors of its patch from shared memory
// once and unpacks it once, the taps come from shared memory ([tap][half][group][4] so the eight groups of a warp read
// 128 contiguous bytes), the MACs are packed f32x2.  Two blocks per SM; the grid is the LARGEST multiple of the channel
// chunks that fits the 2 x 148 resident slots - one block more and a second wave doubles the kernel's time.
// output patch per thread: 4 rows x 2 columns at K = 3 (3 input vectors per output), 2 x 2 at K = 5 (9 per output; the
// 4 x 2 patch needs 64 + 48 fp32 registers for accumulators and the unpacked row and spilled at 128 registers)
// (measured: 1 x 4 at K = 5 and 2 x 2 at K = 3 are within 2 % of these)
struct DwPatch { int R, XO; };
inline DwPatch dw_patch(int K) { return K == 3 ? DwPatch{4, 2} : DwPatch{2, 2}; }

template <int K, int TW, int CG, int DT_R, int DT_TXO>
__global__ void __launch_bounds__(TPB, 2) dw_tile_kernel(const __grid_constant__ CUtensorMap tmx, DwArgs a) {
  dp::pdl_prologue();
  constexpr int QX = TW / DT_TXO;                            // patches per tile row
  constexpr int TH = DT_R * (TPB / CG) / QX;                 // output rows per tile
  static_assert(TH >= DT_R && (TPB / CG) % QX == 0, "thread layout");
  constexpr int IH = TH + K - 1, IW = TW + K - 1;
  constexpr int PXB = CG * 16;                               // bytes per pixel of the chunk
  constexpr uint32_t TILE_BYTES = (uint32_t)IH * IW * PXB;
  constexpr uint32_t TILE_STRIDE = (TILE_BYTES + 127u) & ~127u;
  extern __shared__ __align__(128) unsigned char dsm[];
  unsigned char* tile = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsm) + 127) & ~uintptr_t(127));
  float* s_w = reinterpret_cast<float*>(tile + 2 * TILE_STRIDE);                  // [K*K][2][CG][4]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_w + K * K * CG * 8);             // [2]
  const int C8 = a.C / 8;
  const int cg = threadIdx.x % CG, pt = threadIdx.x / CG;
  const int qx = pt % QX, ry = pt / QX;
  const int c8 = blockIdx.y * CG + cg;
  const bool chan_ok = c8 < C8;
  const int tiles_x = (a.Wo + TW - 1) / TW, tiles_y = (a.Ho + TH - 1) / TH;
  const int ntiles = a.B * tiles_y * tiles_x;
  for (int i = threadIdx.x; i < K * K * CG * 8; i += TPB) {
    const int tap = i / (CG * 8), cc = i - tap * (CG * 8);
    const int g = cc >> 3, j = cc & 7;
    const int ch = blockIdx.y * CG * 8 + cc;
    s_w[tap * (CG * 8) + (j >> 2) * (CG * 4) + g * 4 + (j & 3)] = ch < a.C ? __ldg(a.w + (size_t)tap * a.C + ch) : 0.f;
  }
  if (threadIdx.x == 0) {
    tc::mbar_init(full, 1);
    tc::mbar_init(full + 1, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmx);
  }
  __syncthreads();
  auto issue = [&](int t, int buf) {            // thread 0: one box load = the whole input tile of output tile t
    const int tx = t % tiles_x, r = t / tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    tc::mbar_expect_tx(full + buf, TILE_BYTES);
    tc::tma_load_4d(tile + buf * TILE_STRIDE, &tmx, full + buf, blockIdx.y * CG * 8, tx * TW - a.pad_l, ty * TH - a.pad_t, b);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  float st_s[8], st_q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { st_s[j] = 0.f; st_q[j] = 0.f; }
  const uint32_t w_base = tc::smem_u32(s_w) + (uint32_t)cg * 16u;
  const uint32_t tb0 = tc::smem_u32(tile) + (uint32_t)((DT_R * ry) * IW + DT_TXO * qx) * PXB + (uint32_t)cg * 16u;
  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    // the other buffer was last read in iteration it - 1, and every thread is past that iteration's barrier
    if (threadIdx.x == 0 && t + (int)gridDim.x < ntiles) issue(t + gridDim.x, (it + 1) & 1);
    tc::mbar_wait(full + (it & 1), (it >> 1) & 1);
    const uint32_t tb = tb0 + (uint32_t)(it & 1) * TILE_STRIDE;
    const int tx = t % tiles_x, rr = t / tiles_x;
    const int ty = rr % tiles_y, b = rr / tiles_y;
    float2 acc[DT_R][DT_TXO][4];
#pragma unroll
    for (int r = 0; r < DT_R; ++r)
#pragma unroll
      for (int q = 0; q < DT_TXO; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][q][j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int iy = 0; iy < DT_R + K - 1; ++iy) {
      float2 rv[DT_TXO + K - 1][4];               // this input row of the patch, unpacked once
#pragma unroll
      for (int c = 0; c < DT_TXO + K - 1; ++c) {
        uint32_t u0, u1, u2, u3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3)
                     : "r"(tb + (uint32_t)(iy * IW + c) * PXB));
        const uint32_t uw[4] = {u0, u1, u2, u3};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          rv[c][j] = make_float2(__uint_as_float(uw[j] << 16), __uint_as_float(uw[j] & 0xffff0000u));
      }
#pragma unroll
      for (int r = 0; r < DT_R; ++r) {
        const int ky = iy - r;                     // compile-time after unrolling
        if (ky < 0 || ky >= K) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          float2 w2[4];
          const uint32_t wa = w_base + (uint32_t)((ky * K + kx) * CG * 8) * 4u;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w2[0].x), "=f"(w2[0].y), "=f"(w2[1].x), "=f"(w2[1].y) : "r"(wa));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w2[2].x), "=f"(w2[2].y), "=f"(w2[3].x), "=f"(w2[3].y) : "r"(wa + (uint32_t)(CG * 16)));
#pragma unroll
          for (int q = 0; q < DT_TXO; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][q][j] = __ffma2_rn(rv[q + kx][j], w2[j], acc[r][q][j]);
        }
      }
    }
    // every thread is past its reads of this buffer: the next iteration may refill it
    __syncthreads();
    if (chan_ok) {
#pragma unroll
      for (int r = 0; r < DT_R; ++r) {
        const int oy = ty * TH + DT_R * ry + r;
        if (oy >= a.Ho) continue;
#pragma unroll
        for (int q = 0; q < DT_TXO; ++q) {
          const int ox = tx * TW + DT_TXO * qx + q;
          if (ox >= a.Wo) continue;
          float o[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) { o[2 * j] = acc[r][q][j].x; o[2 * j + 1] = acc[r][q][j].y; }
          const uint4 pk = pack8(o);
          *reinterpret_cast<uint4*>(a.out + (((long long)b * a.Ho + oy) * a.Wo + ox) * a.out_ld + c8 * 8) = pk;
          if (a.stats) {
            float f[8];
            unpack8(pk, f);                        // statistics of the value as stored
#pragma unroll
            for (int j = 0; j < 8; ++j) { st_s[j] += f[j]; st_q[j] = fmaf(f[j], f[j], st_q[j]); }
          }
        }
      }
    }
  }
  if (a.stats) {
    // deterministic block reduction over the pixel threads of each channel group (the tile buffers are free now)
    __syncthreads();
    float* s_red = reinterpret_cast<float*>(tile);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_red[threadIdx.x * 16 + j] = st_s[j]; s_red[threadIdx.x * 16 + 8 + j] = st_q[j]; }
    __syncthreads();
    if (threadIdx.x < CG && blockIdx.y * CG + threadIdx.x < C8) {
      float sacc[8], qacc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { sacc[j] = 0.f; qacc[j] = 0.f; }
      for (int sl = 0; sl < TPB / CG; ++sl) {
        const int tt = sl * CG + threadIdx.x;
#pragma unroll
        for (int j = 0; j < 8; ++j) { sacc[j] += s_red[tt * 16 + j]; qacc[j] += s_red[tt * 16 + 8 + j]; }
      }
      float* dst = a.stats + (size_t)blockIdx.x * 2 * a.C + (size_t)(blockIdx.y * CG + threadIdx.x) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) { dst[j] = sacc[j]; dst[a.C + j] = qacc[j]; }
    }
  }
}

// tile geometry of the stride-1 path: output tile TH x TW, CG channel groups per block
struct DwTilePlan { int TW, CG, TH, ny, nx, R, XO; };
inline DwTilePlan dw_tile_plan(int B, int Ho, int Wo, int C, int K) {
  // candidates (TW, CG) -> TH = R * (256 / CG) / (TW / XO); (8, 8) is the tall narrow tile for the 14 x 18 / 28 x 36 maps
  // Pick the one that computes the fewest padded output vectors (tile overhang in x / y, channel-chunk overhang).
  const int cand[4][2] = {{16, 8}, {16, 16}, {32, 8}, {8, 8}};
  const DwPatch pp = dw_patch(K);
  DwTilePlan best{};
  long long best_cost = -1;
  for (int i = 0; i < 4; ++i) {
    DwTilePlan p;
    p.TW = cand[i][0]; p.CG = cand[i][1]; p.R = pp.R; p.XO = pp.XO;
    p.TH = pp.R * (TPB / p.CG) / (p.TW / pp.XO);
    p.ny = (C / 8 + p.CG - 1) / p.CG;
    const long long tx = (Wo + p.TW - 1) / p.TW, ty = (Ho + p.TH - 1) / p.TH;
    const long long cost = tx * p.TW * ty * p.TH * (long long)p.ny * p.CG;
    const long long ntiles = (long long)B * tx * ty;
    long long nx = (2LL * kNumSMs) / p.ny;                      // two blocks per SM, never a second wave
    if (nx > ntiles) nx = ntiles;
    p.nx = (int)(nx < 1 ? 1 : nx);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = p; }
  }
  return best;
}

// stride-2 data gradient: dx[b,iy,ix,c] = sum over taps with (iy + pad_t - ky) even: w[ky*K+kx][c] * dy[b,(iy+pad_t-ky)/2,...]
template <int K>
__global__ void __launch_bounds__(TPB) dw_dgrad_s2_kernel(const bf16* __restrict__ dy, long long dy_ld, int B, int Ho, int Wo,
                                                          int C, const float* __restrict__ w, int pad_t, int pad_l,
                                                          bf16* __restrict__ dx, long long dx_ld, int Hi, int Wi) {
  dp::pdl_prologue();
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
  dp::pdl_prologue();
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

// ---- stride-2 data gradient from a shared-memory tile --------------------------------------------------------------
// dx[iy,ix] = sum over taps with (iy + pad_t - ky) even of w[ky][kx] * dy[(iy + pad_t - ky) / 2, ...].  For the 2 x 2
// output block at even (iy0, ix0) the dy rows involved are o0 + r, r = 0 .. (K-1)/2, with o0 = ceil((iy0 + pad_t) / 2) -
// (K-1)/2, and output row iy0 + a meets dy row o0 + r through tap ky = K - 1 - e + a - 2r, e = (pad_t & 1) - every tap is
// used exactly once per block, and with the two padding parities as template parameters the tap set is compile-time.
// A persistent block owns a chunk of 8 channel groups and walks 16 x 32 output tiles; the (8 + (K-1)/2) x (16 + (K-1)/2)
// dy tile arrives as one TMA box (zero fill = out-of-range gradient rows / columns), double buffered; a thread owns two
// horizontally adjacent blocks (2 x 4 outputs) per pass, unpacks each dy vector once, MACs are packed f32x2.
constexpr int DG_TH = 16, DG_TW = 32, DG_CG = 8;

template <int K, int EY, int EX>
__global__ void __launch_bounds__(TPB, 2) dw_dgrad_s2_tile_kernel(const __grid_constant__ CUtensorMap tmg, int B, int Hi, int Wi,
                                                                  int C, const float* __restrict__ w, int pad_t, int pad_l,
                                                                  bf16* __restrict__ dx, long long dx_ld) {
  dp::pdl_prologue();
  constexpr int CG = DG_CG, TH = DG_TH, TW = DG_TW;
  constexpr int WR = (K + 1) / 2;                            // dy rows / columns under one 2 x 2 block
  constexpr int NR = TH / 2 + (K - 1) / 2, NC = TW / 2 + (K - 1) / 2;
  constexpr int PXB = CG * 16;
  constexpr uint32_t TILE_BYTES = (uint32_t)NR * NC * PXB;
  constexpr uint32_t TILE_STRIDE = (TILE_BYTES + 127u) & ~127u;
  constexpr int PW = TW / 4;                                 // patches (two blocks) per tile row of blocks
  constexpr int NPATCH = (TH / 2) * PW;
  constexpr int NPT = TPB / CG;                              // pixel threads
  static_assert(NPATCH % NPT == 0, "patches per thread");
  extern __shared__ __align__(128) unsigned char dsm[];
  unsigned char* tile = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsm) + 127) & ~uintptr_t(127));
  float* s_w = reinterpret_cast<float*>(tile + 2 * TILE_STRIDE);                  // [K*K][2][CG][4]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_w + K * K * CG * 8);
  const int C8 = C / 8;
  const int cg = threadIdx.x % CG, pt = threadIdx.x / CG;
  const int c8 = blockIdx.y * CG + cg;
  const bool chan_ok = c8 < C8;
  const int tiles_x = (Wi + TW - 1) / TW, tiles_y = (Hi + TH - 1) / TH;
  const int ntiles = B * tiles_y * tiles_x;
  for (int i = threadIdx.x; i < K * K * CG * 8; i += TPB) {
    const int tap = i / (CG * 8), cc = i - tap * (CG * 8);
    const int g = cc >> 3, j = cc & 7;
    const int ch = blockIdx.y * CG * 8 + cc;
    s_w[tap * (CG * 8) + (j >> 2) * (CG * 4) + g * 4 + (j & 3)] = ch < C ? __ldg(w + (size_t)tap * C + ch) : 0.f;
  }
  if (threadIdx.x == 0) {
    tc::mbar_init(full, 1);
    tc::mbar_init(full + 1, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmg);
  }
  __syncthreads();
  auto issue = [&](int t, int buf) {
    const int tx = t % tiles_x, r = t / tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    const int o0y = ((ty * TH + pad_t + 1) >> 1) - (K - 1) / 2, o0x = ((tx * TW + pad_l + 1) >> 1) - (K - 1) / 2;
    tc::mbar_expect_tx(full + buf, TILE_BYTES);
    tc::tma_load_4d(tile + buf * TILE_STRIDE, &tmg, full + buf, blockIdx.y * CG * 8, o0x, o0y, b);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  const uint32_t w_base = tc::smem_u32(s_w) + (uint32_t)cg * 16u;
  const uint32_t tb0 = tc::smem_u32(tile) + (uint32_t)cg * 16u;
  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    if (threadIdx.x == 0 && t + (int)gridDim.x < ntiles) issue(t + gridDim.x, (it + 1) & 1);
    tc::mbar_wait(full + (it & 1), (it >> 1) & 1);
    const int tx = t % tiles_x, rr = t / tiles_x;
    const int ty = rr % tiles_y, b = rr / tiles_y;
#pragma unroll 1
    for (int pi = pt; pi < NPATCH; pi += NPT) {
      const int bm = pi / PW, bn = (pi - bm * PW) * 2;       // block row, first block column of the patch
      const uint32_t tb = tb0 + (uint32_t)(it & 1) * TILE_STRIDE + (uint32_t)(bm * NC + bn) * PXB;
      float2 acc[2][4][4];                                   // [output row a][output column 2j + b][channel pair]
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[a][q][j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < WR; ++r) {
        float2 v[WR + 1][4];
#pragma unroll
        for (int c = 0; c < WR + 1; ++c) {
          uint32_t u0, u1, u2, u3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3)
                       : "r"(tb + (uint32_t)(r * NC + c) * PXB));
          const uint32_t uw[4] = {u0, u1, u2, u3};
#pragma unroll
          for (int j = 0; j < 4; ++j) v[c][j] = make_float2(__uint_as_float(uw[j] << 16), __uint_as_float(uw[j] & 0xffff0000u));
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int ky = K - 1 - EY + a - 2 * r;             // compile-time after unrolling
          if (ky < 0 || ky >= K) continue;
#pragma unroll
          for (int bb = 0; bb < 2; ++bb)
#pragma unroll
            for (int cc = 0; cc < WR; ++cc) {
              const int kx = K - 1 - EX + bb - 2 * cc;
              if (kx < 0 || kx >= K) continue;
              float2 w2[4];
              const uint32_t wa = w_base + (uint32_t)((ky * K + kx) * CG * 8) * 4u;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w2[0].x), "=f"(w2[0].y), "=f"(w2[1].x), "=f"(w2[1].y) : "r"(wa));
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w2[2].x), "=f"(w2[2].y), "=f"(w2[3].x), "=f"(w2[3].y) : "r"(wa + (uint32_t)(CG * 16)));
#pragma unroll
              for (int jb = 0; jb < 2; ++jb)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[a][2 * jb + bb][j] = __ffma2_rn(v[jb + cc][j], w2[j], acc[a][2 * jb + bb][j]);
            }
        }
      }
      if (chan_ok) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int iy = ty * TH + 2 * bm + a;
          if (iy >= Hi) continue;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ix = tx * TW + 2 * bn + q;
            if (ix >= Wi) continue;
            float o[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) { o[2 * j] = acc[a][q][j].x; o[2 * j + 1] = acc[a][q][j].y; }
            *reinterpret_cast<uint4*>(dx + (((long long)b * Hi + iy) * Wi + ix) * dx_ld + c8 * 8) = pack8(o);
          }
        }
      }
    }
    __syncthreads();                               // every thread is past its reads of this buffer
  }
}

// ---- weight gradient from shared-memory tiles -----------------------------------------------------------------------
// The kernel above reads dy K times and x ~K times through L1 / L2 and multiplies with scalar FMAs.  Here a persistent
// block owns a chunk of WT_CG channel groups and walks TH x 8 output tiles: the dy tile and the input tile under it
// ((TH-1)*S+K rows x (8-1)*S+K columns) arrive as two TMA boxes (zero fill outside either image: out-of-range products
// vanish, no bounds tests) - double buffered at stride 1, single buffered at stride 2 (the tile is four times larger;
// the second block of the SM computes while this one waits).  thread = (channel group, kernel row ky, row slot): it
// walks its output rows with a K-wide sliding window of unpacked input vectors - per output pixel one dy vector, S new
// input vectors, K x 4 packed f32x2 MACs into the K x 8 accumulators it keeps for the whole launch.
// partial[block][tap][c] is folded by dw_wgrad_reduce_kernel.
constexpr int WT_CG = 8, WT_TW = 8;
constexpr int wt_th(int K) { return K == 3 ? 10 : 12; }

template <int K, int S>
__global__ void __launch_bounds__(TPB, 2) dw_wgrad_tile_kernel(const __grid_constant__ CUtensorMap tmx,
                                                               const __grid_constant__ CUtensorMap tmg, int B, int Ho, int Wo,
                                                               int C, int pad_t, int pad_l, float* __restrict__ partial) {
  dp::pdl_prologue();
  constexpr int CG = WT_CG, TW = WT_TW, TH = wt_th(K);
  constexpr int NBUF = S == 1 ? 2 : 1;
  constexpr int NSLOT = (TPB / CG) / K;                      // row slots: 10 (K = 3) / 6 (K = 5)
  static_assert(TH % NSLOT == 0, "rows per slot");
  constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
  constexpr int PXB = CG * 16;
  constexpr uint32_t X_BYTES = (uint32_t)IH * IW * PXB, G_BYTES = (uint32_t)TH * TW * PXB;
  constexpr uint32_t X_STRIDE = (X_BYTES + 127u) & ~127u, BUF_STRIDE = X_STRIDE + ((G_BYTES + 127u) & ~127u);
  static_assert(NBUF * BUF_STRIDE >= (uint32_t)TPB * K * 8 * 4, "reduction scratch fits the tile buffers");
  extern __shared__ __align__(128) unsigned char dsm[];
  unsigned char* tile = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsm) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(tile + NBUF * BUF_STRIDE);
  const int C8 = C / 8;
  const int cg = threadIdx.x % CG, pt = threadIdx.x / CG;
  const int ky = pt % K, slot = pt / K;
  const bool active = slot < NSLOT;
  const int c8 = blockIdx.y * CG + cg;
  const int tiles_x = (Wo + TW - 1) / TW, tiles_y = (Ho + TH - 1) / TH;
  const int ntiles = B * tiles_y * tiles_x;
  if (threadIdx.x == 0) {
    tc::mbar_init(full, 1);
    tc::mbar_init(full + 1, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmx);
    tc::prefetch_tmap(&tmg);
  }
  __syncthreads();
  auto issue = [&](int t, int buf) {
    const int tx = t % tiles_x, r = t / tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    tc::mbar_expect_tx(full + buf, X_BYTES + G_BYTES);
    tc::tma_load_4d(tile + buf * BUF_STRIDE, &tmx, full + buf, blockIdx.y * CG * 8, tx * TW * S - pad_l, ty * TH * S - pad_t, b);
    tc::tma_load_4d(tile + buf * BUF_STRIDE + X_STRIDE, &tmg, full + buf, blockIdx.y * CG * 8, tx * TW, ty * TH, b);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  float2 acc[K][4];
#pragma unroll
  for (int kx = 0; kx < K; ++kx)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[kx][j] = make_float2(0.f, 0.f);
  const uint32_t tb0 = tc::smem_u32(tile) + (uint32_t)cg * 16u;
  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int buf = NBUF == 2 ? (it & 1) : 0;
    if (NBUF == 2 && threadIdx.x == 0 && t + (int)gridDim.x < ntiles) issue(t + gridDim.x, (it + 1) & 1);
    tc::mbar_wait(full + buf, NBUF == 2 ? ((it >> 1) & 1) : (it & 1));
    if (active) {
      const uint32_t xb = tb0 + (uint32_t)buf * BUF_STRIDE;
      const uint32_t gb = xb + X_STRIDE;
#pragma unroll 1
      for (int oy = slot; oy < TH; oy += NSLOT) {
        const uint32_t xrow = xb + (uint32_t)((oy * S + ky) * IW) * PXB;
        const uint32_t grow = gb + (uint32_t)(oy * TW) * PXB;
        float2 win[K][4];                          // sliding window of unpacked input vectors (indices are compile-time)
        auto ldv = [&](uint32_t addr, float2 (&v)[4]) {
          uint32_t u0, u1, u2, u3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3) : "r"(addr));
          const uint32_t uw[4] = {u0, u1, u2, u3};
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = make_float2(__uint_as_float(uw[j] << 16), __uint_as_float(uw[j] & 0xffff0000u));
        };
#pragma unroll
        for (int c = 0; c < K - S; ++c) ldv(xrow + (uint32_t)c * PXB, win[c % K]);
#pragma unroll
        for (int ox = 0; ox < TW; ++ox) {
#pragma unroll
          for (int c = ox * S + K - S; c < ox * S + K; ++c) ldv(xrow + (uint32_t)c * PXB, win[c % K]);   // the S new columns
          float2 g[4];
          ldv(grow + (uint32_t)ox * PXB, g);
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[kx][j] = __ffma2_rn(g[j], win[(ox * S + kx) % K][j], acc[kx][j]);
        }
      }
    }
    __syncthreads();                               // this buffer may be refilled now
    if (NBUF == 1 && threadIdx.x == 0 && t + (int)gridDim.x < ntiles) issue(t + gridDim.x, 0);
  }
  // fixed-order fold over the row slots, one (channel group, ky) row per thread of the first K * CG threads
  float* s_acc = reinterpret_cast<float*>(tile);   // [TPB][K*8]
#pragma unroll
  for (int kx = 0; kx < K; ++kx)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s_acc[threadIdx.x * (K * 8) + kx * 8 + 2 * j] = acc[kx][j].x;
      s_acc[threadIdx.x * (K * 8) + kx * 8 + 2 * j + 1] = acc[kx][j].y;
    }
  __syncthreads();
  if (slot == 0 && c8 < C8) {
#pragma unroll 1
    for (int kx = 0; kx < K; ++kx) {
      float sum[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) sum[j] = 0.f;
      for (int sl = 0; sl < NSLOT; ++sl) {
        const int tt = sl * (K * CG) + threadIdx.x;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] += s_acc[tt * (K * 8) + kx * 8 + j];
      }
      float* dst = partial + ((size_t)blockIdx.x * K * K + ky * K + kx) * C + c8 * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = sum[j];
    }
  }
}

// blocks along x of the tile weight-gradient kernel (= rows of its partials buffer)
inline int wt_blocks(int B, int Ho, int Wo, int C, int K) {
  const int ny = (C / 8 + WT_CG - 1) / WT_CG;
  const long long ntiles = (long long)B * ((Ho + wt_th(K) - 1) / wt_th(K)) * ((Wo + WT_TW - 1) / WT_TW);
  long long nx = (2LL * kNumSMs) / ny;
  if (nx > ntiles) nx = ntiles;
  return (int)(nx < 1 ? 1 : nx);
}

// grad[c][ky][kx] (OIHW with I = 1) (+)= sum_chunks partial[chunk][tap][c]
// block = 32 consecutive (tap, c) columns x 32 chunk lanes: coalesced rows of the partials, fixed-order fp64 fold
__global__ void __launch_bounds__(1024) dw_wgrad_reduce_kernel(const float* __restrict__ partial, int nchunks, int taps, int C,
                                                               float* __restrict__ grad, int accumulate) {
  dp::pdl_prologue();
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

int dp_dwconv_fwd_blocks_s(int B, int Ho, int Wo, int C, int K, int stride) {
  if (stride == 1) return dw_tile_plan(B, Ho, Wo, C, K).nx;
  return dp_dwconv_fwd_blocks(B, Ho, Wo, C);
}

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
  const double dw_work = 8.0 * (double)B * C * ((double)Hi * Wi + (double)Ho * Wo);   // ~4x the bytes: FMA-bound kernels
  if (stride == 1) {
    // shared-memory tile path; stats_partials then has dp_dwconv_fwd_blocks_s(..., 1) rows
    const DwTilePlan p = dw_tile_plan(B, Ho, Wo, C, K);
    CUtensorMap tm;
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)x_ld * 2, (uint64_t)Wi * x_ld * 2, (uint64_t)Hi * Wi * x_ld * 2};
    const uint32_t box[4] = {(uint32_t)(p.CG * 8), (uint32_t)(p.TW + K - 1), (uint32_t)(p.TH + K - 1), 1};
    int rc = dp_make_tmap_bf16(&tm, x, 4, dims, str, box, nullptr, 0 /* no swizzle: a thread group reads whole pixel rows */);
    if (rc) return rc;
    const size_t tile = (((size_t)(p.TH + K - 1) * (p.TW + K - 1) * p.CG * 16) + 127) & ~size_t(127);
    size_t smem = 128 + 2 * tile + (size_t)K * K * p.CG * 8 * 4 + 64;
    if (smem < 128 + (size_t)TPB * 16 * 4) smem = 128 + (size_t)TPB * 16 * 4;
    dim3 grid(p.nx, p.ny);
#define DP_DW_TILE(KK, TWW, CGG, RR, XX)                                                                  \
    do {                                                                                                  \
      cudaError_t e = cudaFuncSetAttribute(dw_tile_kernel<KK, TWW, CGG, RR, XX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return dp_set_error(DP_ERR_CUDA, "dw_tile_kernel smem %zu: %s", smem, cudaGetErrorString(e));       \
      dp::pdl_work(dw_work); dp::launch(dw_tile_kernel<KK, TWW, CGG, RR, XX>, grid, TPB, smem, stream, tm, a);                             \
    } while (0)
    if (K == 3 && p.TW == 8) DP_DW_TILE(3, 8, 8, 4, 2);
    else if (K == 5 && p.TW == 8) DP_DW_TILE(5, 8, 8, 2, 2);
    else if (K == 3 && p.TW == 32) DP_DW_TILE(3, 32, 8, 4, 2);
    else if (K == 3 && p.CG == 16) DP_DW_TILE(3, 16, 16, 4, 2);
    else if (K == 3) DP_DW_TILE(3, 16, 8, 4, 2);
    else if (p.TW == 32) DP_DW_TILE(5, 32, 8, 2, 2);
    else if (p.CG == 16) DP_DW_TILE(5, 16, 16, 2, 2);
    else DP_DW_TILE(5, 16, 8, 2, 2);
#undef DP_DW_TILE
    DP_CHECK_LAUNCH("dw_tile_kernel");
    return DP_OK;
  }
  dim3 grid(dp_dwconv_fwd_blocks(B, Ho, Wo, C), (C / 8 + G - 1) / G);
  const size_t smem = ((size_t)K * K * G * 8 + (stats_partials ? (size_t)TPB * 16 : 0)) * sizeof(float);
  dp::pdl_work(dw_work);
  if (K == 3 && stride == 1) dp::launch(dw_fwd_kernel<3, 1>, grid, TPB, smem, stream, a);
  else if (K == 3) dp::launch(dw_fwd_kernel<3, 2>, grid, TPB, smem, stream, a);
  else if (stride == 1) dp::launch(dw_fwd_kernel<5, 1>, grid, TPB, smem, stream, a);
  else dp::launch(dw_fwd_kernel<5, 2>, grid, TPB, smem, stream, a);
  DP_CHECK_LAUNCH("dw_fwd_kernel");
  return DP_OK;
}

/* data gradient of the stride-2 depthwise convolution: dy (B,Ho,Wo,C) -> dx (B,Hi,Wi,C); w as in dp_dwconv_fwd */
int dp_dwconv_dgrad_s2(const void* dy, long long dy_ld, int B, int Ho, int Wo, int C, const float* w, int K, int pad_t,
                       int pad_l, void* dx, long long dx_ld, int Hi, int Wi, cudaStream_t stream) {
  DP_CHECK_ARG(dy && w && dx && C % 8 == 0 && (K == 3 || K == 5), "dp_dwconv_dgrad_s2: bad arguments");
  if (dy_ld % 8 == 0 && dx_ld % 8 == 0) {
    const int NR = DG_TH / 2 + (K - 1) / 2, NC = DG_TW / 2 + (K - 1) / 2;
    CUtensorMap tmg;
    const uint64_t gd[4] = {(uint64_t)C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    const uint64_t gs[3] = {(uint64_t)dy_ld * 2, (uint64_t)Wo * dy_ld * 2, (uint64_t)Ho * Wo * dy_ld * 2};
    const uint32_t gbox[4] = {(uint32_t)(DG_CG * 8), (uint32_t)NC, (uint32_t)NR, 1};
    int rc = dp_make_tmap_bf16(&tmg, dy, 4, gd, gs, gbox, nullptr, 0);
    if (rc) return rc;
    const size_t tile = (((size_t)NR * NC * DG_CG * 16) + 127) & ~size_t(127);
    const size_t smem = 128 + 2 * tile + (size_t)K * K * DG_CG * 8 * 4 + 64;
    const int ny = (C / 8 + DG_CG - 1) / DG_CG;
    const long long ntiles = (long long)B * ((Hi + DG_TH - 1) / DG_TH) * ((Wi + DG_TW - 1) / DG_TW);
    long long nx = (2LL * kNumSMs) / ny;
    if (nx > ntiles) nx = ntiles;
    if (nx < 1) nx = 1;
    dim3 grid2((unsigned)nx, ny);
    bf16* dxb = reinterpret_cast<bf16*>(dx);
#define DP_DW_DG(KK, EYY, EXX)                                                                                             \
    do {                                                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(dw_dgrad_s2_tile_kernel<KK, EYY, EXX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return dp_set_error(DP_ERR_CUDA, "dw_dgrad_s2_tile_kernel smem %zu: %s", smem, cudaGetErrorString(e)); \
      dp::pdl_work(4.0 * (double)B * C * ((double)Hi * Wi + (double)Ho * Wo)); dp::launch(dw_dgrad_s2_tile_kernel<KK, EYY, EXX>, grid2, TPB, smem, stream, tmg, B, Hi, Wi, C, w, pad_t, pad_l, dxb, dx_ld);  \
    } while (0)
    const int ey = pad_t & 1, ex = pad_l & 1;
    if (K == 3) {
      if (ey && ex) DP_DW_DG(3, 1, 1); else if (ey) DP_DW_DG(3, 1, 0); else if (ex) DP_DW_DG(3, 0, 1); else DP_DW_DG(3, 0, 0);
    } else {
      if (ey && ex) DP_DW_DG(5, 1, 1); else if (ey) DP_DW_DG(5, 1, 0); else if (ex) DP_DW_DG(5, 0, 1); else DP_DW_DG(5, 0, 0);
    }
#undef DP_DW_DG
    DP_CHECK_LAUNCH("dw_dgrad_s2_tile_kernel");
    return DP_OK;
  }
  const int grid = dw_grid((long long)B * Hi * Wi * (C / 8));
  if (K == 3)
    dp::launch(dw_dgrad_s2_kernel<3>, grid, TPB, 0, stream, reinterpret_cast<const bf16*>(dy), dy_ld, B, Ho, Wo, C, w, pad_t, pad_l,
                                                    reinterpret_cast<bf16*>(dx), dx_ld, Hi, Wi);
  else
    dp::launch(dw_dgrad_s2_kernel<5>, grid, TPB, 0, stream, reinterpret_cast<const bf16*>(dy), dy_ld, B, Ho, Wo, C, w, pad_t, pad_l,
                                                    reinterpret_cast<bf16*>(dx), dx_ld, Hi, Wi);
  DP_CHECK_LAUNCH("dw_dgrad_s2_kernel");
  return DP_OK;
}

size_t dp_dwconv_wgrad_workspace(int B, int Ho, int Wo, int C, int K) {
  int n = wg_chunks(B, Ho, Wo, C);
  if (K == 3 || K == 5) { const int t = wt_blocks(B, Ho, Wo, C, K); if (t > n) n = t; }
  return (size_t)n * K * K * C * sizeof(float);
}

/* weight gradient of the depthwise convolution: grad (fp32, [C][1][K][K]) (+)= sum_p dy[p][c] * x[tap(p)][c] */
int dp_dwconv_wgrad(const void* x, long long x_ld, int B, int Hi, int Wi, int C, const void* dy, long long dy_ld, int Ho,
                    int Wo, int K, int stride, int pad_t, int pad_l, float* grad, int accumulate, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream) {
  DP_CHECK_ARG(x && dy && grad && workspace && C % 8 == 0, "dp_dwconv_wgrad: bad arguments");
  DP_CHECK_ARG((K == 3 || K == 5) && (stride == 1 || stride == 2), "dp_dwconv_wgrad: K %d stride %d", K, stride);
  if (workspace_bytes < dp_dwconv_wgrad_workspace(B, Ho, Wo, C, K))
    return dp_set_error(DP_ERR_WORKSPACE, "dp_dwconv_wgrad: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  if (x_ld % 8 == 0 && dy_ld % 8 == 0) {
    const int TH = wt_th(K), S = stride;
    const int IH = (TH - 1) * S + K, IW = (WT_TW - 1) * S + K;
    CUtensorMap tmx, tmg;
    const uint64_t xd[4] = {(uint64_t)C, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
    const uint64_t xs[3] = {(uint64_t)x_ld * 2, (uint64_t)Wi * x_ld * 2, (uint64_t)Hi * Wi * x_ld * 2};
    const uint32_t xbox[4] = {(uint32_t)(WT_CG * 8), (uint32_t)IW, (uint32_t)IH, 1};
    int rc = dp_make_tmap_bf16(&tmx, x, 4, xd, xs, xbox, nullptr, 0);
    if (rc) return rc;
    const uint64_t gd[4] = {(uint64_t)C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    const uint64_t gs[3] = {(uint64_t)dy_ld * 2, (uint64_t)Wo * dy_ld * 2, (uint64_t)Ho * Wo * dy_ld * 2};
    const uint32_t gbox[4] = {(uint32_t)(WT_CG * 8), (uint32_t)WT_TW, (uint32_t)TH, 1};
    rc = dp_make_tmap_bf16(&tmg, dy, 4, gd, gs, gbox, nullptr, 0);
    if (rc) return rc;
    const size_t xb = (((size_t)IH * IW * WT_CG * 16) + 127) & ~size_t(127);
    const size_t gbz = (((size_t)TH * WT_TW * WT_CG * 16) + 127) & ~size_t(127);
    const size_t smem = 128 + (S == 1 ? 2 : 1) * (xb + gbz) + 64;
    const int nx = wt_blocks(B, Ho, Wo, C, K);
    dim3 grid(nx, (C / 8 + WT_CG - 1) / WT_CG);
#define DP_DW_WT(KK, SS)                                                                                                   \
    do {                                                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(dw_wgrad_tile_kernel<KK, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return dp_set_error(DP_ERR_CUDA, "dw_wgrad_tile_kernel smem %zu: %s", smem, cudaGetErrorString(e)); \
      dp::pdl_work(8.0 * (double)B * C * ((double)Hi * Wi + (double)Ho * Wo)); dp::launch(dw_wgrad_tile_kernel<KK, SS>, grid, TPB, smem, stream, tmx, tmg, B, Ho, Wo, C, pad_t, pad_l, partial);            \
    } while (0)
    if (K == 3 && S == 1) DP_DW_WT(3, 1);
    else if (K == 3) DP_DW_WT(3, 2);
    else if (S == 1) DP_DW_WT(5, 1);
    else DP_DW_WT(5, 2);
#undef DP_DW_WT
    DP_CHECK_LAUNCH("dw_wgrad_tile_kernel");
    dp::launch(dw_wgrad_reduce_kernel, dp::ceil_div(K * K * C, 32), 1024, 0, stream, partial, nx, K * K, C, grad, accumulate);
    DP_CHECK_LAUNCH("dw_wgrad_reduce_kernel");
    return DP_OK;
  }
  const int nchunks = wg_chunks(B, Ho, Wo, C);
  const int C8 = C / 8;
  dim3 grid((C8 + WG_C8B - 1) / WG_C8B, nchunks);
  const size_t smem = (size_t)TPB * K * 8 * sizeof(float);
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  const bf16* gb = reinterpret_cast<const bf16*>(dy);
  if (K == 3 && stride == 1) dp::launch(dw_wgrad_kernel<3, 1>, grid, TPB, smem, stream, xb, x_ld, B, Hi, Wi, C, gb, dy_ld, Ho, Wo, pad_t, pad_l, nchunks, partial);
  else if (K == 3) dp::launch(dw_wgrad_kernel<3, 2>, grid, TPB, smem, stream, xb, x_ld, B, Hi, Wi, C, gb, dy_ld, Ho, Wo, pad_t, pad_l, nchunks, partial);
  else if (stride == 1) dp::launch(dw_wgrad_kernel<5, 1>, grid, TPB, smem, stream, xb, x_ld, B, Hi, Wi, C, gb, dy_ld, Ho, Wo, pad_t, pad_l, nchunks, partial);
  else dp::launch(dw_wgrad_kernel<5, 2>, grid, TPB, smem, stream, xb, x_ld, B, Hi, Wi, C, gb, dy_ld, Ho, Wo, pad_t, pad_l, nchunks, partial);
  DP_CHECK_LAUNCH("dw_wgrad_kernel");
  dp::launch(dw_wgrad_reduce_kernel, dp::ceil_div(K * K * C, 32), 1024, 0, stream, partial, nchunks, K * K, C, grad, accumulate);
  DP_CHECK_LAUNCH("dw_wgrad_reduce_kernel");
  return DP_OK;
}

}  // extern "C"
