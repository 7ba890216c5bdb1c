// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces every 3x3 / stride 1 / pad 1 and 1x1 nn.Conv2d of the reference's decoder, fusion blocks
// and heads (reference src/network/blocks.py:149-161, 335-341, 401; midas_net_custom.py:105-113;
// midas_semantics.py:132-143, 195, 203; dpt_depth.py:39-47, 103-106) and, with flipped/transposed
// weights, their data gradients.
//
// GEMM view: M = output pixels, N = output channels, K = taps x input channels.
//   * Activations are NHWC bf16.  An M tile is a th x tw = 128-pixel patch of one image.  For each
//     horizontal tap s one TMA box of (th+2) x tw pixels x KB channels is loaded at (y0-1, x0+s-1);
//     out-of-bounds pixels/channels are zero-filled by TMA, which implements the conv padding.
//     Because tw is a multiple of 8, the three vertical taps r are the SAME shared-memory box
//     addressed r*tw rows further down (a multiple of the 8-row swizzle atom), so one load feeds
//     three UMMA chains (tests/test_umma_probe_gpu.py::test_k_major_row_shifted_start pins this).
//   * Weights are pre-packed [tap][Cout][Cin] bf16 (K-major).  Small layers keep the whole weight
//     set resident in shared memory for the life of the persistent CTA; large ones stream it.
//   * Accumulators live in TMEM, double buffered (2 x BN fp32 columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1.
//   * The k-loop is table driven (ColLoad): each entry is one TMA box (source plane, x/y offset, number of
//     vertical taps that share it, weight tap ids).  The plain 3x3 conv has three entries (one per horizontal
//     tap).  Stride-2 convolutions read the four (row parity, column parity) planes of the input as separate
//     strided tensor maps, which turns them into sums of small stride-1 convolutions; transposed stride-2
//     convolutions are four output-phase launches whose epilogue writes pixel (2y+a, 2x+b).  This puts
//     CrossAttention.spatial_reduction / spatial_upsample (midas_semantics.py:38-61), Dinov2Head.resize_layers[3]
//     (dpt_depth.py:63-68) and their data gradients on the tensor cores as well.
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
//     warps 4..7 = epilogue (TMEM -> registers -> bias / residual / ReLU / BN partial sums -> global).
#include <cstdlib>
#include "common.cuh"
#include "tc.cuh"
#include "bn_fuse.cuh"
#include "../../include/depth_b200.h"

extern unsigned long long* g_wg_dbg;

// Per-role cycle counters (tools/dbg_conv.py).  Compiled out by default: a clock read costs the single-thread MMA
// issuer a dependent-issue slot per use.
#ifdef DP_CONV_TIMING
#define DP_T(x) x
#else
#define DP_T(x)
#endif

namespace {

constexpr int kRoleThreads = 128;   // warp 0 TMA producer, warps 1/3 MMA issuers, warp 2 TMEM allocator
constexpr int kEpiWarps = 8;        // warps 4..11: two per scheduler, splitting the tile's columns
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kRoleThreads + kEpiThreads;
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;

constexpr int kMaxCols = 8;

struct ColLoad {
  int map;        // index of the A tensor map (source plane)
  int dx, dy;     // box origin relative to the tile origin (x0, y0), in plane pixels
  int nr;         // vertical taps sharing this box (box rows = th + nr - 1)
  int wtap[3];    // weight-pack tap id of vertical tap r
  int wslot[3];   // resident-weight slot of vertical tap r
  uint32_t box_bytes;
};

struct Maps {
  CUtensorMap a[kMaxCols];
  CUtensorMap b;
  CUtensorMap o, o2;   // output tiles (TMA-store epilogue)
};

struct ConvArgs {
  int B, H, W, Cout;              // GEMM pixel grid (tile domain) and output channels
  int Ho, Wo, osy, osx, oay, oax; // output tensor dims; GEMM pixel (y,x) -> output pixel (y*osy+oay, x*osx+oax)
  int th, tw, tiles_y, tiles_x;
  int KB, kchunks;
  int BN, n_blocks;
  int stages, resident;
  int ncols, max_nr, nslots;
  int nbuf, niss;                 // TMEM accumulator buffers (2 or 4) and MMA-issuing warps (1 or 2)
  int halo, pitch;                // halo mode: one (th+2) x (tw+2) box per channel chunk serves all nine taps
  int step_nb, step_tx, step_ty, step_n;   // mixed-radix digits of the item stride (gridDim.x)
  int epi_tma;                    // 1: registers -> swizzled smem tile -> TMA store (BN in {16,32,64}); 0: direct stores
  uint32_t out_tile_bytes;        // bytes of one staged output tile (128 pixels x BN bf16)
  int nob;                        // staging ring depth (per output): nob - 2 TMA stores may still be reading smem
  ColLoad cols[kMaxCols];
  uint32_t a_slot_bytes, b_tap_bytes, row_bytes, layout, sbo, idesc;
  uint32_t b_box_bytes;  // bytes TMA actually writes per weight box (slots are rounded up to 1024)
  uint32_t resident_bytes;
  long long total_items;
  bf16* out;  long long out_ld;
  bf16* out2; long long out2_ld;
  const float* bias;
  const bf16* res; long long res_ld;
  const bf16* resb; long long resb_ld;
  int relu, relu2;
  float* stats;  // [gridDim.x][2][Cout] or null
  // fused BatchNorm prologue (warps 2 and 3 rewrite every landed A box in place before the MMAs read it):
  //   A = act(x * pre_ss[c] + pre_ss[pre_c + c]), and exactly 0 outside the image (the conv pads the ACTIVATED tensor)
  const float* pre_ss; int pre_act, pre_c, pre_pad;   // pre_pad = kchunks * KB (table length per row in smem)
  int pH[kMaxCols], pW[kMaxCols];                     // extent of source plane `map` (pixels outside are padding)
  // fused BatchNorm-backward epilogue: `resb` is the pre-activation tensor c of the BN + activation whose output
  // gradient this launch produces; out = acc * [0 < c*aux_ss[n] + aux_ss[Cout+n] < aux_hi]; stats = (sum out, sum out*c)
  const float* aux_ss; float aux_hi; int aux_mode;
  int epi_pipe;                   // software-pipeline a single epilogue operand one tile ahead (epilogue_tma)
  unsigned long long* dbg;
};

struct __align__(8) Barriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[4];
  uint64_t tmem_empty[4];
  uint64_t resident_full;
  uint64_t ready[kMaxStages];   // prologue mode: A box transformed (two arrivals: warps 2 and 3)
  uint32_t tmem_base;
};

// Work items (image, tile row, tile column, N block) are walked with a fixed stride of gridDim.x.  Decoding an item
// index costs three 64-bit divisions (hundreds of cycles on the critical path of every role, every tile), so each
// role decodes once and then advances the mixed-radix digits by the precomputed digits of the stride.
struct TileIter {
  int nb, tx, ty, n;
  __device__ __forceinline__ void init(const ConvArgs& a, unsigned item) {
    nb = (int)(item % (unsigned)a.n_blocks);
    unsigned m = item / (unsigned)a.n_blocks;
    tx = (int)(m % (unsigned)a.tiles_x);
    m /= (unsigned)a.tiles_x;
    ty = (int)(m % (unsigned)a.tiles_y);
    n = (int)(m / (unsigned)a.tiles_y);
  }
  __device__ __forceinline__ void step(const ConvArgs& a) {
    nb += a.step_nb;
    tx += a.step_tx;
    ty += a.step_ty;
    n += a.step_n;
    if (nb >= a.n_blocks) { nb -= a.n_blocks; ++tx; }
    if (tx >= a.tiles_x) { tx -= a.tiles_x; ++ty; }
    if (ty >= a.tiles_y) { ty -= a.tiles_y; ++n; }
  }
};

using dpf::pack_bf16x2;
using dpf::bf16_lo;
using dpf::bf16_hi;
using dpf::transform_box;

// column sums of a 32 (lanes) x 16 (registers) tile: after the call, lane L holds the sum of column col_of(L)
// (lanes L and L^1 hold the same column).  16 shuffles instead of 80.
__device__ __forceinline__ float transpose_reduce16(const float (&v)[16], int lane) {
  float a8[8], a4[4], a2[2], a1;
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float send = h16 ? v[j] : v[j + 8];
    float keep = h16 ? v[j + 8] : v[j];
    a8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float send = h8 ? a8[j] : a8[j + 4];
    float keep = h8 ? a8[j + 4] : a8[j];
    a4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float send = h4 ? a4[j] : a4[j + 2];
    float keep = h4 ? a4[j + 2] : a4[j];
    a2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    float send = h2 ? a2[0] : a2[1];
    float keep = h2 ? a2[1] : a2[0];
    a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}
__device__ __forceinline__ int transpose_reduce16_col(int lane) {
  return ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
}

// BatchNorm-backward epilogue helper: v8 are 8 consecutive output columns (gradient w.r.t. the activated tensor), cw the
// matching 8 bf16 values of the pre-activation tensor c.  The activation is recomputed as in the forward pass
// (m = c * scale + shift in fp32) and the gradient is zeroed where it was clamped; c is returned for the sum of g * c.
__device__ __forceinline__ void aux_mask8(float* v8, const uint4& cw, const float* s_aux, int Cout, int col, float hi,
                                          float (&cx)[8]) {
  const uint32_t w[4] = {cw.x, cw.y, cw.z, cw.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) { cx[2 * k] = bf16_lo(w[k]); cx[2 * k + 1] = bf16_hi(w[k]); }
  if (col < Cout) {                               // Cout is a multiple of 8
    // explicit shared-space loads: through the generic pointer ptxas emitted LD.E (generic) and the epilogue sat on
    // the long scoreboard (31 % of its stall samples)
    float sc[8], sh[8];
    const uint32_t ta = tc::smem_u32(s_aux) + (uint32_t)col * 4u, tb = ta + (uint32_t)Cout * 4u;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[0]), "=f"(sc[1]), "=f"(sc[2]), "=f"(sc[3]) : "r"(ta));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[4]), "=f"(sc[5]), "=f"(sc[6]), "=f"(sc[7]) : "r"(ta + 16u));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sh[0]), "=f"(sh[1]), "=f"(sh[2]), "=f"(sh[3]) : "r"(tb));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sh[4]), "=f"(sh[5]), "=f"(sh[6]), "=f"(sh[7]) : "r"(tb + 16u));
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float m = fmaf(cx[e], sc[e], sh[e]);
      v8[e] = (m > 0.f && m < hi) ? v8[e] : 0.f;
    }
  }
}

// The epilogue's residual / BatchNorm-backward operands are read per tile straight from global memory by the thread
// that owns the pixel.  Issued at the top of a tile they would expose a full DRAM round trip on the epilogue's critical
// path every tile (measured: 64->64 @448x576 0.60 -> 1.32 ms with one operand), so the lines of the tile kPfTiles ahead are
// pulled into L2 first: the loads at the top of a tile then cost an L2 hit.
constexpr int kPfTiles = 2;
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<uint64_t>(p)));
}
__device__ __forceinline__ void prefetch_tile_operands(const ConvArgs& a, const TileIter& tp, int py, int px, int col0) {
  const int y = tp.ty * a.th + py, x = tp.tx * a.tw + px;
  const int oy = y * a.osy + a.oay, ox = x * a.osx + a.oax;
  if (!((y < a.H) && (x < a.W) && (oy < a.Ho) && (ox < a.Wo)) || col0 >= a.Cout) return;
  const long long pix = ((long long)tp.n * a.Ho + oy) * a.Wo + ox;
  if (a.res) prefetch_l2(a.res + pix * a.res_ld + col0);
  if (a.resb) prefetch_l2(a.resb + pix * a.resb_ld + col0);
}

// ---- epilogue A: BN = 2*CW in {16, 32, 64}.  Each of the eight warps owns 32 pixels x CW columns of the tile: one
// tcgen05.ld, TMEM released at once, math in registers, bf16 rows written to a swizzled shared-memory tile that one
// thread hands to TMA (the store clips tile overhang, so no per-pixel predicates and fully coalesced HBM writes).
// BatchNorm partial sums stay in registers across all tiles of the persistent CTA (thread = fixed pixel slot and
// column slice) and are reduced once at the end.
// MODE (compile time, so that the plain path carries none of the operand code: with run-time switches the common plain
// launch lost 20 % - 64->64 @448x576 0.50 -> 0.60 ms - to register pressure and instruction-cache misses):
//   0 no epilogue operand, 1 residual(s) read per tile (L2-prefetched kPfTiles ahead), 2 BatchNorm-backward mode (`resb`
//   is the pre-activation tensor, software-pipelined one tile ahead through registers)
template <int CW, int MODE>
__device__ __forceinline__ void epilogue_tma(const Maps& tm, const ConvArgs& a, Barriers* bars, uint8_t* s_out,
                                             float* s_stats, const float* s_aux, uint32_t tmem, int warp, int lane) {
  const int e = warp - 4, ew = e & 3, half = e >> 2;
  const int m = ew * 32 + lane;
  const int py = m / a.tw, px = m - py * a.tw;
  const int col0 = half * CW;
  constexpr uint32_t kRowBytes = 4u * CW;                  // BN * 2
  constexpr uint32_t kSwz = kRowBytes == 128 ? 7u : (kRowBytes == 64 ? 3u : 1u);
  const bool store_thread = threadIdx.x == kRoleThreads;
  const bool dual = a.out2 != nullptr;
  const bool want_stats = a.stats != nullptr;
  float acc_s[CW], acc_q[CW];
#pragma unroll
  for (int j = 0; j < CW; ++j) { acc_s[j] = 0.f; acc_q[j] = 0.f; }
  uint32_t soff[CW / 8];                                   // swizzled byte offsets of this thread's 16-byte chunks
#pragma unroll
  for (int j = 0; j < CW / 8; ++j) {
    const uint32_t off = (uint32_t)m * kRowBytes + (uint32_t)(col0 * 2 + j * 16);
    soff[j] = off ^ (((off >> 7) & kSwz) << 4);
  }
  const uint32_t s_base = tc::smem_u32(s_out);
  DP_T(long long dbg_ew = 0; long long dbg_ek = 0;)
  int it = 0, ring = 0;
  TileIter ti;
  ti.init(a, blockIdx.x);
  // Epilogue operands (residuals, or the pre-activation tensor of the BatchNorm-backward mode) are read by the thread
  // that owns the pixel.  Loaded at the top of their own tile they put a full memory round trip on the epilogue's
  // critical path every tile (ncu: 23 % of all stall samples on the first use of the loaded value, 64->64 @448x576
  // 0.60 -> 0.95 ms).  With a single operand - the common case - the loads are software-pipelined one tile ahead in the
  // register set the second operand would have used: `r2` receives tile i+1 while `r1` (tile i) is consumed.
  constexpr bool single = MODE == 2;       // MODE 2: the one operand is `resb` (launch() rejects a residual next to it)
  constexpr bool HAS_RES = MODE == 1;
  const bf16* op = a.resb;
  const long long op_ld = a.resb_ld;
  uint4 r1[CW / 8], r2[CW / 8];
  auto load_operand = [&](const TileIter& t, uint4 (&dst)[CW / 8]) {
    const int yy = t.ty * a.th + py, xx = t.tx * a.tw + px;
    const int oyy = yy * a.osy + a.oay, oxx = xx * a.osx + a.oax;
    const bool ok = (yy < a.H) && (xx < a.W) && (oyy < a.Ho) && (oxx < a.Wo);
    const uint4* rp = reinterpret_cast<const uint4*>(op + (((long long)t.n * a.Ho + oyy) * a.Wo + oxx) * op_ld + col0);
#pragma unroll
    for (int j = 0; j < CW / 8; ++j) dst[j] = (ok && col0 + j * 8 < a.Cout) ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
  };
  TileIter tn = ti;                                          // one tile ahead of ti
  if (single && (long long)blockIdx.x < a.total_items) load_operand(tn, r2);
  tn.step(a);
  // operands that are not pipelined through registers (two operands, or plain residuals - for which the register
  // rotation measured no gain) are pulled into L2 kPfTiles tiles ahead instead (64->64 + residual: 1.12 -> 0.95 ms)
  constexpr bool pf = MODE == 1;
  TileIter tp = ti;
  if (pf) {
    for (int k = 0; k < kPfTiles; ++k) {
      if (blockIdx.x + (long long)k * gridDim.x < a.total_items) prefetch_tile_operands(a, tp, py, px, col0);
      tp.step(a);
    }
  }
  for (long long item = blockIdx.x; item < a.total_items; item += gridDim.x, ++it, ti.step(a), tn.step(a)) {
    if (pf) {
      if (item + (long long)kPfTiles * gridDim.x < a.total_items) prefetch_tile_operands(a, tp, py, px, col0);
      tp.step(a);
    }
    const int n = ti.n, y0 = ti.ty * a.th, x0 = ti.tx * a.tw;
    const int buf = it & (a.nbuf - 1);
    const int y = y0 + py, x = x0 + px;
    const int oy = y * a.osy + a.oay, ox = x * a.osx + a.oax;
    const bool valid = (y < a.H) && (x < a.W) && (oy < a.Ho) && (ox < a.Wo);
    const long long pix = ((long long)n * a.Ho + oy) * a.Wo + ox;
    if (single) {
#pragma unroll
      for (int j = 0; j < CW / 8; ++j) r1[j] = r2[j];        // this tile's operand, requested one tile ago
      if (item + gridDim.x < a.total_items) load_operand(tn, r2);
    } else {
      // two operands: fetched here, before waiting for the MMAs (they do not depend on the accumulator)
      if (HAS_RES && a.res) {
        const uint4* rp = reinterpret_cast<const uint4*>(a.res + pix * a.res_ld + col0);
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) r1[j] = (valid && col0 + j * 8 < a.Cout) ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
      }
      if (HAS_RES && a.resb) {
        const uint4* rp = reinterpret_cast<const uint4*>(a.resb + pix * a.resb_ld + col0);
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) r2[j] = (valid && col0 + j * 8 < a.Cout) ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
      }
    }
    DP_T(const long long e0 = clock64();)
    tc::mbar_wait(&bars->tmem_full[buf], (it >> (a.nbuf == 4 ? 2 : 1)) & 1);
    tc::fence_after_sync();
    DP_T(const long long e1 = clock64(); dbg_ew += e1 - e0;)
    uint32_t raw[CW];
    tc::tmem_ld_issue<CW>(tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * a.BN + col0), raw);
    tc::tmem_ld_wait();
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&bars->tmem_empty[buf]);   // accumulator is in registers: free it for the next tile
    float v[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(raw[j]);
    if (a.bias) {
#pragma unroll
      for (int j = 0; j < CW; j += 4) {
        if (col0 + j < a.Cout) {   // Cout is a multiple of 8
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + col0 + j));
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
      }
    }
    if (HAS_RES && a.res) {
#pragma unroll
      for (int j = 0; j < CW / 8; ++j) {
        const uint32_t rr[4] = {r1[j].x, r1[j].y, r1[j].z, r1[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[j * 8 + 2 * k] += bf16_lo(rr[k]); v[j * 8 + 2 * k + 1] += bf16_hi(rr[k]); }
      }
    }
    if (HAS_RES && a.resb) {
#pragma unroll
      for (int j = 0; j < CW / 8; ++j) {
        const uint4 rv = r2[j];
        const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[j * 8 + 2 * k] += bf16_lo(rr[k]); v[j * 8 + 2 * k + 1] += bf16_hi(rr[k]); }
      }
    }
    const uint32_t ring_off = (uint32_t)ring * (dual ? 2u : 1u) * a.out_tile_bytes;
    const uint32_t st = s_base + ring_off;
#pragma unroll
    for (int j = 0; j < CW / 8; ++j) {
      uint32_t q[4], q2[4];
      float cx[8];                      // aux mode: the pre-activation values c of this thread's 8 columns
      if (MODE == 2) aux_mask8(&v[j * 8], r1[j], s_aux, a.Cout, col0 + j * 8, a.aux_hi, cx);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x0v = v[j * 8 + 2 * k], x1v = v[j * 8 + 2 * k + 1];
        const uint32_t plain = pack_bf16x2(x0v, x1v);
        const uint32_t act = pack_bf16x2(fmaxf(x0v, 0.f), fmaxf(x1v, 0.f));
        q[k] = a.relu ? act : plain;
        q2[k] = a.relu2 ? act : plain;
        if (want_stats) {
          // statistics of the value as stored (bf16-rounded), so mean/var describe the tensor the consumer reads
          const float f0 = valid ? bf16_lo(plain) : 0.f, f1 = valid ? bf16_hi(plain) : 0.f;
          acc_s[j * 8 + 2 * k] += f0;
          acc_q[j * 8 + 2 * k] = fmaf(f0, MODE == 2 ? cx[2 * k] : f0, acc_q[j * 8 + 2 * k]);
          acc_s[j * 8 + 2 * k + 1] += f1;
          acc_q[j * 8 + 2 * k + 1] = fmaf(f1, MODE == 2 ? cx[2 * k + 1] : f1, acc_q[j * 8 + 2 * k + 1]);
        }
      }
      tc::st_shared_v4(st + soff[j], q[0], q[1], q[2], q[3]);
      if (dual) tc::st_shared_v4(st + a.out_tile_bytes + soff[j], q2[0], q2[1], q2[2], q2[3]);
    }
    tc::fence_proxy_async();                                  // generic-proxy smem writes -> visible to the TMA engine
    // ring of nob staging tiles, one barrier per tile: before the barrier the store thread makes sure at most
    // nob - 2 earlier stores are still reading shared memory, so the buffer the NEXT tile writes is free.
    if (store_thread) {
      if (a.nob == 2) tc::bulk_wait_read<0>();
      else if (a.nob == 3) tc::bulk_wait_read<1>();
      else tc::bulk_wait_read<2>();
    }
    tc::named_bar_sync(1, kEpiThreads);
    if (store_thread) {
      tc::tma_store_4d(&tm.o, s_out + ring_off, 0, x0, y0, n);
      if (dual) tc::tma_store_4d(&tm.o2, s_out + ring_off + a.out_tile_bytes, 0, x0, y0, n);
      tc::bulk_commit();
    }
    if (++ring == a.nob) ring = 0;
    DP_T(dbg_ek += clock64() - e1;)
  }
  if (store_thread) tc::bulk_wait_read<0>();
  DP_T(if (a.dbg && blockIdx.x == 0 && threadIdx.x == kRoleThreads) { a.dbg[5] = dbg_ew; a.dbg[6] = dbg_ek; a.dbg[7] = it; })
  if (want_stats) {
#pragma unroll
    for (int j = 0; j < CW; ++j) {
      const float s = dp::warp_sum(acc_s[j]);
      const float q = dp::warp_sum(acc_q[j]);
      if (lane == 0 && col0 + j < a.Cout) {
        s_stats[(e * 2 + 0) * a.Cout + col0 + j] = s;
        s_stats[(e * 2 + 1) * a.Cout + col0 + j] = q;
      }
    }
  }
}

// ---- epilogue A1: BN in {16, 32}, single output.  Small tiles are bound by the serial latency of one epilogue pass
// (TMEM load -> math -> staging -> store), not by its work, so the eight warps form two independent groups that take
// alternate tiles (group g <-> MMA issuer g <-> TMEM buffers g, g+2).  Each warp owns 32 pixels x all BN columns, stages
// them in its own swizzled ring slot and issues its own TMA store of a (32 / tw) x tw pixel box: no inter-warp barrier.
template <int BNT, int MODE>      // MODE as in epilogue_tma (here both operand modes read per tile, L2-prefetched)
__device__ __forceinline__ void epilogue_tma_warp(const Maps& tm, const ConvArgs& a, Barriers* bars, uint8_t* s_out,
                                                  float* s_stats, const float* s_aux, uint32_t tmem, int warp, int lane) {
  const int e = warp - 4, ew = e & 3, grp = e >> 2;
  const int m = ew * 32 + lane;
  const int py = m / a.tw, px = m - py * a.tw;
  constexpr uint32_t kRowBytes = 2u * BNT;                  // 32 or 64
  constexpr uint32_t kSwz = kRowBytes == 64 ? 3u : 1u;
  constexpr uint32_t kSlotBytes = 32u * kRowBytes;          // one warp's 32-pixel sub-tile
  constexpr int kRing = 4;
  const bool want_stats = a.stats != nullptr;
  float acc_s[BNT], acc_q[BNT];
#pragma unroll
  for (int j = 0; j < BNT; ++j) { acc_s[j] = 0.f; acc_q[j] = 0.f; }
  uint32_t soff[BNT / 8];
#pragma unroll
  for (int j = 0; j < BNT / 8; ++j) {
    const uint32_t off = (uint32_t)lane * kRowBytes + (uint32_t)(j * 16);
    soff[j] = off ^ (((off >> 7) & kSwz) << 4);
  }
  uint8_t* my_ring = s_out + (size_t)e * kRing * kSlotBytes;
  const uint32_t ring_base = tc::smem_u32(my_ring);
  const int rows_per_warp = 32 / a.tw;                      // tw in {8, 16, 32}
  int it = grp, ring = 0;
  TileIter ti;
  ti.init(a, blockIdx.x + (unsigned)grp * gridDim.x);
  constexpr bool has_op = MODE != 0;
  TileIter tp = ti;                                         // this group's tile kPfTiles ahead (groups take alternate tiles)
  if (has_op) {
    for (int k = 0; k < kPfTiles; ++k) {
      if (blockIdx.x + (long long)(grp + 2 * k) * gridDim.x < a.total_items) prefetch_tile_operands(a, tp, py, px, 0);
      tp.step(a); tp.step(a);
    }
  }
  for (long long item = blockIdx.x + (long long)grp * gridDim.x; item < a.total_items;
       item += 2LL * gridDim.x, it += 2, ti.step(a), ti.step(a)) {
    const int n = ti.n, y0 = ti.ty * a.th, x0 = ti.tx * a.tw;
    const int buf = it & (a.nbuf - 1);
    const int y = y0 + py, x = x0 + px;
    const int oy = y * a.osy + a.oay, ox = x * a.osx + a.oax;
    const bool valid = (y < a.H) && (x < a.W) && (oy < a.Ho) && (ox < a.Wo);
    const long long pix = ((long long)n * a.Ho + oy) * a.Wo + ox;
    if (has_op) {
      if (item + 2LL * kPfTiles * gridDim.x < a.total_items) prefetch_tile_operands(a, tp, py, px, 0);
      tp.step(a); tp.step(a);
    }
    uint4 r1[BNT / 8], r2[BNT / 8];
    if (MODE != 0 && a.res) {
      const uint4* rp = reinterpret_cast<const uint4*>(a.res + pix * a.res_ld);
#pragma unroll
      for (int j = 0; j < BNT / 8; ++j) r1[j] = (valid && j * 8 < a.Cout) ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
    }
    if (MODE != 0 && a.resb) {
      const uint4* rp = reinterpret_cast<const uint4*>(a.resb + pix * a.resb_ld);
#pragma unroll
      for (int j = 0; j < BNT / 8; ++j) r2[j] = (valid && j * 8 < a.Cout) ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
    }
    // the ring slot written below was handed to TMA four of this warp's tiles ago: make sure it has been read
    if (lane == 0) tc::bulk_wait_read<kRing - 1>();
    __syncwarp();
    tc::mbar_wait(&bars->tmem_full[buf], (it >> (a.nbuf == 4 ? 2 : 1)) & 1);
    tc::fence_after_sync();
    uint32_t raw[BNT];
    tc::tmem_ld_issue<BNT>(tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * a.BN), raw);
    tc::tmem_ld_wait();
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&bars->tmem_empty[buf]);
    float v[BNT];
#pragma unroll
    for (int j = 0; j < BNT; ++j) v[j] = __uint_as_float(raw[j]);
    if (a.bias) {
#pragma unroll
      for (int j = 0; j < BNT; j += 4) {
        if (j < a.Cout) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + j));
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
      }
    }
    if (MODE != 0 && a.res) {
#pragma unroll
      for (int j = 0; j < BNT / 8; ++j) {
        const uint32_t rr[4] = {r1[j].x, r1[j].y, r1[j].z, r1[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[j * 8 + 2 * k] += bf16_lo(rr[k]); v[j * 8 + 2 * k + 1] += bf16_hi(rr[k]); }
      }
    }
    if (MODE == 1 && a.resb) {
#pragma unroll
      for (int j = 0; j < BNT / 8; ++j) {
        const uint32_t rr[4] = {r2[j].x, r2[j].y, r2[j].z, r2[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[j * 8 + 2 * k] += bf16_lo(rr[k]); v[j * 8 + 2 * k + 1] += bf16_hi(rr[k]); }
      }
    }
    const uint32_t st = ring_base + (uint32_t)ring * kSlotBytes;
#pragma unroll
    for (int j = 0; j < BNT / 8; ++j) {
      uint32_t q[4];
      float cx[8];
      if (MODE == 2) aux_mask8(&v[j * 8], r2[j], s_aux, a.Cout, j * 8, a.aux_hi, cx);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x0v = v[j * 8 + 2 * k], x1v = v[j * 8 + 2 * k + 1];
        const uint32_t plain = pack_bf16x2(x0v, x1v);
        q[k] = a.relu ? pack_bf16x2(fmaxf(x0v, 0.f), fmaxf(x1v, 0.f)) : plain;
        if (want_stats) {
          const float f0 = valid ? bf16_lo(plain) : 0.f, f1 = valid ? bf16_hi(plain) : 0.f;
          acc_s[j * 8 + 2 * k] += f0;
          acc_q[j * 8 + 2 * k] = fmaf(f0, MODE == 2 ? cx[2 * k] : f0, acc_q[j * 8 + 2 * k]);
          acc_s[j * 8 + 2 * k + 1] += f1;
          acc_q[j * 8 + 2 * k + 1] = fmaf(f1, MODE == 2 ? cx[2 * k + 1] : f1, acc_q[j * 8 + 2 * k + 1]);
        }
      }
      tc::st_shared_v4(st + soff[j], q[0], q[1], q[2], q[3]);
    }
    tc::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tc::tma_store_4d(&tm.o2, my_ring + (size_t)ring * kSlotBytes, 0, x0, y0 + ew * rows_per_warp, n);
      tc::bulk_commit();
    }
    if (++ring == kRing) ring = 0;
  }
  if (lane == 0) tc::bulk_wait_read<0>();
  if (want_stats) {
#pragma unroll
    for (int j = 0; j < BNT; ++j) {
      const float s = dp::warp_sum(acc_s[j]);
      const float q = dp::warp_sum(acc_q[j]);
      if (lane == 0 && j < a.Cout) {
        s_stats[(e * 2 + 0) * a.Cout + j] = s;
        s_stats[(e * 2 + 1) * a.Cout + j] = q;
      }
    }
  }
}

// ---- epilogue A2: BN a multiple of 64 (> 64), single output.  The tile leaves in 64-column sub-blocks: per
// sub-block each of the eight warps takes 32 pixels x 32 columns (one tcgen05.ld), stages bf16 rows in a 128B-swizzled
// [128 px][64 ch] tile and one thread TMA-stores it at channel offset nb*BN + sb*64 (channels >= Cout are clipped by the
// store).  A thread's direct stores would touch 32 B out of every Cout*2-byte pixel row - the wide "expand" pointwise
// layers of the encoder trunk ran at ~1.4 TB/s that way.  BN statistics use the shuffle transpose (columns are too
// many to keep per-thread accumulators).
__device__ __forceinline__ void epilogue_tma_wide(const Maps& tm, const ConvArgs& a, Barriers* bars, uint8_t* s_out,
                                                  float* s_stats, uint32_t tmem, int warp, int lane) {
  const int e = warp - 4, ew = e & 3, half = e >> 2;
  const int m = ew * 32 + lane;
  const int py = m / a.tw, px = m - py * a.tw;
  const bool store_thread = threadIdx.x == kRoleThreads;
  const int et = threadIdx.x - kRoleThreads;                  // 0 .. kEpiThreads - 1
  uint32_t soff[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t off = (uint32_t)m * 128u + (uint32_t)(half * 64 + j * 16);
    soff[j] = off ^ (((off >> 7) & 7u) << 4);
  }
  const uint32_t s_base = tc::smem_u32(s_out);
  const int nsub = a.BN / 64;
  float* grow = a.stats ? a.stats + (size_t)blockIdx.x * 2 * a.Cout : nullptr;    // this CTA's row of the partials
  if (a.stats) {
    for (int i = et; i < 2 * a.Cout; i += kEpiThreads) grow[i] = 0.f;
    tc::named_bar_sync(1, kEpiThreads);
  }
  // s_stats holds the sums of ONE N block: when the CTA moves to another N block (and at the end) the eight warps' sums
  // are folded into the CTA's row of the partials (fixed order; every global slot has one owner thread).  With 1, 2 or 4
  // N blocks a CTA's tiles all share one (148 is a multiple), so the fold runs once.
  auto flush = [&](int fnb) {
    tc::named_bar_sync(1, kEpiThreads);
    for (int i = et; i < 2 * a.BN; i += kEpiThreads) {
      const int which = i / a.BN, c = i - which * a.BN;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kEpiWarps; ++w) { sum += s_stats[(w * 2 + which) * a.BN + c]; s_stats[(w * 2 + which) * a.BN + c] = 0.f; }
      const int gcol = fnb * a.BN + c;
      if (gcol < a.Cout) grow[which * a.Cout + gcol] += sum;
    }
    tc::named_bar_sync(1, kEpiThreads);
  };
  int cur_nb = -1;
  int it = 0, ring = 0;
  TileIter ti;
  ti.init(a, blockIdx.x);
  for (long long item = blockIdx.x; item < a.total_items; item += gridDim.x, ++it, ti.step(a)) {
    const int n = ti.n, y0 = ti.ty * a.th, x0 = ti.tx * a.tw, nb = ti.nb;
    if (a.stats && cur_nb >= 0 && nb != cur_nb) flush(cur_nb);
    cur_nb = nb;
    const int buf = it & (a.nbuf - 1);
    const int y = y0 + py, x = x0 + px;
    const int oy = y * a.osy + a.oay, ox = x * a.osx + a.oax;
    const bool valid = (y < a.H) && (x < a.W) && (oy < a.Ho) && (ox < a.Wo);
    const long long pix = ((long long)n * a.Ho + oy) * a.Wo + ox;
    // sub-blocks that hold real output channels (the last N block of e.g. Cout = 816 = 3 x 256 + 48 has one of four)
    int nsv = (a.Cout - nb * a.BN + 63) / 64;
    nsv = nsv < 0 ? 0 : (nsv > nsub ? nsub : nsv);
    const uint32_t t_base = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * a.BN + half * 32);
    tc::mbar_wait(&bars->tmem_full[buf], (it >> (a.nbuf == 4 ? 2 : 1)) & 1);
    tc::fence_after_sync();
    uint32_t raw[32];
    if (nsv > 0) tc::tmem_ld_issue<32>(t_base, raw);
    else {
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->tmem_empty[buf]);
    }
    for (int sb = 0; sb < nsv; ++sb) {
      const int n0 = nb * a.BN + sb * 64 + half * 32;       // first output channel of this thread's 32 columns
      tc::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
      // the next sub-block's accumulator columns travel while this one is packed, staged, stored and summed
      if (sb + 1 < nsv) tc::tmem_ld_issue<32>(t_base + (uint32_t)((sb + 1) * 64), raw);
      else {
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->tmem_empty[buf]);
      }
      if (n0 < a.Cout) {                                        // warp-uniform; Cout is a multiple of 8
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (n0 + j < a.Cout) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        }
        if (a.res && valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.res + pix * a.res_ld + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n0 + j * 8 < a.Cout) {
              const uint4 r = __ldg(rp + j);
              const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) { v[j * 8 + 2 * k] += bf16_lo(rr[k]); v[j * 8 + 2 * k + 1] += bf16_hi(rr[k]); }
            }
          }
        }
        if (a.resb && valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.resb + pix * a.resb_ld + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n0 + j * 8 < a.Cout) {
              const uint4 r = __ldg(rp + j);
              const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) { v[j * 8 + 2 * k] += bf16_lo(rr[k]); v[j * 8 + 2 * k + 1] += bf16_hi(rr[k]); }
            }
          }
        }
      }
      uint32_t q[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float x0v = v[2 * k], x1v = v[2 * k + 1];
        if (a.relu) { x0v = fmaxf(x0v, 0.f); x1v = fmaxf(x1v, 0.f); }
        q[k] = pack_bf16x2(x0v, x1v);
      }
      if (a.stats && !valid) {        // overhanging pixels are clipped by the store; zero them for the column sums below
#pragma unroll
        for (int k = 0; k < 16; ++k) q[k] = 0u;
      }
      const uint32_t st = s_base + (uint32_t)ring * 16384u;
#pragma unroll
      for (int j = 0; j < 4; ++j) tc::st_shared_v4(st + soff[j], q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
      tc::fence_proxy_async();
      if (store_thread) {
        if (a.nob == 2) tc::bulk_wait_read<0>();
        else tc::bulk_wait_read<1>();
      }
      tc::named_bar_sync(1, kEpiThreads);
      if (store_thread) {
        tc::tma_store_4d(&tm.o, s_out + (size_t)ring * 16384u, nb * a.BN + sb * 64, x0, y0, n);
        tc::bulk_commit();
      }
      if (a.stats) {
        // BatchNorm partial sums from the staged tile (the values as stored): warp e sums pixel rows [16e, 16e+16),
        // lane l the channel pair (2l, 2l+1) - one conflict-free 128-byte row per load, no shuffles.  Slot (e, column)
        // of s_stats has a single owner, so the += needs no atomics.
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t r = (uint32_t)(e * 16 + i);
          const uint32_t addr = st + r * 128u + ((((uint32_t)lane >> 2) ^ (r & 7u)) << 4) + ((uint32_t)lane & 3u) * 4u;
          uint32_t u;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(addr));
          const float f0 = bf16_lo(u), f1 = bf16_hi(u);
          s0 += f0; s1 += f1;
          q0 = fmaf(f0, f0, q0); q1 = fmaf(f1, f1, q1);
        }
        const int col = sb * 64 + 2 * lane;                     // column within the N block
        s_stats[(e * 2 + 0) * a.BN + col] += s0;
        s_stats[(e * 2 + 0) * a.BN + col + 1] += s1;
        s_stats[(e * 2 + 1) * a.BN + col] += q0;
        s_stats[(e * 2 + 1) * a.BN + col + 1] += q1;
      }
      if (++ring == a.nob) ring = 0;
    }
  }
  if (a.stats && cur_nb >= 0) flush(cur_nb);
  if (store_thread) tc::bulk_wait_read<0>();
}

// ---- epilogue B (any BN): 16-column chunks alternate between the two warps that share a TMEM lane quarter; each
// thread writes its pixel's 32 bytes straight to global memory.
__device__ __forceinline__ void epilogue_direct(const ConvArgs& a, Barriers* bars, float* s_stats, uint32_t tmem,
                                                int warp, int lane) {
  const int e = warp - 4, ew = e & 3, half = e >> 2;
  const int m = ew * 32 + lane;
  const int py = m / a.tw, px = m - py * a.tw;
  int it = 0;
  TileIter ti;
  ti.init(a, blockIdx.x);
  for (long long item = blockIdx.x; item < a.total_items; item += gridDim.x, ++it, ti.step(a)) {
    const int n = ti.n, y0 = ti.ty * a.th, x0 = ti.tx * a.tw, nb = ti.nb;
    const int buf = it & (a.nbuf - 1);
    const int y = y0 + py, x = x0 + px;
    const int oy = y * a.osy + a.oay, ox = x * a.osx + a.oax;
    const bool valid = (y < a.H) && (x < a.W) && (oy < a.Ho) && (ox < a.Wo);
    const long long pix = ((long long)n * a.Ho + oy) * a.Wo + ox;
    tc::mbar_wait(&bars->tmem_full[buf], (it >> (a.nbuf == 4 ? 2 : 1)) & 1);
    tc::fence_after_sync();
    const uint32_t t_base = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * a.BN);
    for (int c0 = half * 16; c0 < a.BN; c0 += 32) {
      const int n0 = nb * a.BN + c0;
      if (n0 >= a.Cout) break;  // warp-uniform
      const int nvalid = min(16, a.Cout - n0);
      float v[16];
      tc::tmem_ld16(t_base + c0, v);
      if (a.bias) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nvalid) v[j] += __ldg(a.bias + n0 + j);
      }
      if (a.res && valid) {
        const uint4* rp = reinterpret_cast<const uint4*>(a.res + pix * a.res_ld + n0);
        uint4 r0 = __ldg(rp);
        uint4 r1 = nvalid > 8 ? __ldg(rp + 1) : make_uint4(0, 0, 0, 0);
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[2 * j] += bf16_lo(rr[j]); v[2 * j + 1] += bf16_hi(rr[j]); }
      }
      if (a.resb && valid) {
        const uint4* rp = reinterpret_cast<const uint4*>(a.resb + pix * a.resb_ld + n0);
        uint4 r0 = __ldg(rp);
        uint4 r1 = nvalid > 8 ? __ldg(rp + 1) : make_uint4(0, 0, 0, 0);
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[2 * j] += bf16_lo(rr[j]); v[2 * j + 1] += bf16_hi(rr[j]); }
      }
      if (a.stats) {
        float sv[16], sq[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // statistics of the value as stored (bf16-rounded), so mean/var describe the tensor the consumer reads
          const float q = valid ? __bfloat162float(__float2bfloat16_rn(v[j])) : 0.f;
          sv[j] = q;
          sq[j] = q * q;
        }
        const float cs = transpose_reduce16(sv, lane);
        const float cq = transpose_reduce16(sq, lane);
        if ((lane & 1) == 0) {
          const int col = n0 + transpose_reduce16_col(lane);
          if (col < a.Cout) {
            s_stats[(e * 2 + 0) * a.Cout + col] += cs;
            s_stats[(e * 2 + 1) * a.Cout + col] += cq;
          }
        }
      }
      if (valid) {
        if (a.out2) {
          uint32_t q[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x0v = v[2 * j], x1v = v[2 * j + 1];
            if (a.relu2) { x0v = fmaxf(x0v, 0.f); x1v = fmaxf(x1v, 0.f); }
            q[j] = pack_bf16x2(x0v, x1v);
          }
          uint4* op = reinterpret_cast<uint4*>(a.out2 + pix * a.out2_ld + n0);
          op[0] = make_uint4(q[0], q[1], q[2], q[3]);
          if (nvalid > 8) op[1] = make_uint4(q[4], q[5], q[6], q[7]);
        }
        if (a.out) {
          uint32_t q[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x0v = v[2 * j], x1v = v[2 * j + 1];
            if (a.relu) { x0v = fmaxf(x0v, 0.f); x1v = fmaxf(x1v, 0.f); }
            q[j] = pack_bf16x2(x0v, x1v);
          }
          uint4* op = reinterpret_cast<uint4*>(a.out + pix * a.out_ld + n0);
          op[0] = make_uint4(q[0], q[1], q[2], q[3]);
          if (nvalid > 8) op[1] = make_uint4(q[4], q[5], q[6], q[7]);
        }
      }
    }
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&bars->tmem_empty[buf]);
  }
}

// kPre: the BatchNorm-prologue variant adds a fourth warpgroup (warps 12..15) that transforms the landed A boxes; the
// register file is re-balanced with setmaxnreg so that the epilogue warpgroups keep their 168 registers at 512 threads.
constexpr int kPreThreads = kThreads + 128;
template <bool kPre>
__global__ void __launch_bounds__(kPre ? kPreThreads : kThreads, 1)
conv_tc_kernel(const __grid_constant__ Maps tm, const __grid_constant__ ConvArgs a) {
  dp::pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // layout: [resident weights][output staging tiles][stages x (A | B)][stats][barriers]
  uint8_t* s_res = smem;
  uint8_t* s_out = smem + a.resident_bytes;
  const uint32_t out_bytes = a.epi_tma == 3 ? (uint32_t)(kEpiWarps * 4 * 32 * a.BN * 2)
                             : a.epi_tma == 2 ? (uint32_t)a.nob * 16384u
                                             : (a.epi_tma ? (uint32_t)a.nob * (a.out2 ? 2u : 1u) * a.out_tile_bytes : 0u);
  uint8_t* s_stage = s_out + out_bytes;
  const uint32_t stage_bytes = a.a_slot_bytes + (a.resident ? 0u : (uint32_t)a.max_nr * a.b_tap_bytes);
  float* s_stats = reinterpret_cast<float*>(s_stage + (size_t)a.stages * stage_bytes);
  const int stats_floats = a.stats ? kEpiWarps * 2 * (a.epi_tma == 2 ? a.BN : a.Cout) : 0;   // wide epilogue: per N block, flushed per tile
  float* s_pre = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_stats) + ((stats_floats * 4 + 15) & ~15));
  const int pre_floats = a.pre_ss ? 2 * a.pre_pad : 0;
  float* s_aux = s_pre + pre_floats;              // pre_pad is a multiple of 16: s_aux stays 16-byte aligned
  const int aux_floats = a.aux_mode ? 2 * a.Cout : 0;
  Barriers* bars = reinterpret_cast<Barriers*>(reinterpret_cast<uint8_t*>(s_aux) + ((aux_floats * 4 + 15) & ~15));

  const int warp = tc::warp_idx_uniform();
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) { tc::mbar_init(&bars->full[i], 1); tc::mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&bars->tmem_full[i], 1);
      tc::mbar_init(&bars->tmem_empty[i], a.epi_tma == 3 ? kEpiWarps / 2 : kEpiWarps);   // per-warp mode: one group per buffer
    }
    tc::mbar_init(&bars->resident_full, 1);
    for (int i = 0; i < a.stages; ++i) tc::mbar_init(&bars->ready[i], 4);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tm.a[0]);
    tc::prefetch_tmap(&tm.b);
    if (a.epi_tma) tc::prefetch_tmap(a.epi_tma == 3 ? &tm.o2 : &tm.o);
  }
  if (warp == 2) tc::tmem_alloc(&bars->tmem_base, kTmemCols);
  for (int i = threadIdx.x; i < stats_floats; i += blockDim.x) s_stats[i] = 0.f;
  dp::pdl_wait();   // everything above overlaps the previous kernel's tail; global memory is touched from here on
  for (int i = threadIdx.x; i < pre_floats; i += blockDim.x) {      // [scale | shift], zero beyond the real channels
    const int which = i / a.pre_pad, c = i - which * a.pre_pad;
    s_pre[i] = c < a.pre_c ? __ldg(a.pre_ss + which * a.pre_c + c) : 0.f;
  }
  for (int i = threadIdx.x; i < aux_floats; i += blockDim.x) s_aux[i] = __ldg(a.aux_ss + i);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = bars->tmem_base;
  DP_T(const long long t_start = clock64();)
  // kPre: 512 threads x 128 registers at launch; each role branch re-sizes its warpgroup's share first thing:
  // (56 + 168 + 168 + 104) x 128 threads = 63488 <= 65536
  if (warp == 0) {
    if constexpr (kPre) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // ================= TMA producer (one elected lane) =================
    if (tc::elect_one_sync()) {
    if (a.resident) {
      tc::mbar_expect_tx(&bars->resident_full, (uint32_t)(a.nslots * a.kchunks) * a.b_box_bytes);
      if (a.halo) {
        for (int tap = 0; tap < 9; ++tap)
          for (int kc = 0; kc < a.kchunks; ++kc)
            tc::tma_load_3d(s_res + (size_t)(tap * a.kchunks + kc) * a.b_tap_bytes, &tm.b, &bars->resident_full,
                            kc * a.KB, 0, tap);
      } else {
        for (int c = 0; c < a.ncols; ++c)
          for (int r = 0; r < a.cols[c].nr; ++r)
            for (int kc = 0; kc < a.kchunks; ++kc)
              tc::tma_load_3d(s_res + (size_t)(a.cols[c].wslot[r] * a.kchunks + kc) * a.b_tap_bytes, &tm.b,
                              &bars->resident_full, kc * a.KB, 0, a.cols[c].wtap[r]);
      }
    }
    uint32_t stage = 0, phase = 0;
    TileIter ti;
    ti.init(a, blockIdx.x);
    for (long long item = blockIdx.x; item < a.total_items; item += gridDim.x, ti.step(a)) {
      const int n = ti.n, y0 = ti.ty * a.th, x0 = ti.tx * a.tw, nb = ti.nb;
      for (int c = 0; c < a.ncols; ++c) {
        const ColLoad& col = a.cols[c];
        for (int kc = 0; kc < a.kchunks; ++kc) {
          tc::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sA = s_stage + (size_t)stage * stage_bytes;
          tc::mbar_expect_tx(&bars->full[stage], col.box_bytes + (a.resident ? 0u : (uint32_t)col.nr * a.b_box_bytes));
          tc::tma_load_4d(sA, &tm.a[col.map], &bars->full[stage], kc * a.KB, x0 + col.dx, y0 + col.dy, n);
          if (!a.resident) {
            uint8_t* sB = sA + a.a_slot_bytes;
            for (int r = 0; r < col.nr; ++r)
              tc::tma_load_3d(sB + (size_t)r * a.b_tap_bytes, &tm.b, &bars->full[stage], kc * a.KB, nb * a.BN,
                              col.wtap[r]);
          }
          if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    }
  } else if (warp == 1 || (warp == 3 && a.niss == 2)) {
    if constexpr (kPre) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (tc::elect_one_sync()) {
    // ================= MMA issuers (one thread each; two warps alternate tiles) =================
    // A lone thread needs ~50 cycles of scalar work per tcgen05.mma, more than a 128xNx16 UMMA with N <= 64 takes on
    // the tensor pipe, so two issuers work on alternate tiles with their own TMEM accumulators.
    if (a.resident) { tc::mbar_wait(&bars->resident_full, 0); tc::fence_after_sync(); }
    const int wiss = warp == 1 ? 0 : 1;
    const int kk_n = a.KB / 16;
    const uint32_t b_hi = tc::desc_hi(a.sbo, a.layout);                                   // also A's hi outside halo mode
    const uint32_t a_hi_halo = tc::desc_hi((uint32_t)a.pitch * a.row_bytes, a.layout);
    const uint32_t res_base = tc::smem_u32(s_res);
    const int ksteps = a.ncols * a.kchunks;
    DP_T(long long dbg_te = 0; long long dbg_wf = 0; long long dbg_is = 0;)
    int it = wiss;
    // stage ring position of this issuer's next tile: tile `it` starts at global k-step it * ksteps
    uint32_t stage = 0, phase = 0;
    for (int k = 0; k < wiss * ksteps; ++k)
      if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
    const int nbuf_mask = a.nbuf - 1, nbuf_shift = a.nbuf == 4 ? 2 : 1;
    for (long long item = blockIdx.x + (long long)wiss * gridDim.x; item < a.total_items;
         item += (long long)a.niss * gridDim.x, it += a.niss) {
      const int buf = it & nbuf_mask;
      DP_T(const long long q0 = clock64();)
      tc::mbar_wait(&bars->tmem_empty[buf], ((it >> nbuf_shift) & 1) ^ 1);
      tc::fence_after_sync();
      DP_T(dbg_te += clock64() - q0;)
      const uint32_t d_tmem = tmem + (uint32_t)(buf * a.BN);
      uint32_t accumulate = 0;
      for (int c = 0; c < a.ncols; ++c) {
        const ColLoad& col = a.cols[c];
        for (int kc = 0; kc < a.kchunks; ++kc) {
          DP_T(const long long q1 = clock64();)
          tc::mbar_wait(kPre ? &bars->ready[stage] : &bars->full[stage], phase);
          tc::fence_after_sync();
          DP_T(const long long q2 = clock64(); dbg_wf += q2 - q1;)
          const uint32_t a_base = tc::smem_u32(s_stage + (size_t)stage * stage_bytes);
          if (a.halo) {
            // all nine taps read the same halo box: tap (r,s) starts (r*pitch + s) pixel rows in; the 8-pixel patch
            // rows are `pitch` rows apart (SBO).  The swizzle is a function of the absolute smem address, so neither
            // offset needs to be atom aligned (tests/test_umma_probe_gpu.py::test_k_major_halo_addressing).
            const uint32_t a_lo0 = tc::desc_lo(a_base, 16);
            const uint32_t b_lo0 = tc::desc_lo(res_base + (uint32_t)kc * a.b_tap_bytes, 16);
            const uint32_t row16 = a.row_bytes >> 4, btap16 = (a.b_tap_bytes * (uint32_t)a.kchunks) >> 4;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
              for (int s3 = 0; s3 < 3; ++s3) {
                const uint32_t al = a_lo0 + (uint32_t)(r * a.pitch + s3) * row16;
                const uint32_t bl = b_lo0 + (uint32_t)(r * 3 + s3) * btap16;
                tc::umma_bf16_lohi(d_tmem, al, a_hi_halo, bl, b_hi, a.idesc, accumulate);
                accumulate = 1;
                if (kk_n > 1) tc::umma_bf16_lohi(d_tmem, al + 2, a_hi_halo, bl + 2, b_hi, a.idesc, 1);
                if (kk_n > 2) {
                  tc::umma_bf16_lohi(d_tmem, al + 4, a_hi_halo, bl + 4, b_hi, a.idesc, 1);
                  tc::umma_bf16_lohi(d_tmem, al + 6, a_hi_halo, bl + 6, b_hi, a.idesc, 1);
                }
              }
            }
          } else {
            const uint32_t b_base = a.resident ? res_base + (uint32_t)kc * a.b_tap_bytes : a_base + a.a_slot_bytes;
            for (int r = 0; r < col.nr; ++r) {
              const uint32_t al = tc::desc_lo(a_base + (uint32_t)(r * a.tw) * a.row_bytes, 16);
              const uint32_t bl = tc::desc_lo(a.resident ? b_base + (uint32_t)(col.wslot[r] * a.kchunks) * a.b_tap_bytes
                                                         : b_base + (uint32_t)r * a.b_tap_bytes, 16);
              tc::umma_bf16_lohi(d_tmem, al, b_hi, bl, b_hi, a.idesc, accumulate);
              accumulate = 1;
              if (kk_n > 1) tc::umma_bf16_lohi(d_tmem, al + 2, b_hi, bl + 2, b_hi, a.idesc, 1);
              if (kk_n > 2) {
                tc::umma_bf16_lohi(d_tmem, al + 4, b_hi, bl + 4, b_hi, a.idesc, 1);
                tc::umma_bf16_lohi(d_tmem, al + 6, b_hi, bl + 6, b_hi, a.idesc, 1);
              }
            }
          }
          tc::umma_commit(&bars->empty[stage]);  // frees the smem slot when these MMAs have read it
          DP_T(dbg_is += clock64() - q2;)
          if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
        }
      }
      tc::umma_commit(&bars->tmem_full[buf]);  // accumulator complete -> epilogue
      for (int k = 0; k < (a.niss - 1) * ksteps; ++k)   // skip the other issuer's tile in the stage ring
        if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
    }
    DP_T(if (a.dbg && blockIdx.x == 0 && wiss == 0) { a.dbg[0] = dbg_te; a.dbg[1] = dbg_wf; a.dbg[2] = dbg_is; a.dbg[4] = it; })
    }
  } else if (warp < 4) {
    if constexpr (kPre) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");     // idle warps of the role warpgroup
  } else if (kPre && warp >= 12) {
    // ================= prologue transform (BatchNorm + activation on the landed A boxes) =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    const int t64 = (warp - 12) * 32 + lane;
    uint32_t stage = 0, phase = 0;
    TileIter ti;
    ti.init(a, blockIdx.x);
    for (long long item = blockIdx.x; item < a.total_items; item += gridDim.x, ti.step(a)) {
      const int y0 = ti.ty * a.th, x0 = ti.tx * a.tw;
      for (int c = 0; c < a.ncols; ++c) {
        const ColLoad& col = a.cols[c];
        const int boxW = a.halo ? a.tw + 2 : a.tw;
        const int npx = (a.halo ? a.th + 2 : a.th + col.nr - 1) * boxW;
        for (int kc = 0; kc < a.kchunks; ++kc) {
          tc::mbar_wait(&bars->full[stage], phase);
          const uint32_t base = tc::smem_u32(s_stage + (size_t)stage * stage_bytes);
          if (a.row_bytes == 128)
            transform_box<128, 128>(base, npx, boxW, x0 + col.dx, y0 + col.dy, a.pW[col.map], a.pH[col.map], s_pre, a.pre_pad,
                               kc * a.KB, a.pre_act, t64);
          else if (a.row_bytes == 64)
            transform_box<64, 128>(base, npx, boxW, x0 + col.dx, y0 + col.dy, a.pW[col.map], a.pH[col.map], s_pre, a.pre_pad,
                              kc * a.KB, a.pre_act, t64);
          else
            transform_box<32, 128>(base, npx, boxW, x0 + col.dx, y0 + col.dy, a.pW[col.map], a.pH[col.map], s_pre, a.pre_pad,
                              kc * a.KB, a.pre_act, t64);
          tc::fence_proxy_async();              // generic-proxy writes -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars->ready[stage]);
          if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ================= epilogue: TMEM -> registers -> (smem -> TMA store | global) =================
    if constexpr (kPre) asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    const int mode = a.aux_mode ? 2 : ((a.res || a.resb) ? 1 : 0);
#define DP_EPI(FN, N)                                                                       \
  do {                                                                                      \
    if (mode == 0) FN<N, 0>(tm, a, bars, s_out, s_stats, s_aux, tmem, warp, lane);          \
    else if (mode == 1) FN<N, 1>(tm, a, bars, s_out, s_stats, s_aux, tmem, warp, lane);     \
    else FN<N, 2>(tm, a, bars, s_out, s_stats, s_aux, tmem, warp, lane);                    \
  } while (0)
    if (a.epi_tma == 3) {
      if (a.BN == 32) DP_EPI(epilogue_tma_warp, 32);
      else DP_EPI(epilogue_tma_warp, 16);
    } else if (a.epi_tma == 2) {
      epilogue_tma_wide(tm, a, bars, s_out, s_stats, tmem, warp, lane);
    } else if (a.epi_tma) {
      if (a.BN == 64) DP_EPI(epilogue_tma, 32);
      else if (a.BN == 32) DP_EPI(epilogue_tma, 16);
      else DP_EPI(epilogue_tma, 8);
#undef DP_EPI
    } else {
      epilogue_direct(a, bars, s_stats, tmem, warp, lane);
    }
    if (a.stats && a.epi_tma != 2) {       // (the wide epilogue has flushed its sums tile by tile)
      tc::named_bar_sync(1, kEpiThreads);  // the epilogue warps only
      float* dst = a.stats + (size_t)blockIdx.x * 2 * a.Cout;
      for (int i = threadIdx.x - kRoleThreads; i < 2 * a.Cout; i += kEpiThreads) {
        const int which = i / a.Cout, col = i - which * a.Cout;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kEpiWarps; ++w) s += s_stats[(w * 2 + which) * a.Cout + col];
        dst[i] = s;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  DP_T(if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[3] = (unsigned long long)(clock64() - t_start);)
  if (warp == 2) tc::tmem_dealloc(tmem, kTmemCols);
}

struct Plane {   // a (possibly strided) pixel-grid view of an NHWC bf16 tensor; strides in elements
  const void* base;
  long long ld_px, ld_row, ld_img;
  int Hp, Wp;
};
struct ColSpec { int plane, dx, dy, nr, wtap[3]; };
struct OutMap { int Ho, Wo, osy, osx, oay, oax; };
struct Epilogue {
  const float* bias;
  const void* res; long long res_ld;
  const void* res2; long long res2_ld;
  int relu; void* out; long long out_ld; void* out2; long long out2_ld; int relu2;
  float* stats;
  // fused BatchNorm (dp_conv_fuse_t): prologue scale/shift/activation of the input, backward mask + sums of the output
  const float* pre_ss = nullptr; int pre_act = 0;
  const void* mask_x = nullptr; long long mask_ld = 0; const float* mask_ss = nullptr; int mask_act = 0;
};

struct Plan {
  ConvArgs a;
  size_t smem;
  int grid;
};

// geometry that does not depend on pointers: tile shape, K/N blocking, stages, grid
int make_plan(Plan& p, int B, int Hg, int Wg, int Cin, int Cout, const ColSpec* cols, int ncols, int want_stats,
              int allow_halo = 0, int n_out = 1, int pre = 0, int aux = 0) {
  ConvArgs& a = p.a;
  if (ncols < 1 || ncols > kMaxCols) return dp_set_error(DP_ERR_UNSUPPORTED, "conv_tc: %d column loads", ncols);
  a.B = B; a.H = Hg; a.W = Wg; a.Cout = Cout;
  // patch shape: th*tw == 128, tw a multiple of 8 (swizzle-atom alignment of the vertical tap offsets)
  const int cand[5][2] = {{8, 16}, {16, 8}, {4, 32}, {2, 64}, {1, 128}};
  long long best = -1;
  for (int i = 0; i < 5; ++i) {
    long long t = (long long)dp::ceil_div(Hg, cand[i][0]) * dp::ceil_div(Wg, cand[i][1]);
    if (best < 0 || t < best) { best = t; a.th = cand[i][0]; a.tw = cand[i][1]; }
  }
  a.KB = Cin > 32 ? 64 : (Cin > 16 ? 32 : 16);
  a.kchunks = dp::ceil_div(Cin, a.KB);
  // N blocking: every N block re-reads the activation tile, so pointwise (1x1) layers - bandwidth-bound, small weight
  // tiles - take up to 256 columns per block (two TMEM accumulators); 3x3 layers stream three weight taps per stage
  // and stay at <= 160 columns so that the pipeline keeps several stages.
  const bool pointwise = (ncols == 1 && cols[0].nr == 1);
  const int bn_cap = pointwise ? 256 : 160;
  a.n_blocks = dp::ceil_div(Cout, bn_cap);
  a.BN = ((dp::ceil_div(Cout, a.n_blocks) + 15) / 16) * 16;
  if (pointwise && a.BN > 64) a.BN = ((a.BN + 63) / 64) * 64;     // wide TMA-store epilogue works in 64-column sub-blocks
  a.row_bytes = a.KB * 2;
  // halo mode (plain 3x3, whole weight set resident): 16x8 patches, one 18x10-pixel box per channel chunk
  a.halo = 0;
  a.pitch = 0;
  if (allow_halo && a.n_blocks == 1) {
    const size_t w9 = (size_t)9 * a.kchunks * (((size_t)a.BN * a.row_bytes + 1023) & ~size_t(1023));
    if (w9 <= 100 * 1024) { a.halo = 1; a.th = 16; a.tw = 8; a.pitch = a.tw + 2; }
  }
  a.tiles_y = dp::ceil_div(Hg, a.th);
  a.tiles_x = dp::ceil_div(Wg, a.tw);
  a.layout = tc::swizzle_for_row_bytes(a.row_bytes);
  a.sbo = 8 * a.row_bytes;
  a.idesc = tc::make_idesc_bf16(128, a.BN, 0, 0);
  a.nbuf = 4 * a.BN <= kTmemCols ? 4 : 2;
  a.niss = a.nbuf == 4 ? 2 : 1;
  a.b_box_bytes = (uint32_t)(a.BN * a.row_bytes);
  a.b_tap_bytes = (a.b_box_bytes + 1023u) & ~1023u;
  a.ncols = ncols;
  a.max_nr = 0;
  a.nslots = 0;
  uint32_t max_box = 0;
  if (a.halo) {
    a.ncols = ncols = 1;
    ColLoad& d = a.cols[0];
    d.map = 0; d.dx = -1; d.dy = -1; d.nr = 3;
    for (int r = 0; r < 3; ++r) { d.wtap[r] = 0; d.wslot[r] = 0; }
    d.box_bytes = (uint32_t)((a.th + 2) * (a.tw + 2)) * a.row_bytes;
    max_box = d.box_bytes;
    a.max_nr = 3;
    a.nslots = 9;
  } else
  for (int c = 0; c < ncols; ++c) {
    ColLoad& d = a.cols[c];
    d.map = c; d.dx = cols[c].dx; d.dy = cols[c].dy; d.nr = cols[c].nr;
    if (d.nr < 1 || d.nr > 3) return dp_set_error(DP_ERR_UNSUPPORTED, "conv_tc: %d vertical taps in one box", d.nr);
    for (int r = 0; r < 3; ++r) { d.wtap[r] = r < d.nr ? cols[c].wtap[r] : 0; d.wslot[r] = r < d.nr ? a.nslots++ : 0; }
    d.box_bytes = (uint32_t)((a.th + d.nr - 1) * a.tw) * a.row_bytes;
    if (d.box_bytes > max_box) max_box = d.box_bytes;
    if (d.nr > a.max_nr) a.max_nr = d.nr;
  }
  a.a_slot_bytes = (max_box + 1023u) & ~1023u;
  const size_t all_w = (size_t)a.nslots * a.kchunks * a.b_tap_bytes;
  a.resident = (a.n_blocks == 1 && all_w <= 100 * 1024) ? 1 : 0;
  a.resident_bytes = a.resident ? (uint32_t)all_w : 0u;
  a.pre_pad = a.kchunks * a.KB;
  const size_t side_bytes = (pre ? (size_t)2 * a.pre_pad * 4 : 0) + (aux ? ((size_t)2 * Cout * 4 + 15) & ~size_t(15) : 0);
  size_t stats_bytes = (want_stats ? ((size_t)kEpiWarps * 2 * Cout * 4 + 15) & ~size_t(15) : 0) + side_bytes;
  // the wide epilogue keeps BatchNorm partial sums for one N block only and adds them to the CTA's row of the global
  // partials after every tile ([8 warps][2][Cout] floats would be 52 KB at Cout = 816 and push those layers off it)
  const size_t stats_bytes_wide = (want_stats ? (size_t)kEpiWarps * 2 * a.BN * 4 : 0) + side_bytes;
  // TMA-store epilogue: one N block of 16/32/64 columns; two staging tiles per output
  a.epi_tma = (a.n_blocks == 1 && (a.BN == 16 || a.BN == 32 || a.BN == 64) && n_out >= 1) ? 1 : 0;
  a.out_tile_bytes = (uint32_t)(128 * a.BN * 2);
  a.nob = n_out >= 2 ? 2 : (a.BN == 64 ? 3 : 4);
  size_t out_bytes = a.epi_tma ? (size_t)a.nob * n_out * a.out_tile_bytes : 0;
  // per-warp mode: two epilogue groups on alternate tiles (needs the two-issuer / four-buffer MMA side: single k-step
  // tiles) and 32-pixel sub-tiles that are whole rows of the tile
  if (a.epi_tma == 1 && n_out == 1 && (a.BN == 16 || a.BN == 32) && a.nbuf == 4 && a.ncols * a.kchunks == 1 && a.tw <= 32) {
    a.epi_tma = 3;
    out_bytes = (size_t)kEpiWarps * 4 * 32 * a.BN * 2;
  }
  if (!a.epi_tma && n_out == 1 && a.BN > 64 && a.BN % 64 == 0) {
    // wide TMA-store epilogue: three (or two) 16 KB staging tiles, as long as the load pipeline keeps >= 3 stages
    const size_t stage_b = a.a_slot_bytes + (a.resident ? 0 : (size_t)a.max_nr * a.b_tap_bytes);
    const size_t base_fixed = 1024 + a.resident_bytes + stats_bytes_wide + sizeof(Barriers) + 64;
    for (int nob = 3; nob >= 2 && !a.epi_tma; --nob)
      if (base_fixed + (size_t)nob * 16384 + 3 * stage_b <= 220 * 1024) { a.epi_tma = 2; a.nob = nob; out_bytes = (size_t)nob * 16384; }
    if (a.epi_tma == 2) stats_bytes = stats_bytes_wide;
  }
  const size_t fixed = 1024 + a.resident_bytes + out_bytes + stats_bytes + sizeof(Barriers) + 64;
  const size_t budget = 220 * 1024;
  const size_t stage = a.a_slot_bytes + (a.resident ? 0 : (size_t)a.max_nr * a.b_tap_bytes);
  if (fixed + 2 * stage > budget)
    return dp_set_error(DP_ERR_UNSUPPORTED, "conv_tc: shape needs %zu B of shared memory", fixed + 2 * stage);
  int st = (int)((budget - fixed) / stage);
  a.stages = st > kMaxStages ? kMaxStages : st;
  // Two issuers alternate tiles (issuer w takes tiles w, w+2, ...).  The stage ring is shared, and an mbarrier parity wait
  // is only sound for a waiter that observes EVERY phase of the barrier it waits on: `try_wait.parity p` answers "the
  // phase with parity p has completed", which is also true of the phase two laps back.  With an arbitrary ring depth the
  // two issuers meet on the same full[] barriers in alternating laps, so an issuer that runs ahead asks for lap L+1 of
  // a stage whose lap L (the other issuer's) has not landed yet, gets "complete" from lap L-1, and issues UMMAs on a
  // slot TMA is still writing - the intermittent fault seen in round 1 on the EfficientNet 1x1 layers (Cin 144/192, three
  // k-steps per tile).  The ring is therefore cut to a multiple of 2 * ksteps stages: tile t then always occupies ring
  // block t mod (stages / ksteps), even blocks belong to issuer 0 and odd blocks to issuer 1, and every full/empty
  // barrier has exactly one producer and one consumer for the life of the kernel (the single k-step case is the same
  // rule with ksteps = 1).  Shapes whose ring cannot hold two tiles run with one issuer.
  {
    const int ksteps = a.ncols * a.kchunks;
    static const int multi_ok = []() { const char* e = getenv("DP_TWO_ISSUER_MULTI"); return e ? atoi(e) : 1; }();
    if (a.niss == 2) {
      const int per = 2 * ksteps;
      if (ksteps > 4 || a.stages < per || (ksteps > 1 && !multi_ok)) a.niss = 1;
      else a.stages = (a.stages / per) * per;
    }
  }
  p.smem = fixed + (size_t)a.stages * stage;
  a.total_items = (long long)B * a.tiles_y * a.tiles_x * a.n_blocks;
  p.grid = (int)(a.total_items < dp::kNumSMs ? a.total_items : dp::kNumSMs);
  {
    int g = p.grid;
    a.step_nb = g % a.n_blocks; g /= a.n_blocks;
    a.step_tx = g % a.tiles_x; g /= a.tiles_x;
    a.step_ty = g % a.tiles_y; g /= a.tiles_y;
    a.step_n = g;
  }
  if (a.total_items >= (1LL << 31)) return dp_set_error(DP_ERR_UNSUPPORTED, "conv_tc: %lld work items", a.total_items);
  return DP_OK;
}

int launch(const Plane* planes, const ColSpec* cols, int ncols, int B, int Hg, int Wg, int Cin, const void* w_packed,
           int Cin_p, int ntaps, int Cout, const Epilogue& ep, const OutMap& om, cudaStream_t stream, int allow_halo = 0) {
  Plan p;
  int rc = make_plan(p, B, Hg, Wg, Cin, Cout, cols, ncols, ep.stats != nullptr, allow_halo,
                     ep.out ? (ep.out2 ? 2 : 1) : 0, ep.pre_ss != nullptr, ep.mask_x != nullptr);
  if (rc) return rc;
  if (ep.mask_x && !(p.a.epi_tma == 1 || p.a.epi_tma == 3))
    return dp_set_error(DP_ERR_UNSUPPORTED, "conv_tc: the BatchNorm-backward epilogue needs a single 16/32/64-column N block "
                        "(Cout %d)", Cout);
  if (ep.mask_x && (ep.res || ep.res2 || ep.relu || ep.out2))
    return dp_set_error(DP_ERR_INVALID, "conv_tc: the BatchNorm-backward epilogue excludes residuals / relu / dual output");
  if (p.a.halo) ncols = 1;
  ConvArgs& a = p.a;
  a.Ho = om.Ho; a.Wo = om.Wo; a.osy = om.osy; a.osx = om.osx; a.oay = om.oay; a.oax = om.oax;
  a.out = reinterpret_cast<bf16*>(ep.out); a.out_ld = ep.out_ld;
  a.out2 = reinterpret_cast<bf16*>(ep.out2); a.out2_ld = ep.out2_ld;
  a.bias = ep.bias;
  a.res = reinterpret_cast<const bf16*>(ep.res); a.res_ld = ep.res_ld;
  a.resb = reinterpret_cast<const bf16*>(ep.res2); a.resb_ld = ep.res2_ld;
  a.relu = ep.relu; a.relu2 = ep.relu2;
  a.stats = ep.stats;
  a.pre_ss = ep.pre_ss; a.pre_act = ep.pre_act; a.pre_c = Cin;
  for (int c = 0; c < kMaxCols; ++c) { a.pH[c] = planes[cols[c < ncols ? c : 0].plane].Hp; a.pW[c] = planes[cols[c < ncols ? c : 0].plane].Wp; }
  a.epi_pipe = ep.mask_x ? 1 : 0;     // measured: the rotation pays in the BatchNorm-backward mode only (1.34 -> 1.03 ms)
  a.aux_mode = ep.mask_x ? 1 : 0;
  a.aux_ss = ep.mask_ss;
  a.aux_hi = ep.mask_act == 2 ? 6.f : __builtin_inff();
  if (ep.mask_x) { a.resb = reinterpret_cast<const bf16*>(ep.mask_x); a.resb_ld = ep.mask_ld; }
  a.dbg = g_wg_dbg;
  Maps tm;
  for (int c = 0; c < ncols; ++c) {
    const Plane& pl = planes[cols[c].plane];
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)pl.Wp, (uint64_t)pl.Hp, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)pl.ld_px * 2, (uint64_t)pl.ld_row * 2, (uint64_t)pl.ld_img * 2};
    uint32_t box[4] = {(uint32_t)a.KB, (uint32_t)a.tw, (uint32_t)(a.th + cols[c].nr - 1), 1};
    if (a.halo) { box[1] = (uint32_t)(a.tw + 2); box[2] = (uint32_t)(a.th + 2); }
    rc = dp_make_tmap_bf16(&tm.a[c], pl.base, 4, dims, str, box, nullptr, a.row_bytes);
    if (rc) return rc;
  }
  for (int c = ncols; c < kMaxCols; ++c) tm.a[c] = tm.a[0];
  {
    uint64_t dims[3] = {(uint64_t)Cin_p, (uint64_t)Cout, (uint64_t)ntaps};
    uint64_t str[2] = {(uint64_t)Cin_p * 2, (uint64_t)Cout * Cin_p * 2};
    uint32_t box[3] = {(uint32_t)a.KB, (uint32_t)a.BN, 1};
    rc = dp_make_tmap_bf16(&tm.b, w_packed, 3, dims, str, box, nullptr, a.row_bytes);
    if (rc) return rc;
  }
  tm.o = tm.a[0];
  tm.o2 = tm.a[0];
  if (a.epi_tma) {
    // output tiles leave through TMA: the (possibly strided, for the transposed-conv output phases) pixel grid of the
    // output tensor that GEMM pixel (y, x) maps to; overhanging pixels / channels are clipped by the store
    const uint64_t Wv = (uint64_t)((om.Wo - om.oax + om.osx - 1) / om.osx), Hv = (uint64_t)((om.Ho - om.oay + om.osy - 1) / om.osy);
    uint64_t dims[4] = {(uint64_t)Cout, Wv, Hv, (uint64_t)B};
    uint32_t box[4] = {(uint32_t)(a.epi_tma == 2 ? 64 : a.BN), (uint32_t)a.tw, (uint32_t)a.th, 1};
    if (a.epi_tma == 3) {      // per-warp stores: (32 / tw) x tw pixel boxes of the single output, kept in tm.o2
      uint32_t wbox[4] = {(uint32_t)a.BN, (uint32_t)a.tw, (uint32_t)(32 / a.tw), 1};
      uint64_t str[3] = {(uint64_t)om.osx * a.out_ld * 2, (uint64_t)om.osy * om.Wo * a.out_ld * 2,
                         (uint64_t)om.Ho * om.Wo * a.out_ld * 2};
      rc = dp_make_tmap_bf16(&tm.o2, a.out + ((long long)om.oay * om.Wo + om.oax) * a.out_ld, 4, dims, str, wbox, nullptr,
                             a.BN * 2);
      if (rc) return rc;
    }
    for (int k = 0; k < 2 && a.epi_tma != 3; ++k) {
      bf16* base = k == 0 ? a.out : a.out2;
      const long long ld = k == 0 ? a.out_ld : a.out2_ld;
      if (!base) continue;
      uint64_t str[3] = {(uint64_t)om.osx * ld * 2, (uint64_t)om.osy * om.Wo * ld * 2, (uint64_t)om.Ho * om.Wo * ld * 2};
      rc = dp_make_tmap_bf16(k == 0 ? &tm.o : &tm.o2, base + ((long long)om.oay * om.Wo + om.oax) * ld, 4, dims, str, box,
                             nullptr, a.epi_tma == 2 ? 128 : a.BN * 2);
      if (rc) return rc;
    }
  }
  auto kern = a.pre_ss ? conv_tc_kernel<true> : conv_tc_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) return dp_set_error(DP_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", p.smem, cudaGetErrorString(e));
  {
    const double px = (double)B * Hg * Wg;
    const double bytes = 2.0 * px * (Cin + Cout * (1 + (ep.res ? 1 : 0) + (ep.res2 || ep.mask_x ? 1 : 0) + (ep.out2 ? 1 : 0)));
    const double flops = 2.0 * px * Cin * Cout * ntaps;
    dp::pdl_work(bytes > flops / 200.0 ? bytes : flops / 200.0);
  }
  dp::launch(kern, p.grid, a.pre_ss ? kPreThreads : kThreads, p.smem, stream, tm, a);
  DP_CHECK_LAUNCH("conv_tc_kernel");
  return DP_OK;
}

int plain_cols(ColSpec* cols, int KS) {
  const int pad = KS / 2;
  for (int s = 0; s < KS; ++s) {
    cols[s].plane = 0; cols[s].dx = s - pad; cols[s].dy = -pad; cols[s].nr = KS;
    for (int r = 0; r < 3; ++r) cols[s].wtap[r] = r < KS ? r * KS + s : 0;
  }
  return KS;
}

inline int floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

// stride-2 conv (conv rule i = 2*o - pad + k) as column loads over the four parity planes of the input
int down2_cols(ColSpec* cols, int K, int pad) {
  int n = 0;
  for (int kx = 0; kx < K; ++kx) {
    const int px = (kx - pad) & 1, dx = floordiv2(kx - pad);
    for (int py = 0; py < 2; ++py) {
      ColSpec c;
      c.plane = py * 2 + px; c.dx = dx; c.nr = 0; c.dy = 0;
      for (int ky = 0; ky < K; ++ky) {
        if (((ky - pad) & 1) != py) continue;
        const int dy = floordiv2(ky - pad);
        if (c.nr == 0) c.dy = dy;
        else if (dy != c.dy + c.nr) return -1;
        if (c.nr >= 3) return -1;
        c.wtap[c.nr++] = ky * K + kx;
      }
      if (c.nr == 0) continue;
      for (int r = c.nr; r < 3; ++r) c.wtap[r] = 0;
      if (n >= kMaxCols) return -1;
      cols[n++] = c;
    }
  }
  return n;
}

// one output phase (a, b) of a transposed stride-2 conv (rule i = (o + pad - k)/2): o = 2*y + a
int up2_cols(ColSpec* cols, int K, int pad, int pa, int pb) {
  int n = 0;
  for (int kx = K - 1; kx >= 0; --kx) {
    if (((pb + pad - kx) & 1) != 0) continue;
    ColSpec c;
    c.plane = 0; c.dx = floordiv2(pb + pad - kx); c.nr = 0; c.dy = 0;
    for (int ky = K - 1; ky >= 0; --ky) {       // decreasing k -> increasing input row
      if (((pa + pad - ky) & 1) != 0) continue;
      const int dy = floordiv2(pa + pad - ky);
      if (c.nr == 0) c.dy = dy;
      else if (dy != c.dy + c.nr) return -1;
      if (c.nr >= 3) return -1;
      c.wtap[c.nr++] = ky * K + kx;
    }
    if (c.nr == 0) continue;
    for (int r = c.nr; r < 3; ++r) c.wtap[r] = 0;
    if (n >= kMaxCols) return -1;
    cols[n++] = c;
  }
  return n;
}

}  // namespace

extern "C" {

/* number of CTAs (= rows of the BN-statistics partials buffer) dp_conv2d_tc launches for this shape */
int dp_conv2d_tc_grid(int B, int H, int W, int Cin, int Cout, int KS) {
  Plan p;
  ColSpec cols[kMaxCols];
  const int n = plain_cols(cols, KS);
  if (make_plan(p, B, H, W, Cin, Cout, cols, n, 0, KS == 3)) return -1;
  return p.grid;
}

int dp_conv2d_tc(const void* x, long long x_ld, int B, int H, int W, int Cin, const void* w_packed, int Cin_p,
                 int Cout, int KS, const float* bias, const void* residual, long long res_ld, const void* residual2,
                 long long res2_ld, int relu, void* out, long long out_ld, void* out2, long long out2_ld, int relu2,
                 float* stats_partials, cudaStream_t stream) {
  return dp_conv2d_tc_fused(x, x_ld, B, H, W, Cin, w_packed, Cin_p, Cout, KS, bias, residual, res_ld, residual2, res2_ld,
                            relu, out, out_ld, out2, out2_ld, relu2, stats_partials, nullptr, stream);
}

int dp_conv2d_tc_caps(int B, int H, int W, int Cin, int Cout, int KS) {
  Plan p;
  ColSpec cols[kMaxCols];
  const int n = plain_cols(cols, KS);
  if (make_plan(p, B, H, W, Cin, Cout, cols, n, 1, KS == 3, 1, 1, 1)) return 0;
  return DP_CONV_CAP_PROLOGUE | ((p.a.epi_tma == 1 || p.a.epi_tma == 3) ? DP_CONV_CAP_BN_BACKWARD : 0);
}

int dp_conv2d_tc_fused(const void* x, long long x_ld, int B, int H, int W, int Cin, const void* w_packed, int Cin_p,
                       int Cout, int KS, const float* bias, const void* residual, long long res_ld, const void* residual2,
                       long long res2_ld, int relu, void* out, long long out_ld, void* out2, long long out2_ld, int relu2,
                       float* stats_partials, const dp_conv_fuse_t* fuse, cudaStream_t stream) {
  DP_CHECK_ARG(x && w_packed && (out || out2), "dp_conv2d_tc: null pointer");
  DP_CHECK_ARG(KS == 3 || KS == 1, "dp_conv2d_tc: kernel size %d (only 1 and 3, stride 1)", KS);
  DP_CHECK_ARG(Cin % 8 == 0 && Cout % 8 == 0 && Cin_p % 8 == 0 && Cin_p >= Cin,
               "dp_conv2d_tc: channels must be multiples of 8 (Cin %d Cin_p %d Cout %d)", Cin, Cin_p, Cout);
  DP_CHECK_ARG(x_ld % 8 == 0 && (!out || out_ld % 8 == 0) && (!out2 || out2_ld % 8 == 0) &&
               (!residual || res_ld % 8 == 0) && (!residual2 || res2_ld % 8 == 0),
               "dp_conv2d_tc: pixel strides must be multiples of 8 elements");
  DP_CHECK_ARG(B > 0 && H > 0 && W > 0, "dp_conv2d_tc: bad shape");
  Plane pl{x, x_ld, (long long)W * x_ld, (long long)H * W * x_ld, H, W};
  ColSpec cols[kMaxCols];
  const int n = plain_cols(cols, KS);
  Epilogue ep{bias, residual, res_ld, residual2, res2_ld, relu, out, out_ld, out2, out2_ld, relu2, stats_partials};
  if (fuse) {
    DP_CHECK_ARG(!fuse->mask_x || (fuse->mask_scale_shift && stats_partials && fuse->mask_ld % 8 == 0),
                 "dp_conv2d_tc_fused: mask_x needs mask_scale_shift, stats_partials and a 16-byte aligned pixel stride");
    ep.pre_ss = fuse->pre_scale_shift; ep.pre_act = fuse->pre_act;
    ep.mask_x = fuse->mask_x; ep.mask_ld = fuse->mask_ld; ep.mask_ss = fuse->mask_scale_shift; ep.mask_act = fuse->mask_act;
  }
  OutMap om{H, W, 1, 1, 0, 0};
  return launch(&pl, cols, n, B, H, W, Cin, w_packed, Cin_p, KS * KS, Cout, ep, om, stream, KS == 3);
}

/* Stride-2 convolution (conv rule i = 2*o - pad + k, K x K taps) on the tensor cores: nn.Conv2d(k3,s2,p1) forward
 * (midas_semantics.py:39-45, dpt_depth.py:63-68) and the data gradient of nn.ConvTranspose2d(k4,s2,p1)
 * (midas_semantics.py:52-58).  x: (B,Hi,Wi,Cin) NHWC bf16; w_packed bf16 [K*K][Cout][Cin_p]; out: (B,Ho,Wo,Cout). */
int dp_conv2d_tc_down2(const void* x, long long x_ld, int B, int Hi, int Wi, int Cin, const void* w_packed, int Cin_p,
                       int Cout, int K, int pad, const float* bias, int relu, void* out, long long out_ld, int Ho,
                       int Wo, float* stats_partials, cudaStream_t stream) {
  DP_CHECK_ARG(x && w_packed && out, "dp_conv2d_tc_down2: null pointer");
  DP_CHECK_ARG(Cin % 8 == 0 && Cout % 8 == 0 && Cin_p >= Cin && x_ld % 8 == 0 && out_ld % 8 == 0,
               "dp_conv2d_tc_down2: channels / strides must be multiples of 8");
  ColSpec cols[kMaxCols];
  const int n = down2_cols(cols, K, pad);
  if (n <= 0) return dp_set_error(DP_ERR_UNSUPPORTED, "dp_conv2d_tc_down2: K %d pad %d not supported", K, pad);
  Plane pl[4];
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      Plane& q = pl[py * 2 + px];
      q.base = xb + ((long long)py * Wi + px) * x_ld;
      q.ld_px = 2 * x_ld; q.ld_row = 2LL * Wi * x_ld; q.ld_img = (long long)Hi * Wi * x_ld;
      q.Hp = (Hi - py + 1) / 2; q.Wp = (Wi - px + 1) / 2;
      if (q.Hp < 1) q.Hp = 1;
      if (q.Wp < 1) q.Wp = 1;
    }
  Epilogue ep{bias, nullptr, 0, nullptr, 0, relu, out, out_ld, nullptr, 0, 0, stats_partials};
  OutMap om{Ho, Wo, 1, 1, 0, 0};
  return launch(pl, cols, n, B, Ho, Wo, Cin, w_packed, Cin_p, K * K, Cout, ep, om, stream);
}

int dp_conv2d_tc_down2_grid(int B, int Ho, int Wo, int Cin, int Cout, int K, int pad) {
  Plan p;
  ColSpec cols[kMaxCols];
  const int n = down2_cols(cols, K, pad);
  if (n <= 0 || make_plan(p, B, Ho, Wo, Cin, Cout, cols, n, 0)) return -1;
  return p.grid;
}

/* Transposed stride-2 convolution (rule i = (o + pad - k)/2 when even) as four output-phase launches:
 * nn.ConvTranspose2d(k4,s2,p1) forward (midas_semantics.py:52-58) and the data gradient of nn.Conv2d(k3,s2,p1).
 * x: (B,Hi,Wi,Cin); w_packed bf16 [K*K][Cout][Cin_p]; out: (B,Ho,Wo,Cout). */
int dp_conv2d_tc_up2(const void* x, long long x_ld, int B, int Hi, int Wi, int Cin, const void* w_packed, int Cin_p,
                     int Cout, int K, int pad, const float* bias, int relu, void* out, long long out_ld, int Ho, int Wo,
                     cudaStream_t stream) {
  DP_CHECK_ARG(x && w_packed && out, "dp_conv2d_tc_up2: null pointer");
  DP_CHECK_ARG(Cin % 8 == 0 && Cout % 8 == 0 && Cin_p >= Cin && x_ld % 8 == 0 && out_ld % 8 == 0,
               "dp_conv2d_tc_up2: channels / strides must be multiples of 8");
  Plane pl{x, x_ld, (long long)Wi * x_ld, (long long)Hi * Wi * x_ld, Hi, Wi};
  for (int pa = 0; pa < 2; ++pa)
    for (int pb = 0; pb < 2; ++pb) {
      const int Hg = (Ho - pa + 1) / 2, Wg = (Wo - pb + 1) / 2;
      if (Hg <= 0 || Wg <= 0) continue;
      ColSpec cols[kMaxCols];
      const int n = up2_cols(cols, K, pad, pa, pb);
      if (n < 0) return dp_set_error(DP_ERR_UNSUPPORTED, "dp_conv2d_tc_up2: K %d pad %d not supported", K, pad);
      if (n == 0) return dp_set_error(DP_ERR_UNSUPPORTED, "dp_conv2d_tc_up2: phase (%d,%d) has no taps", pa, pb);
      Epilogue ep{bias, nullptr, 0, nullptr, 0, relu, out, out_ld, nullptr, 0, 0, nullptr};
      OutMap om{Ho, Wo, 2, 2, pa, pb};
      int rc = launch(&pl, cols, n, B, Hg, Wg, Cin, w_packed, Cin_p, K * K, Cout, ep, om, stream);
      if (rc) return rc;
    }
  return DP_OK;
}

}  // extern "C"
