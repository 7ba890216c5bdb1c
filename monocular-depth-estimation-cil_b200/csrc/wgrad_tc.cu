// Weight gradient of the 3x3/s1/p1 and 1x1 convolutions on tcgen05 tensor cores.
//
// dW[tap][co][ci] = sum over pixels p of dY[p][co] * X[p + tap][ci]   (what autograd's
// convolution_backward computes for the nn.Conv2d layers listed in conv_tc.cu).
//
// GEMM view: M = co, N = ci, K = pixels.  Both operands are NHWC, i.e. "MN-major" (the contracted
// pixel index is the strided one), which UMMA consumes directly from 128B/64B/32B-swizzled TMA boxes
// (tests/test_umma_probe_gpu.py pins the MN-major descriptor conventions).  One CTA owns one
// horizontal tap s, one block of output channels and one <=64-wide chunk of input channels; it
// walks its share of the 128-pixel patches, issuing for each patch 3 (vertical taps) x 8 (K=16
// pixel slices) UMMAs into three TMEM accumulators that share the dY tile and read the same X halo
// box at r*tw rows offset.  Split-K partials go to a workspace and are reduced deterministically
// (no atomics) by wgrad_reduce_kernel, which also un-packs to the OIHW fp32 layout of param.grad.
#include "common.cuh"
#include "tc.cuh"
#include "bn_fuse.cuh"
#include "../../include/depth_b200.h"

#ifdef DP_CONV_TIMING
#define DP_T(x) x
#else
#define DP_T(x)
#endif

namespace {

constexpr int kThreads = 192;  // warp 0 producer, warp 1 MMA (+TMEM alloc), warps 2..5 epilogue
constexpr int kMaxStages = 8;

constexpr int kMaxKW = 4;

struct WgGroup {      // one TMA box of the tapped operand: vertical taps ky[0..nr) share it
  int map, dx, dy, nr;
  int ky[3];
  uint32_t box_bytes, slot_off;
  uint32_t idesc, tmem_col;   // one UMMA covers all nr taps: N = nr*NC, columns start at tmem_col
};

struct WgMaps {
  CUtensorMap p;                  // plain operand (M side)
  CUtensorMap t[kMaxKW * 2];      // tapped operand boxes, one per (horizontal tap, group)
};

struct WgArgs {
  int B, H, W, Cout, Cin, KH, KW;   // H, W: GEMM pixel grid = grid of the plain operand; Cout = its channels (M side)
  int th, tw, tiles_y, tiles_x;
  int MC, m_chunks, M, co_blocks;   // M chunk width (<=64 channels), chunks per UMMA, UMMA M, blocks over Cout
  int NC, ci_chunks;                // N chunk width (<=64 channels), chunks over Cin (tapped operand channels)
  int psplit, stages;
  int halo, pitch;                  // halo mode: one (th+2) x (tw+2) box of the tapped operand serves all nine taps
  int step_tx, step_ty, step_n;     // digits of the tile stride (psplit) in the (tiles_x, tiles_y, B) radix
  int ngrp[kMaxKW];
  WgGroup grp[kMaxKW][2];
  long long tiles_total;
  uint32_t a_box_bytes, a_slot_bytes, stage_bytes;
  uint32_t rowA, rowB, layoutA, layoutB, idesc, a_lbo;
  // fused BatchNorm prologue on the tapped operand: X = act(c * pre_ss[ci] + pre_ss[Cin + ci]), zero outside the plane
  // (the epilogue warps, idle until the accumulators are complete, rewrite every landed X box in place)
  const float* pre_ss; int pre_act;
  int pH[kMaxKW * 2], pW[kMaxKW * 2];
  float* partial;  // [psplit][KH*KW][Cout][Cin]
  unsigned long long* dbg;  // optional cycle counters (diagnostics): [0]=producer wait, [1]=mma wait, [2]=mma issue, [3]=total, [4]=tiles
};

struct __align__(8) WgBars {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t done;
  uint64_t ready[kMaxStages];     // prologue mode: X boxes transformed (four arrivals: warps 2..5)
  uint32_t tmem_base;
  __align__(16) float pre[2][64]; // this CTA's scale / shift slice (NC <= 64 channels)
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ WgMaps tm, const __grid_constant__ WgArgs a) {
  dp::pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  WgBars* bars = reinterpret_cast<WgBars*>(smem + (size_t)a.stages * a.stage_bytes);
  const int warp = tc::warp_idx_uniform(), lane = threadIdx.x & 31;

  int id = blockIdx.x;
  const int ps = id % a.psplit; id /= a.psplit;
  const int cc = id % a.ci_chunks; id /= a.ci_chunks;
  const int cb = id % a.co_blocks; id /= a.co_blocks;
  const int s = id;  // horizontal tap (always 0 in halo mode: the CTA walks all three itself)

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) { tc::mbar_init(&bars->full[i], 1); tc::mbar_init(&bars->empty[i], 1); }
    tc::mbar_init(&bars->done, 1);
    for (int i = 0; i < a.stages; ++i) tc::mbar_init(&bars->ready[i], 4);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tm.p);
    tc::prefetch_tmap(&tm.t[0]);
  }
  if (warp == 1) tc::tmem_alloc(&bars->tmem_base, 512);
  dp::pdl_wait();   // global memory is touched from here on
  if (a.pre_ss && threadIdx.x >= 64) {
    const int i = threadIdx.x - 64;                 // 128 threads: [scale | shift] x 64 channels
    const int which = i >> 6, c = cc * a.NC + (i & 63);
    bars->pre[which][i & 63] = ((i & 63) < a.NC && c < a.Cin) ? __ldg(a.pre_ss + which * a.Cin + c) : 0.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = bars->tmem_base;
  const int real_chunks = a.MC < 64 ? 1 : a.m_chunks;  // chunks actually loaded (narrow layers alias chunk 0)

  DP_T(const long long t_start = clock64();)
  if (warp == 0) {
    if (tc::elect_one_sync()) {
    uint32_t stage = 0, phase = 0;
    DP_T(long long w_acc = 0;)
    // tile walk with stride psplit: decode once, then advance the (tx, ty, n) digits by the stride's digits
    int tx, ty, n;
    {
      unsigned m = (unsigned)ps;
      tx = (int)(m % (unsigned)a.tiles_x); m /= (unsigned)a.tiles_x;
      ty = (int)(m % (unsigned)a.tiles_y);
      n = (int)(m / (unsigned)a.tiles_y);
    }
    for (long long t = ps; t < a.tiles_total; t += a.psplit) {
      const int y0 = ty * a.th, x0 = tx * a.tw;
      DP_T(const long long c0 = clock64();)
      tc::mbar_wait(&bars->empty[stage], phase ^ 1);
      DP_T(w_acc += clock64() - c0;)
      uint8_t* sA = smem + (size_t)stage * a.stage_bytes;
      uint8_t* sX = sA + (size_t)real_chunks * a.a_slot_bytes;
      uint32_t txb = (uint32_t)real_chunks * a.a_box_bytes;
      for (int g = 0; g < a.ngrp[s]; ++g) txb += a.grp[s][g].box_bytes;
      tc::mbar_expect_tx(&bars->full[stage], txb);
      for (int j = 0; j < real_chunks; ++j)
        tc::tma_load_4d(sA + (size_t)j * a.a_slot_bytes, &tm.p, &bars->full[stage], cb * a.M + j * a.MC, x0, y0, n);
      for (int g = 0; g < a.ngrp[s]; ++g) {
        const WgGroup& G = a.grp[s][g];
        tc::tma_load_4d(sX + G.slot_off, &tm.t[G.map], &bars->full[stage], cc * a.NC, x0 + G.dx, y0 + G.dy, n);
      }
      if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
      tx += a.step_tx; ty += a.step_ty; n += a.step_n;
      if (tx >= a.tiles_x) { tx -= a.tiles_x; ++ty; }
      if (ty >= a.tiles_y) { ty -= a.tiles_y; ++n; }
    }
    DP_T(if (a.dbg && blockIdx.x == 0) a.dbg[0] = (unsigned long long)w_acc;)
    }
  } else if (warp == 1) {
    if (tc::elect_one_sync()) {
    uint32_t stage = 0, phase = 0;
    uint32_t accumulate = 0;
    DP_T(long long w_acc = 0; long long i_acc = 0; long long ntile = 0;)
    const uint32_t a_hi = tc::desc_hi(8 * a.rowA, a.layoutA), b_hi = tc::desc_hi(8 * a.rowB, a.layoutB);
    const uint32_t a_step = (16u * a.rowA) >> 4, b_step = (16u * a.rowB) >> 4;   // 16 pixel rows per UMMA K-step
    for (long long t = ps; t < a.tiles_total; t += a.psplit) {
      DP_T(const long long c0 = clock64();)
      tc::mbar_wait(a.pre_ss ? &bars->ready[stage] : &bars->full[stage], phase);
      tc::fence_after_sync();
      DP_T(const long long c1 = clock64(); w_acc += c1 - c0; ++ntile;)
      const uint32_t a_base = tc::smem_u32(smem + (size_t)stage * a.stage_bytes);
      const uint32_t x_base = a_base + (uint32_t)real_chunks * a.a_slot_bytes;
      if (a.halo) {
        // one halo box: horizontal tap sx starts sx pixel rows in, vertical taps are `pitch` rows apart (LBO), and one
        // K=16 slice is one 16-pixel tile row, i.e. `pitch` box rows further down per step.  The swizzle is a function
        // of the absolute shared-memory address, so none of these offsets needs to be atom aligned.
        const WgGroup& G = a.grp[0][0];
        const uint32_t al0 = tc::desc_lo(a_base, a.a_lbo);
        const uint32_t row16 = a.rowB >> 4, b_row_step = ((uint32_t)a.pitch * a.rowB) >> 4;
        const uint32_t bl00 = tc::desc_lo(x_base + G.slot_off, (uint32_t)a.pitch * a.rowB);
#pragma unroll
        for (int sx = 0; sx < 3; ++sx) {
          const uint32_t d_col = tmem + (uint32_t)(sx * 3 * a.NC);
          const uint32_t bl0 = bl00 + (uint32_t)sx * row16;
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16)
            tc::umma_bf16_lohi(d_col, al0 + (uint32_t)k16 * a_step, a_hi, bl0 + (uint32_t)k16 * b_row_step, b_hi, G.idesc,
                               k16 > 0 ? 1u : accumulate);
        }
      } else
      for (int g = 0; g < a.ngrp[s]; ++g) {
        const WgGroup& G = a.grp[s][g];
        // the nr vertical taps of a group are the same box read r*tw pixel rows further down: with the MN-major
        // descriptor's LBO = tw*rowB they become nr consecutive N-chunks of ONE UMMA (N = nr*NC), whose
        // accumulator columns [ky0*NC, (ky0+nr)*NC) are exactly the per-tap accumulators (contiguous ky).
        const uint32_t al0 = tc::desc_lo(a_base, a.a_lbo), bl0 = tc::desc_lo(x_base + G.slot_off, (uint32_t)a.tw * a.rowB);
        const uint32_t d_col = tmem + G.tmem_col;
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16)
          tc::umma_bf16_lohi(d_col, al0 + (uint32_t)k16 * a_step, a_hi, bl0 + (uint32_t)k16 * b_step, b_hi, G.idesc,
                             k16 > 0 ? 1u : accumulate);
      }
      accumulate = 1;
      tc::umma_commit(&bars->empty[stage]);
      DP_T(i_acc += clock64() - c1;)
      if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
    }
    tc::umma_commit(&bars->done);
    DP_T(if (a.dbg && blockIdx.x == 0) { a.dbg[1] = (unsigned long long)w_acc; a.dbg[2] = (unsigned long long)i_acc; a.dbg[4] = (unsigned long long)ntile; })
    }
  } else if (warp >= 2) {
    if (a.pre_ss) {
      // prologue transform of the tapped operand, tile by tile behind the producer
      const int t128 = threadIdx.x - 64;
      uint32_t stage = 0, phase = 0;
      int tx, ty;
      {
        unsigned m = (unsigned)ps;
        tx = (int)(m % (unsigned)a.tiles_x); m /= (unsigned)a.tiles_x;
        ty = (int)(m % (unsigned)a.tiles_y);
      }
      for (long long t = ps; t < a.tiles_total; t += a.psplit) {
        const int y0 = ty * a.th, x0 = tx * a.tw;
        tc::mbar_wait(&bars->full[stage], phase);
        const uint32_t x_base = tc::smem_u32(smem + (size_t)stage * a.stage_bytes) + (uint32_t)real_chunks * a.a_slot_bytes;
        for (int g = 0; g < a.ngrp[s]; ++g) {
          const WgGroup& G = a.grp[s][g];
          const int boxW = a.halo ? a.pitch : a.tw;
          const int npx = (a.halo ? a.th + 2 : a.th + G.nr - 1) * boxW;
          const uint32_t base = x_base + G.slot_off;
          if (a.rowB == 128)
            dpf::transform_box<128, 128>(base, npx, boxW, x0 + G.dx, y0 + G.dy, a.pW[G.map], a.pH[G.map], &bars->pre[0][0], 64, 0, a.pre_act, t128);
          else if (a.rowB == 64)
            dpf::transform_box<64, 128>(base, npx, boxW, x0 + G.dx, y0 + G.dy, a.pW[G.map], a.pH[G.map], &bars->pre[0][0], 64, 0, a.pre_act, t128);
          else
            dpf::transform_box<32, 128>(base, npx, boxW, x0 + G.dx, y0 + G.dy, a.pW[G.map], a.pH[G.map], &bars->pre[0][0], 64, 0, a.pre_act, t128);
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->ready[stage]);
        if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
        tx += a.step_tx; ty += a.step_ty;
        if (tx >= a.tiles_x) { tx -= a.tiles_x; ++ty; }
        if (ty >= a.tiles_y) ty -= a.tiles_y;
      }
    }
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    tc::mbar_wait(&bars->done, 0);
    tc::fence_after_sync();
    // accumulator row -> TMEM lane: M=128: lane = row; M=64: lane = (row/16)*32 + row%16
    int row = -1;
    if (a.M == 128) row = q * 32 + lane;
    else if (lane < 16) row = q * 16 + lane;
    const int co = cb * a.M + row;
    const bool row_ok = row >= 0 && row < (a.MC < 64 ? a.MC : a.M) && co < a.Cout;
    const bool has_work = ps < a.tiles_total;  // a split with no tiles left its TMEM untouched: write zeros
    const int ns = a.halo ? 3 : 1;
    for (int sx = 0; sx < ns; ++sx)
    for (int gr = 0; gr < a.ngrp[s] * 3; ++gr) {
      const WgGroup& G = a.grp[s][gr / 3];
      const int r = gr % 3;
      if (r >= G.nr) continue;  // warp-uniform
      const int tap = G.ky[r] * a.KW + (a.halo ? sx : s);
      const uint32_t col_base = G.tmem_col + (uint32_t)(sx * 3 * a.NC);
      for (int c0 = 0; c0 < a.NC; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + col_base + (uint32_t)(r * a.NC + c0), v);
        if (row_ok) {
          float* dst = a.partial + (((size_t)ps * a.KH * a.KW + tap) * a.Cout + co) * a.Cin + cc * a.NC + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cc * a.NC + c0 + j < a.Cin) dst[j] = has_work ? v[j] : 0.f;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  DP_T(if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[3] = (unsigned long long)(clock64() - t_start);)
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// sum the split-K partials and write OIHW fp32: grad[co][ci][r][s] (+)= sum_ps partial[ps][tap][co][ci]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int psplit, int taps, int Cout, int Cin,
                                    float* __restrict__ grad, int accumulate) {
  dp::pdl_prologue();
  const long long total = (long long)taps * Cout * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < psplit; ++p) s += partial[(size_t)p * total + i];
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int tap = (int)(i / ((long long)Cin * Cout));
    const size_t o = ((size_t)co * Cin + ci) * taps + tap;
    grad[o] = accumulate ? grad[o] + s : s;
  }
}

}  // namespace
unsigned long long* g_wg_dbg = nullptr;   // diagnostics only (dp_debug_set_buffer); never set on the product path
namespace {

struct WgPlan {
  WgArgs a;
  size_t smem;
  int grid;
};

struct WgPlaneT {  // strided pixel-grid view of the tapped operand
  const void* base;
  long long ld_px, ld_row, ld_img;
  int Hp, Wp;
};

inline int wg_floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

// stride: 1 (plain conv: T index = p - pad + k) or 2 (T index = 2p - pad + k, through parity planes)
int wg_plan(WgPlan& p, int B, int Hg, int Wg, int Cp, int Ct, int K, int pad, int stride) {
  WgArgs& a = p.a;
  if (K < 1 || K > kMaxKW || (stride != 1 && stride != 2))
    return dp_set_error(DP_ERR_UNSUPPORTED, "wgrad_tc: K %d stride %d", K, stride);
  a.B = B; a.H = Hg; a.W = Wg; a.Cin = Ct; a.Cout = Cp; a.KH = K; a.KW = K;
  const int cand[5][2] = {{8, 16}, {16, 8}, {4, 32}, {2, 64}, {1, 128}};
  long long best = -1;
  for (int i = 0; i < 5; ++i) {
    long long t = (long long)dp::ceil_div(Hg, cand[i][0]) * dp::ceil_div(Wg, cand[i][1]);
    if (best < 0 || t < best) { best = t; a.th = cand[i][0]; a.tw = cand[i][1]; }
  }
  // halo mode (plain 3x3 with few tapped-operand channels, i.e. the bandwidth-bound full-resolution layers): every
  // operand tile is fetched once and all nine taps accumulate in one CTA (9 x NC <= 512 TMEM columns)
  a.halo = (stride == 1 && K == 3 && pad == 1 && Ct <= 64) ? 1 : 0;
  a.pitch = 0;
  if (a.halo) { a.th = 8; a.tw = 16; a.pitch = a.tw + 2; }
  a.tiles_y = dp::ceil_div(Hg, a.th);
  a.tiles_x = dp::ceil_div(Wg, a.tw);
  a.tiles_total = (long long)B * a.tiles_y * a.tiles_x;
  a.MC = Cp > 32 ? 64 : (Cp > 16 ? 32 : 16);
  a.M = Cp > 64 ? 128 : 64;
  a.m_chunks = a.M / a.MC;
  a.co_blocks = dp::ceil_div(Cp, a.M);
  a.NC = (Ct >= 64 && !a.halo) ? 64 : (Ct > 16 ? 32 : 16);
  a.ci_chunks = dp::ceil_div(Ct, a.NC);
  a.rowA = a.MC * 2; a.rowB = a.NC * 2;
  a.layoutA = tc::swizzle_for_row_bytes(a.rowA);
  a.layoutB = tc::swizzle_for_row_bytes(a.rowB);
  a.idesc = tc::make_idesc_bf16(a.M, a.NC, 1, 1);
  a.a_box_bytes = 128u * a.rowA;
  a.a_slot_bytes = (a.a_box_bytes + 1023u) & ~1023u;
  // tapped-operand groups per horizontal tap
  uint32_t max_x = 0;
  for (int kx = 0; kx < K; ++kx) {
    a.ngrp[kx] = 0;
    uint32_t off = 0;
    if (stride == 1) {
      WgGroup& G = a.grp[kx][0];
      G.map = kx * 2; G.dx = kx - pad; G.dy = -pad; G.nr = K;
      if (K > 3) return dp_set_error(DP_ERR_UNSUPPORTED, "wgrad_tc: stride-1 K %d", K);
      for (int r = 0; r < 3; ++r) G.ky[r] = r < K ? r : 0;
      G.box_bytes = (uint32_t)((a.th + K - 1) * a.tw) * a.rowB;
      if (a.halo) {
        G.map = 0; G.dx = -pad;
        G.box_bytes = (uint32_t)((a.th + 2) * a.pitch) * a.rowB;
      }
      G.slot_off = 0;
      off = (G.box_bytes + 1023u) & ~1023u;
      a.ngrp[kx] = 1;
    } else {
      for (int py = 0; py < 2; ++py) {
        WgGroup G;
        G.nr = 0; G.dy = 0; G.dx = wg_floordiv2(kx - pad); G.map = kx * 2 + py;
        for (int ky = 0; ky < K; ++ky) {
          if (((ky - pad) & 1) != py) continue;
          const int dy = wg_floordiv2(ky - pad);
          if (G.nr == 0) G.dy = dy;
          else if (dy != G.dy + G.nr) return dp_set_error(DP_ERR_UNSUPPORTED, "wgrad_tc: taps not contiguous");
          if (G.nr >= 3) return dp_set_error(DP_ERR_UNSUPPORTED, "wgrad_tc: too many taps per group");
          G.ky[G.nr++] = ky;
        }
        if (G.nr == 0) continue;
        for (int r = G.nr; r < 3; ++r) G.ky[r] = 0;
        G.box_bytes = (uint32_t)((a.th + G.nr - 1) * a.tw) * a.rowB;
        G.slot_off = off;
        off += (G.box_bytes + 1023u) & ~1023u;
        a.grp[kx][a.ngrp[kx]++] = G;
      }
    }
    if (off > max_x) max_x = off;
  }
  for (int kx = 0; kx < K; ++kx) {
    int slot = 0;
    for (int g = 0; g < a.ngrp[kx]; ++g) {
      WgGroup& G = a.grp[kx][g];
      G.tmem_col = (uint32_t)(slot * a.NC);
      G.idesc = tc::make_idesc_bf16(a.M, G.nr * a.NC, 1, 1);
      slot += G.nr;
    }
    if (slot * a.NC * (a.halo ? 3 : 1) > 512)
      return dp_set_error(DP_ERR_UNSUPPORTED, "wgrad_tc: accumulators exceed TMEM allocation");
  }
  const int real_chunks = a.MC < 64 ? 1 : a.m_chunks;
  a.a_lbo = a.MC < 64 ? 0u : a.a_slot_bytes;
  a.stage_bytes = real_chunks * a.a_slot_bytes + max_x;
  const int base_ctas = (a.halo ? 1 : K) * a.co_blocks * a.ci_chunks;
  // split-K factor.  Cost model (cycles at ~1.9 GHz): every CTA pays a fixed price (launch, TMEM allocation, the
  // epilogue that writes taps x M x NC fp32 partials) plus a per-tile price (UMMA time on the pipe, or the issue /
  // barrier latency floor for narrow layers); the partials are written once and read once by the reduce kernel.
  double tile_cyc = 0.0;
  for (int g = 0; g < a.ngrp[0]; ++g) {
    const double pipe = (a.M / 128.0) * (a.grp[0][g].nr * a.NC) / 2.0;   // cycles per K=16 UMMA
    tile_cyc += 8.0 * (pipe > 20.0 ? pipe : 20.0);
  }
  if (a.halo) tile_cyc *= 3.0;
  tile_cyc += 250.0;
  const double fixed_cyc = 5000.0 + 0.5 * K * a.M * a.NC;
  const double part_bytes = (double)K * K * Cp * Ct * 4.0;
  int ps = 1;
  double best_cost = 1e30;
  const int ps_max = (int)(a.tiles_total < 4 * dp::kNumSMs ? a.tiles_total : 4 * dp::kNumSMs);
  for (int c = 1; c <= ps_max; ++c) {
    const int waves = (base_ctas * c + dp::kNumSMs - 1) / dp::kNumSMs;
    const double per_cta = (double)((a.tiles_total + c - 1) / c) * tile_cyc + fixed_cyc;
    const double cost = waves * per_cta / 1.9e9 + 2.0 * c * part_bytes / 5.0e12;
    if (cost < best_cost) { best_cost = cost; ps = c; }
  }
  a.psplit = ps;
  {
    int g = ps;
    a.step_tx = g % a.tiles_x; g /= a.tiles_x;
    a.step_ty = g % a.tiles_y; g /= a.tiles_y;
    a.step_n = g;
  }
  long long per = (a.tiles_total + ps - 1) / ps;
  int st = (int)((200 * 1024) / a.stage_bytes);
  if (st > kMaxStages) st = kMaxStages;
  if (st > per) st = (int)per;
  if (st < 1) st = 1;
  a.stages = st;
  p.smem = 1024 + (size_t)st * a.stage_bytes + sizeof(WgBars) + 64;
  p.grid = base_ctas * ps;
  return DP_OK;
}

int wg_launch(WgPlan& p, const void* P, long long p_ld, const WgPlaneT* planes /*[kMaxKW*2], indexed by map id*/, int B,
              float* grad, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream,
              const float* pre_ss = nullptr, int pre_act = 0) {
  WgArgs& a = p.a;
  a.pre_ss = pre_ss; a.pre_act = pre_act;
  const int K = a.KH;
  const size_t need = (size_t)a.psplit * K * K * a.Cout * a.Cin * sizeof(float);
  if (workspace_bytes < need) return dp_set_error(DP_ERR_WORKSPACE, "wgrad_tc: workspace %zu < %zu", workspace_bytes, need);
  a.partial = reinterpret_cast<float*>(workspace);
  a.dbg = g_wg_dbg;
  for (int i = 0; i < kMaxKW * 2; ++i) { a.pH[i] = planes[i].Hp; a.pW[i] = planes[i].Wp; }
  WgMaps tm;
  {
    uint64_t dims[4] = {(uint64_t)a.Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)p_ld * 2, (uint64_t)a.W * p_ld * 2, (uint64_t)a.H * a.W * p_ld * 2};
    uint32_t box[4] = {(uint32_t)a.MC, (uint32_t)a.tw, (uint32_t)a.th, 1};
    int rc = dp_make_tmap_bf16(&tm.p, P, 4, dims, str, box, nullptr, a.rowA);
    if (rc) return rc;
  }
  bool have_first = false;
  CUtensorMap first;
  for (int kx = 0; kx < K; ++kx)
    for (int g = 0; g < a.ngrp[kx]; ++g) {
      const WgGroup& G = a.grp[kx][g];
      const WgPlaneT* pl = &planes[G.map];   // planes are indexed like the tensor maps: [kx*2 + group parity]
      uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)pl->Wp, (uint64_t)pl->Hp, (uint64_t)B};
      uint64_t str[3] = {(uint64_t)pl->ld_px * 2, (uint64_t)pl->ld_row * 2, (uint64_t)pl->ld_img * 2};
      uint32_t box[4] = {(uint32_t)a.NC, (uint32_t)(a.halo ? a.pitch : a.tw), (uint32_t)(a.th + G.nr - 1), 1};
      int rc = dp_make_tmap_bf16(&tm.t[G.map], pl->base, 4, dims, str, box, nullptr, a.rowB);
      if (a.halo) { first = tm.t[G.map]; have_first = true; kx = K; break; }   // the single halo box serves every tap
      if (rc) return rc;
      if (!have_first) { first = tm.t[G.map]; have_first = true; }
    }
  for (int i = 0; i < kMaxKW * 2; ++i) {
    bool used = false;
    for (int kx = 0; kx < K; ++kx)
      for (int g = 0; g < a.ngrp[kx]; ++g) used |= (a.grp[kx][g].map == i);
    if (!used) tm.t[i] = first;
  }
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) return dp_set_error(DP_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  {
    const double px = (double)B * a.H * a.W;
    const double bytes = 2.0 * px * (a.Cin + a.Cout), flops = 2.0 * px * a.Cin * a.Cout * K * K;
    dp::pdl_work(bytes > flops / 200.0 ? bytes : flops / 200.0);
  }
  dp::launch(wgrad_tc_kernel, p.grid, kThreads, p.smem, stream, tm, a);
  DP_CHECK_LAUNCH("wgrad_tc_kernel");
  const long long total = (long long)K * K * a.Cout * a.Cin;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4 * dp::kNumSMs) blocks = 4 * dp::kNumSMs;
  dp::launch(wgrad_reduce_kernel, blocks, 256, 0, stream, a.partial, a.psplit, K * K, a.Cout, a.Cin, grad, accumulate);
  DP_CHECK_LAUNCH("wgrad_reduce_kernel");
  return DP_OK;
}

}  // namespace

extern "C" {

/* diagnostics: device buffer of >= 8 u64 that block 0 of the next wgrad launches fills with cycle counters; NULL = off */
void dp_debug_set_buffer(void* p) { g_wg_dbg = reinterpret_cast<unsigned long long*>(p); }

size_t dp_conv2d_wgrad_tc_workspace(int B, int H, int W, int Cin, int Cout, int KS) {
  WgPlan p;
  if (wg_plan(p, B, H, W, Cout, Cin, KS, KS / 2, 1)) return 0;
  return (size_t)p.a.psplit * KS * KS * Cout * Cin * sizeof(float);
}

int dp_conv2d_wgrad_tc(const void* x, long long x_ld, const void* dy, long long dy_ld, int B, int H, int W, int Cin,
                       int Cout, int KS, float* grad_oihw, int accumulate, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  return dp_conv2d_wgrad_tc_fused(x, x_ld, dy, dy_ld, B, H, W, Cin, Cout, KS, grad_oihw, accumulate, workspace,
                                  workspace_bytes, nullptr, 0, stream);
}

int dp_conv2d_wgrad_tc_fused(const void* x, long long x_ld, const void* dy, long long dy_ld, int B, int H, int W, int Cin,
                             int Cout, int KS, float* grad_oihw, int accumulate, void* workspace, size_t workspace_bytes,
                             const float* pre_scale_shift, int pre_act, cudaStream_t stream) {
  DP_CHECK_ARG(x && dy && grad_oihw && workspace, "dp_conv2d_wgrad_tc: null pointer");
  DP_CHECK_ARG(KS == 3 || KS == 1, "dp_conv2d_wgrad_tc: kernel size %d", KS);
  DP_CHECK_ARG(Cin % 8 == 0 && Cout % 8 == 0 && x_ld % 8 == 0 && dy_ld % 8 == 0,
               "dp_conv2d_wgrad_tc: channels / strides must be multiples of 8");
  WgPlan p;
  int rc = wg_plan(p, B, H, W, Cout, Cin, KS, KS / 2, 1);
  if (rc) return rc;
  WgPlaneT planes[kMaxKW * 2];
  for (int i = 0; i < kMaxKW * 2; ++i) planes[i] = WgPlaneT{x, x_ld, (long long)W * x_ld, (long long)H * W * x_ld, H, W};
  return wg_launch(p, dy, dy_ld, planes, B, grad_oihw, accumulate, workspace, workspace_bytes, stream, pre_scale_shift,
                   pre_act);
}

/* grad[cp][ct][ky][kx] (+)= sum over plain-grid pixels p of P[p][cp] * T[2p - pad + k][ct]  (K x K taps, stride 2):
 *   nn.Conv2d(k3,s2,p1):          P = dY (B,Ho,Wo,O), T = X  (B,Hi,Wi,I)  -> weight.grad [O][I][3][3]
 *   nn.ConvTranspose2d(k4,s2,p1): P = X  (B,Hi,Wi,I), T = dY (B,Ho,Wo,O)  -> weight.grad [I][O][4][4] */
size_t dp_conv2d_wgrad_tc_s2_workspace(int B, int Hp, int Wp, int Cp, int Ct, int K, int pad) {
  WgPlan p;
  if (wg_plan(p, B, Hp, Wp, Cp, Ct, K, pad, 2)) return 0;
  return (size_t)p.a.psplit * K * K * Cp * Ct * sizeof(float);
}

int dp_conv2d_wgrad_tc_s2(const void* P, long long p_ld, int Hp, int Wp, int Cp, const void* T, long long t_ld, int Ht,
                          int Wt, int Ct, int B, int K, int pad, float* grad, int accumulate, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream) {
  DP_CHECK_ARG(P && T && grad && workspace, "dp_conv2d_wgrad_tc_s2: null pointer");
  DP_CHECK_ARG(Cp % 8 == 0 && Ct % 8 == 0 && p_ld % 8 == 0 && t_ld % 8 == 0,
               "dp_conv2d_wgrad_tc_s2: channels / strides must be multiples of 8");
  WgPlan p;
  int rc = wg_plan(p, B, Hp, Wp, Cp, Ct, K, pad, 2);
  if (rc) return rc;
  // planes[kx*2 + py]: parity plane (py, px(kx)) of T
  WgPlaneT planes[kMaxKW * 2];
  for (int i = 0; i < kMaxKW * 2; ++i) planes[i] = WgPlaneT{T, t_ld, (long long)Wt * t_ld, (long long)Ht * Wt * t_ld, Ht, Wt};
  const bf16* tb = reinterpret_cast<const bf16*>(T);
  for (int kx = 0; kx < K; ++kx)
    for (int py = 0; py < 2; ++py) {
      const int px = (kx - pad) & 1;
      WgPlaneT& q = planes[kx * 2 + py];
      q.base = tb + ((long long)py * Wt + px) * t_ld;
      q.ld_px = 2 * t_ld; q.ld_row = 2LL * Wt * t_ld; q.ld_img = (long long)Ht * Wt * t_ld;
      q.Hp = (Ht - py + 1) / 2; q.Wp = (Wt - px + 1) / 2;
      if (q.Hp < 1) q.Hp = 1;
      if (q.Wp < 1) q.Wp = 1;
    }
  return wg_launch(p, P, p_ld, planes, B, grad, accumulate, workspace, workspace_bytes, stream);
}

}  // extern "C"
