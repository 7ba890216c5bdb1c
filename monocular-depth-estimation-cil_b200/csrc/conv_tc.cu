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
//   * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
//     warps 4..7 = epilogue (TMEM -> registers -> bias / residual / ReLU / BN partial sums -> global).
#include "common.cuh"
#include "tc.cuh"
#include "../../include/depth_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;

struct ConvArgs {
  int B, H, W, Cout;
  int KS, pad;
  int th, tw, tiles_y, tiles_x;
  int KB, kchunks;
  int BN, n_blocks;
  int stages, resident;
  uint32_t a_stage_bytes, b_tap_bytes, b_stage_bytes, row_bytes, layout, sbo, idesc;
  uint32_t a_box_bytes, b_box_bytes;  // bytes TMA actually writes per box (slots are rounded up to 1024)
  uint32_t resident_bytes;
  long long total_items;
  bf16* out;  long long out_ld;
  bf16* out2; long long out2_ld;
  const float* bias;
  const bf16* res; long long res_ld;
  const bf16* resb; long long resb_ld;
  int relu, relu2;
  float* stats;  // [gridDim.x][2][Cout] or null
};

struct __align__(8) Barriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t resident_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void decode_item(const ConvArgs& a, long long item, int& n, int& y0, int& x0, int& nb) {
  nb = (int)(item % a.n_blocks);
  long long m = item / a.n_blocks;
  int tx = (int)(m % a.tiles_x);
  m /= a.tiles_x;
  int ty = (int)(m % a.tiles_y);
  n = (int)(m / a.tiles_y);
  y0 = ty * a.th;
  x0 = tx * a.tw;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

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

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // layout: [resident weights][stages x (A | B)][stats][barriers]
  uint8_t* s_res = smem;
  uint8_t* s_stage = smem + a.resident_bytes;
  const uint32_t stage_bytes = a.a_stage_bytes + (a.resident ? 0u : a.b_stage_bytes);
  float* s_stats = reinterpret_cast<float*>(s_stage + (size_t)a.stages * stage_bytes);
  const int stats_floats = a.stats ? 4 * 2 * a.Cout : 0;
  Barriers* bars = reinterpret_cast<Barriers*>(reinterpret_cast<uint8_t*>(s_stats) + ((stats_floats * 4 + 15) & ~15));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) { tc::mbar_init(&bars->full[i], 1); tc::mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&bars->tmem_full[i], 1); tc::mbar_init(&bars->tmem_empty[i], 4); }
    tc::mbar_init(&bars->resident_full, 1);
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
  }
  if (warp == 2) tc::tmem_alloc(&bars->tmem_base, kTmemCols);
  for (int i = threadIdx.x; i < stats_floats; i += kThreads) s_stats[i] = 0.f;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = bars->tmem_base;
  const int ksteps = a.KS * a.kchunks;  // one k-step = (horizontal tap s, channel chunk kc)

  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    if (a.resident) {
      tc::mbar_expect_tx(&bars->resident_full, (uint32_t)(a.KS * a.KS * a.kchunks) * a.b_box_bytes);
      for (int s = 0; s < a.KS; ++s)
        for (int kc = 0; kc < a.kchunks; ++kc)
          for (int r = 0; r < a.KS; ++r)
            tc::tma_load_3d(s_res + (size_t)((s * a.kchunks + kc) * a.KS + r) * a.b_tap_bytes, &tmB,
                            &bars->resident_full, kc * a.KB, 0, r * a.KS + s);
    }
    uint32_t stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < a.total_items; item += gridDim.x) {
      int n, y0, x0, nb;
      decode_item(a, item, n, y0, x0, nb);
      for (int s = 0; s < a.KS; ++s) {
        for (int kc = 0; kc < a.kchunks; ++kc) {
          tc::mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sA = s_stage + (size_t)stage * stage_bytes;
          tc::mbar_expect_tx(&bars->full[stage], a.a_box_bytes + (a.resident ? 0u : (uint32_t)a.KS * a.b_box_bytes));
          tc::tma_load_4d(sA, &tmA, &bars->full[stage], kc * a.KB, x0 + s - a.pad, y0 - a.pad, n);
          if (!a.resident) {
            uint8_t* sB = sA + a.a_stage_bytes;
            for (int r = 0; r < a.KS; ++r)
              tc::tma_load_3d(sB + (size_t)r * a.b_tap_bytes, &tmB, &bars->full[stage], kc * a.KB, nb * a.BN,
                              r * a.KS + s);
          }
          if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================= MMA issuer (single thread) =================
    if (a.resident) { tc::mbar_wait(&bars->resident_full, 0); tc::fence_after_sync(); }
    uint32_t stage = 0, phase = 0;
    int it = 0;
    const int kk_n = a.KB / 16;
    for (long long item = blockIdx.x; item < a.total_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      tc::mbar_wait(&bars->tmem_empty[buf], ((it >> 1) & 1) ^ 1);
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem + (uint32_t)(buf * a.BN);
      uint32_t accumulate = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        tc::mbar_wait(&bars->full[stage], phase);
        tc::fence_after_sync();
        const uint32_t a_base = tc::smem_u32(s_stage + (size_t)stage * stage_bytes);
        const uint32_t b_base = a.resident ? tc::smem_u32(s_res) + (uint32_t)(ks * a.KS) * a.b_tap_bytes
                                           : a_base + a.a_stage_bytes;
        for (int r = 0; r < a.KS; ++r) {
          const uint32_t ar = a_base + (uint32_t)(r * a.tw) * a.row_bytes;
          const uint32_t br = b_base + (uint32_t)r * a.b_tap_bytes;
          for (int kk = 0; kk < kk_n; ++kk) {
            const uint64_t da = tc::make_smem_desc(ar + kk * 32, 16, a.sbo, a.layout);
            const uint64_t db = tc::make_smem_desc(br + kk * 32, 16, a.sbo, a.layout);
            tc::umma_bf16(d_tmem, da, db, a.idesc, accumulate);
            accumulate = 1;
          }
        }
        tc::umma_commit(&bars->empty[stage]);  // frees the smem slot when these MMAs have read it
        if (++stage == (uint32_t)a.stages) { stage = 0; phase ^= 1; }
      }
      tc::umma_commit(&bars->tmem_full[buf]);  // accumulator complete -> epilogue
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> global =================
    const int ew = warp - 4;  // == warp % 4: TMEM lanes [32*ew, 32*ew+32)
    const int m = ew * 32 + lane;
    const int py = m / a.tw, px = m - py * a.tw;
    int it = 0;
    for (long long item = blockIdx.x; item < a.total_items; item += gridDim.x, ++it) {
      int n, y0, x0, nb;
      decode_item(a, item, n, y0, x0, nb);
      const int buf = it & 1;
      const int y = y0 + py, x = x0 + px;
      const bool valid = (y < a.H) && (x < a.W);
      const long long pix = ((long long)n * a.H + y) * a.W + x;
      tc::mbar_wait(&bars->tmem_full[buf], (it >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t t_base = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * a.BN);
      for (int c0 = 0; c0 < a.BN; c0 += 16) {
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
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rr[j]);
            v[2 * j] += __low2float(h);
            v[2 * j + 1] += __high2float(h);
          }
        }
        if (a.resb && valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(a.resb + pix * a.resb_ld + n0);
          uint4 r0 = __ldg(rp);
          uint4 r1 = nvalid > 8 ? __ldg(rp + 1) : make_uint4(0, 0, 0, 0);
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rr[j]);
            v[2 * j] += __low2float(h);
            v[2 * j + 1] += __high2float(h);
          }
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
              s_stats[(ew * 2 + 0) * a.Cout + col] += cs;
              s_stats[(ew * 2 + 1) * a.Cout + col] += cq;
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
    if (a.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
      float* dst = a.stats + (size_t)blockIdx.x * 2 * a.Cout;
      for (int i = threadIdx.x - 128; i < 2 * a.Cout; i += 128) {
        const int which = i / a.Cout, col = i - which * a.Cout;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) s += s_stats[(w * 2 + which) * a.Cout + col];
        dst[i] = s;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem, kTmemCols);
}

struct Plan {
  ConvArgs a;
  size_t smem;
  int grid;
};

int make_plan(Plan& p, int B, int H, int W, int Cin, int Cout, int KS, int want_stats) {
  ConvArgs& a = p.a;
  a.B = B; a.H = H; a.W = W; a.Cout = Cout; a.KS = KS; a.pad = KS / 2;
  // patch shape: th*tw == 128, tw a multiple of 8 (swizzle-atom alignment of the vertical tap offsets)
  const int cand[5][2] = {{8, 16}, {16, 8}, {4, 32}, {2, 64}, {1, 128}};
  long long best = -1;
  for (int i = 0; i < 5; ++i) {
    long long t = (long long)dp::ceil_div(H, cand[i][0]) * dp::ceil_div(W, cand[i][1]);
    if (best < 0 || t < best) { best = t; a.th = cand[i][0]; a.tw = cand[i][1]; }
  }
  a.tiles_y = dp::ceil_div(H, a.th);
  a.tiles_x = dp::ceil_div(W, a.tw);
  a.KB = Cin > 32 ? 64 : (Cin > 16 ? 32 : 16);
  a.kchunks = dp::ceil_div(Cin, a.KB);
  a.n_blocks = dp::ceil_div(Cout, 160);
  a.BN = ((dp::ceil_div(Cout, a.n_blocks) + 15) / 16) * 16;
  a.row_bytes = a.KB * 2;
  a.layout = tc::swizzle_for_row_bytes(a.row_bytes);
  a.sbo = 8 * a.row_bytes;
  a.idesc = tc::make_idesc_bf16(128, a.BN, 0, 0);
  a.a_box_bytes = (uint32_t)((a.th + 2 * a.pad) * a.tw * a.row_bytes);
  a.b_box_bytes = (uint32_t)(a.BN * a.row_bytes);
  a.a_stage_bytes = (a.a_box_bytes + 1023u) & ~1023u;
  a.b_tap_bytes = (a.b_box_bytes + 1023u) & ~1023u;
  a.b_stage_bytes = a.b_tap_bytes * KS;
  const size_t all_w = (size_t)KS * KS * a.kchunks * a.b_tap_bytes;
  a.resident = (a.n_blocks == 1 && all_w <= 100 * 1024) ? 1 : 0;
  a.resident_bytes = a.resident ? (uint32_t)all_w : 0u;
  const size_t stats_bytes = want_stats ? ((size_t)4 * 2 * Cout * 4 + 15) & ~size_t(15) : 0;
  const size_t fixed = 1024 + a.resident_bytes + stats_bytes + sizeof(Barriers) + 64;
  const size_t budget = 220 * 1024;
  const size_t stage = a.a_stage_bytes + (a.resident ? 0 : a.b_stage_bytes);
  if (fixed + 2 * stage > budget)
    return dp_set_error(DP_ERR_UNSUPPORTED, "conv_tc: shape needs %zu B of shared memory", fixed + 2 * stage);
  int st = (int)((budget - fixed) / stage);
  a.stages = st > kMaxStages ? kMaxStages : st;
  p.smem = fixed + (size_t)a.stages * stage;
  a.total_items = (long long)B * a.tiles_y * a.tiles_x * a.n_blocks;
  p.grid = (int)(a.total_items < dp::kNumSMs ? a.total_items : dp::kNumSMs);
  return DP_OK;
}

}  // namespace

extern "C" {

/* number of CTAs (= rows of the BN-statistics partials buffer) dp_conv2d_tc launches for this shape */
int dp_conv2d_tc_grid(int B, int H, int W, int Cin, int Cout, int KS) {
  Plan p;
  if (make_plan(p, B, H, W, Cin, Cout, KS, 0)) return -1;
  return p.grid;
}

int dp_conv2d_tc(const void* x, long long x_ld, int B, int H, int W, int Cin, const void* w_packed, int Cin_p,
                 int Cout, int KS, const float* bias, const void* residual, long long res_ld, const void* residual2,
                 long long res2_ld, int relu, void* out, long long out_ld, void* out2, long long out2_ld, int relu2,
                 float* stats_partials, cudaStream_t stream) {
  DP_CHECK_ARG(x && w_packed && (out || out2), "dp_conv2d_tc: null pointer");
  DP_CHECK_ARG(KS == 3 || KS == 1, "dp_conv2d_tc: kernel size %d (only 1 and 3, stride 1)", KS);
  DP_CHECK_ARG(Cin % 8 == 0 && Cout % 8 == 0 && Cin_p % 8 == 0 && Cin_p >= Cin,
               "dp_conv2d_tc: channels must be multiples of 8 (Cin %d Cin_p %d Cout %d)", Cin, Cin_p, Cout);
  DP_CHECK_ARG(x_ld % 8 == 0 && (!out || out_ld % 8 == 0) && (!out2 || out2_ld % 8 == 0) &&
               (!residual || res_ld % 8 == 0) && (!residual2 || res2_ld % 8 == 0),
               "dp_conv2d_tc: pixel strides must be multiples of 8 elements");
  DP_CHECK_ARG(B > 0 && H > 0 && W > 0, "dp_conv2d_tc: bad shape");
  Plan p;
  int rc = make_plan(p, B, H, W, Cin, Cout, KS, stats_partials != nullptr);
  if (rc) return rc;
  ConvArgs& a = p.a;
  a.out = reinterpret_cast<bf16*>(out); a.out_ld = out_ld;
  a.out2 = reinterpret_cast<bf16*>(out2); a.out2_ld = out2_ld;
  a.bias = bias;
  a.res = reinterpret_cast<const bf16*>(residual); a.res_ld = res_ld;
  a.resb = reinterpret_cast<const bf16*>(residual2); a.resb_ld = res2_ld;
  a.relu = relu; a.relu2 = relu2;
  a.stats = stats_partials;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)x_ld * 2, (uint64_t)W * x_ld * 2, (uint64_t)H * W * x_ld * 2};
    uint32_t box[4] = {(uint32_t)a.KB, (uint32_t)a.tw, (uint32_t)(a.th + 2 * a.pad), 1};
    rc = dp_make_tmap_bf16(&tmA, x, 4, dims, str, box, nullptr, a.row_bytes);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin_p, (uint64_t)Cout, (uint64_t)(KS * KS)};
    uint64_t str[2] = {(uint64_t)Cin_p * 2, (uint64_t)Cout * Cin_p * 2};
    uint32_t box[3] = {(uint32_t)a.KB, (uint32_t)a.BN, 1};
    rc = dp_make_tmap_bf16(&tmB, w_packed, 3, dims, str, box, nullptr, a.row_bytes);
    if (rc) return rc;
  }
  cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) return dp_set_error(DP_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", p.smem, cudaGetErrorString(e));
  conv_tc_kernel<<<p.grid, kThreads, p.smem, stream>>>(tmA, tmB, a);
  DP_CHECK_LAUNCH("conv_tc_kernel");
  return DP_OK;
}

}  // extern "C"
