// CrossAttention token path (reference midas_semantics.py:63-127) for dim = 32 (8 heads x 4):
//   * LayerNorm + Linear fused per token (norm_q/q, norm_k/k, norm_v/v, norm_out/proj), forward and backward;
//   * the window loop of midas_semantics.py:93-112 in its exact "last writer" closed form: the reference
//     overwrites out[range] window after window, so token i ends up attending over the key range of the LAST
//     window whose flattened range contains i, and only that window's result receives gradient.  The host
//     passes the resulting (query run, key range) segments; one flash-style kernel evaluates them
//     (4.3 M score entries per image instead of the reference's 21.9 M), and two deterministic kernels
//     (query-centric for dq, key-centric for dk/dv) give the backward without atomics.
#include "common.cuh"
#include "../../include/depth_b200.h"

namespace {

using namespace dp;
constexpr int D = 32;       // embedding dim == warp size
constexpr int NH = 8, HD = 4;

template <typename T> __device__ __forceinline__ float ldv(const T* p);
template <> __device__ __forceinline__ float ldv<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldv<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stv(T* p, float v);
template <> __device__ __forceinline__ void stv<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stv<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// one warp per token, lane = channel
template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) ln_linear_fwd_kernel(const InT* __restrict__ x, long long x_ld, size_t ntok,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, const float* __restrict__ W /*[out][in]*/,
                                                            const float* __restrict__ bias, OutT* __restrict__ out,
                                                            long long out_ld) {
  __shared__ float sW[D][D + 1];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) sW[i / D][i % D] = W[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const float g = gamma[lane], b = beta[lane], bo = bias ? bias[lane] : 0.f;
  for (size_t t = warp; t < ntok; t += nwarps) {
    const float v = ldv(x + t * x_ld + lane);
    const float mean = warp_sum(v) * (1.f / D);
    const float dv = v - mean;
    const float var = warp_sum(dv * dv) * (1.f / D);
    const float y = dv * rsqrtf(var + eps) * g + b;
    float acc = bo;
#pragma unroll
    for (int c = 0; c < D; ++c) acc += sW[lane][c] * __shfl_sync(0xffffffffu, y, c);
    stv(out + t * out_ld + lane, acc);
  }
}

// backward: dx (same dtype family as x: bf16 in -> bf16 grad, fp32 in -> fp32 grad), per-block partials of
// [dW 32x32 | dbias 32 | dgamma 32 | dbeta 32] = 1120 floats
constexpr int LNL_NPAR = D * D + 3 * D;
template <typename InT, typename GT>
__global__ void __launch_bounds__(256) ln_linear_bwd_kernel(const InT* __restrict__ x, long long x_ld, size_t ntok,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, const float* __restrict__ W,
                                                            const GT* __restrict__ dout, long long do_ld,
                                                            InT* __restrict__ dx, long long dx_ld,
                                                            float* __restrict__ partial) {
  __shared__ float sred[8][LNL_NPAR];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  float Wcol[D];   // W[j][lane] for all j
  float dWrow[D];  // dW[lane][c] for all c
#pragma unroll
  for (int j = 0; j < D; ++j) { Wcol[j] = W[j * D + lane]; dWrow[j] = 0.f; }
  float dbias = 0.f, dgam = 0.f, dbet = 0.f;
  const float g = gamma[lane], b = beta[lane];
  for (size_t t = warp; t < ntok; t += nwarps) {
    const float v = ldv(x + t * x_ld + lane);
    const float mean = warp_sum(v) * (1.f / D);
    const float dvv = v - mean;
    const float var = warp_sum(dvv * dvv) * (1.f / D);
    const float rstd = rsqrtf(var + eps);
    const float xhat = dvv * rstd;
    const float y = xhat * g + b;
    const float go = ldv(dout + t * do_ld + lane);  // dout[lane]
    dbias += go;
    float dy = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const float goj = __shfl_sync(0xffffffffu, go, j);
      const float yj = __shfl_sync(0xffffffffu, y, j);
      dy += goj * Wcol[j];        // dy[lane] = sum_j dout[j] W[j][lane]
      dWrow[j] += go * yj;        // dW[lane][j] += dout[lane] * y[j]
    }
    dgam += dy * xhat;
    dbet += dy;
    const float dxh = dy * g;
    const float m1 = warp_sum(dxh) * (1.f / D);
    const float m2 = warp_sum(dxh * xhat) * (1.f / D);
    if (dx) stv(dx + t * dx_ld + lane, rstd * (dxh - m1 - xhat * m2));
  }
#pragma unroll
  for (int j = 0; j < D; ++j) sred[wib][lane * D + j] = dWrow[j];
  sred[wib][D * D + lane] = dbias;
  sred[wib][D * D + D + lane] = dgam;
  sred[wib][D * D + 2 * D + lane] = dbet;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < LNL_NPAR; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += sred[w][i];
    partial[(size_t)blockIdx.x * LNL_NPAR + i] = s;
  }
}

__global__ void lnl_reduce_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ dW,
                                  float* __restrict__ dbias, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                  int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= LNL_NPAR) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += (double)partial[(size_t)b * LNL_NPAR + i];
  float* dst;
  int k;
  if (i < D * D) { dst = dW; k = i; }
  else if (i < D * D + D) { dst = dbias; k = i - D * D; }
  else if (i < D * D + 2 * D) { dst = dgamma; k = i - D * D - D; }
  else { dst = dbeta; k = i - D * D - 2 * D; }
  if (dst) dst[k] = accumulate ? dst[k] + (float)s : (float)s;
}

// ---- segmented attention ---------------------------------------------------------------------------------------
// items: int4 {q0, nq (<=32), k_lo, k_hi}; block = 8 warps (warp = head), lane = query
constexpr int KCH = 128;  // keys staged per chunk

// The three attention kernels are issue-bound (head dim 4: ~15-25 instructions per (query, key, head) score), so the
// arithmetic is packed: dot products and accumulations as f32x2 FMAs, scores kept in log2 units (the softmax scale and
// log2(e) are folded into the query once per thread) so that every exponential is a bare MUFU.EX2.
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
struct F4 { float2 lo, hi; };                              // a float4 as two packed pairs: (x, y), (z, w)
__device__ __forceinline__ F4 ld4(const float* p) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  return F4{make_float2(t.x, t.y), make_float2(t.z, t.w)};
}
__device__ __forceinline__ float dot4(const F4& a, const F4& b) {
  const float2 t = __ffma2_rn(a.hi, b.hi, __fmul2_rn(a.lo, b.lo));
  return t.x + t.y;
}
__device__ __forceinline__ void axpy4(F4& acc, float s, const F4& v) {
  const float2 ss = make_float2(s, s);
  acc.lo = __ffma2_rn(ss, v.lo, acc.lo);
  acc.hi = __ffma2_rn(ss, v.hi, acc.hi);
}

__global__ void __launch_bounds__(256) attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                       const float* __restrict__ v, int N, float scale,
                                                       const int4* __restrict__ items, float* __restrict__ out,
                                                       float* __restrict__ lse /*[B][N][NH]*/) {
  __shared__ float sK[KCH][D];
  __shared__ float sV[KCH][D];
  const int4 it = items[blockIdx.x];
  const int b = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = (size_t)b * N;
  const bool active = lane < it.y;
  const int qi = it.x + lane;
  F4 qv{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  if (active) qv = ld4(q + (base + qi) * D + h * HD);
  {
    const float2 sc2 = make_float2(scale * kLog2e, scale * kLog2e);   // scores in log2 units
    qv.lo = __fmul2_rn(qv.lo, sc2); qv.hi = __fmul2_rn(qv.hi, sc2);
  }
  float m = -INFINITY, l = 0.f;
  F4 acc{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  for (int k0 = it.z; k0 < it.w; k0 += KCH) {
    const int nk = min(KCH, it.w - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < nk * (D / 4); e += blockDim.x) {
      const int j = e / (D / 4), c4 = e % (D / 4);
      reinterpret_cast<float4*>(&sK[j][0])[c4] = __ldg(reinterpret_cast<const float4*>(k + (base + k0 + j) * D) + c4);
      reinterpret_cast<float4*>(&sV[j][0])[c4] = __ldg(reinterpret_cast<const float4*>(v + (base + k0 + j) * D) + c4);
    }
    __syncthreads();
    // online softmax over groups of four keys: one rescale (exp) per group instead of one per key - the loop is bound
    // by the MUFU rate (two exponentials per score otherwise)
    for (int j = 0; j < nk; j += 4) {
      float sc[4];
      F4 vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = j + u < nk ? j + u : nk - 1;
        const F4 kk = ld4(&sK[jj][h * HD]);
        vv[u] = ld4(&sV[jj][h * HD]);
        const float s = dot4(qv, kk);
        sc[u] = j + u < nk ? s : -INFINITY;
      }
      const float mn = fmaxf(fmaxf(m, fmaxf(sc[0], sc[1])), fmaxf(sc[2], sc[3]));
      const float corr = ex2(m - mn);
      l *= corr;
      const float2 c2 = make_float2(corr, corr);
      acc.lo = __fmul2_rn(acc.lo, c2); acc.hi = __fmul2_rn(acc.hi, c2);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float p = ex2(sc[u] - mn);
        l += p;
        axpy4(acc, p, vv[u]);
      }
      m = mn;
    }
  }
  if (active) {
    const float inv = 1.f / l;
    *reinterpret_cast<float4*>(out + (base + qi) * D + h * HD) =
        make_float4(acc.lo.x * inv, acc.lo.y * inv, acc.hi.x * inv, acc.hi.y * inv);
    lse[(base + qi) * NH + h] = (m + __log2f(l)) * kLn2;      // natural-log units, as before
  }
}

// dq (query-centric): also writes delta[i][h] = dout_i . out_i
__global__ void __launch_bounds__(256) attn_bwd_q_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                         const float* __restrict__ v, const float* __restrict__ out,
                                                         const float* __restrict__ dout, const float* __restrict__ lse,
                                                         int N, float scale, const int4* __restrict__ items,
                                                         float* __restrict__ dq, float* __restrict__ delta) {
  __shared__ float sK[KCH][D];
  __shared__ float sV[KCH][D];
  const int4 it = items[blockIdx.x];
  const int b = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = (size_t)b * N;
  const bool active = lane < it.y;
  const int qi = it.x + lane;
  F4 qv{make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, go = qv, ov = qv;
  float L = 0.f;
  if (active) {
    qv = ld4(q + (base + qi) * D + h * HD);
    go = ld4(dout + (base + qi) * D + h * HD);
    ov = ld4(out + (base + qi) * D + h * HD);
    L = lse[(base + qi) * NH + h] * kLog2e;                   // log2 units
  }
  const float dl = dot4(go, ov);
  const float2 sc2 = make_float2(scale * kLog2e, scale * kLog2e);
  const F4 q2{__fmul2_rn(qv.lo, sc2), __fmul2_rn(qv.hi, sc2)};   // scores in log2 units
  F4 acc{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  for (int k0 = it.z; k0 < it.w; k0 += KCH) {
    const int nk = min(KCH, it.w - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < nk * (D / 4); e += blockDim.x) {
      const int j = e / (D / 4), c4 = e % (D / 4);
      reinterpret_cast<float4*>(&sK[j][0])[c4] = __ldg(reinterpret_cast<const float4*>(k + (base + k0 + j) * D) + c4);
      reinterpret_cast<float4*>(&sV[j][0])[c4] = __ldg(reinterpret_cast<const float4*>(v + (base + k0 + j) * D) + c4);
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < nk; ++j) {
      const F4 kk = ld4(&sK[j][h * HD]);
      const F4 vv = ld4(&sV[j][h * HD]);
      const float p = ex2(dot4(q2, kk) - L);
      const float ds = p * (dot4(go, vv) - dl);              // the softmax scale is applied once, after the loop
      axpy4(acc, ds, kk);
    }
  }
  if (active) {
    *reinterpret_cast<float4*>(dq + (base + qi) * D + h * HD) =
        make_float4(acc.lo.x * scale, acc.lo.y * scale, acc.hi.x * scale, acc.hi.y * scale);
    delta[(base + qi) * NH + h] = dl;
  }
}

// dk, dv (key-centric): block = 32 consecutive keys x 8 heads; loops over every segment whose key range meets them.
// segs: int4 {q_lo, q_hi, k_lo, k_hi} (un-chunked query runs)
constexpr int QCH = 64;
__global__ void __launch_bounds__(256) attn_bwd_kv_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                          const float* __restrict__ v, const float* __restrict__ dout,
                                                          const float* __restrict__ lse, const float* __restrict__ delta,
                                                          int N, float scale, const int4* __restrict__ segs, int nseg,
                                                          float* __restrict__ dk, float* __restrict__ dv) {
  __shared__ float sQ[QCH][D];
  __shared__ float sG[QCH][D];
  __shared__ float sL[QCH][NH];
  __shared__ float sDl[QCH][NH];
  const int b = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = (size_t)b * N;
  const int j0 = blockIdx.x * 32, j = j0 + lane;
  const bool kvalid = j < N;
  F4 kk{make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, vv = kk;
  if (kvalid) {
    kk = ld4(k + (base + j) * D + h * HD);
    vv = ld4(v + (base + j) * D + h * HD);
  }
  const float2 sc2 = make_float2(scale * kLog2e, scale * kLog2e);
  const F4 k2{__fmul2_rn(kk.lo, sc2), __fmul2_rn(kk.hi, sc2)};   // scores in log2 units
  F4 adk{make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, adv = adk;
  for (int sidx = 0; sidx < nseg; ++sidx) {
    const int4 sg = segs[sidx];
    if (sg.w <= j0 || sg.z >= j0 + 32) continue;  // block-uniform: key ranges do not meet
    const bool mine = kvalid && j >= sg.z && j < sg.w;
    for (int q0 = sg.x; q0 < sg.y; q0 += QCH) {
      const int nq = min(QCH, sg.y - q0);
      __syncthreads();
      for (int e = threadIdx.x; e < nq * (D / 4); e += blockDim.x) {
        const int i = e / (D / 4), c4 = e % (D / 4);
        reinterpret_cast<float4*>(&sQ[i][0])[c4] = __ldg(reinterpret_cast<const float4*>(q + (base + q0 + i) * D) + c4);
        reinterpret_cast<float4*>(&sG[i][0])[c4] = __ldg(reinterpret_cast<const float4*>(dout + (base + q0 + i) * D) + c4);
      }
      for (int e = threadIdx.x; e < nq * NH; e += blockDim.x) {
        sL[e / NH][e % NH] = lse[(base + q0) * NH + e] * kLog2e;     // log2 units
        sDl[e / NH][e % NH] = delta[(base + q0) * NH + e];
      }
      __syncthreads();
      if (mine) {
#pragma unroll 4
        for (int i = 0; i < nq; ++i) {
          const F4 qq = ld4(&sQ[i][h * HD]);
          const F4 gg = ld4(&sG[i][h * HD]);
          const float p = ex2(dot4(qq, k2) - sL[i][h]);
          const float ds = p * (dot4(gg, vv) - sDl[i][h]);    // the softmax scale is applied once, after the loops
          axpy4(adv, p, gg);
          axpy4(adk, ds, qq);
        }
      }
    }
  }
  if (kvalid) {
    *reinterpret_cast<float4*>(dk + (base + j) * D + h * HD) =
        make_float4(adk.lo.x * scale, adk.lo.y * scale, adk.hi.x * scale, adk.hi.y * scale);
    *reinterpret_cast<float4*>(dv + (base + j) * D + h * HD) = make_float4(adv.lo.x, adv.lo.y, adv.hi.x, adv.hi.y);
  }
}

constexpr int kLnlBlocks = 2 * kNumSMs;

}  // namespace

extern "C" {

int dp_lnl_blocks(void) { return kLnlBlocks; }
int dp_lnl_partial_floats(void) { return LNL_NPAR; }

/* LayerNorm(32) + Linear(32->32).  x: [ntok][x_ld] bf16 (x_is_f32=0) or fp32; out fp32 (out_is_bf16=0) or bf16. */
int dp_ln_linear_fwd(const void* x, long long x_ld, int x_is_f32, size_t ntok, int dim, const float* gamma,
                     const float* beta, float eps, const float* W, const float* bias, void* out, long long out_ld,
                     int out_is_bf16, cudaStream_t stream) {
  DP_CHECK_ARG(x && gamma && beta && W && out, "dp_ln_linear_fwd: null pointer");
  if (dim != D) return dp_set_error(DP_ERR_UNSUPPORTED, "dp_ln_linear_fwd: dim %d (only 32 = features 64)", dim);
  int blocks = (int)((ntok + 7) / 8);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (blocks < 1) blocks = 1;
  if (!x_is_f32 && !out_is_bf16)
    ln_linear_fwd_kernel<bf16, float><<<blocks, 256, 0, stream>>>((const bf16*)x, x_ld, ntok, gamma, beta, eps, W, bias, (float*)out, out_ld);
  else if (x_is_f32 && out_is_bf16)
    ln_linear_fwd_kernel<float, bf16><<<blocks, 256, 0, stream>>>((const float*)x, x_ld, ntok, gamma, beta, eps, W, bias, (bf16*)out, out_ld);
  else if (x_is_f32 && !out_is_bf16)
    ln_linear_fwd_kernel<float, float><<<blocks, 256, 0, stream>>>((const float*)x, x_ld, ntok, gamma, beta, eps, W, bias, (float*)out, out_ld);
  else
    ln_linear_fwd_kernel<bf16, bf16><<<blocks, 256, 0, stream>>>((const bf16*)x, x_ld, ntok, gamma, beta, eps, W, bias, (bf16*)out, out_ld);
  DP_CHECK_LAUNCH("ln_linear_fwd_kernel");
  return DP_OK;
}

/* dx has x's dtype; dout is fp32 (dout_is_bf16=0) or bf16.  partial: float[dp_lnl_blocks()][dp_lnl_partial_floats()].
 * dW [32][32], dbias (may be NULL), dgamma, dbeta get the reduced parameter gradients. */
int dp_ln_linear_bwd(const void* x, long long x_ld, int x_is_f32, size_t ntok, int dim, const float* gamma,
                     const float* beta, float eps, const float* W, const void* dout, long long do_ld, int dout_is_bf16,
                     void* dx, long long dx_ld, float* partial, float* dW, float* dbias, float* dgamma, float* dbeta,
                     int accumulate, cudaStream_t stream) {
  DP_CHECK_ARG(x && gamma && beta && W && dout && partial && dW && dgamma && dbeta, "dp_ln_linear_bwd: null pointer");
  if (dim != D) return dp_set_error(DP_ERR_UNSUPPORTED, "dp_ln_linear_bwd: dim %d (only 32)", dim);
  if (!x_is_f32 && !dout_is_bf16)
    ln_linear_bwd_kernel<bf16, float><<<kLnlBlocks, 256, 0, stream>>>((const bf16*)x, x_ld, ntok, gamma, beta, eps, W, (const float*)dout, do_ld, (bf16*)dx, dx_ld, partial);
  else if (x_is_f32 && dout_is_bf16)
    ln_linear_bwd_kernel<float, bf16><<<kLnlBlocks, 256, 0, stream>>>((const float*)x, x_ld, ntok, gamma, beta, eps, W, (const bf16*)dout, do_ld, (float*)dx, dx_ld, partial);
  else if (x_is_f32 && !dout_is_bf16)
    ln_linear_bwd_kernel<float, float><<<kLnlBlocks, 256, 0, stream>>>((const float*)x, x_ld, ntok, gamma, beta, eps, W, (const float*)dout, do_ld, (float*)dx, dx_ld, partial);
  else
    ln_linear_bwd_kernel<bf16, bf16><<<kLnlBlocks, 256, 0, stream>>>((const bf16*)x, x_ld, ntok, gamma, beta, eps, W, (const bf16*)dout, do_ld, (bf16*)dx, dx_ld, partial);
  DP_CHECK_LAUNCH("ln_linear_bwd_kernel");
  lnl_reduce_kernel<<<dp::ceil_div(LNL_NPAR, 128), 128, 0, stream>>>(partial, kLnlBlocks, dW, dbias, dgamma, dbeta, accumulate);
  DP_CHECK_LAUNCH("lnl_reduce_kernel");
  return DP_OK;
}

/* q,k,v,out: fp32 [B][N][32]; items: device int4[nitems] {q0, nq<=32, k_lo, k_hi}; lse: fp32 [B][N][8] */
int dp_attn_fwd(const float* q, const float* k, const float* v, int B, int N, float scale, const void* items,
                int nitems, float* out, float* lse, cudaStream_t stream) {
  DP_CHECK_ARG(q && k && v && items && out && lse && nitems > 0, "dp_attn_fwd: bad arguments");
  attn_fwd_kernel<<<dim3(nitems, B), 256, 0, stream>>>(q, k, v, N, scale, (const int4*)items, out, lse);
  DP_CHECK_LAUNCH("attn_fwd_kernel");
  return DP_OK;
}

/* delta: fp32 scratch [B][N][8]; segs: device int4[nseg] {q_lo, q_hi, k_lo, k_hi} */
int dp_attn_bwd(const float* q, const float* k, const float* v, const float* out, const float* dout, const float* lse,
                int B, int N, float scale, const void* items, int nitems, const void* segs, int nseg, float* dq,
                float* dk, float* dv, float* delta, cudaStream_t stream) {
  DP_CHECK_ARG(q && k && v && out && dout && lse && items && segs && dq && dk && dv && delta, "dp_attn_bwd: null pointer");
  attn_bwd_q_kernel<<<dim3(nitems, B), 256, 0, stream>>>(q, k, v, out, dout, lse, N, scale, (const int4*)items, dq, delta);
  DP_CHECK_LAUNCH("attn_bwd_q_kernel");
  attn_bwd_kv_kernel<<<dim3(dp::ceil_div(N, 32), B), 256, 0, stream>>>(q, k, v, dout, lse, delta, N, scale,
                                                                        (const int4*)segs, nseg, dk, dv);
  DP_CHECK_LAUNCH("attn_bwd_kv_kernel");
  return DP_OK;
}

}  // extern "C"
