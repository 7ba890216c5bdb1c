// Fused, single-pass loss and metric reductions (HBM-bound by design).
//
// Reference arithmetic being replaced (all /root/reference/src):
//   util.py:129-156  scale_invariant_loss      util.py:90-127   silog_loss
//   util.py:24-44    gradient_loss             util.py:46-88    edge_aware_loss
//   util.py:210-219  absolute_relative_error   util.py:183-207  delta_thres
//   main.py:51-89    combined_loss             main.py:254-392  evaluate_model metric set
//
// One pass over (pred, target[, rgb]) produces per-sample raw moments in fp64 (dp_depth_moments);
// tiny combine kernels turn moments into the reference's scalars; the backward is one stencil pass.
// Algorithmic HBM bytes: 8 B/px forward (pred+target fp32 once), 12 B/px backward (re-read + grad
// write); the edge term adds 12 B/px of RGB per pass plus a 12 B/px min/max pre-pass.
#include "common.cuh"
#include "tc.cuh"
#include <cstdlib>
#include "../../include/depth_b200.h"

namespace {

using namespace dp;

constexpr int NMOM = DP_NMOM;
constexpr int TPB = 256;

struct MomArgs {
  const float* pred;
  const float* target;
  const float* rgb;
  int B, H, W, chunks;
  unsigned flags;
  float eps;
  const float* mm_partials;  // [mm_n][2] per-block (min,max) of the rgb gradient magnitude
  int mm_n;
  double* partials;  // [B][chunks][NMOM]
};

// RGB gradient magnitude at (y,x): sqrt(mean_c dx^2 + mean_c dy^2), right/bottom zero padded (util.py:58-67)
__device__ __forceinline__ float grad_mag(const float* rgb_b, int H, int W, int y, int x) {
  const size_t plane = (size_t)H * W;
  float sx = 0.f, sy = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* r = rgb_b + c * plane + (size_t)y * W + x;
    float v = __ldg(r);
    float dx = (x + 1 < W) ? fabsf(v - __ldg(r + 1)) : 0.f;
    float dy = (y + 1 < H) ? fabsf(v - __ldg(r + W)) : 0.f;
    sx += dx * dx;
    sy += dy * dy;
  }
  return sqrtf(sx / 3.0f + sy / 3.0f);
}

__global__ void __launch_bounds__(TPB) rgb_minmax_kernel(const float* __restrict__ rgb, int B, int H, int W,
                                                         float* __restrict__ out /*[grid][2]*/) {
  const size_t total = (size_t)B * H * W;
  float mn = INFINITY, mx = -INFINITY;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int b = (int)(i / ((size_t)H * W));
    int rem = (int)(i - (size_t)b * H * W);
    int y = rem / W, x = rem - y * W;
    float g = grad_mag(rgb + (size_t)b * 3 * H * W, H, W, y, x);
    mn = fminf(mn, g);
    mx = fmaxf(mx, g);
  }
  __shared__ float smn[32], smx[32];
  mn = warp_min(mn);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x < 32) {
    int nw = blockDim.x >> 5;
    mn = threadIdx.x < nw ? smn[threadIdx.x] : INFINITY;
    mx = threadIdx.x < nw ? smx[threadIdx.x] : -INFINITY;
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = mn; out[2 * blockIdx.x + 1] = mx; }
  }
}

__device__ __forceinline__ void reduce_minmax(const float* parts, int n, float& gmin, float& gmax) {
  __shared__ float s_mm[2];
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    mn = fminf(mn, parts[2 * i]);
    mx = fmaxf(mx, parts[2 * i + 1]);
  }
  __shared__ float smn[32], smx[32];
  mn = warp_min(mn);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
    s_mm[0] = mn; s_mm[1] = mx;
  }
  __syncthreads();
  gmin = s_mm[0];
  gmax = s_mm[1];
}

// One thread handles VEC consecutive pixels of one row.  Moments accumulate in fp64 per thread.
template <int VEC, bool STENCIL, bool EDGE>
__global__ void __launch_bounds__(TPB) moments_kernel(MomArgs a) {
  const int b = blockIdx.y;
  const int H = a.H, W = a.W;
  const int rows_per = (H + a.chunks - 1) / a.chunks;
  const int r0 = blockIdx.x * rows_per;
  const int r1 = min(H, r0 + rows_per);
  const float* __restrict__ P = a.pred + (size_t)b * H * W;
  const float* __restrict__ T = a.target + (size_t)b * H * W;
  const float* __restrict__ RGB = EDGE ? a.rgb + (size_t)b * 3 * H * W : nullptr;
  const unsigned flags = a.flags;
  const float eps = a.eps;

  float gmin = 0.f, ginv = 0.f;
  if (EDGE) {
    float gmax;
    reduce_minmax(a.mm_partials, a.mm_n, gmin, gmax);
    ginv = gmax - gmin + 1e-6f;  // util.py:70 (divide, not multiply-by-reciprocal, to match rounding)
  }

  double acc[NMOM];
#pragma unroll
  for (int k = 0; k < NMOM; ++k) acc[k] = 0.0;

  const int WV = (W + VEC - 1) / VEC;
  const int nitems = (r1 > r0) ? (r1 - r0) * WV : 0;
  for (int it = threadIdx.x; it < nitems; it += TPB) {
    const int y = r0 + it / WV;
    const int x0 = (it % WV) * VEC;
    float pv[VEC + 1], tv[VEC + 1], pd[VEC], td[VEC];
    const size_t off = (size_t)y * W + x0;
    if (VEC == 4) {
      float4 p4 = ldg4(P + off), t4 = ldg4(T + off);
      pv[0] = p4.x; pv[1] = p4.y; pv[2] = p4.z; pv[3] = p4.w;
      tv[0] = t4.x; tv[1] = t4.y; tv[2] = t4.z; tv[3] = t4.w;
    } else {
      pv[0] = __ldg(P + off);
      tv[0] = __ldg(T + off);
    }
    if (STENCIL) {
      const bool has_r = x0 + VEC < W;
      pv[VEC] = has_r ? __ldg(P + off + VEC) : 0.f;
      tv[VEC] = has_r ? __ldg(T + off + VEC) : 0.f;
      if (y + 1 < H) {
        if (VEC == 4) {
          float4 p4 = ldg4(P + off + W), t4 = ldg4(T + off + W);
          pd[0] = p4.x; pd[1] = p4.y; pd[2] = p4.z; pd[3] = p4.w;
          td[0] = t4.x; td[1] = t4.y; td[2] = t4.z; td[3] = t4.w;
        } else {
          pd[0] = __ldg(P + off + W);
          td[0] = __ldg(T + off + W);
        }
      }
    }
    // fp32 partial sums over the VEC pixels of this item, promoted to fp64 once per item
    float s1 = 0.f, s2 = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f, gx = 0.f, gy = 0.f, ex = 0.f, ey = 0.f, ar = 0.f,
          ab = 0.f, sq = 0.f, v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int x = x0 + j;
      if (VEC == 1 || x < W) {
        const float p = pv[j], t = tv[j];
        if (flags & (DP_F_SI | DP_F_SILOG)) {
          const float d = logf(p + eps) - logf(t + eps);
          s1 += d;
          s2 += d * d;
          if (t > 0.f) { m0 += 1.f; m1 += d; m2 += d * d; }
        }
        if (flags & DP_F_ABSREL) ar += fabsf(t - p) / (t + 1e-6f);
        if (flags & DP_F_M4) {
          const float e = fabsf(p - t);
          ab += e;
          sq += e * e;
          if (t > 1e-6f) {
            const float dv = logf(p > 1e-6f ? p : 1e-6f) - logf(t);
            v0 += 1.f; v1 += dv; v2 += dv * dv;
          }
        }
        if (STENCIL) {
          float e_x = 0.f, e_y = 0.f;
          if (x + 1 < W) e_x = fabsf(fabsf(p - pv[j + 1]) - fabsf(t - tv[j + 1]));
          if (y + 1 < H) e_y = fabsf(fabsf(p - pd[j]) - fabsf(t - td[j]));
          gx += e_x;
          gy += e_y;
          if (EDGE) {
            const float g = (grad_mag(RGB, H, W, y, x) - gmin) / ginv;
            ex += g * e_x;
            ey += g * e_y;
          }
        }
      }
    }
    acc[DP_M_S1] += s1; acc[DP_M_S2] += s2; acc[DP_M_M0] += m0; acc[DP_M_M1] += m1; acc[DP_M_M2] += m2;
    acc[DP_M_GX] += gx; acc[DP_M_GY] += gy; acc[DP_M_EX] += ex; acc[DP_M_EY] += ey; acc[DP_M_AR] += ar;
    acc[DP_M_AB] += ab; acc[DP_M_SQ] += sq; acc[DP_M_V0] += v0; acc[DP_M_V1] += v1; acc[DP_M_V2] += v2;
  }
  __shared__ double red[NMOM * 32];
  block_sum<NMOM>(acc, red);
  if (threadIdx.x == 0) {
    double* o = a.partials + ((size_t)b * a.chunks + blockIdx.x) * NMOM;
#pragma unroll
    for (int k = 0; k < NMOM; ++k) o[k] = acc[k];
  }
}

__global__ void moments_finalize_kernel(const double* __restrict__ partials, int chunks, double* __restrict__ out) {
  const int b = blockIdx.x;
  const int k = threadIdx.x;
  if (k >= NMOM) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += partials[((size_t)b * chunks + c) * NMOM + k];
  out[(size_t)b * NMOM + k] = s;
}

// ---- combine: moments -> the reference's scalars ------------------------------------------------
struct LossW {
  float w_si, w_silog, vf, w_grad, beta;
  int sqroot;
};

__global__ void loss_combine_kernel(const double* __restrict__ mom, int B, int H, int W, LossW w, unsigned flags,
                                    float* __restrict__ out /*[DP_NLOSS]*/, float* __restrict__ per_sample /*[B] or null*/) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double n = (double)H * W;
  double si = 0.0, M0 = 0.0, M1 = 0.0, M2 = 0.0, GX = 0.0, GY = 0.0, EX = 0.0, EY = 0.0, AR = 0.0;
  for (int b = 0; b < B; ++b) {
    const double* m = mom + (size_t)b * NMOM;
    double v = m[DP_M_S2] / n - (m[DP_M_S1] * m[DP_M_S1]) / (n * n);
    float vf32 = (float)v;
    if (w.sqroot) vf32 = sqrtf(vf32);
    if (per_sample) per_sample[b] = vf32;
    si += (double)vf32;
    M0 += m[DP_M_M0]; M1 += m[DP_M_M1]; M2 += m[DP_M_M2];
    GX += m[DP_M_GX]; GY += m[DP_M_GY]; EX += m[DP_M_EX]; EY += m[DP_M_EY]; AR += m[DP_M_AR];
  }
  const float si_loss = (float)(si / B);
  const double md = M1 / M0;
  const float silog = (float)(M2 / M0 - (double)w.vf * md * md);
  const float grad = (float)(GX / ((double)B * H * (W - 1))) + (float)(GY / ((double)B * (H - 1) * W));
  const float edge_raw = (float)(EX / ((double)B * n)) + (float)(EY / ((double)B * n));
  const float absrel = (float)(AR / ((double)B * n));
  const float t_si = si_loss * w.w_si;
  const float t_silog = (flags & DP_F_SILOG) ? silog * w.w_silog : 0.f;
  const float t_grad = (flags & DP_F_GRAD) ? grad * w.w_grad : 0.f;
  const float t_edge = (flags & DP_F_EDGE) ? w.beta * edge_raw : 0.f;
  out[DP_L_TOTAL] = t_si + t_silog + t_grad + t_edge;
  out[DP_L_SI] = t_si;
  out[DP_L_SILOG] = t_silog;
  out[DP_L_GRAD] = t_grad;
  out[DP_L_EDGE] = t_edge;
  out[DP_L_ABSREL] = absrel;
  out[DP_L_SI_RAW] = si_loss;
  out[DP_L_SILOG_RAW] = silog;
}

// ---- backward of combined loss w.r.t. pred --------------------------------------------------------
struct BwdArgs {
  const float* pred;
  const float* target;
  const float* rgb;
  const double* mom;       // [B][NMOM]
  const float* mm_partials;
  int mm_n;
  const float* grad_out;   // device scalar (dL/dtotal) or null => 1
  const float* si_scale;   // per-sample multiplier of the SI term or null
  float* grad_pred;
  int B, H, W;
  unsigned flags;
  float eps;
  LossW w;
};

template <bool STENCIL, bool EDGE>
__global__ void __launch_bounds__(TPB) loss_bwd_kernel(BwdArgs a) {
  const int H = a.H, W = a.W, B = a.B;
  const double n = (double)H * W;
  float gmin = 0.f, ginv = 1.f;
  if (EDGE) {
    float gmax;
    reduce_minmax(a.mm_partials, a.mm_n, gmin, gmax);
    ginv = gmax - gmin + 1e-6f;
  }
  __shared__ double s_tot[3];
  if (threadIdx.x == 0) {
    double M0 = 0, M1 = 0;
    for (int b = 0; b < B; ++b) { M0 += a.mom[(size_t)b * NMOM + DP_M_M0]; M1 += a.mom[(size_t)b * NMOM + DP_M_M1]; }
    s_tot[0] = M0; s_tot[1] = M1;
  }
  __syncthreads();
  const double M0 = s_tot[0], M1 = s_tot[1];
  const float go = a.grad_out ? __ldg(a.grad_out) : 1.f;
  const size_t total = (size_t)B * H * W;
  const float inv_nx = (float)(1.0 / ((double)B * H * (W - 1)));
  const float inv_ny = (float)(1.0 / ((double)B * (H - 1) * W));
  const float inv_n = (float)(1.0 / ((double)B * n));
  for (size_t i = (size_t)blockIdx.x * TPB + threadIdx.x; i < total; i += (size_t)gridDim.x * TPB) {
    const int b = (int)(i / ((size_t)H * W));
    const int rem = (int)(i - (size_t)b * H * W);
    const int y = rem / W, x = rem - y * W;
    const float* P = a.pred + (size_t)b * H * W;
    const float* T = a.target + (size_t)b * H * W;
    const float p = __ldg(P + rem), t = __ldg(T + rem);
    float g = 0.f;
    if (a.flags & (DP_F_SI | DP_F_SILOG)) {
      const float d = logf(p + a.eps) - logf(t + a.eps);
      double gd = 0.0;
      if ((a.flags & DP_F_SI) && a.w.w_si != 0.f) {
        const double S1 = a.mom[(size_t)b * NMOM + DP_M_S1];
        const double sc = a.si_scale ? (double)__ldg(a.si_scale + b) : 1.0;
        gd += sc * (double)a.w.w_si * (2.0 * d / n - 2.0 * S1 / (n * n)) / B;
      }
      if ((a.flags & DP_F_SILOG) && a.w.w_silog != 0.f && t > 0.f)
        gd += (double)a.w.w_silog * (2.0 * d / M0 - 2.0 * (double)a.w.vf * M1 / (M0 * M0));
      g += (float)(gd / (double)(p + a.eps));
    }
    if (STENCIL) {
      const float wg = a.w.w_grad, wb = EDGE ? a.w.beta : 0.f;
      const float* RGB = EDGE ? a.rgb + (size_t)b * 3 * H * W : nullptr;
      float acc = 0.f;
      // pair (x, x+1): this pixel is the left element
      if (x + 1 < W) {
        const float pr = __ldg(P + rem + 1), tr = __ldg(T + rem + 1);
        const float e = fabsf(p - pr) - fabsf(t - tr);
        float wgt = wg * inv_nx;
        if (EDGE) wgt += wb * inv_n * ((grad_mag(RGB, H, W, y, x) - gmin) / ginv);
        acc += wgt * sgnf(e) * sgnf(p - pr);
      }
      if (x > 0) {  // pair (x-1, x): this pixel is the right element
        const float pl = __ldg(P + rem - 1), tl = __ldg(T + rem - 1);
        const float e = fabsf(pl - p) - fabsf(tl - t);
        float wgt = wg * inv_nx;
        if (EDGE) wgt += wb * inv_n * ((grad_mag(RGB, H, W, y, x - 1) - gmin) / ginv);
        acc -= wgt * sgnf(e) * sgnf(pl - p);
      }
      if (y + 1 < H) {
        const float pdn = __ldg(P + rem + W), tdn = __ldg(T + rem + W);
        const float e = fabsf(p - pdn) - fabsf(t - tdn);
        float wgt = wg * inv_ny;
        if (EDGE) wgt += wb * inv_n * ((grad_mag(RGB, H, W, y, x) - gmin) / ginv);
        acc += wgt * sgnf(e) * sgnf(p - pdn);
      }
      if (y > 0) {
        const float pu = __ldg(P + rem - W), tu = __ldg(T + rem - W);
        const float e = fabsf(pu - p) - fabsf(tu - t);
        float wgt = wg * inv_ny;
        if (EDGE) wgt += wb * inv_n * ((grad_mag(RGB, H, W, y - 1, x) - gmin) / ginv);
        acc -= wgt * sgnf(e) * sgnf(pu - p);
      }
      g += acc;
    }
    a.grad_pred[i] = go * g;
  }
}

// ---- delta-threshold pixel counts -------------------------------------------------------------------
struct CntArgs {
  const float* pred;
  const float* target;
  const double* mom;  // S1 per sample (aligned mode)
  int B, H, W, chunks, nthr, aligned;
  float eps_div;
  float thr[DP_MAX_THR];
  unsigned long long* partials;  // [B][chunks][DP_MAX_THR]
};

template <int VEC>
__global__ void __launch_bounds__(TPB) delta_counts_kernel(CntArgs a) {
  const int b = blockIdx.y;
  const size_t n = (size_t)a.H * a.W;
  const float* __restrict__ P = a.pred + (size_t)b * n;
  const float* __restrict__ T = a.target + (size_t)b * n;
  float s = 1.f;
  if (a.aligned) s = expf((float)(-a.mom[(size_t)b * NMOM + DP_M_S1] / (double)n));  // util.py:200
  const size_t nv = n / VEC;
  const size_t per = (nv + a.chunks - 1) / a.chunks;
  const size_t i0 = blockIdx.x * per, i1 = min(nv, i0 + per);
  unsigned cnt[DP_MAX_THR];
#pragma unroll
  for (int k = 0; k < DP_MAX_THR; ++k) cnt[k] = 0;
  for (size_t i = i0 + threadIdx.x; i < i1; i += TPB) {
    float pv[VEC], tv[VEC];
    if (VEC == 4) {
      float4 p4 = ldg4(P + i * 4), t4 = ldg4(T + i * 4);
      pv[0] = p4.x; pv[1] = p4.y; pv[2] = p4.z; pv[3] = p4.w;
      tv[0] = t4.x; tv[1] = t4.y; tv[2] = t4.z; tv[3] = t4.w;
    } else {
      pv[0] = __ldg(P + i);
      tv[0] = __ldg(T + i);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float al = pv[j] * s;
      const float r1 = al / (tv[j] + a.eps_div);   // util.py:204 (eps_div = 0) / main.py:318 (1e-6)
      const float r2 = tv[j] / (al + a.eps_div);
#pragma unroll
      for (int k = 0; k < DP_MAX_THR; ++k)
        if (k < a.nthr && r1 < a.thr[k] && r2 < a.thr[k]) cnt[k]++;   // NaN / inf compare false, as torch.max + lt
    }
  }
  __shared__ unsigned sc[DP_MAX_THR][32];
#pragma unroll
  for (int k = 0; k < DP_MAX_THR; ++k) {
    unsigned v = cnt[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sc[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < DP_MAX_THR) {
    unsigned long long tot = 0;
    for (int w = 0; w < TPB / 32; ++w) tot += sc[threadIdx.x][w];
    a.partials[((size_t)b * a.chunks + blockIdx.x) * DP_MAX_THR + threadIdx.x] = tot;
  }
}

__global__ void counts_finalize_kernel(const unsigned long long* __restrict__ partials, int chunks, int nthr,
                                       unsigned long long* __restrict__ counts /*[B][nthr]*/) {
  const int b = blockIdx.x, k = threadIdx.x;
  if (k >= nthr) return;
  unsigned long long s = 0;
  for (int c = 0; c < chunks; ++c) s += partials[((size_t)b * chunks + c) * DP_MAX_THR + k];
  counts[(size_t)b * nthr + k] = s;
}

// evaluation.py:157-166 scalars for one batch from moments + counts (one block; fixed summation order)
__global__ void __launch_bounds__(256) metrics_combine_kernel(const double* __restrict__ mom,
                                                              const unsigned long long* __restrict__ counts, int B, int H,
                                                              int W, int nthr, float* __restrict__ out) {
  __shared__ double red[(2 + DP_MAX_THR) * 32];
  const double n = (double)H * W;
  double v[2 + DP_MAX_THR];
#pragma unroll
  for (int k = 0; k < 2 + DP_MAX_THR; ++k) v[k] = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double* m = mom + (size_t)b * NMOM;
    v[0] += (double)sqrtf((float)(m[DP_M_S2] / n - (m[DP_M_S1] * m[DP_M_S1]) / (n * n)));
    v[1] += m[DP_M_AR];
#pragma unroll
    for (int k = 0; k < DP_MAX_THR; ++k)
      if (k < nthr) v[2 + k] += (double)(float)((double)counts[(size_t)b * nthr + k] / n);
  }
  block_sum<2 + DP_MAX_THR>(v, red);
  if (threadIdx.x == 0) {
    out[0] = (float)(v[0] / B);
    out[1] = (float)(v[1] / ((double)B * n));
    for (int k = 0; k < nthr; ++k) out[2 + k] = (float)(v[2 + k] / B);
  }
}


// ---- fused evaluation metrics: SI-RMSE + AbsRel + aligned delta counts in ONE launch -----------------------------
// evaluation.py:157-166 calls scale_invariant_loss(sqroot=True), absolute_relative_error and three delta_thres.
// delta needs the per-sample scale exp(mean(log t - log p)) before any pixel can be classified, i.e. two sweeps over
// each sample.  A thread-block cluster of kEvalCluster CTAs owns one sample: sweep 1 accumulates the moments (fp32
// over 4 pixels, fp64 beyond), the CTAs exchange their partials through distributed shared memory, and sweep 2
// re-reads the CTA's own slice (2 MB per sample: an L2 hit) to count.  HBM traffic: 8 B/px, once.
constexpr int kEvalCluster = 8;

struct EvalArgs {
  const float* pred;
  const float* target;
  int B, nthr;
  long long n;         // pixels per sample
  float eps;
  float thr[DP_MAX_THR];
  double* moments;                 // [B][NMOM]: S1, S2, AR filled
  unsigned long long* counts;      // [B][nthr]
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double* local_ptr, uint32_t rank) {
  uint32_t sa = (uint32_t)__cvta_generic_to_shared(local_ptr), ra;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sa), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_dsmem_u64(const unsigned long long* local_ptr, uint32_t rank) {
  uint32_t sa = (uint32_t)__cvta_generic_to_shared(local_ptr), ra;
  unsigned long long v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sa), "r"(rank));
  asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(ra) : "memory");
  return v;
}

// FAST = false: IEEE logf / division, the arithmetic of the reference's torch ops (default).
// FAST = true : MUFU lg2 / rcp (2^-22-class errors: SI-RMSE / AbsRel within ~1e-6 relative of the exact path, delta counts
//               within a few pixels per million) - the variant that is bandwidth- rather than issue-bound.
template <int VEC, bool FAST, int NT>      // NT: compile-time threshold count (3 = evaluation.py's set), 0 = a.nthr at run time
__global__ void __cluster_dims__(kEvalCluster, 1, 1) __launch_bounds__(TPB) eval_fused_kernel(EvalArgs a) {
  const int b = blockIdx.x / kEvalCluster;
  const uint32_t rank = cluster_ctarank();
  const long long n = a.n;
  const float* __restrict__ P = a.pred + (size_t)b * n;
  const float* __restrict__ T = a.target + (size_t)b * n;
  const long long nv = n / VEC;
  const long long per = (nv + kEvalCluster - 1) / kEvalCluster;
  const long long i0 = (long long)rank * per, i1 = min(nv, i0 + per);
  const float eps = a.eps;

  __shared__ double s_mom[3];                       // this CTA's S1, S2, AR
  __shared__ unsigned long long s_cnt[DP_MAX_THR];  // this CTA's counts
  __shared__ double red[3 * 32];

  // ---- sweep 1: moments ----
  double acc[3] = {0.0, 0.0, 0.0};
  constexpr int U = 1;     // items per round (more in flight costs occupancy: measured slower)
  for (long long i = i0 + threadIdx.x; i < i1; i += (long long)U * TPB) {
    float pv[U][VEC], tv[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + (long long)u * TPB;
      if (iu < i1) {
        if (VEC == 4) {
          const float4 p4 = ldg4(P + iu * 4), t4 = ldg4(T + iu * 4);
          pv[u][0] = p4.x; pv[u][1] = p4.y; pv[u][2] = p4.z; pv[u][3] = p4.w;
          tv[u][0] = t4.x; tv[u][1] = t4.y; tv[u][2] = t4.z; tv[u][3] = t4.w;
        } else {
          pv[u][0] = __ldg(P + iu);
          tv[u][0] = __ldg(T + iu);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + (long long)u * TPB >= i1) continue;
      float s1 = 0.f, s2 = 0.f, ar = 0.f;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float d, q;
        if (FAST) {
          d = (__log2f(pv[u][j] + eps) - __log2f(tv[u][j] + eps)) * 0.69314718055994531f;
          q = __fdividef(fabsf(tv[u][j] - pv[u][j]), tv[u][j] + 1e-6f);
        } else {
          d = logf(pv[u][j] + eps) - logf(tv[u][j] + eps);             // util.py:143
          q = fabsf(tv[u][j] - pv[u][j]) / (tv[u][j] + 1e-6f);         // util.py:218
        }
        s1 += d;
        s2 += d * d;
        ar += q;
      }
      acc[0] += s1; acc[1] += s2; acc[2] += ar;
    }
  }
  block_sum<3>(acc, red);
  if (threadIdx.x == 0) { s_mom[0] = acc[0]; s_mom[1] = acc[1]; s_mom[2] = acc[2]; }
  cluster_sync_all();
  double S1 = 0.0;
  for (uint32_t r = 0; r < kEvalCluster; ++r) S1 += ld_dsmem_f64(&s_mom[0], r);   // fixed order: same value in every CTA
  const float s = expf((float)(-S1 / (double)n));                                  // util.py:200

  // ---- sweep 2: aligned delta counts over the same slice (L2-resident) ----
  unsigned cnt[DP_MAX_THR];
#pragma unroll
  for (int k = 0; k < DP_MAX_THR; ++k) cnt[k] = 0;
  for (long long i = i0 + threadIdx.x; i < i1; i += (long long)U * TPB) {
    float pv[U][VEC], tv[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + (long long)u * TPB;
      if (iu < i1) {
        if (VEC == 4) {
          const float4 p4 = ldg4(P + iu * 4), t4 = ldg4(T + iu * 4);
          pv[u][0] = p4.x; pv[u][1] = p4.y; pv[u][2] = p4.z; pv[u][3] = p4.w;
          tv[u][0] = t4.x; tv[u][1] = t4.y; tv[u][2] = t4.z; tv[u][3] = t4.w;
        } else {
          pv[u][0] = __ldg(P + iu);
          tv[u][0] = __ldg(T + iu);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + (long long)u * TPB >= i1) continue;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float al = pv[u][j] * s;
        float r1, r2;
        if (FAST) {   // plain reciprocal-multiply: 0/0, x/0 and inf cases still compare false below
          r1 = al * __frcp_rn(tv[u][j]);
          r2 = tv[u][j] * __frcp_rn(al);
        } else {
          r1 = al / tv[u][j];          // util.py:204: no epsilon in the divisions
          r2 = tv[u][j] / al;
        }
#pragma unroll
        for (int k = 0; k < (NT ? NT : DP_MAX_THR); ++k)
          if ((NT || k < a.nthr) && r1 < a.thr[k] && r2 < a.thr[k]) cnt[k]++;   // NaN / inf compare false, as torch.max + lt
      }
    }
  }
  __shared__ unsigned sc[DP_MAX_THR][TPB / 32];
#pragma unroll
  for (int k = 0; k < DP_MAX_THR; ++k) {
    unsigned v = cnt[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sc[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < DP_MAX_THR) {
    unsigned long long tot = 0;
    for (int w = 0; w < TPB / 32; ++w) tot += sc[threadIdx.x][w];
    s_cnt[threadIdx.x] = tot;
  }
  cluster_sync_all();
  if (rank == 0) {
    if (threadIdx.x < 3) {
      double v = 0.0;
      for (uint32_t r = 0; r < kEvalCluster; ++r) v += ld_dsmem_f64(&s_mom[threadIdx.x], r);
      const int slot = threadIdx.x == 0 ? DP_M_S1 : (threadIdx.x == 1 ? DP_M_S2 : DP_M_AR);
      a.moments[(size_t)b * NMOM + slot] = v;
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + a.nthr) {
      const int k = threadIdx.x - 32;
      unsigned long long v = 0;
      for (uint32_t r = 0; r < kEvalCluster; ++r) v += ld_dsmem_u64(&s_cnt[k], r);
      a.counts[(size_t)b * a.nthr + k] = v;
    }
  }
  cluster_sync_all();   // keep every CTA's shared memory alive until rank 0 has read it
}

// ---- streaming evaluation kernel (default path): each input byte crosses HBM once and is classified from shared memory ----
// One CTA per SM.  A group of G co-resident CTAs owns one sample at a time (the launch is cooperative, so every CTA of
// a group is resident); the groups walk the batch in parallel.  Each CTA keeps its slice of (pred, target) in one of
// kEvSlots shared-memory slots, filled by 1-D bulk copies (cp.async.bulk, completion on mbarriers).  Roles, iteration i:
//     consumers (28 warps) : sweep 1 of sample i (moments, as the chunks land)  ->  sweep 2 of sample i-1 (delta counts
//                            from shared memory, releasing each chunk as soon as it has been classified)
//     exchanger (1 warp)   : posts the CTA's moment partials of sample i to global memory, collects the S1 partials of
//                            the other CTAs of the group, adds the G values in rank order and hands the per-sample scale
//                            to the consumers - one sweep ahead of where they need it
//     producer  (1 warp)   : refills every released chunk with the sample kEvSlots iterations ahead
// DRAM traffic is the algorithmic 8 B/px (ncu: 1.344 GB for 1.342 GB of input); the second sweep touches neither HBM nor L2.
constexpr int kEvCW = 28;                   // consumer warps
constexpr int kEvCT = kEvCW * 32;           // consumer threads
constexpr int kEvThreads = kEvCT + 64;      // + producer warp + exchange warp
constexpr int kEvChunkPx = kEvCT * 4;       // pixels per chunk: one float4 of each operand per consumer thread
constexpr int kEvMaxChunks = 8;             // chunks per slot
constexpr int kEvSlots = 2;                 // shared-memory slots = samples in flight per CTA (exact / MUFU arithmetic)
// The lean arithmetic is no longer issue-bound, and with two slots the kernel is then bound by the latency of one
// sample's round trip (load -> sweep 1 -> scale exchange through global memory -> sweep 2): two slots x 86 KB in flight
// per 7 us is 55 % of an SM's HBM share.  It therefore runs with four smaller slots and classifies TWO samples behind the
// moments sweep, so the exchange has two sweeps' time to complete and the producer stays two samples ahead.
constexpr int kEvSlotsLean = 4, kEvLagLean = 2;   // measured alternatives at 448x576: 5 slots (3 groups of 49) 374, 8 slots (2 groups of 74) 213, lag 1 367 Gpx/s
constexpr int kEvCWLean = 14, kEvVPTLean = 4;   // 14 consumer warps x 4 float4 per chunk = one chunk per slice at 448x576 (measured: 28 x 1 -> 395, 14 x 2 -> 470, 14 x 4 -> 480, 7 x 4 -> 390 Gpx/s)
constexpr int kEvPrefetch = 0;                  // samples of L2 prefetch beyond the slot ring: off (measured 0 -> 398, 1 -> 404, 3 -> 390 Gpx/s, but ncu showed 17 % more DRAM reads with 1: prefetched lines evicted before use)
// static shared memory the plan leaves room for: 2.4 KB in the instantiations with a compile-time threshold count
// (1 or 3), 3.4 KB with the run-time count (count arrays sized for DP_MAX_THR)
constexpr int kEvStaticFixed = 5120, kEvStaticRuntime = 6144;
__host__ inline int eval_static_allowance(int nthr) { return (nthr == 1 || nthr == 3) ? kEvStaticFixed : kEvStaticRuntime; }

struct EvsArgs {
  const float* pred;
  const float* target;
  int B, nthr, G, ngroups, per;  // per: pixels per slice (multiple of 4)
  int pf_dist;                   // L2 prefetch distance in samples (0 = off)
  long long n;
  float eps;
  float thr[DP_MAX_THR];
  unsigned long long* ll;        // [B][G][2] flagged S1 words, zero before launch
  double* mom_part;              // [B][G][4]: S1, S2, AR
  unsigned long long* cnt_part;  // [B][G][DP_MAX_THR]
};

__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
// L2 prefetch of a contiguous range (no shared-memory slot, no completion to wait for): lets HBM stream several samples
// ahead of the slot ring, so the slot loads themselves are served from L2 (bytes in flight are bounded by the L2, not
// by the 223 KB of shared memory; DRAM traffic stays the algorithmic 8 B/px - every line is fetched once).
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(kEvCT) : "memory"); }
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The exchange warp of the streaming kernels.  Per sample: wait for the consumers' per-warp moment partials, post the CTA's
// sums, collect the S1 partials of all G CTAs of the group, add them in rank order (the same value in every CTA) and
// publish the per-sample scale exp(-S1/n) (util.py:200) to the consumers.
// S1 travels as two 8-byte words {flag = 1 : 32 bits of the double}: an aligned 8-byte store is a single transaction, so
// data and flag land together - no fence, no counter, one round trip to post and one polling read to collect.  The
// words are zeroed by the entry point before the launch.
template <bool LOG2_UNITS, int RB, int LAG, int NTS, int CW>
__device__ __forceinline__ void exchange_loop(unsigned long long* ll, double* mom_part, int G, int ngroups, int group,
                                              int rank, int nit, long long n, uint64_t* red_full, uint64_t* scale_full,
                                              const double* s_red /*[RB][3*kEvCW]*/, float* s_scale /*[RB]*/,
                                              const unsigned* s_neg /*[RB][kEvCW]*/, unsigned* s_flag /*[RB]*/,
                                              unsigned* s_cnt /*[RB][NTS]*/, uint64_t* done_bar,
                                              unsigned long long* cnt_part, int nthr) {
  // The warp is software-pipelined: iteration `it` folds and POSTS this CTA's partials of sample it, then COLLECTS the
  // group's partials of sample it - 1, which every CTA posted one sample-time ago - so the poll normally succeeds on its
  // first read and the warp's time per sample is one global round trip, not two.  The warp's per-sample time bounds the
  // whole kernel (throughput = groups x pixels per sample / exchange time), so the folds are butterflies over lanes
  // (fixed order: the same bits in every CTA of the group), not serial chains on one lane.
  const int lane = threadIdx.x & 31;
  auto collect = [&](int it) {
    const int b = group + it * ngroups;
    const int rb = it % RB;
    double S1 = 0.0;
    const unsigned long long* gw = ll + (size_t)b * G * 2;
    for (int g0 = 0; g0 < G; g0 += 32) {
      const int g = g0 + lane;
      unsigned long long w0 = 1ull << 32, w1 = 1ull << 32;
      unsigned spins = 0;
      long long t0 = 0;
      for (;;) {
        if (g < G)
          asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(gw + (size_t)g * 2) : "memory");
        const bool ok = (w0 >> 32) == 1ull && (w1 >> 32) == 1ull;
        if (__all_sync(0xffffffffu, ok)) break;
        if ((++spins & 255u) == 0) {        // bounded: a protocol bug traps instead of hanging
          const long long now = clock64();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 4000000000LL) __trap();
        }
      }
      const double v = g < G ? __hiloint2double((int)(unsigned)w0, (int)(unsigned)w1) : 0.0;
      S1 += warp_sum(v);                    // xor butterfly: every lane of every CTA adds the same pairs in the same order
    }
    if (lane == 0) {
      s_scale[rb] = expf((float)(-S1 / (double)n));
      tc::mbar_arrive(&scale_full[rb]);
    }
    __syncwarp();
  };
  // Per-sample delta counts: every consumer warp adds its count to s_cnt[sample % RB] (shared-memory atomics, no CTA
  // barrier).  A warp starts sweep 1 of sample `it` only after its sweep 2 of sample it - 1 - LAG, so once red_full[it]
  // has completed the counts of that older sample are final: they are flushed here, and the ring slot is zeroed before
  // scale_full of the sample that will reuse it is signalled.
  int flushed = 0;
  auto flush = [&](int j) {
    const int b = group + j * ngroups;
    if (lane < NTS) {
      const unsigned v = s_cnt[(j % RB) * NTS + lane];
      s_cnt[(j % RB) * NTS + lane] = 0u;
      if (lane < nthr) cnt_part[((size_t)b * G + rank) * DP_MAX_THR + lane] = v;
    }
    __syncwarp();
  };
  for (int it = 0; it < nit; ++it) {
    const int b = group + it * ngroups;
    const int rb = it % RB, rph = (it / RB) & 1;
    tc::mbar_wait(&red_full[rb], rph);
    while (flushed <= it - 1 - LAG) flush(flushed++);
    {
      const unsigned nb = __reduce_or_sync(0xffffffffu, lane < CW ? s_neg[rb * CW + lane] : 0u);
      if (lane == 0) s_flag[rb] = nb >> 31;      // a negative operand somewhere in this CTA's slice of the sample
    }
    const double* r = s_red + rb * 3 * CW;
    double m0 = lane < CW ? r[lane] : 0.0, m1 = lane < CW ? r[CW + lane] : 0.0, m2 = lane < CW ? r[2 * CW + lane] : 0.0;
    m0 = warp_sum(m0); m1 = warp_sum(m1); m2 = warp_sum(m2);
    if (lane == 0) {
      if (LOG2_UNITS) { m0 *= 0.6931471805599453; m1 *= 0.6931471805599453 * 0.6931471805599453; }
      unsigned long long* w = ll + ((size_t)b * G + rank) * 2;
      const unsigned long long w0 = (1ull << 32) | (unsigned)__double2hiint(m0);
      const unsigned long long w1 = (1ull << 32) | (unsigned)__double2loint(m0);
      asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"(w0), "l"(w1) : "memory");
      double* mp = mom_part + ((size_t)b * G + rank) * 4;      // for eval_gather_kernel
      __stcg(mp + 0, m0); __stcg(mp + 1, m1); __stcg(mp + 2, m2);
    }
    __syncwarp();
    // classification lags two samples (RB = 3): collect one sample late, off the critical path.  With a lag of one the
    // consumers need this sample's scale as soon as their next sweep 1 ends: collect at once.
    if (RB > 2) { if (it >= 1) collect(it - 1); }
    else collect(it);
  }
  if (RB > 2 && nit >= 1) collect(nit - 1);
  tc::mbar_wait(done_bar, 0);                    // every consumer warp has finished its last sweep 2
  while (flushed < nit) flush(flushed++);
}

// util.py:204-205 for one pixel, both quotients as written there
template <int NTS, bool RT>
__device__ __forceinline__ void delta_px_generic(float al, float t, const float* thr, int nthr, unsigned (&cnt)[NTS]) {
  const float r1 = __fdiv_rn(al, t), r2 = __fdiv_rn(t, al);
#pragma unroll
  for (int k = 0; k < NTS; ++k)
    if ((!RT || k < nthr) && r1 < thr[k] && r2 < thr[k]) cnt[k]++;
}

// FAST   : MUFU lg2 / rcp instead of IEEE logf / division (as eval_fused_kernel).
// ONEDIV : every threshold is > 1, so of the two quotients al/t and t/al only the one with the larger numerator can
//          reach a threshold (the other is <= 1 by monotonicity of rounding): one division per pixel, same counts.
//          Needs both operands non-negative or of mixed sign; a pixel group with a negative pair takes the generic path.
// LEAN   : the default arithmetic.  Algebraically the same quantities with the transcendental work halved:
//            sweep 1  r = rcp(t + eps);  d = lg2((p + eps) * r)  [= lg2(p + eps) - lg2(t + eps)],  |t - p| * r  (AbsRel shares
//                     the reciprocal: both epsilons are 1e-6), ~10 instructions per pixel;
//            sweep 2  with hi = max(p*s, t), lo = min(p*s, t):  max(a/t, t/a) < thr  <=>  hi < thr * lo  for non-negative
//                     operands (thr <= 1 gives false on both sides; 0/0, x/0 give false on both sides): no division at
//                     all, ~12 instructions per pixel.
//            A CTA whose slice holds a negative value (sign bits OR-ed during sweep 1) or whose sample scale is not finite
//            classifies that slice with the exact two-quotient code, so the semantics of util.py:204-205 hold for any
//            input; what differs from the exact arithmetic is rounding only (measured: SI-RMSE / AbsRel ~1e-7 relative,
//            counts a few pixels per million - contract: 1e-5 / 0.01 % of pixels).
// CW consumer warps, each thread taking VPT float4 of each operand per chunk (CW * VPT = 28: the chunk stays 3584 pixels).
// The per-sample bookkeeping (barrier waits, cross-lane folds, count atomics) is paid per THREAD, so the lean arithmetic -
// whose pixel work is down to ~36 instructions - runs 14 fatter warps (two float4 per chunk) instead of 28.
template <bool FAST, bool ONEDIV, int NT, bool LEAN = false, int NS_ = kEvSlots, int LAG_ = 1, int CW = kEvCW, int VPT = 1>
__global__ void __launch_bounds__((CW + 2) * 32, 1) eval_stream_kernel(EvsArgs a) {
  constexpr int CT = CW * 32;                      // consumer threads
  constexpr int NS = NS_;
  constexpr int LAG = LAG_;                        // sweep 2 runs LAG samples behind sweep 1
  constexpr int RB = LAG + 1;                      // ring of per-sample reduction / scale buffers
  constexpr int NTS = NT ? NT : DP_MAX_THR;
  constexpr int CH = CW * VPT * 128;             // pixels per chunk: VPT float4 of each operand per consumer thread
  extern __shared__ __align__(128) unsigned char ev_smem[];
  __shared__ uint64_t full[NS][kEvMaxChunks], empty[NS][kEvMaxChunks];
  __shared__ uint64_t red_full[RB], scale_full[RB];
  __shared__ double s_red[RB][3 * CW];
  __shared__ float s_scale[RB];
  __shared__ unsigned s_neg[RB][CW];        // LEAN: OR of the operands' bit patterns per consumer warp (sign bit = negative seen)
  __shared__ unsigned s_flag[RB];              // ... folded by the exchange warp: 1 = the slice holds a negative operand
  __shared__ unsigned s_cnt[RB][NTS];          // per-sample delta counts of this CTA (atomics from the consumer warps)
  __shared__ uint64_t done_bar;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int group = blockIdx.x / a.G, rank = blockIdx.x - group * a.G;
  const long long n = a.n;
  const long long px0 = (long long)rank * a.per;
  const long long left = n - px0;
  const int len = left <= 0 ? 0 : (left > a.per ? a.per : (int)left);   // multiple of 4 (n and per are)
  const int nch = (len + CH - 1) / CH;
  const int nit = group < a.B ? (a.B - group + a.ngroups - 1) / a.ngroups : 0;   // samples of this group
  const size_t slot_floats = (size_t)a.per * 2;                                  // pred slice, then target slice

  if (tid == 0) {
    for (int s = 0; s < NS; ++s)
      for (int c = 0; c < kEvMaxChunks; ++c) { tc::mbar_init(&full[s][c], 1); tc::mbar_init(&empty[s][c], CW); }
    for (int i = 0; i < RB; ++i) { tc::mbar_init(&red_full[i], CW); tc::mbar_init(&scale_full[i], 1); }
    tc::mbar_init(&done_bar, CW);
    tc::fence_barrier_init();
  }
  if (tid < RB * NTS) (&s_cnt[0][0])[tid] = 0u;
  if (tid < RB * CW) (&s_neg[0][0])[tid] = 0u;
  __syncthreads();

  if (warp == CW) {   // ---- producer warp: one lane streams the slices in ----
    if (lane == 0) {
      int slot = 0, ph = 0;
      const int pfd = a.pf_dist;                   // samples of L2 prefetch distance beyond the slot ring
      if (pfd > 0 && len > 0)
        for (int k = 0; k < NS + pfd && k < nit; ++k) {
          const int bk = group + k * a.ngroups;
          bulk_prefetch_l2(a.pred + (size_t)bk * n + px0, (uint32_t)len * 4u);
          bulk_prefetch_l2(a.target + (size_t)bk * n + px0, (uint32_t)len * 4u);
        }
      for (int it = 0; it < nit; ++it) {
        const int b = group + it * a.ngroups;
        const float* P = a.pred + (size_t)b * n + px0;
        const float* T = a.target + (size_t)b * n + px0;
        if (pfd > 0 && len > 0 && it + NS + pfd < nit) {
          const int bk = group + (it + NS + pfd) * a.ngroups;
          bulk_prefetch_l2(a.pred + (size_t)bk * n + px0, (uint32_t)len * 4u);
          bulk_prefetch_l2(a.target + (size_t)bk * n + px0, (uint32_t)len * 4u);
        }
        float* Ps = reinterpret_cast<float*>(ev_smem) + slot * slot_floats;
        float* Ts = Ps + a.per;
        for (int c = 0; c < nch; ++c) {
          if (it >= NS) tc::mbar_wait(&empty[slot][c], ph ^ 1);
          const int cl = min(CH, len - c * CH);
          tc::mbar_expect_tx(&full[slot][c], (uint32_t)cl * 8u);
          bulk_g2s(Ps + c * CH, P + c * CH, (uint32_t)cl * 4u, &full[slot][c]);
          bulk_g2s(Ts + c * CH, T + c * CH, (uint32_t)cl * 4u, &full[slot][c]);
        }
        if (++slot == NS) { slot = 0; ph ^= 1; }
      }
    }
    return;
  }

  if (warp == CW + 1) {   // ---- exchange warp: per-sample scale through global memory, off the consumers' path ----
    exchange_loop<FAST || LEAN, RB, LAG, NTS, CW>(a.ll, a.mom_part, a.G, a.ngroups, group, rank, nit, n, red_full, scale_full,
                                              &s_red[0][0], s_scale, &s_neg[0][0], s_flag, &s_cnt[0][0], &done_bar, a.cnt_part,
                                              NT ? NT : a.nthr);
    return;
  }

  // ---- consumer warps ----
  const int len4 = len >> 2;
  const float eps = a.eps;
  int slot1 = 0, ph1 = 0, slot2 = 0;
  for (int it = 0; it < nit + LAG; ++it) {
    if (it < nit) {
      // sweep 1 of sample `it`: moments (fp32 over 4 pixels, fp64 beyond) as the slice lands
      const float4* P4 = reinterpret_cast<const float4*>(reinterpret_cast<float*>(ev_smem) + slot1 * slot_floats) + tid;
      const float4* T4 = reinterpret_cast<const float4*>(reinterpret_cast<float*>(ev_smem) + slot1 * slot_floats + a.per) + tid;
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
      float f0 = 0.f, f1 = 0.f, f2 = 0.f;          // LEAN: fp32 over this thread's <= 32 pixels, fp64 across lanes / warps / CTAs
      unsigned negbits = 0;
      for (int c = 0; c < nch; ++c) {
        tc::mbar_wait(&full[slot1][c], ph1);
#pragma unroll
        for (int v = 0; v < VPT; ++v)
        if ((c * VPT + v) * CT + tid < len4) {   // VPT float4 of each operand per thread per chunk
          const float4 p4 = P4[(c * VPT + v) * CT], t4 = T4[(c * VPT + v) * CT];
          const float pv[4] = {p4.x, p4.y, p4.z, p4.w}, tv[4] = {t4.x, t4.y, t4.z, t4.w};
          float s1 = 0.f, s2 = 0.f, ar = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float d, e;
            if (LEAN) {   // one reciprocal serves the log ratio and AbsRel; d in log2 units
              const float r = rcp_approx(tv[q] + eps);
              d = lg2_approx((pv[q] + eps) * r);
              e = fabsf(tv[q] - pv[q]) * r;
              negbits |= __float_as_uint(pv[q]) | __float_as_uint(tv[q]);
            } else if (FAST) {   // d in log2 units; scaled by ln 2 once per CTA
              d = lg2_approx(pv[q] + eps) - lg2_approx(tv[q] + eps);
              e = fabsf(tv[q] - pv[q]) * rcp_approx(tv[q] + 1e-6f);
            } else {
              d = logf(pv[q] + eps) - logf(tv[q] + eps);                 // util.py:143
              e = __fdiv_rn(fabsf(tv[q] - pv[q]), tv[q] + 1e-6f);        // util.py:218
            }
            s1 += d;
            s2 += d * d;
            ar += e;
          }
          if (LEAN) { f0 += s1; f1 += s2; f2 += ar; }
          else { acc0 += s1; acc1 += s2; acc2 += ar; }
        }
      }
      if (LEAN) { acc0 = f0; acc1 = f1; acc2 = f2; }
      acc0 = warp_sum(acc0); acc1 = warp_sum(acc1); acc2 = warp_sum(acc2);
      if (LEAN) negbits = __reduce_or_sync(0xffffffffu, negbits);
      if (lane == 0) {
        double* r = s_red[it % RB];
        r[warp] = acc0; r[CW + warp] = acc1; r[2 * CW + warp] = acc2;
        // published with the moments: the consumers read it after scale_full, which the exchange warp signals only
        // after every warp's arrival here (release / acquire chain through the two mbarriers)
        if (LEAN) s_neg[it % RB][warp] = negbits;
        tc::mbar_arrive(&red_full[it % RB]);
      }
      if (++slot1 == NS) { slot1 = 0; ph1 ^= 1; }
    }
    if (it >= LAG) {
      // sweep 2 of sample `it - LAG`: scale-aligned delta counts from shared memory; finished chunks go back to the producer
      const int j = it - LAG;
      const int b = group + j * a.ngroups;
      tc::mbar_wait(&scale_full[j % RB], (j / RB) & 1);
      const float s = s_scale[j % RB];
      const float4* P4 = reinterpret_cast<const float4*>(reinterpret_cast<float*>(ev_smem) + slot2 * slot_floats) + tid;
      const float4* T4 = reinterpret_cast<const float4*>(reinterpret_cast<float*>(ev_smem) + slot2 * slot_floats + a.per) + tid;
      unsigned cnt[NTS];
#pragma unroll
      for (int k = 0; k < NTS; ++k) cnt[k] = 0;
      // LEAN: a negative operand in this CTA's slice, or a NaN / inf scale -> exact two-quotient classification
      const bool lean_ok = LEAN && s_flag[j % RB] == 0u && (s - s == 0.f);
      for (int c = 0; c < nch; ++c) {
#pragma unroll
        for (int v = 0; v < VPT; ++v)
        if ((c * VPT + v) * CT + tid < len4) {
          const float4 p4 = P4[(c * VPT + v) * CT], t4 = T4[(c * VPT + v) * CT];
          const float al[4] = {p4.x * s, p4.y * s, p4.z * s, p4.w * s}, tv[4] = {t4.x, t4.y, t4.z, t4.w};
          bool generic = !ONEDIV;
          if (LEAN && lean_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              // hi < thr * lo  <=>  fma(-thr, lo, hi) < 0: one FFMA, and the count is the sign bit (a NaN from
              // inf - inf is the canonical positive NaN: counted as "not below", like the reference's comparison)
              const float hi = fmaxf(al[q], tv[q]), lo = fminf(al[q], tv[q]);
#pragma unroll
              for (int k = 0; k < NTS; ++k)
                if (NT || k < a.nthr) cnt[k] += __float_as_uint(fmaf(-a.thr[k], lo, hi)) >> 31;
            }
            generic = false;
          } else if (LEAN) {
            generic = true;
          } else if (ONEDIV) {
            const float mx = fminf(fminf(fmaxf(al[0], tv[0]), fmaxf(al[1], tv[1])),
                                   fminf(fmaxf(al[2], tv[2]), fmaxf(al[3], tv[3])));
            generic = mx < 0.f;   // some pair is negative in both operands
          }
          if (generic) {
#pragma unroll
            for (int q = 0; q < 4; ++q) delta_px_generic<NTS, NT == 0>(al[q], tv[q], a.thr, a.nthr, cnt);
          } else if (!LEAN) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const bool sel = al[q] > tv[q];
              const float num = sel ? al[q] : tv[q], den = sel ? tv[q] : al[q];
              const float r = FAST ? num * rcp_approx(den) : __fdiv_rn(num, den);
#pragma unroll
              for (int k = 0; k < NTS; ++k)
                if ((NT || k < a.nthr) && r < a.thr[k]) cnt[k]++;    // NaN / inf compare false, as torch.max + lt
            }
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[slot2][c]);
      }
#pragma unroll
      for (int k = 0; k < NTS; ++k) {
        const unsigned v = __reduce_add_sync(0xffffffffu, cnt[k]);
        if (lane == 0 && (NT || k < a.nthr)) atomicAdd(&s_cnt[j % RB][k], v);    // flushed by the exchange warp
      }
      (void)b;
      if (++slot2 == NS) slot2 = 0;
    }
  }
  __syncwarp();
  if (lane == 0) tc::mbar_arrive(&done_bar);     // this warp's counts are all in s_cnt
}

// partials of the streaming kernels -> per-sample moments / counts and the per-sample terms of evaluation.py:157-166.
// One warp per sample: lanes take the CTAs of the group (all partial loads of a sample in flight at once), a butterfly
// in a fixed pattern adds them.
__global__ void __launch_bounds__(1024) eval_gather_kernel(const double* __restrict__ mom_part,
                                                           const unsigned long long* __restrict__ cnt_part, int B, int G,
                                                           double n, int nthr, double* __restrict__ moments,
                                                           unsigned long long* __restrict__ counts,
                                                           double* __restrict__ terms /*[B][2 + DP_MAX_THR]*/) {
  dp::pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
  unsigned long long c[DP_MAX_THR];
#pragma unroll
  for (int k = 0; k < DP_MAX_THR; ++k) c[k] = 0;
  for (int g = lane; g < G; g += 32) {
    const double* mp = mom_part + ((size_t)b * G + g) * 4;
    m0 += mp[0]; m1 += mp[1]; m2 += mp[2];
#pragma unroll
    for (int k = 0; k < DP_MAX_THR; ++k)
      if (k < nthr) c[k] += cnt_part[((size_t)b * G + g) * DP_MAX_THR + k];
  }
  m0 = warp_sum(m0); m1 = warp_sum(m1); m2 = warp_sum(m2);
#pragma unroll
  for (int k = 0; k < DP_MAX_THR; ++k) {
    if (k < nthr) {
      c[k] = __reduce_add_sync(0xffffffffu, (unsigned)c[k]);   // a sample has < 2^32 pixels
    }
  }
  if (lane == 0) {
    moments[(size_t)b * NMOM + DP_M_S1] = m0;
    moments[(size_t)b * NMOM + DP_M_S2] = m1;
    moments[(size_t)b * NMOM + DP_M_AR] = m2;
    double* t = terms + (size_t)b * (2 + DP_MAX_THR);
    t[0] = (double)sqrtf((float)(m1 / n - (m0 * m0) / (n * n)));   // util.py:152-154 (sqroot=True)
    t[1] = m2;
    for (int k = 0; k < nthr; ++k) {
      counts[(size_t)b * nthr + k] = c[k];
      t[2 + k] = (double)(float)((double)c[k] / n);                // util.py:205 per-sample mean
    }
  }
}

// per-sample terms -> the batch means evaluation.py:157-166 reports (one block, fixed order)
__global__ void __launch_bounds__(256) eval_means_kernel(const double* __restrict__ terms, int B, double n, int nthr,
                                                         float* __restrict__ out) {
  dp::pdl_prologue();
  __shared__ double red[(2 + DP_MAX_THR) * 32];
  double v[2 + DP_MAX_THR];
#pragma unroll
  for (int k = 0; k < 2 + DP_MAX_THR; ++k) v[k] = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double* t = terms + (size_t)b * (2 + DP_MAX_THR);
#pragma unroll
    for (int k = 0; k < 2 + DP_MAX_THR; ++k)
      if (k < 2 + nthr) v[k] += t[k];
  }
  block_sum<2 + DP_MAX_THR>(v, red);
  if (threadIdx.x == 0) {
    out[0] = (float)(v[0] / B);
    out[1] = (float)(v[1] / ((double)B * n));
    for (int k = 0; k < nthr; ++k) out[2 + k] = (float)(v[2 + k] / B);
  }
}

// util.py:159-181 per_pixel_scale_invariant_loss: out = (d - mean d)^2 with d = log p - log t (no epsilon) per image;
// the mean comes from the moments pass (S1 of dp_depth_moments run with eps = 0).
__global__ void __launch_bounds__(TPB) per_pixel_si_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                           const double* __restrict__ mom, long long n, int B,
                                                           float* __restrict__ out) {
  const long long total = n * B;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < total; i += (long long)gridDim.x * TPB) {
    const int b = (int)(i / n);
    const float mean = (float)(mom[(size_t)b * NMOM + DP_M_S1] / (double)n);
    const float d = logf(__ldg(pred + i)) - logf(__ldg(target + i));
    out[i] = (d - mean) * (d - mean);
  }
}

struct EvalPlan {
  bool ok;
  int G, ngroups, per;
  size_t dyn;
};

// Pure function of the shape and of the device's shared memory per SM / SM count, shared by the workspace query and
// the launch: one CTA per SM, kEvSlots slices resident per CTA, and the CTAs-per-sample count G that wastes the fewest
// thread slots (slices that are whole chunks) and SMs (groups * G close to the SM count).
inline EvalPlan eval_plan_for(long long n, int B, int smem_sm, int sms, int nthr, int slots = kEvSlots,
                              int chunk_px = kEvChunkPx) {
  EvalPlan p{};
  if (n % 4 != 0 || n <= 0 || B <= 0 || sms <= 0) return p;
  const long long budget = ((long long)smem_sm - 1024 - eval_static_allowance(nthr)) / slots;   // bytes per slot
  if (budget < 4096) return p;
  const long long gmin = (n * 8 + budget - 1) / budget;
  double best = -1.0;
  for (long long G = gmin; G <= sms && G < gmin + 32; ++G) {
    long long per = (n + G - 1) / G;
    per = (per + 3) & ~3LL;
    const long long nch = (per + chunk_px - 1) / chunk_px;
    if (per * 8 > budget || nch > kEvMaxChunks) continue;
    long long groups = sms / G;
    if (groups > B) groups = B;
    const double eff = (double)n / (double)(nch * chunk_px * G) * (double)(groups * G) / (double)sms;
    if (eff > best + 1e-9) {
      best = eff;
      p.G = (int)G; p.per = (int)per; p.ngroups = (int)groups;
    }
  }
  if (best < 0.0) return p;
  p.dyn = (size_t)p.per * 8 * slots;
  p.ok = true;
  return p;
}

inline EvalPlan eval_plan(long long n, int B, int nthr, int slots = kEvSlots, int chunk_px = kEvChunkPx) {
  int smem_sm = 0, sms = 0, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return EvalPlan{};
  if (cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) != cudaSuccess) return EvalPlan{};
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return EvalPlan{};
  return eval_plan_for(n, B, smem_sm, sms, nthr, slots, chunk_px);
}

inline size_t eval_ws_ll_bytes(int B, int G) { return (((size_t)B * G * 2 * sizeof(unsigned long long)) + 255) & ~(size_t)255; }

inline int pick_chunks(int B, int H) {
  int c = (4 * kNumSMs + B - 1) / B;
  if (c < 1) c = 1;
  if (c > H) c = H;
  return c;
}
constexpr int kMinMaxBlocks = 4 * kNumSMs;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

size_t dp_depth_moments_workspace(int B, int H, int W) {
  (void)W;
  size_t mom = (size_t)B * pick_chunks(B, H) * NMOM * sizeof(double);
  size_t cnt = (size_t)B * pick_chunks(B, H) * DP_MAX_THR * sizeof(unsigned long long);
  return (mom > cnt ? mom : cnt) + 256;
}

size_t dp_rgb_minmax_bytes(void) { return (size_t)kMinMaxBlocks * 2 * sizeof(float); }

int dp_rgb_gradmag_minmax(const float* rgb, int B, int H, int W, float* minmax_partials, cudaStream_t stream) {
  DP_CHECK_ARG(rgb && minmax_partials && B > 0 && H > 0 && W > 0, "dp_rgb_gradmag_minmax: bad arguments");
  rgb_minmax_kernel<<<kMinMaxBlocks, TPB, 0, stream>>>(rgb, B, H, W, minmax_partials);
  DP_CHECK_LAUNCH("rgb_minmax_kernel");
  return DP_OK;
}

int dp_depth_moments(const float* pred, const float* target, const float* rgb, const float* minmax_partials,
                     int B, int H, int W, unsigned flags, float eps, double* moments, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream) {
  DP_CHECK_ARG(pred && target && moments && workspace, "dp_depth_moments: null pointer");
  DP_CHECK_ARG(B > 0 && H > 0 && W > 0, "dp_depth_moments: bad shape %d %d %d", B, H, W);
  if (flags & DP_F_EDGE) DP_CHECK_ARG(rgb && minmax_partials, "dp_depth_moments: edge term needs rgb + minmax partials");
  if (workspace_bytes < dp_depth_moments_workspace(B, H, W))
    return dp_set_error(DP_ERR_WORKSPACE, "dp_depth_moments: workspace %zu < %zu", workspace_bytes,
                        dp_depth_moments_workspace(B, H, W));
  MomArgs a;
  a.pred = pred; a.target = target; a.rgb = rgb; a.B = B; a.H = H; a.W = W;
  a.chunks = pick_chunks(B, H);
  a.flags = flags; a.eps = eps; a.mm_partials = minmax_partials; a.mm_n = kMinMaxBlocks;
  a.partials = reinterpret_cast<double*>(workspace);
  const bool vec = (W % 4 == 0) && aligned16(pred) && aligned16(target);
  const bool edge = flags & DP_F_EDGE;
  const bool stencil = edge || (flags & DP_F_GRAD);
  dim3 grid(a.chunks, B);
  if (vec) {
    if (edge) moments_kernel<4, true, true><<<grid, TPB, 0, stream>>>(a);
    else if (stencil) moments_kernel<4, true, false><<<grid, TPB, 0, stream>>>(a);
    else moments_kernel<4, false, false><<<grid, TPB, 0, stream>>>(a);
  } else {
    if (edge) moments_kernel<1, true, true><<<grid, TPB, 0, stream>>>(a);
    else if (stencil) moments_kernel<1, true, false><<<grid, TPB, 0, stream>>>(a);
    else moments_kernel<1, false, false><<<grid, TPB, 0, stream>>>(a);
  }
  DP_CHECK_LAUNCH("moments_kernel");
  moments_finalize_kernel<<<B, 32, 0, stream>>>(a.partials, a.chunks, moments);
  DP_CHECK_LAUNCH("moments_finalize_kernel");
  return DP_OK;
}

int dp_loss_combine(const double* moments, int B, int H, int W, unsigned flags, float w_si, float w_silog,
                    float variance_focus, float w_grad, float beta, int sqroot, float* out, float* per_sample,
                    cudaStream_t stream) {
  DP_CHECK_ARG(moments && out && B > 0, "dp_loss_combine: bad arguments");
  LossW w{w_si, w_silog, variance_focus, w_grad, beta, sqroot};
  loss_combine_kernel<<<1, 32, 0, stream>>>(moments, B, H, W, w, flags, out, per_sample);
  DP_CHECK_LAUNCH("loss_combine_kernel");
  return DP_OK;
}

int dp_loss_backward(const float* pred, const float* target, const float* rgb, const float* minmax_partials,
                     const double* moments, const float* grad_out, const float* si_sample_scale, int B, int H, int W,
                     unsigned flags, float eps, float w_si, float w_silog, float variance_focus, float w_grad, float beta, float* grad_pred,
                     cudaStream_t stream) {
  DP_CHECK_ARG(pred && target && moments && grad_pred, "dp_loss_backward: null pointer");
  BwdArgs a;
  a.pred = pred; a.target = target; a.rgb = rgb; a.mom = moments; a.mm_partials = minmax_partials;
  a.mm_n = kMinMaxBlocks; a.grad_out = grad_out; a.si_scale = si_sample_scale; a.grad_pred = grad_pred; a.B = B; a.H = H; a.W = W;
  a.flags = flags; a.eps = eps;
  a.w = LossW{w_si, w_silog, variance_focus, w_grad, beta, 0};
  const bool edge = (flags & DP_F_EDGE) && beta != 0.f;
  const bool stencil = edge || ((flags & DP_F_GRAD) && w_grad != 0.f);
  if (edge) DP_CHECK_ARG(rgb && minmax_partials, "dp_loss_backward: edge term needs rgb + minmax partials");
  if (!(flags & DP_F_GRAD)) a.w.w_grad = 0.f;
  const size_t total = (size_t)B * H * W;
  int blocks = (int)((total + TPB - 1) / TPB);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (edge) loss_bwd_kernel<true, true><<<blocks, TPB, 0, stream>>>(a);
  else if (stencil) loss_bwd_kernel<true, false><<<blocks, TPB, 0, stream>>>(a);
  else loss_bwd_kernel<false, false><<<blocks, TPB, 0, stream>>>(a);
  DP_CHECK_LAUNCH("loss_bwd_kernel");
  return DP_OK;
}

int dp_delta_counts(const float* pred, const float* target, const double* moments, int B, int H, int W,
                    const float* thresholds, int nthr, int aligned, float eps_div, unsigned long long* counts,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DP_CHECK_ARG(pred && target && counts && workspace && thresholds, "dp_delta_counts: null pointer");
  DP_CHECK_ARG(nthr >= 1 && nthr <= DP_MAX_THR, "dp_delta_counts: nthr %d out of [1,%d]", nthr, DP_MAX_THR);
  DP_CHECK_ARG(!aligned || moments, "dp_delta_counts: aligned mode needs moments");
  if (workspace_bytes < dp_depth_moments_workspace(B, H, W))
    return dp_set_error(DP_ERR_WORKSPACE, "dp_delta_counts: workspace too small");
  CntArgs a;
  a.pred = pred; a.target = target; a.mom = moments; a.B = B; a.H = H; a.W = W;
  a.chunks = pick_chunks(B, H);
  a.nthr = nthr; a.aligned = aligned; a.eps_div = eps_div;
  for (int k = 0; k < DP_MAX_THR; ++k) a.thr[k] = k < nthr ? thresholds[k] : 0.f;
  a.partials = reinterpret_cast<unsigned long long*>(workspace);
  dim3 grid(a.chunks, B);
  const bool vec = (((size_t)H * W) % 4 == 0) && aligned16(pred) && aligned16(target);
  if (vec) delta_counts_kernel<4><<<grid, TPB, 0, stream>>>(a);
  else delta_counts_kernel<1><<<grid, TPB, 0, stream>>>(a);
  DP_CHECK_LAUNCH("delta_counts_kernel");
  counts_finalize_kernel<<<B, 32, 0, stream>>>(a.partials, a.chunks, nthr, counts);
  DP_CHECK_LAUNCH("counts_finalize_kernel");
  return DP_OK;
}

int dp_per_pixel_si(const float* pred, const float* target, const double* moments, int B, int H, int W, float* out,
                    cudaStream_t stream) {
  DP_CHECK_ARG(pred && target && moments && out && B > 0 && H > 0 && W > 0, "dp_per_pixel_si: bad arguments");
  const long long n = (long long)H * W;
  long long blocks = (n * B + TPB - 1) / TPB;
  if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
  per_pixel_si_kernel<<<(int)blocks, TPB, 0, stream>>>(pred, target, moments, n, B, out);
  DP_CHECK_LAUNCH("per_pixel_si_kernel");
  return DP_OK;
}

int dp_metrics_combine(const double* moments, const unsigned long long* counts, int B, int H, int W, int nthr,
                       float* out, cudaStream_t stream) {
  DP_CHECK_ARG(moments && counts && out, "dp_metrics_combine: null pointer");
  metrics_combine_kernel<<<1, 256, 0, stream>>>(moments, counts, B, H, W, nthr, out);
  DP_CHECK_LAUNCH("metrics_combine_kernel");
  return DP_OK;
}


static size_t eval_ws_bytes(int B, const EvalPlan& p) {
  return eval_ws_ll_bytes(B, p.G) + (size_t)B * p.G * 4 * sizeof(double) +
         (size_t)B * p.G * DP_MAX_THR * sizeof(unsigned long long) + (size_t)B * (2 + DP_MAX_THR) * sizeof(double);
}

int dp_eval_metrics_plan(long long pixels, int B, int nthr, int smem_per_sm, int sms, int* ctas_per_sample, int* groups,
                         int* slice_pixels, size_t* dynamic_smem) {
  const EvalPlan p = eval_plan_for(pixels, B, smem_per_sm, sms, nthr);
  if (!p.ok) return 0;
  if (ctas_per_sample) *ctas_per_sample = p.G;
  if (groups) *groups = p.ngroups;
  if (slice_pixels) *slice_pixels = p.per;
  if (dynamic_smem) *dynamic_smem = p.dyn;
  return 1;
}

size_t dp_eval_metrics_workspace(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 256;
  size_t need = 256;
  for (int nthr = 1; nthr <= 2; ++nthr)        // the two decompositions (compile-time / run-time threshold count) ...
    for (int cfg : {0, 1}) {                       // ... of the exact and the lean plan
      const EvalPlan p = eval_plan((long long)H * W, B, nthr, cfg ? kEvSlotsLean : kEvSlots, cfg ? kEvCWLean * kEvVPTLean * 128 : kEvChunkPx);
      if (p.ok && eval_ws_bytes(B, p) > need) need = eval_ws_bytes(B, p);
    }
  return need;
}

/* evaluation.py:157-166: a streaming kernel (8 B/px of HBM traffic, pixels classified from shared memory) and the
 * parallel combine; shapes it cannot take (pixel count not a multiple of 4, unaligned bases, slices larger than the
 * chip's shared memory) go through the cluster kernel.  out[0]=SI-RMSE, out[1]=AbsRel, out[2+k]=delta_k (batch means). */
int dp_eval_metrics(const float* pred, const float* target, int B, int H, int W, const float* thresholds, int nthr,
                    float eps, int fast_math, double* moments, unsigned long long* counts, float* out,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DP_CHECK_ARG(pred && target && thresholds && moments && counts && out, "dp_eval_metrics: null pointer");
  DP_CHECK_ARG(B > 0 && H > 0 && W > 0, "dp_eval_metrics: bad shape %d %d %d", B, H, W);
  DP_CHECK_ARG(nthr >= 1 && nthr <= DP_MAX_THR, "dp_eval_metrics: nthr %d out of [1,%d]", nthr, DP_MAX_THR);
  const long long n = (long long)H * W;
  const bool vec = (n % 4 == 0) && aligned16(pred) && aligned16(target);

  const bool lean = fast_math == 2 && eps == 1e-6f;
  EvalPlan plan = vec ? eval_plan(n, B, nthr, lean ? kEvSlotsLean : kEvSlots, lean ? kEvCWLean * kEvVPTLean * 128 : kEvChunkPx)
                      : EvalPlan{};
  bool lean_plan = lean && plan.ok;
  if (vec && lean && !plan.ok) plan = eval_plan(n, B, nthr);      // shapes too small / odd for four slots: MUFU path
  if (plan.ok) {
    DP_CHECK_ARG(workspace, "dp_eval_metrics: null workspace");
    if (workspace_bytes < eval_ws_bytes(B, plan))
      return dp_set_error(DP_ERR_WORKSPACE, "dp_eval_metrics: workspace %zu < %zu", workspace_bytes, eval_ws_bytes(B, plan));
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    unsigned long long* ll = reinterpret_cast<unsigned long long*>(ws);
    double* mom_part = reinterpret_cast<double*>(ws + eval_ws_ll_bytes(B, plan.G));
    unsigned long long* cnt_part =
        reinterpret_cast<unsigned long long*>(ws + eval_ws_ll_bytes(B, plan.G) + (size_t)B * plan.G * 4 * sizeof(double));
    EvsArgs a;
    a.pred = pred; a.target = target; a.B = B; a.nthr = nthr; a.G = plan.G; a.ngroups = plan.ngroups; a.per = plan.per;
    a.n = n; a.eps = eps; a.ll = ll; a.mom_part = mom_part; a.cnt_part = cnt_part;
    a.pf_dist = kEvPrefetch;
    bool onediv = true;
    for (int k = 0; k < DP_MAX_THR; ++k) {
      a.thr[k] = k < nthr ? thresholds[k] : 0.f;
      if (k < nthr && !(thresholds[k] > 1.0f)) onediv = false;
    }
#define DP_EVS_PICK(F, O)                                                                       \
  (nthr == 3 ? (const void*)eval_stream_kernel<F, O, 3>                                         \
             : (nthr == 1 ? (const void*)eval_stream_kernel<F, O, 1> : (const void*)eval_stream_kernel<F, O, 0>))
#define DP_EVS_LEAN                                                                                                  \
  (nthr == 3 ? (const void*)eval_stream_kernel<false, false, 3, true, kEvSlotsLean, kEvLagLean, kEvCWLean, kEvVPTLean>       \
             : (nthr == 1 ? (const void*)eval_stream_kernel<false, false, 1, true, kEvSlotsLean, kEvLagLean, kEvCWLean, kEvVPTLean> \
                          : (const void*)eval_stream_kernel<false, false, 0, true, kEvSlotsLean, kEvLagLean, kEvCWLean, kEvVPTLean>))
    const void* fn;
    // mode 2 shares one reciprocal between the SI term (eps) and AbsRel (1e-6, util.py:218): needs eps == 1e-6
    if (lean_plan) fn = DP_EVS_LEAN;
    else if (fast_math) fn = onediv ? DP_EVS_PICK(true, true) : DP_EVS_PICK(true, false);
    else fn = onediv ? DP_EVS_PICK(false, true) : DP_EVS_PICK(false, false);
#undef DP_EVS_LEAN
#undef DP_EVS_PICK
    void* kargs[1] = {&a};
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.dyn);
    int resident = 0;
    const int ev_threads = lean_plan ? (kEvCWLean + 2) * 32 : kEvThreads;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, fn, ev_threads, plan.dyn);
    if (e != cudaSuccess)
      return dp_set_error(DP_ERR_CUDA, "dp_eval_metrics: launch setup failed: %s", cudaGetErrorString(e));
    if (resident < 1)
      return dp_set_error(DP_ERR_UNSUPPORTED, "dp_eval_metrics: a CTA with %zu B of shared memory is not resident", plan.dyn);
    e = cudaMemsetAsync(ll, 0, (size_t)B * plan.G * 2 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return dp_set_error(DP_ERR_CUDA, "dp_eval_metrics: memset failed: %s", cudaGetErrorString(e));
    e = cudaLaunchCooperativeKernel(fn, dim3(plan.ngroups * plan.G), dim3(ev_threads), kargs, plan.dyn, stream);
    dp_count_launch(1);
    if (e != cudaSuccess)
      return dp_set_error(DP_ERR_CUDA, "eval_stream_kernel launch failed: %s", cudaGetErrorString(e));
    double* terms = reinterpret_cast<double*>(cnt_part + (size_t)B * plan.G * DP_MAX_THR);
    dp::launch(eval_gather_kernel, (B + 31) / 32, 1024, 0, stream, mom_part, cnt_part, B, plan.G, (double)n, nthr, moments, counts, terms);
    DP_CHECK_LAUNCH("eval_gather_kernel");
    dp::launch(eval_means_kernel, 1, 256, 0, stream, terms, B, (double)n, nthr, out);
    DP_CHECK_LAUNCH("eval_means_kernel");
    return DP_OK;
  }

  EvalArgs a;
  a.pred = pred; a.target = target; a.B = B; a.nthr = nthr; a.n = n; a.eps = eps;
  for (int k = 0; k < DP_MAX_THR; ++k) a.thr[k] = k < nthr ? thresholds[k] : 0.f;
  a.moments = moments; a.counts = counts;
  const dim3 grid(B * kEvalCluster);
#define DP_EVAL_LAUNCH(V, F)                                                        \
  do {                                                                              \
    if (nthr == 3) eval_fused_kernel<V, F, 3><<<grid, TPB, 0, stream>>>(a);         \
    else if (nthr == 1) eval_fused_kernel<V, F, 1><<<grid, TPB, 0, stream>>>(a);    \
    else eval_fused_kernel<V, F, 0><<<grid, TPB, 0, stream>>>(a);                   \
  } while (0)
  if (vec && fast_math) DP_EVAL_LAUNCH(4, true);
  else if (vec) DP_EVAL_LAUNCH(4, false);
  else if (fast_math) DP_EVAL_LAUNCH(1, true);
  else DP_EVAL_LAUNCH(1, false);
#undef DP_EVAL_LAUNCH
  DP_CHECK_LAUNCH("eval_fused_kernel");
  metrics_combine_kernel<<<1, 256, 0, stream>>>(moments, counts, B, H, W, nthr, out);
  DP_CHECK_LAUNCH("metrics_combine_kernel");
  return DP_OK;
}

}  // extern "C"
