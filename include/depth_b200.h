/* depth_b200.h - C ABI of libdepth_b200.so (hand-written sm_100a kernels for the dense-prediction hot path).
 *
 * The reference (HairongLuo/monocular-depth-estimation-cil) is pure Python/PyTorch and has no FFI; its seam is
 * the Python call signatures in src/util.py, src/main.py:51-89 and the nn.Module classes in src/network/.
 * Every entry point below names the reference arithmetic it replaces.  The Python host side
 * (monocular-depth-estimation-cil_b200/) binds these with ctypes and mirrors the reference signatures.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the comment says "host";
 *   - the caller owns all memory including workspaces (query the *_workspace() functions);
 *   - nothing allocates, synchronises or keeps mutable global state; work is enqueued on `stream`;
 *   - return 0 on success, a negative DP_ERR_* code otherwise; dp_last_error() gives the thread-local message;
 *   - re-entrant across host threads and streams (autograd's backward thread calls in concurrently).
 * Activations handled by the conv/BN/resize entry points are NHWC bf16; loss/metric inputs are the
 * reference's NCHW fp32 (B,1,H,W) tensors.
 */
#ifndef DEPTH_B200_H
#define DEPTH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define DP_ABI_VERSION 3

int dp_abi_version(void);
const char* dp_last_error(void);
/* number of kernel launches issued through this library by the calling process (bench.py's gpu_launches) */
unsigned long long dp_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Loss / metric reductions
 * ---------------------------------------------------------------------------------------------- */
/* per-sample raw moments, double[B][DP_NMOM] */
#define DP_NMOM 16
#define DP_M_S1 0  /* sum d,   d = log(p+eps)-log(t+eps)              util.py:143-150 */
#define DP_M_S2 1  /* sum d^2                                                          */
#define DP_M_M0 2  /* #pixels with t>0                                 util.py:108     */
#define DP_M_M1 3  /* sum d over t>0                                   util.py:114-121 */
#define DP_M_M2 4  /* sum d^2 over t>0                                                 */
#define DP_M_GX 5  /* sum ||dx p|-|dx t||  over (H, W-1)               util.py:35-41   */
#define DP_M_GY 6  /* sum ||dy p|-|dy t||  over (H-1, W)               util.py:36-42   */
#define DP_M_EX 7  /* sum g*||dx p|-|dx t||, g = normalised RGB grad   util.py:85      */
#define DP_M_EY 8  /*                                                   util.py:86      */
#define DP_M_AR 9  /* sum |t-p|/(t+1e-6)                               util.py:218     */
#define DP_M_AB 10 /* sum |p-t|                                        main.py:291     */
#define DP_M_SQ 11 /* sum (p-t)^2                                      main.py:292     */
#define DP_M_V0 12 /* #pixels with t>1e-6                              main.py:304     */
#define DP_M_V1 13 /* sum dv, dv = log(max(p,1e-6)) - log(t) over t>1e-6  main.py:311-318 */
#define DP_M_V2 14 /* sum dv^2                                                          */

/* which terms a pass evaluates */
#define DP_F_SI 1u
#define DP_F_SILOG 2u
#define DP_F_GRAD 4u
#define DP_F_EDGE 8u
#define DP_F_ABSREL 16u
#define DP_F_M4 32u

/* dp_loss_combine output slots, float[DP_NLOSS] */
#define DP_NLOSS 8
#define DP_L_TOTAL 0     /* main.py:82   */
#define DP_L_SI 1        /* si * alpha   */
#define DP_L_SILOG 2
#define DP_L_GRAD 3
#define DP_L_EDGE 4
#define DP_L_ABSREL 5    /* util.py:218  */
#define DP_L_SI_RAW 6    /* unweighted (sqroot applied if requested: evaluation.py:157) */
#define DP_L_SILOG_RAW 7

#define DP_MAX_THR 8

size_t dp_depth_moments_workspace(int B, int H, int W);
size_t dp_rgb_minmax_bytes(void);

/* util.py:58-70: per-block (min,max) of the RGB gradient magnitude; rgb is (B,3,H,W) fp32; out is dp_rgb_minmax_bytes() */
int dp_rgb_gradmag_minmax(const float* rgb, int B, int H, int W, float* minmax_partials, cudaStream_t stream);

/* One pass over pred/target (B,1,H,W fp32) [and rgb]: fills moments[B][DP_NMOM] for the terms in `flags`.
 * Replaces the elementwise+reduction chains of util.py:24-156,210-219 and main.py:291-318. */
int dp_depth_moments(const float* pred, const float* target, const float* rgb, const float* minmax_partials,
                     int B, int H, int W, unsigned flags, float eps, double* moments,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* moments -> scalars of main.py:66-82 (weights applied as there); per_sample (float[B], may be NULL) gets the
 * per-sample SI terms (sqrt'ed when sqroot!=0, evaluation.py:157). */
int dp_loss_combine(const double* moments, int B, int H, int W, unsigned flags, float w_si, float w_silog,
                    float variance_focus, float w_grad, float beta, int sqroot, float* out, float* per_sample,
                    cudaStream_t stream);

/* d(total loss)/d(pred): what autograd derives from main.py:66-82; grad_out is a device scalar or NULL (=1);
 * si_sample_scale (float[B] or NULL) multiplies the SI term per sample (1/(2*sqrt(v_b)) for sqroot=True). */
int dp_loss_backward(const float* pred, const float* target, const float* rgb, const float* minmax_partials,
                     const double* moments, const float* grad_out, const float* si_sample_scale,
                     int B, int H, int W, unsigned flags, float eps,
                     float w_si, float w_silog, float variance_focus, float w_grad, float beta, float* grad_pred,
                     cudaStream_t stream);

/* util.py:200-205 (aligned=1, eps_div=0: scale from moments[.][DP_M_S1]) or main.py:318-321 (aligned=0, eps_div=1e-6).
 * thresholds: HOST pointer to nthr floats.  counts: device u64[B][nthr]. */
int dp_delta_counts(const float* pred, const float* target, const double* moments, int B, int H, int W,
                    const float* thresholds, int nthr, int aligned, float eps_div, unsigned long long* counts,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* util.py:159-181 per_pixel_scale_invariant_loss: out[b][i] = (d_i - mean_b d)^2, d = log p - log t (no epsilon);
 * `moments` from dp_depth_moments(flags = DP_F_SI, eps = 0).  pred, target, out: fp32 (B,H,W). */
int dp_per_pixel_si(const float* pred, const float* target, const double* moments, int B, int H, int W, float* out,
                    cudaStream_t stream);

/* evaluation.py:157-166 for one batch: out[0]=SI-RMSE, out[1]=AbsRel, out[2+k]=delta_k */
int dp_metrics_combine(const double* moments, const unsigned long long* counts, int B, int H, int W, int nthr,
                       float* out, cudaStream_t stream);

/* evaluation.py:157-166 for one batch in one streaming pass (8 B/px of HBM traffic): groups of co-resident CTAs keep
 * their slices of (pred, target) in shared memory (1-D bulk copies on mbarriers), accumulate the SI / AbsRel moments
 * (util.py:129-156, 210-219), exchange the per-sample scale through `workspace` and count the scale-aligned delta
 * thresholds (util.py:183-207) from shared memory; then a parallel combine.  Shapes the streaming kernel cannot take
 * (H*W not a multiple of 4, bases not 16-byte aligned) run the thread-block-cluster kernel (second sweep from L2).
 * thresholds: HOST pointer.  moments: device double[B][DP_NMOM] (S1, S2, AR filled).  counts: device u64[B][nthr].
 * out: device float[2+nthr] = SI-RMSE, AbsRel, delta_k (batch means).
 * workspace: device, >= dp_eval_metrics_workspace(B,H,W) bytes (contents irrelevant on entry).
 * fast_math = 0: IEEE logf / division (the reference's arithmetic, kept as the checker); 1: MUFU lg2 / rcp per operand;
 * 2: the lean arithmetic - one shared reciprocal and one lg2 per pixel in the moments sweep, division-free
 * classification hi < thr * lo in the counting sweep, exact two-quotient code for slices that hold a negative value or
 * a non-finite scale.  Modes 1 and 2 differ from mode 0 by rounding only (SI-RMSE / AbsRel ~1e-7 relative, delta counts
 * a few pixels per million; contract 1e-5 relative / 0.01 % of pixels). */
size_t dp_eval_metrics_workspace(int B, int H, int W);
/* The streaming kernel's decomposition for `pixels` per sample and `nthr` thresholds on a device with `smem_per_sm`
 * bytes of shared memory and `sms` SMs (host arithmetic only, no CUDA call): CTAs per sample, sample groups in flight, pixels per CTA slice and
 * dynamic shared memory per CTA.  Returns 1, or 0 when the shape goes through the cluster kernel instead. */
int dp_eval_metrics_plan(long long pixels, int B, int nthr, int smem_per_sm, int sms, int* ctas_per_sample, int* groups,
                         int* slice_pixels, size_t* dynamic_smem);
int dp_eval_metrics(const float* pred, const float* target, int B, int H, int W, const float* thresholds, int nthr,
                    float eps, int fast_math, double* moments, unsigned long long* counts, float* out,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Convolutions (NHWC bf16 activations, fp32 accumulation in TMEM)
 * ---------------------------------------------------------------------------------------------- */
/* 3x3/stride 1/pad 1 or 1x1 convolution as an implicit GEMM on tcgen05 tensor cores.
 * Replaces nn.Conv2d forward at blocks.py:149-161,335-341,401; midas_net_custom.py:106-110;
 * midas_semantics.py:132-143,195; dpt_depth.py:39-47,103-106 - and, called with the transposed+flipped
 * weight pack, the data gradient autograd derives for them.
 *   x         NHWC bf16, pixel stride x_ld elements (a channel slice of a wider buffer is fine)
 *   w_packed  bf16 [KS*KS][Cout][Cin_p]  (tap = r*KS+s; Cin_p >= Cin, zero padded)
 *   epilogue  y = acc (+ bias[c]) (+ residual[pixel][c]) (+ residual2[pixel][c]);
 *             out  = relu ? max(y,0) : y   (may be NULL)
 *             out2 = relu2 ? max(y,0) : y  (may be NULL)
 *   stats_partials  NULL or float[dp_conv2d_tc_grid()][2][Cout]: per-CTA sum / sum of squares of y over pixels
 *                   (train-mode BatchNorm statistics, midas_semantics.py:133,136,142,196) */
int dp_conv2d_tc_grid(int B, int H, int W, int Cin, int Cout, int KS);
int dp_conv2d_tc(const void* x, long long x_ld, int B, int H, int W, int Cin, const void* w_packed, int Cin_p,
                 int Cout, int KS, const float* bias, const void* residual, long long res_ld, const void* residual2,
                 long long res2_ld, int relu, void* out, long long out_ld, void* out2, long long out2_ld, int relu2,
                 float* stats_partials, cudaStream_t stream);

/* The same convolution with BatchNorm fused on either side of the GEMM (reference conv -> BN -> ReLU -> conv chains:
 * midas_semantics.py:132-150 ResidualBlock, :195-203 fusion_head / depth_head).
 *   prologue: the A operand is act(x * scale[c] + shift[c]) - the train/eval BatchNorm2d + ReLU / ReLU6 that precedes
 *     the conv in the reference - applied to every TMA-landed tile in shared memory by two transform warps before the
 *     tensor core reads it; pixels outside the image stay exactly zero (nn.Conv2d pads the ACTIVATED tensor).  The
 *     activated tensor is never written to HBM.
 *   backward mask: for a data-gradient launch whose output is the gradient w.r.t. a = act(c * scale + shift), the
 *     epilogue recomputes the activation from c (mask_x, the pre-activation tensor, (B,H,W,Cout) bf16), zeroes the
 *     clamped positions, and stats_partials receives [grid][2][Cout] = (sum g, sum g * c) over the stored values: the two
 *     batch reductions of BatchNorm's backward, which therefore needs no reduction pass of its own.
 * dp_conv2d_tc_caps: bit mask of what this shape supports (the mask epilogue needs a single N block of <= 64 columns). */
#define DP_CONV_CAP_PROLOGUE 1
#define DP_CONV_CAP_BN_BACKWARD 2
typedef struct dp_conv_fuse {
  const float* pre_scale_shift;   /* [2][Cin] fp32 (dp_bn_finalize / dp_bn_eval_coeffs layout) or NULL */
  int pre_act;                    /* 0 affine only, 1 ReLU, 2 ReLU6 */
  const void* mask_x;             /* NULL or pre-activation tensor c, NHWC bf16, Cout channels */
  long long mask_ld;
  const float* mask_scale_shift;  /* [2][Cout] */
  int mask_act;                   /* 1 ReLU, 2 ReLU6 */
} dp_conv_fuse_t;
int dp_conv2d_tc_caps(int B, int H, int W, int Cin, int Cout, int KS);
int dp_conv2d_tc_fused(const void* x, long long x_ld, int B, int H, int W, int Cin, const void* w_packed, int Cin_p,
                       int Cout, int KS, const float* bias, const void* residual, long long res_ld, const void* residual2,
                       long long res2_ld, int relu, void* out, long long out_ld, void* out2, long long out2_ld, int relu2,
                       float* stats_partials, const dp_conv_fuse_t* fuse, cudaStream_t stream);

/* Stride-2 convolution (conv rule i = 2*o - pad + k, K x K taps) on the tensor cores, reading the four parity planes
 * of the input as strided TMA tensors: nn.Conv2d(k3,s2,p1) forward (midas_semantics.py:39-45, dpt_depth.py:63-68) and
 * the data gradient of nn.ConvTranspose2d(k4,s2,p1) (midas_semantics.py:52-58).
 * x (B,Hi,Wi,Cin) NHWC bf16; w_packed bf16 [K*K][Cout][Cin_p]; out (B,Ho,Wo,Cout); stats as in dp_conv2d_tc. */
int dp_conv2d_tc_down2_grid(int B, int Ho, int Wo, int Cin, int Cout, int K, int pad);
int dp_conv2d_tc_down2(const void* x, long long x_ld, int B, int Hi, int Wi, int Cin, const void* w_packed, int Cin_p,
                       int Cout, int K, int pad, const float* bias, int relu, void* out, long long out_ld, int Ho,
                       int Wo, float* stats_partials, cudaStream_t stream);
/* Transposed stride-2 convolution (rule i = (o + pad - k)/2 when even) as four output-phase launches whose epilogue
 * writes pixel (2y+a, 2x+b): nn.ConvTranspose2d(k4,s2,p1) forward and the data gradient of nn.Conv2d(k3,s2,p1). */
int dp_conv2d_tc_up2(const void* x, long long x_ld, int B, int Hi, int Wi, int Cin, const void* w_packed, int Cin_p,
                     int Cout, int K, int pad, const float* bias, int relu, void* out, long long out_ld, int Ho, int Wo,
                     cudaStream_t stream);

/* Weight gradient of the same convolutions (autograd's convolution_backward w.r.t. weight), tcgen05 GEMM over
 * pixels with split-K partials reduced deterministically.  x, dy: NHWC bf16; grad_oihw: fp32 [Cout][Cin][KS][KS]
 * (overwritten, or added to when accumulate != 0). */
size_t dp_conv2d_wgrad_tc_workspace(int B, int H, int W, int Cin, int Cout, int KS);
int dp_conv2d_wgrad_tc(const void* x, long long x_ld, const void* dy, long long dy_ld, int B, int H, int W, int Cin,
                       int Cout, int KS, float* grad_oihw, int accumulate, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream);

/* Same, with x = act(c * scale[ci] + shift[ci]) computed on the fly from the pre-BatchNorm tensor c passed as `x`
 * (pre_scale_shift [2][Cin] fp32, pre_act 0 / 1 ReLU / 2 ReLU6; zero padding applies to the activated tensor): the weight
 * gradient of the second conv of a conv -> BN -> ReLU -> conv chain without the activated tensor ever being stored. */
int dp_conv2d_wgrad_tc_fused(const void* x, long long x_ld, const void* dy, long long dy_ld, int B, int H, int W, int Cin,
                             int Cout, int KS, float* grad_oihw, int accumulate, void* workspace, size_t workspace_bytes,
                             const float* pre_scale_shift, int pre_act, cudaStream_t stream);

/* grad[cp][ct][ky][kx] (+)= sum over plain-grid pixels p of P[p][cp] * T[2p - pad + k][ct]  (K x K taps, stride 2):
 *   nn.Conv2d(k3,s2,p1):          P = dY (B,Ho,Wo,O), T = X  (B,Hi,Wi,I)  -> weight.grad [O][I][3][3]
 *   nn.ConvTranspose2d(k4,s2,p1): P = X  (B,Hi,Wi,I), T = dY (B,Ho,Wo,O)  -> weight.grad [I][O][4][4] */
size_t dp_conv2d_wgrad_tc_s2_workspace(int B, int Hp, int Wp, int Cp, int Ct, int K, int pad);
int dp_conv2d_wgrad_tc_s2(const void* P, long long p_ld, int Hp, int Wp, int Cp, const void* T, long long t_ld, int Ht,
                          int Wt, int Ct, int B, int K, int pad, float* grad, int accumulate, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Layout / dtype boundaries and weight packing (csrc/layout.cu)
 * ---------------------------------------------------------------------------------------------- */
/* reference modules speak NCHW fp32; (B,C,H,W) fp32 <-> (B,H,W,ld) bf16.  Channels C..ld-1 of dst are left untouched,
 * except for C <= 8 with dst_ld == 8 (the RGB input of the stem convolution), where they are written as zeros. */
int dp_nchw_f32_to_nhwc_bf16(const float* src, int B, int C, int H, int W, void* dst, long long dst_ld, cudaStream_t stream);
int dp_nhwc_bf16_to_nchw_f32(const void* src, long long src_ld, int B, int C, int H, int W, float* dst, cudaStream_t stream);
int dp_cast_f32_to_bf16(const float* src, void* dst, size_t n, cudaStream_t stream);
int dp_cast_bf16_to_f32(const void* src, float* dst, size_t n, cudaStream_t stream);
/* fp32 weight [D0][D1][KH][KW] (nn.Conv2d: D0=out, D1=in; nn.ConvTranspose2d: D0=in, D1=out) -> bf16 [KH*KW][A][ld]
 * with (A, b) = swap ? (D1, d0) : (D0, d1) and the tap order reversed when flip != 0.
 *   conv forward operand:  swap 0 flip 0;   stride-1 data gradient (correlation with dY): swap 1 flip 1;
 *   gather-form data gradient / transposed-conv forward: swap 1 flip 0;  transposed-conv data gradient: swap 0 flip 0 */
int dp_pack_conv_weight(const float* w, int D0, int D1, int KH, int KW, int swap, int flip, void* dst, int ld,
                        cudaStream_t stream);
/* the same for n weights in one launch: `descs_device` is a DEVICE array of n descriptors (fields as the arguments of
 * dp_pack_conv_weight); blocks_per_weight 256-thread blocks stride over each weight. */
typedef struct dp_pack_desc {
  const void* src;      /* fp32 [D0][D1][KH][KW] */
  void* dst;            /* bf16 [KH*KW][A][ld] */
  int D0, D1, KH, KW, swap, flip, ld, pad_;
} dp_pack_desc_t;
int dp_pack_conv_weights_batched(const dp_pack_desc_t* descs_device, int n, int blocks_per_weight, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Bandwidth-bound NHWC bf16 kernels (csrc/elementwise.cu)
 * ---------------------------------------------------------------------------------------------- */
/* out = g_raw + g_relu * (y > 0): gradient of a tensor consumed both raw and through nn.ReLU (blocks.py:361-374) */
int dp_add_relu_bwd(const void* g_raw, const void* g_relu, const void* y, void* out, size_t n, cudaStream_t stream);
int dp_relu_bf16(const void* x, void* out, size_t n, cudaStream_t stream);
int dp_add_bf16(const void* a, const void* b, const void* c, void* out, size_t n, cudaStream_t stream);
/* dst[p][0..C) = src[p][0..C): channel concat (torch.cat dim=1, midas_semantics.py:250) and slice copies */
int dp_copy_channels(const void* src, long long src_ld, void* dst, long long dst_ld, size_t npix, int C,
                     cudaStream_t stream);
/* F.interpolate(mode="bilinear") forward / backward (blocks.py:226-238,432-434; dpt_depth.py:147; midas_semantics.py:243) */
int dp_resize_bilinear_nhwc(const void* src, long long src_ld, int B, int Hi, int Wi, int C, void* dst, long long dst_ld,
                            int Ho, int Wo, int align_corners, cudaStream_t stream);
int dp_resize_bilinear_nhwc_bwd(const void* gout, long long g_ld, int B, int Hi, int Wi, int C, void* gin,
                                long long gin_ld, int Ho, int Wo, int align_corners, cudaStream_t stream);
/* fp32 planes: the RGB->DINOv2 resize (midas_semantics.py:233) and the prediction resize (util.py:308-313) */
int dp_resize_bilinear_planes_f32(const float* src, int planes, int Hi, int Wi, float* dst, int Ho, int Wo,
                                  int align_corners, cudaStream_t stream);
/* per-channel sums over pixels: mode 0 sum x; 1 sum x, sum x^2; 2 sum g, sum g*x with g = dy*(mask>0); 3 = 2 with the
 * ReLU6 mask (0 < mask < 6).
 * partial: float[dp_chan_reduce_blocks()][2][C]; dp_sum_partials folds partials into out[rows][C] */
int dp_chan_reduce_blocks(void);
/* mask_ss (modes 2/3, optional): BN scale/shift [2][C] of the forward pass - the activation mask is then recomputed
 * from x (x*scale+shift in fp32) instead of being read from the activated output; `mask`, if given, is then the residual added before the activation. */
int dp_chan_reduce(int mode, const void* x, long long x_ld, const void* dy, long long dy_ld, const void* mask,
                   long long m_ld, const float* mask_ss, size_t npix, int C, float* partial, cudaStream_t stream);
int dp_sum_partials(const float* partial, int nparts, int rows, int C, float* out, int accumulate, cudaStream_t stream);
/* nn.BatchNorm2d (midas_semantics.py:40-61,133-151,196): train-mode finalize from (sum, sumsq) partials incl. running
 * statistics update (momentum, unbiased variance, num_batches_tracked += 1); eval-mode coefficients; apply; backward */
int dp_bn_finalize(const float* partial, int nparts, int C, double count, const float* gamma, const float* beta,
                   float eps, float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                   float* scale_shift, float* save_mean_invstd, cudaStream_t stream);
int dp_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, int C, float* scale_shift, float* save_mean_invstd, cudaStream_t stream);
int dp_bn_apply(const void* x, long long x_ld, const float* scale_shift, const void* x2, long long x2_ld,
                const float* scale_shift2, const void* res, long long res_ld, size_t npix, int C, int relu, void* y,
                long long y_ld, cudaStream_t stream);
int dp_bn_bwd_apply(const void* dy, long long dy_ld, const void* mask, long long m_ld, const float* mask_ss,
                    const void* x, long long x_ld,
                    const float* red, const float* save_mean_invstd, const float* gamma, double count, int train,
                    size_t npix, int C, void* dx, long long dx_ld, void* gmask, long long gm_ld, float* dgamma,
                    float* dbeta, int accumulate, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Direct convolutions for strided / transposed layers and the C->1 heads (csrc/conv_direct.cu)
 * ---------------------------------------------------------------------------------------------- */
/* gather-form conv: conv rule iy = oy*stride - pad + ky, or transposed rule iy = (oy + pad - ky)/stride.
 * Serves nn.Conv2d(stride 2) and nn.ConvTranspose2d forward and both data gradients
 * (midas_semantics.py:38-61; dpt_depth.py:49-69).  w_packed bf16 [KH*KW][Co][Ci]. */
int dp_conv_gather(const void* in, long long in_ld, int B, int Hi, int Wi, int Ci, const void* w_packed,
                   const float* bias, void* out, long long out_ld, int Ho, int Wo, int Co, int KH, int KW, int stride,
                   int pad, int transposed, int relu, cudaStream_t stream);
size_t dp_conv_wgrad_direct_workspace(int B, int Hp, int Wp, int Cp, int Ct, int KH, int KW);
int dp_conv_wgrad_direct(const void* P, long long p_ld, int Hp, int Wp, int Cp, const void* T, long long t_ld, int Ht,
                         int Wt, int Ct, int B, int KH, int KW, int stride, int pad, int perm, float* out,
                         int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* C->1 conv (3x3 pad 1 or 1x1) + bias + optional ReLU, fp32 (B,H,W) output: midas_semantics.py:203-204,
 * midas_net_custom.py:110-111, dpt_depth.py:282-283.  w fp32 [1][C][KS][KS]. */
int dp_head_conv_fwd(const void* x, long long x_ld, int B, int H, int W, int C, int KS, const float* w,
                     const float* bias, int relu, float* out, cudaStream_t stream);
size_t dp_head_conv_bwd_workspace(int C, int KS);
int dp_head_conv_bwd(const float* dout, const float* out, int relu, const void* x, long long x_ld, int B, int H, int W,
                     int C, int KS, const float* w, void* dx, long long dx_ld, float* dw, float* db, int accumulate,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------
 * CrossAttention token path (csrc/attention.cu; midas_semantics.py:63-127), dim 32 = 8 heads x 4
 * ---------------------------------------------------------------------------------------------- */
int dp_lnl_blocks(void);
int dp_lnl_partial_floats(void);
int dp_ln_linear_fwd(const void* x, long long x_ld, int x_is_f32, size_t ntok, int dim, const float* gamma,
                     const float* beta, float eps, const float* W, const float* bias, void* out, long long out_ld,
                     int out_is_bf16, cudaStream_t stream);
int dp_ln_linear_bwd(const void* x, long long x_ld, int x_is_f32, size_t ntok, int dim, const float* gamma,
                     const float* beta, float eps, const float* W, const void* dout, long long do_ld, int dout_is_bf16,
                     void* dx, long long dx_ld, float* partial, float* dW, float* dbias, float* dgamma, float* dbeta,
                     int accumulate, cudaStream_t stream);
/* window loop of midas_semantics.py:93-112 in last-writer form; items/segs: device int4 arrays (see attention.cu) */
int dp_attn_fwd(const float* q, const float* k, const float* v, int B, int N, float scale, const void* items,
                int nitems, float* out, float* lse, cudaStream_t stream);
int dp_attn_bwd(const float* q, const float* k, const float* v, const float* out, const float* dout, const float* lse,
                int B, int N, float scale, const void* items, int nitems, const void* segs, int nseg, float* dq,
                float* dk, float* dv, float* delta, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Depthwise convolutions of the EfficientNet-Lite3 trunk (hub model consumed at blocks.py:166-186)
 * ---------------------------------------------------------------------------------------------- */
/* K in {3,5}, stride in {1,2}; explicit top/left padding (symmetric or TF-"SAME"); w fp32 [K*K][C] tap-major.
 * stats_partials: null or float[dp_dwconv_fwd_blocks()][2][C] = per-block (sum, sumsq) of the stored output. */
int dp_dwconv_fwd_blocks(int B, int Ho, int Wo, int C);
/* rows of the statistics partials for a given stride: stride 1 runs the shared-memory tile kernel (one TMA box per
 * 8 x TW output tile, persistent blocks per channel chunk), stride 2 the register-window kernel */
int dp_dwconv_fwd_blocks_s(int B, int Ho, int Wo, int C, int K, int stride);
int dp_dwconv_fwd(const void* x, long long x_ld, int B, int Hi, int Wi, int C, const float* w, int K, int stride,
                  int pad_t, int pad_l, void* out, long long out_ld, int Ho, int Wo, float* stats_partials,
                  cudaStream_t stream);
int dp_dwconv_dgrad_s2(const void* dy, long long dy_ld, int B, int Ho, int Wo, int C, const float* w, int K, int pad_t,
                       int pad_l, void* dx, long long dx_ld, int Hi, int Wi, cudaStream_t stream);
size_t dp_dwconv_wgrad_workspace(int B, int Ho, int Wo, int C, int K);
int dp_dwconv_wgrad(const void* x, long long x_ld, int B, int Hi, int Wi, int C, const void* dy, long long dy_ld, int Ho,
                    int Wo, int K, int stride, int pad_t, int pad_l, float* grad, int accumulate, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Diagnostics
 * ---------------------------------------------------------------------------------------------- */
/* device buffer (>= 8 u64) that block 0 of subsequent wgrad launches fills with cycle counters; NULL switches it off */
void dp_debug_set_buffer(void* p);

#ifdef __cplusplus
}
#endif
#endif /* DEPTH_B200_H */
