/*
 * segsde_b200.h — C-ABI of the B200-native monodepth / joint-segmentation hot path.
 *
 * The reference (lhoyer/improving_segmentation_with_selfsupervised_depth) has no FFI: its
 * hot path is the Python surface of models/ and loss/ executing stock ATen/cuDNN ops.  Each
 * entry point below replaces the ATen op sequence at the cited reference location; the Python
 * mirror of models/ and loss/ (package improving_segmentation_with_selfsupervised_depth_b200)
 * binds them with ctypes.  See INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - activations are "NHWC views": channel stride 1, explicit element strides for n/h/w
 *     (so halo-padded buffers and channel slices of concat buffers are expressible);
 *   - image-like loss inputs (colour frames, disparities) stay in the reference's NCHW planar
 *     layout (loader contract, sequence_segmentation_loader.py:183-250);
 *   - weights are O,kh,kw,I ("OHWI") fp32 = an OIHW torch tensor in channels_last memory format;
 *   - every call takes the cudaStream_t to launch on (as void*), never allocates, never syncs;
 *   - return value: 0 = ok, <0 = argument error (SEGSDE_E_*), >0 = cudaError_t of the launch.
 */
#ifndef SEGSDE_B200_H
#define SEGSDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEGSDE_OK 0
#define SEGSDE_E_ARG (-1)       /* bad shape / null pointer / unsupported flag combination */
#define SEGSDE_E_ALIGN (-2)     /* pointer or stride not aligned as the kernel requires */
#define SEGSDE_E_UNSUPPORTED (-3)
#define SEGSDE_E_WORKSPACE (-4) /* workspace too small */

/* NHWC view: element (n,h,w,c) lives at ptr[n*sn + h*sh + w*sw + c]. */
typedef struct {
  void* ptr;
  int32_t n, h, w, c;
  int64_t sn, sh, sw;
} segsde_nhwc_t;

/* activation codes (conv epilogue / act backward) */
enum { SEGSDE_ACT_NONE = 0, SEGSDE_ACT_RELU = 1, SEGSDE_ACT_ELU = 2, SEGSDE_ACT_SIGMOID = 3 };
/* padding modes */
enum { SEGSDE_PAD_ZERO = 0, SEGSDE_PAD_REFLECT = 1 };

const char* segsde_version(void);
/* Human-readable text for a return code (static storage). */
const char* segsde_error_string(int code);
/* Number of kernel launches issued through this library since load (all streams). */
int64_t segsde_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Convolution (replaces nn.Conv2d / Conv3x3+ReflectionPad2d / upsample+cat feeding a conv:
 * models/monodepth_layers.py:127-142, models/depth_decoder.py:93-101, torchvision resnet blocks
 * used by models/resnet_encoder.py:90-101, models/pose_decoder.py:30-33, models/model_parts.py:9-25)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t kh, kw, stride, pad, dil;
  int32_t pad_mode;     /* SEGSDE_PAD_* */
  int32_t up1;          /* 1: source 1 is nearest-upsampled x2 on the fly (depth_decoder.py:93-96) */
  int32_t act;          /* SEGSDE_ACT_* fused after bias */
  int32_t nchw_norm_in; /* 1: x1.ptr is an NCHW planar image and (x-0.45)/0.225 is applied on load
                           (resnet_encoder.py:92); x1 strides are then ignored */
  int32_t stride_w;     /* 0 = same as `stride`; otherwise the horizontal stride while `stride` is the vertical one
                           (tensor-core fwd / wgrad only — the stem's row-band view, segsde_stem_pack) */
} segsde_conv_desc_t;

/* y = act(conv(cat(x1[up], x2), w) + bias).  x2 may be NULL (c=0).  bias may be NULL.
 * w: [Cout][kh][kw][C1+C2] fp32.  Generic CUDA-core path: any shape.  1x1 / stride 1 / single source with
 * 2 <= Cout <= 32, C % 4 == 0 and >= 4096 pixels (the 19-class segmentation heads,
 * joint_segmentation_depth_decoder.py:106-107) takes the few-output-channel kernels of conv_fewcout.cu in all three
 * directions (fwd / dgrad / wgrad below); results are fp32 either way. */
int segsde_conv2d_fwd(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2, const float* w,
                      const float* bias, const segsde_nhwc_t* y, const segsde_conv_desc_t* d,
                      void* stream);
/* dx{1,2} = conv_transpose(dy, w), folded through reflect padding / upsampling / the concat split.
 * Either dx may be NULL (source needs no gradient). dx buffers must be zero-filled by the caller
 * when d->pad_mode==REFLECT or d->up1 (folding accumulates with atomics). */
int segsde_conv2d_dgrad(const segsde_nhwc_t* dy, const float* w, const segsde_nhwc_t* dx1,
                        const segsde_nhwc_t* dx2, const segsde_conv_desc_t* d, void* stream);
/* dw += sum_pixels dy (x) patch(x) ; dbias += sum_pixels dy.  dw/dbias must be zero-filled (or hold
 * a running gradient) — split-K partial sums are accumulated with atomics. dbias may be NULL. */
int segsde_conv2d_wgrad(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2, const segsde_nhwc_t* dy,
                        float* dw, float* dbias, const segsde_conv_desc_t* d, void* stream);

/* Tensor-core (tcgen05/TMEM/TMA, TF32 in / FP32 accumulate) implicit-GEMM path.
 * Same contract as the generic entry points, restricted shapes; returns SEGSDE_E_UNSUPPORTED when
 * the shape is outside the family so the caller can route to the generic path. */
int segsde_conv2d_fwd_tc(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2, const float* w,
                         const float* bias, const segsde_nhwc_t* y, const segsde_conv_desc_t* d,
                         void* stream);
/* fwd_tc that also accumulates the BatchNorm batch statistics of its (pre-activation) output in the epilogue:
 * stats[0..C) += sum y, stats[C..2C) += sum y^2 (fp64, zero-filled by the caller; same layout as segsde_bn_stats
 * with shift 0, consumed by segsde_bn_finalize).  stats == NULL behaves like segsde_conv2d_fwd_tc. */
int segsde_conv2d_fwd_tc_stats(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2, const float* w, const float* bias,
                               const segsde_nhwc_t* y, const segsde_conv_desc_t* d, double* stats, void* stream);
int segsde_conv2d_dgrad_tc(const segsde_nhwc_t* dy, const float* w, const segsde_nhwc_t* dx1,
                           const segsde_nhwc_t* dx2, const segsde_conv_desc_t* d, void* stream);
int segsde_conv2d_wgrad_tc(const segsde_nhwc_t* x1, const segsde_nhwc_t* x2,
                           const segsde_nhwc_t* dy, float* dw, float* dbias,
                           const segsde_conv_desc_t* d, void* stream);
/* Helpers of the tensor-core route.
 * pad_prep: y[N, Hc+2p, Wc+2p, C] = ReflectionPad2d(p)(nearest-x2-upsample?(x)), (Hc,Wc) = (H,W) << up;
 * pad_fold: its adjoint (gradient w.r.t. x from the gradient w.r.t. the padded tensor);
 * weight_transpose_flip: wt[ci-c_begin][kh-1-r][kw-1-s][co] = w[co][r][s][ci], which makes
 *   dgrad(dy, w) == fprop(dy, wt) with pad' = dil*(k-1) - pad;
 * act_bwd_bias: dz = dy * act'(y) (dz may be NULL) and dbias[c] += sum_pixels dz (dbias may be NULL). */
int segsde_pad_prep(const segsde_nhwc_t* x, const segsde_nhwc_t* y, int up, int pad, void* stream);
int segsde_pad_fold(const segsde_nhwc_t* dyp, const segsde_nhwc_t* dx, int up, int pad, void* stream);
int segsde_weight_transpose_flip(const float* w, float* wt, int cout, int kh, int kw, int ctot, int c_begin,
                                 int c_count, void* stream);
/* nearest x2 upsampling + ReflectionPad2d(1) + 3x3 convolution (monodepth_layers.py:ConvBlock after depth_decoder.py:93-96)
 * as four 2x2 "phase" convolutions on the replicate-padded LOW-RES input: output pixel (2i+a, 2j+b) = 2x2 taps of
 * xp[:, a:, b:] with weights pre-summed over the 3x3 taps that fall on the same low-res pixel (a=0: {w0, w1+w2},
 * a=1: {w0+w1, w2}; columns alike).  pad_replicate: xp [n,h+2,w+2,c] from x.  weight_phase_up: wp[cout][2][2][c_count]
 * of phase (a,b) from w [cout][3][3][ctot].  weight_phase_up_fold: the adjoint, dw += fold(dwp[2][2][cout][2][2][c_count]).
 * phase_up_fold: dx [n,h,w,c] from the four gradients g_ab [n,h+1,w+1,c] w.r.t. the phase views (adjoint of the views +
 * of the replicate padding). */
int segsde_pad_replicate(const segsde_nhwc_t* x, const segsde_nhwc_t* xp, void* stream);
int segsde_weight_phase_up(const float* w, float* wp, int cout, int ctot, int c_begin, int c_count, int a, int b, void* stream);
int segsde_weight_phase_up_fold(const float* dwp, float* dw, int cout, int ctot, int c_begin, int c_count, void* stream);
int segsde_phase_up_fold(const float* g00, const float* g01, const float* g10, const float* g11, const segsde_nhwc_t* dx,
                         void* stream);
/* Weights of phase (a, b) in {0,1}^2 of the dgrad of a 3x3 / stride-2 / pad-1 convolution as a (1+a) x (1+b)-tap stride-1
 * convolution of dy that writes dx[:, a::2, b::2]: wt[ci][th][tw][co] = w[co][r][s][c_begin + ci] with row taps
 * a=0: {r=1}; a=1: {r=2 (offset 0), r=0 (offset +1)} and columns alike.  w: [cout][3][3][ctot] (OHWI). */
int segsde_weight_phase_s2(const float* w, float* wt, int cout, int ctot, int c_begin, int c_count, int a, int b,
                           void* stream);
int segsde_act_bwd_bias(const segsde_nhwc_t* y, const segsde_nhwc_t* dy, const segsde_nhwc_t* dz, int act,
                        float* dbias, void* stream);
/* Stem as a GEMM: cols[n,oh,ow,(r*kw+s)*C + c] = (x[n,c,oh*stride-pad+r,ow*stride-pad+s]-0.45)/0.225, zero outside
 * the image and for k >= kh*kw*C (kpad: K rounded up to a multiple of 32).  x1/x2: NCHW planar frames.
 * copy_rows: dst[r][c] = c < ncopy ? src[r][c] : 0 (weight matrix <-> its K-padded form). */
int segsde_stem_im2col(const float* x1, const float* x2, int c1, int c2, int n, int h, int w, int kh, int kw,
                       int stride, int pad, int kpad, float* cols, void* stream);
int segsde_copy_rows(const float* src, int ld_src, float* dst, int ld_dst, int rows, int ncopy, void* stream);
/* Stem without im2col (resnet_encoder.py:92-93, 7x7 / stride 2 / pad 3 on 3- or 6-channel frames).
 * stem_pack: xp[n][h+2*halo][wp][P] = (x-0.45)/0.225 of the NCHW frames x1 (c1 ch) ++ x2 (c2 ch, may be NULL) at
 * [.., y+halo, x+halo, c], zero in the halo, for x >= w+halo and for c >= c1+c2 (P = 4 or 8 channels per pixel).
 * An output pixel (oy, ox) then reads, for each kernel row ky, the 8*P contiguous floats that start at padded
 * pixel (2*oy+ky, 2*ox): the convolution is a (kh x 1)-tap implicit GEMM over the OVERLAPPING view
 * {c = 8*P, w = W/2 with stride 2*P floats, h = H+2*halo} with vertical stride 2 (segsde_conv_desc_t.stride_w = 1),
 * K = kh*8*P — the activation is read once instead of being expanded 12x by im2col.
 * stem_pack_w: dir 0: wp[co][ky][kx*P+c] = w[co][ky][kx][c] (zero for kx = 7 or c >= cin); dir 1: the inverse
 * gather (w <- wp) used on the weight gradient.  w: [cout][kh][kw][cin] (OHWI), kw <= 8. */
int segsde_stem_pack(const float* x1, const float* x2, int c1, int c2, int n, int h, int w, int halo, int wp, int P,
                     float* xp, void* stream);
int segsde_stem_pack_w(float* w, float* wp, int cout, int kh, int kw, int cin, int P, int dir, void* stream);
/* Disparity heads (C -> 1, 3x3, pad 1) on the tensor cores: y = act(bias + sum_t z[resolve(p + tap_t)][t]) over the
 * 9 tap planes z = 1x1conv(x) (first 9 of >= 9 channels), and the adjoint stencil gcol[q][t] (32 channels, 9 used)
 * of dy that turns dgrad / wgrad into 1x1 GEMMs. */
int segsde_head_stencil_fwd(const segsde_nhwc_t* z, const segsde_nhwc_t* y, const float* bias, int act, int reflect,
                            int pad, void* stream);
int segsde_head_gcol(const segsde_nhwc_t* dy, const segsde_nhwc_t* gcol, int reflect, int pad, void* stream);
/* 1 if this process can run the tensor-core path (driver entry point for tensor maps found). */
int segsde_tc_available(void);
/* The tensor-core kernel the calling thread launched last (0 = none yet) — test / profiling aid: which member of the
 * family a shape was routed to. */
#define SEGSDE_TC_KERNEL_CONV 1      /* tc_conv_kernel: one TMA box per tap (1x1, strided, narrow layers) */
#define SEGSDE_TC_KERNEL_ROWHALO 2   /* tc_conv3x3_kernel: 3x3 / stride 1, >= 128 px wide, row-halo reuse */
#define SEGSDE_TC_KERNEL_WGRAD 3     /* tc_wgrad_kernel */
#define SEGSDE_TC_KERNEL_WGRAD3X3 4  /* tc_wgrad3x3_kernel: halo-reuse weight gradient */
#define SEGSDE_TC_KERNEL_CONV256 5  /* tc_conv_kernel with 256-pixel tiles (two accumulators share each weight tile) */
int segsde_tc_last_kernel(void);

/* y = act(x) as a standalone pass (ConvBlock with BatchNorm: conv -> BN -> ELU). */
int segsde_act_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, int act, void* stream);
/* dz = dy * act'(y) computed from the activation OUTPUT y (ReLU/ELU/sigmoid). In-place allowed. */
int segsde_act_bwd(const segsde_nhwc_t* y, const segsde_nhwc_t* dy, const segsde_nhwc_t* dz, int act,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * BatchNorm2d (torchvision blocks, ASPP model_parts.py:11,23, seg head
 * joint_segmentation_depth_decoder.py:43).  Train mode = batch statistics (biased var for
 * normalisation, unbiased for the running update, momentum as torch).
 * ------------------------------------------------------------------------------------------- */
/* sums: 3C doubles, zero-filled: [0,C) = sum (x-s), [C,2C) = sum (x-s)^2, [2C,3C) = per-channel shift s
 * (0 on the generic path, x at pixel 0 on the fast path) — consumed by segsde_bn_finalize. */
int segsde_bn_stats(const segsde_nhwc_t* x, double* sums, void* stream);
/* mean/invstd from sums; running stats update when running_mean != NULL. count = N*H*W. */
int segsde_bn_finalize(const double* sums, int c, int64_t count, float eps, float momentum,
                       float* mean, float* invstd, float* running_mean, float* running_var,
                       void* stream);
/* invstd from running_var (eval mode): mean=running_mean, invstd=1/sqrt(var+eps). */
int segsde_bn_eval_prepare(const float* running_mean, const float* running_var, int c, float eps,
                           float* mean, float* invstd, void* stream);
/* y = act((x-mean)*invstd*gamma + beta [+ residual]); act in {NONE, RELU}. residual may be NULL. */
int segsde_bn_apply(const segsde_nhwc_t* x, const float* mean, const float* invstd,
                    const float* gamma, const float* beta, const segsde_nhwc_t* residual,
                    const segsde_nhwc_t* y, int act, void* stream);
/* Train-mode forward in one launch: segsde_bn_finalize (batch mean / invstd from `sums`, written to mean / invstd for
 * the backward pass, running buffers updated when given) followed by segsde_bn_apply.  Every CTA derives the
 * per-channel scale / shift from `sums` itself (fp64) into shared memory, CTA 0 also publishes mean / invstd and
 * updates the running statistics (torch.nn.BatchNorm2d semantics: unbiased variance, momentum). */
int segsde_bn_apply_train(const segsde_nhwc_t* x, const double* sums, int64_t count, float eps, float momentum,
                          const float* gamma, const float* beta, const segsde_nhwc_t* residual,
                          const segsde_nhwc_t* y, int act, float* mean, float* invstd, float* running_mean,
                          float* running_var, void* stream);
/* Backward, step 1: red[0..C)=sum dz, red[C..2C)=sum dz*xhat (fp64, zero-filled) where dz = dy * relu'(out).
 * act==RELU: the mask comes from the saved output y, or — y == NULL, layer WITHOUT a residual — is recomputed from x as
 * (x - mean) * (invstd * gamma) + beta > 0 (the forward kernels' exact expression): one full tensor read less. */
int segsde_bn_bwd_reduce(const segsde_nhwc_t* x, const segsde_nhwc_t* y, const segsde_nhwc_t* dy,
                         const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                         double* red, void* stream);
/* Backward, step 2: dx (train: full formula; eval: dz*gamma*invstd), dres = dz (may be NULL),
 * dgamma/dbeta (+=, may be NULL) from red.  y == NULL with act==RELU as above (not together with dres). */
int segsde_bn_bwd_apply(const segsde_nhwc_t* x, const segsde_nhwc_t* y, const segsde_nhwc_t* dy,
                        const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                        int training, const double* red, int64_t count, const segsde_nhwc_t* dx,
                        const segsde_nhwc_t* dres, float* dgamma, float* dbeta, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pooling / resampling / layout
 * ------------------------------------------------------------------------------------------- */
/* MaxPool2d(3, stride 2, pad 1) (resnet_encoder.py:96). idx: uint8 argmax tap per output element. */
int segsde_maxpool3x3s2_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, uint8_t* idx, void* stream);
int segsde_maxpool3x3s2_bwd(const segsde_nhwc_t* dy, const uint8_t* idx, const segsde_nhwc_t* dx,
                            void* stream); /* dx zero-filled by caller */
/* Spatial mean: y[n,c] = scale * mean_hw x (ASPPPooling, pose_decoder.py:53). y: [N,1,1,C] view. */
int segsde_spatial_mean_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, float scale, void* stream);
int segsde_spatial_mean_bwd(const segsde_nhwc_t* dy, const segsde_nhwc_t* dx, float scale, void* stream);
/* y[n,h,w,c] = x[n,0,0,c] (bilinear resize from 1x1) and its adjoint (sum over h,w). */
int segsde_broadcast_hw_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, void* stream);
int segsde_broadcast_hw_bwd(const segsde_nhwc_t* dy, const segsde_nhwc_t* dx, void* stream);
/* Strided NHWC copy (channel concat into a slice view; layout normalisation). */
int segsde_copy_nhwc(const segsde_nhwc_t* x, const segsde_nhwc_t* y, void* stream);
/* y = a + b ; y = a * sigmoid(b) (SelfAttention gate, model_parts.py:43-45) and its backward. */
int segsde_add(const segsde_nhwc_t* a, const segsde_nhwc_t* b, const segsde_nhwc_t* y, void* stream);
int segsde_gate_fwd(const segsde_nhwc_t* f, const segsde_nhwc_t* a, const segsde_nhwc_t* y, void* stream);
int segsde_gate_bwd(const segsde_nhwc_t* f, const segsde_nhwc_t* a, const segsde_nhwc_t* dy,
                    const segsde_nhwc_t* df, const segsde_nhwc_t* da, void* stream);
/* NCHW planar (arbitrary strides given in elements) <-> NHWC view. */
int segsde_nchw_to_nhwc(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                        const segsde_nhwc_t* y, void* stream);
int segsde_nhwc_to_nchw(const segsde_nhwc_t* x, float* y, int64_t sn, int64_t sc, int64_t sh,
                        int64_t sw, void* stream);
/* Bilinear resize, F.interpolate semantics (align_corners 0/1), multi-channel NHWC. */
int segsde_bilinear_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, int align_corners, void* stream);
int segsde_bilinear_bwd(const segsde_nhwc_t* dy, const segsde_nhwc_t* dx, int align_corners,
                        void* stream); /* dx zero-filled by caller */
/* Dropout (model_parts.py:25, joint_segmentation_depth_decoder.py:45): mask is generated from a
 * counter-based Philox stream (seed, offset) unless mask_in != NULL (replay mode: 0/1 floats).
 * y = x * mask / (1-p); mask_out (uint8, may be NULL) is what backward needs. channelwise=1 is
 * Dropout2d (monodepth_layers.py:119). */
int segsde_dropout_fwd(const segsde_nhwc_t* x, const segsde_nhwc_t* y, float p, uint64_t seed,
                       uint64_t offset, const float* mask_in, uint8_t* mask_out, int channelwise,
                       void* stream);
int segsde_dropout_bwd(const segsde_nhwc_t* dy, const uint8_t* mask, float p, int channelwise,
                       const segsde_nhwc_t* dx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pose head (models/pose_decoder.py:53-58, models/monodepth_layers.py:30-105)
 * ------------------------------------------------------------------------------------------- */
/* vec: [B,6] = (axisangle xyz, translation xyz) -> M: [B,16] row-major 4x4; invert as reference.
 * jac (may be NULL): [B,12,6] d M[:3,:4] / d vec, consumed by the backward entry. */
int segsde_pose_matrix_fwd(const float* vec, int b, int invert, float* M, float* jac, void* stream);
int segsde_pose_matrix_bwd(const float* jac, const float* dM, int b, float* dvec, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Monodepth photometric loss (loss/monodepth_loss.py:64-192, models/monodepth_layers.py:18-27,
 * 145-199, 224-254): bilinear disparity upsampling, disp->depth, backproject, project,
 * grid_sample(border, align_corners=True), SSIM(3x3, reflect)+L1, identity auto-mask with
 * tie-break noise, per-pixel min, mean — ONE fused launch for all scales (the target / source frames,
 * the identity candidates and the target window statistics are shared by the scales); optional unit
 * gradients w.r.t. every scale's disparity map and the two pose matrices in the same pass.
 * ------------------------------------------------------------------------------------------- */
#define SEGSDE_REPROJ_NO_SSIM 1
#define SEGSDE_REPROJ_AVG 2
#define SEGSDE_REPROJ_NO_AUTOMASK 4
#define SEGSDE_REPROJ_MAX_SCALES 4

typedef struct {
  const float* tgt;       /* [B,3,H,W] ("color",0,0) */
  const float* src[2];    /* [B,3,H,W] ("color",f,0), f = frame_ids[1:] */
  const float* K;         /* [B,4,4] ("K",0) */
  const float* inv_K;     /* [B,4,4] ("inv_K",0) */
  const float* T[2];      /* [B,4,4] cam_T_cam per source frame */
  int32_t B, H, W, F, S;  /* S = number of scales (1..SEGSDE_REPROJ_MAX_SCALES) */
  const float* disp[SEGSDE_REPROJ_MAX_SCALES];   /* [B,1,hs,ws] sigmoid disparity of scale s */
  int32_t hs[SEGSDE_REPROJ_MAX_SCALES], ws[SEGSDE_REPROJ_MAX_SCALES];
  const float* noise[SEGSDE_REPROJ_MAX_SCALES];  /* [B,Fi,H,W] tie-break noise already scaled by 1e-5, or NULL = Philox */
  uint64_t seed, offset;  /* Philox stream of scale s = (seed, offset + s) when noise[s] == NULL */
  float min_depth, max_depth;
  int32_t flags;          /* SEGSDE_REPROJ_* */
  /* outputs */
  float* loss_partial;    /* [S][segsde_reproj_num_partials] per-warp sums of the min-loss */
  float* ident_sel[SEGSDE_REPROJ_MAX_SCALES];    /* [B,H,W] 1.0 where a reprojection candidate won, or NULL */
  float* gdisp[SEGSDE_REPROJ_MAX_SCALES];        /* [B,1,hs,ws] d(mean min-loss of scale s)/d disp_s, accumulated
                                                    (zero-filled); all NULL = forward only */
  float* gT_partial;      /* [S][F][B][tiles][12] per-warp d/dP partials (only with gdisp) */
} segsde_reproj_args_t;

int segsde_reproj_num_partials(int B, int H, int W);  /* = B * tiles */
int segsde_reproj_tiles(int H, int W);                /* warps per sample: column strips x row bands */
int segsde_reproj_fused(const segsde_reproj_args_t* a, void* stream);
/* loss[s] = sum(partials[s])/(B*H*W); gT[s][f][b] (4x4) = K[:3,:]^T * sum_tiles gP. Deterministic order.
 * loss_out: [S]; gT: [S][F][B][16]. */
int segsde_reproj_finalize(const float* loss_partial, int n_partial, int64_t count, int S, float* loss_out,
                           const float* gT_partial, const float* K, int B, int tiles, int F,
                           float* gT, void* stream);
/* Optional materialisation of the reference's side outputs (monodepth_loss.py:78-98):
 * depth [B,1,H,W], sample [B,H,W,2], colour [B,3,H,W] for one source frame. Any may be NULL. */
int segsde_reproj_materialize(const float* src, const float* disp, const float* K,
                              const float* inv_K, const float* T, int B, int H, int W, int hs,
                              int ws, float min_depth, float max_depth, float* depth,
                              float* sample, float* color, void* stream);
/* generate_depth_test_pred (monodepth_loss.py:54-62): upsample + disp_to_depth. */
int segsde_disp_to_depth_up(const float* disp, int B, int hs, int ws, int H, int W, float min_depth,
                            float max_depth, float* depth, void* stream);

/* Edge-aware smoothness (monodepth_layers.py:208-221 with the mean normalisation of
 * monodepth_loss.py:182-184). Two launches: (1) per-sample disparity sums, (2) loss + gradient. */
int segsde_smooth_mean(const float* disp, int B, int h, int w, float* mean_out /*[B]*/, void* stream);
/* acc: [2 + B] fp32 zero-filled: acc[0]=mean_x term, acc[1]=mean_y term, acc[2+b]=sum_i ghat_i d_i.
 * ghat (may be NULL): [B,h,w] gradient w.r.t. the normalised disparity. */
int segsde_smooth_fused(const float* disp, const float* img, const float* mean, int B, int h, int w,
                        float* acc, float* ghat, void* stream);
/* gdisp += wgt * (ghat/(m+eps) - dot_b/((m+eps)^2 hw)) ; dot from acc. */
int segsde_smooth_grad_finalize(const float* ghat, const float* mean, const float* acc, int B, int h,
                                int w, float wgt, float* gdisp, void* stream);

/* acc[0], acc[1] of segsde_smooth_fused are already divided by their element counts, so
 * out[s] = reproj[s] + smooth_w_host[s]*(sacc[s*stride]+sacc[s*stride+1]); out[S] = mean_s out[s]
 * (monodepth_loss.py:186-191). smooth_w_host: S HOST floats (disparity_smoothness / 2^s). */
int segsde_mono_combine(const float* reproj, const float* sacc, int S, int sacc_stride,
                        const float* smooth_w_host, float* out, void* stream);
/* y (+)= (ca*a[0] + cb*b[0]) * x, a/b device scalars (upstream gradients; either may be NULL). */
int segsde_scale_by_dev(const float* x, const float* a, float ca, const float* b, float cb, float* y,
                        int accumulate, int64_t n, void* stream);

/* y[i] = a * x[i] (+ y[i] if accumulate) — used to scale stored unit gradients. */
int segsde_axpby(const float* x, float a, float* y, int accumulate, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Segmentation cross-entropy (loss/loss.py:17-37): ignore_index 250, mean over valid pixels,
 * optional per-pixel weights (mean over all pixels).  logits: NHWC view; target int64 [N,H,W].
 * acc: [4] fp32 zero-filled -> acc[0] = sum of (weighted) NLL, acc[1] = number of valid pixels,
 * acc[2] = 1 when pixel_w contains a NaN (the weighting is then skipped, loss.py:31-32),
 * acc[3] = number of labels outside [0,C) that are not ignore_index (they contribute nothing; torch
 * raises a device assert for them).
 * ------------------------------------------------------------------------------------------- */
int segsde_ce_fwd(const segsde_nhwc_t* logits, const int64_t* target, const float* pixel_w,
                  int ignore_index, float* acc, void* stream);
/* dlogits = gscale_dev[0] * (softmax - onehot) * w ; gscale is a device scalar so that no host
 * sync is needed to divide by the valid count.  acc: the forward call's acc (its NaN flag decides
 * whether pixel_w applies); required when pixel_w is given. */
int segsde_ce_bwd(const segsde_nhwc_t* logits, const int64_t* target, const float* pixel_w,
                  int ignore_index, const float* gscale_dev, const float* acc,
                  const segsde_nhwc_t* dlogits, void* stream);

/* Standalone layer forms (models/monodepth_layers.py:145-254), NCHW planar fp32 as in the reference. */
int segsde_backproject(const float* depth, const float* inv_K, int B, int H, int W, float* out /*[B,4,HW]*/,
                       void* stream);
int segsde_project3d(const float* points /*[B,4,HW]*/, const float* K, const float* T, int B, int H, int W,
                     float eps, float* out /*[B,H,W,2]*/, void* stream);
int segsde_ssim_map(const float* x, const float* y, int planes, int H, int W, float* out, void* stream);
int segsde_upsample2x_nearest(const segsde_nhwc_t* x, const segsde_nhwc_t* y, void* stream);

/* out[0] = acc[0]/den, inv[0] = 1/den, inv[1] = 0 with den = const_den > 0 ? const_den : acc[1]
 * (the "mean over valid pixels" division of F.cross_entropy without a host sync). */
int segsde_ratio(const float* acc, float const_den, float* out, float* inv, void* stream);

/* berHu pseudo-depth loss (loss/loss.py:5-15; train.py:494): a = |target - input| * mask (log(1+.) of both first when
 * apply_log), C = threshold * max(a), loss = mean(a <= C ? a : (a^2 + C^2) / (2C)).  The maximum stays on the device
 * (maxbits: its float bit pattern, zero-initialised by the caller; sum: zero-initialised fp64 accumulator) — the
 * reference's .item() round trip is not needed.  mask may be NULL (all ones).  bwd: dinput = d loss / d input * gloss[0]
 * with C held constant, as in the reference (its C is a Python float). */
int segsde_berhu_fwd(const float* input, const float* target, const float* mask, int64_t n, int apply_log,
                     float threshold, unsigned int* maxbits, double* sum, float* loss, void* stream);
int segsde_berhu_bwd(const float* input, const float* target, const float* mask, int64_t n, int apply_log,
                     float threshold, const unsigned int* maxbits, const float* gloss, float* dinput, void* stream);
/* Per-pixel normalised entropy of the softmax over the channel axis of NCHW logits (loss/loss.py:40-47):
 * entropy[n,h,w] = -sum_c p log2(p + 1e-30) / log2(C); normalize: (e - min) / (max - min) over the whole tensor
 * (minmax: two words initialised to {0x7f7fffff, 0} by the caller). */
int segsde_pixel_entropy(const float* logits, int n, int c, int h, int w, int normalize, float* entropy,
                         unsigned int* minmax, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Step-level ops around the model / loss call (Trainer methods in the reference's train.py)
 * ------------------------------------------------------------------------------------------- */
/* T1 feature-distance loss, torch.dist(a, b, p=2) (train.py:480-484): dist = sqrt(sum (a-b)^2); sum: zero-filled fp64
 * scratch.  bwd: da = g[0] * (a - b) / dist, db = -da (either may be NULL); zero when dist == 0. */
int segsde_feature_distance_fwd(const float* a, const float* b, int64_t n, double* sum, float* dist, void* stream);
int segsde_feature_distance_bwd(const float* a, const float* b, int64_t n, const float* dist, const float* g, float* da,
                                float* db, void* stream);
/* T2 DepthMix.  sample_minmax_normalize: out[b] = (d[b] - min_b) / (max_b - min_b) per sample (train.py:688-692);
 * minmax: [B][2] words initialised to {0xffffffff, 0}.  depthcomp_mask (train.py:585-604, pairs (i, (i+1) % B)):
 * mask[i] = [d_i >= d_other - margin] * [d_i >= foreground_threshold] as int64 0/1 (or fp32 0/1 into mask_f32; exactly
 * one of the two outputs is given); threshold_dev (nullable): one threshold per sample on the device (the reference
 * draws `ft` per image, train.py:594-598), overriding the host scalar; compare = 0 drops the comparison with the
 * other sample (mode "depth", train.py:605-615 + loader/transformmasks.py:33-42).  mix (loader/transformsgpu.py:
 * 33-47, the mask.shape[0] == data.shape[0] branch): out[i] = m[i] * x[i] + (1 - m[i]) * x[(i+1) % B], mask [B,H*W]
 * int64 or fp32 broadcast over channels; x / out addressed by element strides (sample, channel, pixel). */
int segsde_sample_minmax_normalize(const float* d, int b, int64_t hw, unsigned int* minmax, float* out, void* stream);
int segsde_depthcomp_mask(const float* d, int b, int64_t hw, float margin, float foreground_threshold,
                          const float* threshold_dev, int compare, int64_t* mask_i64, float* mask_f32, void* stream);
int segsde_mix(const float* x, float* out, const int64_t* mask_i64, const float* mask_f32, int b, int c, int64_t hw,
               int64_t x_sn, int64_t x_sc, int64_t x_sp, int64_t o_sn, int64_t o_sc, int64_t o_sp, void* stream);
/* Teacher softmax over the class axis (train.py:667, `torch.softmax(logits_u_w.detach(), dim=1)`): x / out addressed by
 * (sample, channel, pixel) element strides, so planar and channels-last tensors both work without a copy. */
int segsde_softmax_channels(const float* x, float* out, int b, int c, int64_t hw, int64_t x_sn, int64_t x_sc, int64_t x_sp,
                            int64_t o_sn, int64_t o_sc, int64_t o_sp, void* stream);
/* T3 pseudo labels (train.py:644-648): label = argmax_c prob (first maximum), ignore_index where the maximum is 0;
 * count (zero-filled) += #pixels with max >= threshold; pixel_weight (nullable) = weight_scale * count / (B*H*W) at
 * every pixel — the confidence weight stays on the device (the reference reads it with .item()).  prob addressed by
 * element strides (sample, channel, pixel); max_prob nullable. */
int segsde_pseudo_label(const float* prob, int b, int c, int64_t hw, int64_t sn, int64_t sc, int64_t sp, float threshold,
                        int64_t ignore_index, int64_t* label, float* max_prob, unsigned long long* count,
                        float weight_scale, float* pixel_weight, void* stream);
/* T4 EMA teacher update (train.py:346-358) as a multi-tensor kernel: dst[i] = alpha * dst[i] + beta * src[i] for a
 * list of dense fp32 tensors (HOST arrays of device pointers / element counts); one launch per 48 tensors / 320 chunks
 * instead of one per parameter. */
int segsde_multi_axpby(int ntensors, float* const* dst, const float* const* src, const int64_t* numel, float alpha,
                       float beta, void* stream);
/* Optimizer step and gradient clipping as multi-tensor kernels.  The reference builds
 * torch.optim.Adam / SGD through utils/optimizers.py:7-30 (train.py:291-295) and clips with
 * torch.nn.utils.clip_grad_norm_ (train.py:516-524); same arithmetic as torch's single-tensor implementations.
 * All pointer lists are HOST arrays of device pointers to dense fp32 tensors of numel[i] elements.
 *   adam: g += wd*p; m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 *   sgd : g += wd*p; buf = first_step ? g : mom*buf + (1-damp)*g; g = nesterov ? g + mom*buf : buf; p -= lr*g
 *         (momentum_buf may be NULL when momentum == 0)
 *   clip: total_norm = sqrt(sum g^2) over all tensors -> total_norm[0]; coef[0] = min(1, max_norm/(total_norm+1e-6));
 *         g *= coef (sum: zero-filled fp64 scratch) */
int segsde_multi_adam(int n, float* const* p, float* const* g, float* const* exp_avg, float* const* exp_avg_sq,
                      const int64_t* numel, float lr, float beta1, float beta2, float eps, float weight_decay,
                      int64_t step, void* stream);
int segsde_multi_sgd(int n, float* const* p, float* const* g, float* const* momentum_buf, const int64_t* numel, float lr,
                     float momentum, float dampening, float weight_decay, int nesterov, int first_step, void* stream);
int segsde_multi_clip_grad_norm(int n, float* const* g, const int64_t* numel, float max_norm, double* sum,
                                float* total_norm, float* coef, void* stream);
/* Mixed-precision loss scaling (torch.cuda.amp.GradScaler as used at train.py:486-530).  multi_unscale: g *= 1 / scale[0]
 * in place, found_inf[0] = 1.0 when any gradient element is inf / NaN (found_inf zero-filled by the caller; device
 * scalars, no host round trip).  amp_update_scale: scale *= backoff on overflow, *= growth after growth_interval clean
 * steps (growth_tracker: device int). */
int segsde_multi_unscale(int n, float* const* g, const int64_t* numel, const float* scale, float* found_inf,
                         void* stream);
int segsde_amp_update_scale(float* scale, int* growth_tracker, const float* found_inf, float growth_factor,
                            float backoff_factor, int growth_interval, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Validation / inference path and label-selection scoring (rank 4)
 * ------------------------------------------------------------------------------------------- */
/* hist[gt][pred] += 1 over all pixels with 0 <= gt < n_classes (evaluation/metrics.py:12-25, train.py:846-850): pred is
 * the first arg-max of the logits (addressed by sample / channel / pixel element strides) or, when `pred` is given,
 * taken from it.  hist: [n_classes][n_classes] int64, accumulated. n_classes <= 32. */
int segsde_confusion_update(const float* logits, const int64_t* pred, const int64_t* gt, int b, int c, int64_t hw, int64_t sn,
                            int64_t sc, int64_t sp, int n_classes, int64_t* hist, void* stream);
/* Eval-mode BatchNorm folded into the preceding convolution: w_out[o][k] = w[o][k] * gamma[o] / sqrt(var[o] + eps),
 * b_out[o] = beta[o] + (conv_bias[o] - mean[o]) * gamma[o] / sqrt(var[o] + eps); w: [cout][k] (OHWI flattened). */
int segsde_bn_fold(const float* w, const float* conv_bias, const float* gamma, const float* beta, const float* mean,
                   const float* var, float eps, int cout, int k, float* w_out, float* b_out, void* stream);
/* F.adaptive_avg_pool2d / adaptive_max_pool2d on NCHW planar data (label_selection.py:398-426). */
int segsde_adaptive_pool(const float* x, int nc, int h, int w, int oh, int ow, int is_max, float* y, void* stream);
/* torch.cdist(f, f, p) for f: [n][d] (label_selection.py:603-606). */
int segsde_pairwise_distance(const float* f, int n, int64_t d, float p, float* out, void* stream);
/* iterative_farthest_point (label_selection.py:617-640) in one launch: repeatedly adds the sample with the largest
 * distance to the selected set.  is_current: [n] 0/1, updated; new_idx / new_dist: [n_new]; count[0] = samples added;
 * scratch: [n] floats. */
int segsde_farthest_point(const float* dist, int n, int* is_current, int n_new, int64_t* new_idx, float* new_dist, int* count,
                          float* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * On-device input pipeline pieces, NCHW planar fp32
 * ------------------------------------------------------------------------------------------- */
/* kornia.filters.GaussianBlur2d((ky,kx), sigma) with reflect border (loader/transformsgpu.py:21-30) as two separable
 * passes; taps_y / taps_x: normalised 1-D Gaussians on the device, ky / kx odd <= 255; tmp: scratch like x. */
int segsde_gaussian_blur(const float* x, float* tmp, float* y, int planes, int h, int w, int ky, int kx, const float* taps_y,
                         const float* taps_x, void* stream);
/* The four kornia ColorJitter primitives (loader/transformsgpu.py:10-18) fused: brightness (additive), contrast
 * (multiplicative), saturation (HSV S scale), hue (HSV H shift, radians), each clamped to [0,1], applied in
 * order4 (a permutation of 0=b,1=c,2=s,3=h).  x / y: [b][3][hw]. */
int segsde_color_jitter(const float* x, float* y, int b, int64_t hw, float brightness, float contrast, float saturation,
                        float hue, const int* order4, void* stream);
/* Scales 1..3 of the loader's image pyramid as exact 2^s x 2^s box means (F.interpolate(mode="area")), one launch;
 * h, w multiples of 8; any output may be NULL. */
int segsde_area_pyramid(const float* x, int planes, int h, int w, float* y1, float* y2, float* y3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGSDE_B200_H */
