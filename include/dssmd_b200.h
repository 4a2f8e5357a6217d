/*
 * dssmd_b200.h -- C ABI of libdssmd_b200.so: the DSSMD correlation cost volume and
 * disparity warp as hand-written sm_100a CUDA kernels.
 *
 * The upstream project (Magicboomliu/DSSMD) is pure Python/PyTorch and has no FFI
 * of its own; each entry point below replaces one upstream Python function (or
 * the autograd backward PyTorch derives for it) and cites it.  Paths are
 * relative to the upstream tree.  dssmd_b200/_lib.py is the ctypes binding;
 * INTEGRATION.md shows the stub an upstream maintainer would add.
 *
 * Conventions
 *  - All tensors are dense, contiguous NCHW device buffers owned by the caller.
 *    The library never allocates, frees or retains a pointer after return.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Every call only enqueues work on that stream; nothing synchronises.
 *  - Re-entrant and thread-safe (nn.DataParallel calls from one host thread per
 *    GPU, upstream trainfiles/dssmd_traner_new.py:131-135): no global mutable
 *    state other than the thread-local last-error string.  The caller makes the
 *    buffers' device current (cudaSetDevice) before the call.
 *  - Return value: 0 on success, a DSSMD_ERR_* code otherwise;
 *    dssmd_last_error() then describes the failure.  Never aborts or throws.
 */
#ifndef DSSMD_B200_H_
#define DSSMD_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DSSMD_ABI_VERSION 1

#if defined(__GNUC__)
#define DSSMD_API __attribute__((visibility("default")))
#else
#define DSSMD_API
#endif

/* element types (`dtype` arguments) */
#define DSSMD_F32 0
#define DSSMD_BF16 1
#define DSSMD_F16 2
#define DSSMD_F64 3

/* grid_sample padding modes (`padding_mode` arguments) */
#define DSSMD_PAD_ZEROS 0
#define DSSMD_PAD_BORDER 1

/* error codes */
#define DSSMD_OK 0
#define DSSMD_ERR_INVALID_ARGUMENT 1 /* null pointer, non-positive size, unknown enum */
#define DSSMD_ERR_UNSUPPORTED 2      /* dtype/shape combination not implemented */
#define DSSMD_ERR_CUDA 3             /* a CUDA runtime call or launch failed */

DSSMD_API int dssmd_version(void);
/* Thread-local, valid until the next failing call on the same thread. */
DSSMD_API const char *dssmd_last_error(void);

/* ---- correlation cost volume ------------------------------------------------
 * Replaces build_corr(img_left, img_right, max_disp),
 * models/UtilsNet/submoudles.py:301-311 (and the 'correlation' branch of
 * CostVolume.forward, models/UtilsNet/cost_volume.py:39-47):
 *   out[b,d,y,x] = (1/C) * sum_c L[b,c,y,x] * R[b,c,y,x-d]   for x >= d, else 0
 * L, R: [B,C,H,W]; out: [B,D,H,W] (every element is written, no pre-zeroing
 * needed); D = max_disp channels d = 0..D-1.  dtype F32/BF16/F16 accumulate in
 * fp32, F64 in fp64. */
DSSMD_API int dssmd_corr_fwd(const void *L, const void *R, void *out,
                   int B, int C, int H, int W, int D, int dtype, void *stream);

/* Backward of the above (the graph autograd builds from submoudles.py:304-308):
 *   gL[b,c,y,x ] = (1/C) * sum_{d<=min(D-1,x)}      g[b,d,y,x]    * R[b,c,y,x-d]
 *   gR[b,c,y,x'] = (1/C) * sum_{d<=min(D-1,W-1-x')} g[b,d,y,x'+d] * L[b,c,y,x'+d]
 * g: [B,D,H,W]; gL, gR: [B,C,H,W], fully overwritten.  Either of gL / gR may be
 * NULL to skip that gradient. */
DSSMD_API int dssmd_corr_bwd(const void *g, const void *L, const void *R, void *gL, void *gR,
                   int B, int C, int H, int W, int D, int dtype, void *stream);

/* Correlation written straight into / read straight out of the concat buffer of its consumer.
 * Upstream: in_conv3b = torch.cat((conv_redir(conv3_l), build_corr(conv3_l, conv3_r, max_disp // 8)), dim=1),
 * models/DSSMD.py:129-131 (same pattern models/DSSMD_new.py, SSMDNet.py, SDNet.py).
 * cat / g_cat: [B, cat_channels, H, W]; the D correlation channels are channels
 * [channel_offset, channel_offset + D) of it; every other channel is left untouched (fwd) / ignored (bwd).
 * Same arithmetic, kernels and results as dssmd_corr_fwd / dssmd_corr_bwd. */
DSSMD_API int dssmd_corr_fwd_cat(const void *L, const void *R, void *cat, int B, int C, int H, int W, int D,
                       int cat_channels, int channel_offset, int dtype, void *stream);
DSSMD_API int dssmd_corr_bwd_cat(const void *g_cat, const void *L, const void *R, void *gL, void *gR,
                       int B, int C, int H, int W, int D, int cat_channels, int channel_offset, int dtype, void *stream);

/* ---- disparity warp -----------------------------------------------------------
 * Replaces disp_warp(img, disp, padding_mode), utils/disparity_warper.py:83-106
 * (byte-identical copy at models/UtilsNet/stereo_op.py:41-63), including its
 * helpers meshgrid (:60-80) and normalize_coords (:48-58) and the two
 * F.grid_sample calls (:100, :103; bilinear, align_corners=False):
 *   warped[b,c,y,x] = bilinear(img[b,c], ix, iy),  mask = 1[sum of in-bounds
 *   bilinear weights at the un-clamped (ix,iy) >= 0.9999]
 * img, warped, mask: [B,C,H,W]; disp: [B,1,H,W].  `mask` may be NULL (skipped).
 * `neg_flag` (one int, may be NULL) is set to 1 with a plain store if any disparity
 * is negative or NaN -- the condition of upstream's `assert disp.min() >= 0` (:93).
 * It may live in device memory or in pinned (device-accessible) host memory; the
 * caller zeroes it beforehand and raises after the launch has completed.  A pinned
 * flag needs only an event wait on `stream`, no device->host copy.
 * dtype: F32 or F64. */
DSSMD_API int dssmd_warp_fwd(const void *img, const void *disp, void *warped, void *mask, int *neg_flag,
                   int B, int C, int H, int W, int padding_mode, int dtype, void *stream);

/* Backward of disp_warp (aten::grid_sampler_2d_backward composed with the
 * coordinate arithmetic of :95-99); the mask carries no gradient.
 * g: [B,C,H,W] gradient of `warped`; g_img: [B,C,H,W]; g_disp: [B,1,H,W]; both
 * fully overwritten, either may be NULL to skip it.
 * Image-like fp32 shapes (C <= 4, W % 4 == 0, 4 <= W <= 2048, H >= 2, every pointer
 * 16-byte aligned) take a one-pass kernel that needs NO workspace: pass NULL / 0.
 * `workspace` is an optional caller-owned device scratch buffer of at least
 * dssmd_warp_bwd_workspace_bytes(...) bytes (contents undefined afterwards) for the
 * other fp32 shapes (wider rows): with it a two-kernel path runs (horizontal scatter
 * into the workspace, vertical mix), without it a slower generic kernel.
 * Accuracy of g_img: per row a 2^-30 fixed-point quantum of the row's sum of |g|
 * (deterministic); rows with an outlier, unbalanced channels or non-finite values
 * take exact / fp32 paths. */
#include <stddef.h>
DSSMD_API size_t dssmd_warp_bwd_workspace_bytes(int B, int C, int H, int W, int dtype);
DSSMD_API int dssmd_warp_bwd(const void *g, const void *img, const void *disp, void *g_img, void *g_disp,
                   void *workspace, size_t workspace_bytes,
                   int B, int C, int H, int W, int padding_mode, int dtype, void *stream);

/* ---- bilinear resize (LRwarp_error's pre-scaling) -------------------------------
 * Replaces F.interpolate(img, size=(Ho,Wo), mode='bilinear', align_corners=False)
 * as called by LRwarp_error, utils/disparity_warper.py:110-113.
 * in: [B,C,Hi,Wi]; out: [B,C,Ho,Wo].  dtype: F32 or F64. */
DSSMD_API int dssmd_resize_bilinear_fwd(const void *in, void *out, int B, int C, int Hi, int Wi, int Ho, int Wo,
                              int dtype, void *stream);
/* g_out: [B,C,Ho,Wo]; g_in: [B,C,Hi,Wi], fully overwritten. */
DSSMD_API int dssmd_resize_bilinear_bwd(const void *g_out, void *g_in, int B, int C, int Hi, int Wi, int Ho, int Wo,
                              int dtype, void *stream);

/* ---- fused left/right warp residual ----------------------------------------------
 * Replaces LRwarp_error(imgL, disp, imgR), utils/disparity_warper.py:109-115
 * (== models/UtilsNet/stereo_op.py:67-73) in one launch: both images are
 * bilinearly resized to disp's [H,W] when wider (W0 > W), imgL is warped by disp
 * with 'border' padding, and  out = imgR' - warped  is written.
 * imgL, imgR: [B,C,H0,W0]; disp: [B,1,H,W]; out: [B,C,H,W].  dtype: F32 or F64. */
DSSMD_API int dssmd_lrwarp_error_fwd(const void *imgL, const void *disp, const void *imgR, void *out, int *neg_flag,
                           int B, int C, int H0, int W0, int H, int W, int dtype, void *stream);

/* ---- fused L1 photometric term on the warp residual --------------------------------
 * The primitive of a photometric / reconstruction loss.  Upstream ships
 * only LRwarp_error (utils/disparity_warper.py:109-115) and no loss that calls it; this entry point
 * computes, for images already at disp's resolution,
 *   err = imgR - warp(imgL, disp)                     ('border' padding, as LRwarp_error)
 *   row_sums[b*H + y] = sum_{c,x} |err[b,c,y,x]|      (fp32, fixed summation order)
 *   g_disp_unit[b,0,y,x] = d(sum |err|) / d(disp[b,0,y,x])        (may be NULL)
 * in one pass and without writing err: sum(row_sums) / (B*C*H*W) is
 * LRwarp_error(imgL, disp, imgR).abs().mean(), and g_disp_unit times the upstream scalar / (B*C*H*W)
 * is its gradient w.r.t. disp (what autograd derives through aten::abs, the subtraction and
 * grid_sampler_2d_backward).  The images carry no gradient.  `neg_flag` as in dssmd_warp_fwd.
 * imgL, imgR: [B,C,H,W]; disp, g_disp_unit: [B,1,H,W]; row_sums: [B*H].  dtype: F32;
 * DSSMD_ERR_UNSUPPORTED for shapes outside C <= 4, W % 4 == 0, H >= 2, 16-byte aligned tensors. */
DSSMD_API int dssmd_photometric_l1_fwd(const void *imgL, const void *imgR, const void *disp, void *g_disp_unit, void *row_sums,
                             int *neg_flag, int B, int C, int H, int W, int dtype, void *stream);

/* ---- haze synthesis (input side of the training step) ----------------------------
 * One launch for the elementwise chain upstream runs per batch (trainfiles/dssmd_traner_new.py:250-256):
 *   convert_disp_to_depth  DeBug/inference.py:57-64    depth = (focal_length * baseline) / disp
 *   depth2trans            DeBug/inference.py:126-134  trans = exp((-depth) * beta)
 *   recover_haze_images    DeBug/inference.py:95-112   haze  = (clear * trans) + airlight * (1 - trans)
 * src: [B,1,H,W], a disparity map (src_is_disparity != 0) or a depth map (then baseline / focal_length are unused);
 * clear, haze: [B,C,H,W]; trans, depth: [B,1,H,W]; baseline, focal_length, beta, airlight: [B] device arrays of the
 * compute dtype.  Any of haze / trans / depth may be NULL (not wanted).  clear and src have image_dtype, the outputs
 * compute_dtype -- upstream's type promotion: (F32, F32), (F32, F64) (float32 images with the float64 scalars the data
 * loaders produce), (F64, F64).  Every operation is rounded individually, in upstream's order. */
DSSMD_API int dssmd_haze_fwd(const void *clear, const void *src, int src_is_disparity, const void *baseline, const void *focal_length,
                   const void *beta, const void *airlight, void *haze, void *trans, void *depth,
                   int B, int C, int H, int W, int image_dtype, int compute_dtype, void *stream);

/* ---- ASM reconstruction loss ------------------------------------------------------
 * Replaces RecoveredCleanImagesLoss.forward(transmission_map, airlight, haze_image, clean_image),
 * losses/DSSMDLoss.py:81-111 (V2 :115-145), called per level of the transmission pyramid by
 * trainfiles/dssmd_traner_new.py:298-302, together with its autograd backward:
 *   rec = (haze_s - A * (1 - t)) / t;   loss = mean | clamp(rec, 0, 1) - clean_s |
 * with haze_s / clean_s the bilinear (align_corners=False) downsample of the images to the size of t when they are
 * W0 / w times larger (integer factor, H0 == h * W0 / w).
 * trans: [B,1,h,w]; airlight: [B]; haze, clean: [B,C,H0,W0]; g_trans_unit: [B,1,h,w] = d(sum |diff|)/d(trans), may be
 * NULL; partial: [B, nblocks, 2] floats with nblocks = dssmd_asm_recovered_l1_blocks(B, h, w): per block the sum of
 * |diff| and of d|diff|/d(airlight).  loss = sum(partial[..., 0]) / (B*C*h*w); d loss / d airlight[b] =
 * sum(partial[b, :, 1]) / (B*C*h*w).  One launch, deterministic. */
DSSMD_API int dssmd_asm_recovered_l1_blocks(int B, int h, int w);
DSSMD_API int dssmd_asm_recovered_l1_fwd(const void *trans, const void *airlight, const void *haze, const void *clean,
                               void *g_trans_unit, void *partial, int B, int C, int H0, int W0, int h, int w,
                               int dtype, void *stream);
/* The same launch followed by the reduction of the partials on the device (one small block, fixed order), so that the
 * caller needs no further arithmetic: out: 1 + B floats, out[0] = the loss (mean |diff| over B*C*h*w elements, what
 * upstream's forward returns), out[1 + b] = sum(partial[b, :, 1]) (d(sum |diff|)/d airlight[b]; scale by 1/(B*C*h*w)). */
DSSMD_API int dssmd_asm_recovered_l1_loss(const void *trans, const void *airlight, const void *haze, const void *clean,
                               void *g_trans_unit, void *partial, void *out, int B, int C, int H0, int W0, int h, int w,
                               int dtype, void *stream);

/* ---- attention along the epipolar line ---------------------------------------------
 * Replaces the core of STTM_Single.forward, models/BidNet/STTM_Simple.py:43-57 (q, k, v are the outputs of its three
 * 1x1 convolutions; used at full resolution by models/BidNet/BidNet.py:115, 239-240):
 *   out[b,c,y,x] = sum_x' softmax_x'( scale * sum_c' q[b,c',y,x] k[b,c',y,x'] ) * v[b,c,y,x']
 * q, k, v, out: [B,C,H,W] NCHW fp32 (no permutes, the [B*H, W, W] attention matrix is never written);
 * lse: [B*H, W] row log-sum-exp of the scaled scores, saved for the backward (may be NULL for inference).
 * C in {16, 32, 64, 128}, W % 4 == 0, B*H <= 65535, 16-byte aligned tensors; anything else: DSSMD_ERR_UNSUPPORTED. */
DSSMD_API int dssmd_epipolar_attn_fwd(const void *q, const void *k, const void *v, void *out, void *lse,
                            int B, int C, int H, int W, float scale, int dtype, void *stream);
/* Backward of the above: g_q, g_k, g_v: [B,C,H,W] (g_q may be NULL, g_k / g_v only together); delta: scratch of
 * B*H*W + 4*B*H floats ([B*H, W] sums sum_c g_out * out, then one power-of-two scale per image row for q, k, v and
 * g_out; written by the first of the three launches).  Deterministic, no atomics. */
DSSMD_API int dssmd_epipolar_attn_bwd(const void *q, const void *k, const void *v, const void *out, const void *g_out,
                            const void *lse, void *delta, void *g_q, void *g_k, void *g_v,
                            int B, int C, int H, int W, float scale, int dtype, void *stream);

/* ---- self test ------------------------------------------------------------------------
 * The lean warp kernels divide pixel coordinates by W - 1 with a reciprocal hoisted out of the pixel loop
 * (csrc/warp_lean.cuh: div_rn_by) and rely on that being IEEE round-to-nearest division, bit for bit -- the
 * validity mask is compared exactly with upstream's.  This entry point checks it on the device against
 * __fdiv_rn for every divisor W - 1 with W in [W_lo, W_hi] and every numerator k / 64, k in [-16384, 64 * W_hi]
 * (a superset of the x - disp values the kernels see at 1/64-pixel resolution), plus the same numerators
 * perturbed by one ulp either way.  *mismatches (device or pinned-host memory, 8 bytes, zeroed by the
 * caller) receives the number of differing quotients. */
DSSMD_API int dssmd_selftest_division(int W_lo, int W_hi, unsigned long long *mismatches, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DSSMD_B200_H_ */
