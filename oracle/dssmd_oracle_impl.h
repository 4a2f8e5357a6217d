/*
 * dssmd_oracle_impl.h -- type-generic body of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Included twice by dssmd_oracle.c, once with
 * REAL=float / SUF=f32 and once with REAL=double / SUF=f64.  Nothing in the
 * product package (dssmd_b200/) may include, link or call this.
 *
 * Every function restates one reference function; the reference file:line it
 * follows is cited above it (paths relative to the upstream DSSMD tree).
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* ------------------------------------------------------------------------
 * Correlation forward.
 * Restates build_corr, models/UtilsNet/submoudles.py:301-311 (same math in
 * CostVolume 'correlation', models/UtilsNet/cost_volume.py:39-47):
 *     out = zeros[B, D, H, W]
 *     out[:, d, :, d:] = mean_c( L[:, :, :, d:] * R[:, :, :, :W-d] )
 * mean = (sum over c, in channel order) / C, accumulated in REAL.
 * Channels d >= W stay all-zero (the reference's slices are empty there).
 * ---------------------------------------------------------------------- */
void FN(oracle_corr_fwd)(const REAL *L, const REAL *R, REAL *out,
                         int B, int C, int H, int W, int D)
{
    const long HW = (long)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y) {
            const REAL *Lr = L + (long)b * C * HW + (long)y * W;
            const REAL *Rr = R + (long)b * C * HW + (long)y * W;
            REAL *o = out + (long)b * D * HW + (long)y * W;
            for (int d = 0; d < D; ++d) {
                REAL *od = o + (long)d * HW;
                for (int x = 0; x < W; ++x) od[x] = (REAL)0;
                for (int c = 0; c < C; ++c) {
                    const REAL *lc = Lr + (long)c * HW;
                    const REAL *rc = Rr + (long)c * HW;
                    for (int x = d; x < W; ++x) od[x] += lc[x] * rc[x - d];
                }
                for (int x = d; x < W; ++x) od[x] = od[x] / (REAL)C;
            }
        }
}

/* ------------------------------------------------------------------------
 * Correlation backward (what autograd derives from submoudles.py:304-308:
 * CopySlices <- MeanBackward1 <- MulBackward0 <- 2x SliceBackward0).
 *     gL[b,c,y,x ] = (1/C) * sum_{d=0}^{min(D-1,x)}      g[b,d,y,x]    * R[b,c,y,x-d]
 *     gR[b,c,y,x'] = (1/C) * sum_{d=0}^{min(D-1,W-1-x')} g[b,d,y,x'+d] * L[b,c,y,x'+d]
 * MeanBackward divides the incoming gradient by C first (g/C), then the
 * MulBackward multiplies by the other operand and the per-d contributions are
 * summed in d order by the engine; this follows that order.
 * ---------------------------------------------------------------------- */
void FN(oracle_corr_bwd)(const REAL *g, const REAL *L, const REAL *R,
                         REAL *gL, REAL *gR,
                         int B, int C, int H, int W, int D)
{
    const long HW = (long)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y) {
            const REAL *gr = g + (long)b * D * HW + (long)y * W;
            for (int c = 0; c < C; ++c) {
                const long off = ((long)b * C + c) * HW + (long)y * W;
                const REAL *lc = L + off, *rc = R + off;
                REAL *glc = gL + off, *grc = gR + off;
                for (int x = 0; x < W; ++x) {
                    REAL a = (REAL)0, r = (REAL)0;
                    for (int d = 0; d < D && d <= x; ++d)
                        a += (gr[(long)d * HW + x] / (REAL)C) * rc[x - d];
                    for (int d = 0; d < D && x + d < W; ++d)
                        r += (gr[(long)d * HW + x + d] / (REAL)C) * lc[x + d];
                    glc[x] = a;
                    grc[x] = r;
                }
            }
        }
}

/* ------------------------------------------------------------------------
 * Sampling coordinate of the warp, in the reference's exact op order.
 * utils/disparity_warper.py:95-99 (meshgrid + offset + normalize_coords
 * :48-58) followed by F.grid_sample's un-normalisation for
 * align_corners=False (the call at :100/:103 passes no align_corners;
 * ATen/native/GridSampler.h:34 `((coord + 1) * size - 1) / 2`, which the
 * compiled kernels evaluate as one FMA  (coord+1) * (size/2) - 0.5):
 *     t  = x - disp
 *     gx = 2 * (t / (W-1)) - 1
 *     ix = fma(gx + 1, W/2, -0.5)
 * ---------------------------------------------------------------------- */
static inline REAL FN(oracle_unnorm)(REAL pix, int size)
{
    REAL g = (REAL)2 * (pix / (REAL)(size - 1)) - (REAL)1;
    return FMA(g + (REAL)1, (REAL)size / (REAL)2, (REAL)-0.5);
}

/* ------------------------------------------------------------------------
 * Warp forward.  Restates disp_warp, utils/disparity_warper.py:83-106
 * (byte-identical copy at models/UtilsNet/stereo_op.py:41-63).
 *   padding_mode: 0 = 'zeros', 1 = 'border'   (grid_sample, bilinear).
 *   warped : [B,C,H,W] bilinear sample of img at (ix, iy)
 *   mask   : [B,C,H,W] 1 where the zero-padded bilinear sample of an all-ones
 *            image is >= 0.9999 (compared in REAL), else 0    (:102-105)
 * Returns 1 when the reference's `assert disp.min() >= 0` (:93) would fire
 * (any negative or NaN disparity), 0 otherwise; outputs are still written.
 * ---------------------------------------------------------------------- */
int FN(oracle_warp_fwd)(const REAL *img, const REAL *disp, REAL *warped, REAL *mask,
                        int B, int C, int H, int W, int padding_mode)
{
    const long HW = (long)H * W;
    int bad = 0;
#pragma omp parallel for collapse(2) schedule(static) reduction(| : bad)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y) {
            const REAL iy_raw = FN(oracle_unnorm)((REAL)y, H);
            for (int x = 0; x < W; ++x) {
                const REAL dv = disp[(long)b * HW + (long)y * W + x];
                if (!(dv >= (REAL)0)) bad |= 1;
                const REAL ix_raw = FN(oracle_unnorm)((REAL)x - dv, W);

                /* ---- mask: zeros padding on the un-clamped coordinate ---- */
                {
                    REAL x0 = FLOOR(ix_raw), y0 = FLOOR(iy_raw);
                    REAL w = ix_raw - x0, e = (REAL)1 - w;
                    REAL n = iy_raw - y0, s = (REAL)1 - n;
                    long xi = (long)x0, yi = (long)y0;
                    int inx0 = xi >= 0 && xi < W, inx1 = xi + 1 >= 0 && xi + 1 < W;
                    int iny0 = yi >= 0 && yi < H, iny1 = yi + 1 >= 0 && yi + 1 < H;
                    REAL m = (REAL)0;
                    if (inx0 && iny0) m += s * e;
                    if (inx1 && iny0) m += s * w;
                    if (inx0 && iny1) m += n * e;
                    if (inx1 && iny1) m += n * w;
                    REAL mv = (m >= (REAL)0.9999) ? (REAL)1 : (REAL)0;
                    if (!(ix_raw == ix_raw)) mv = (REAL)0; /* NaN coordinate */
                    for (int c = 0; c < C; ++c)
                        mask[((long)b * C + c) * HW + (long)y * W + x] = mv;
                }

                /* ---- image sample ---- */
                REAL ix = ix_raw, iy = iy_raw;
                if (padding_mode == 1) {
                    ix = FMIN((REAL)(W - 1), FMAX(ix, (REAL)0));
                    iy = FMIN((REAL)(H - 1), FMAX(iy, (REAL)0));
                }
                REAL x0 = FLOOR(ix), y0 = FLOOR(iy);
                REAL w = ix - x0, e = (REAL)1 - w;
                REAL n = iy - y0, s = (REAL)1 - n;
                long xi = (long)x0, yi = (long)y0;
                int inx0 = xi >= 0 && xi < W, inx1 = xi + 1 >= 0 && xi + 1 < W;
                int iny0 = yi >= 0 && yi < H, iny1 = yi + 1 >= 0 && yi + 1 < H;
                for (int c = 0; c < C; ++c) {
                    const REAL *p = img + ((long)b * C + c) * HW;
                    REAL v = (REAL)0;
                    if (inx0 && iny0) v += p[yi * W + xi] * (s * e);
                    if (inx1 && iny0) v += p[yi * W + xi + 1] * (s * w);
                    if (inx0 && iny1) v += p[(yi + 1) * W + xi] * (n * e);
                    if (inx1 && iny1) v += p[(yi + 1) * W + xi + 1] * (n * w);
                    warped[((long)b * C + c) * HW + (long)y * W + x] = v;
                }
            }
        }
    return bad;
}

/* ------------------------------------------------------------------------
 * Warp backward: what autograd derives for disp_warp
 * (aten::grid_sampler_2d_backward; clip_coordinates_set_grad,
 * ATen/native/GridSampler.h:66-83).  The mask carries no gradient.
 *   g_img : scatter-add of g * tap-weight into the in-bounds taps
 *   g_disp: -(2/(W-1)) * (W/2) * clipmult * sum_c g * d(sample)/d(ix)
 * g_img must hold B*C*H*W elements, g_disp B*H*W; both are overwritten.
 * Rows are processed sequentially per (b) so the scatter needs no atomics.
 * ---------------------------------------------------------------------- */
void FN(oracle_warp_bwd)(const REAL *g, const REAL *img, const REAL *disp,
                         REAL *g_img, REAL *g_disp,
                         int B, int C, int H, int W, int padding_mode)
{
    const long HW = (long)H * W;
    for (long i = 0; i < (long)B * C * HW; ++i) g_img[i] = (REAL)0;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y) {
            const REAL iy_raw = FN(oracle_unnorm)((REAL)y, H);
            for (int x = 0; x < W; ++x) {
                const REAL dv = disp[(long)b * HW + (long)y * W + x];
                REAL ix = FN(oracle_unnorm)((REAL)x - dv, W);
                REAL iy = iy_raw;
                REAL gmult = (REAL)W / (REAL)2; /* d ix / d gx */
                if (padding_mode == 1) {
                    if (ix <= (REAL)0) { ix = (REAL)0; gmult = (REAL)0; }
                    else if (ix >= (REAL)(W - 1)) { ix = (REAL)(W - 1); gmult = (REAL)0; }
                    iy = FMIN((REAL)(H - 1), FMAX(iy, (REAL)0));
                }
                REAL x0 = FLOOR(ix), y0 = FLOOR(iy);
                REAL w = ix - x0, e = (REAL)1 - w;
                REAL n = iy - y0, s = (REAL)1 - n;
                long xi = (long)x0, yi = (long)y0;
                int inx0 = xi >= 0 && xi < W, inx1 = xi + 1 >= 0 && xi + 1 < W;
                int iny0 = yi >= 0 && yi < H, iny1 = yi + 1 >= 0 && yi + 1 < H;
                REAL gix = (REAL)0;
                for (int c = 0; c < C; ++c) {
                    const long pl = ((long)b * C + c) * HW;
                    const REAL go = g[pl + (long)y * W + x];
                    const REAL *p = img + pl;
                    REAL *q = g_img + pl;
                    if (inx0 && iny0) { q[yi * W + xi] += go * (s * e);           gix -= p[yi * W + xi] * s * go; }
                    if (inx1 && iny0) { q[yi * W + xi + 1] += go * (s * w);       gix += p[yi * W + xi + 1] * s * go; }
                    if (inx0 && iny1) { q[(yi + 1) * W + xi] += go * (n * e);     gix -= p[(yi + 1) * W + xi] * n * go; }
                    if (inx1 && iny1) { q[(yi + 1) * W + xi + 1] += go * (n * w); gix += p[(yi + 1) * W + xi + 1] * n * go; }
                }
                /* chain: ix <- gx (gmult) <- t (2/(W-1)) <- disp (-1) */
                g_disp[(long)b * HW + (long)y * W + x] =
                    -(((REAL)2 / (REAL)(W - 1)) * (gmult * gix));
            }
        }
}

/* ------------------------------------------------------------------------
 * Bilinear resize, align_corners=False, as used by LRwarp_error
 * (utils/disparity_warper.py:110-113: F.interpolate(img, size=disp.shape[-2:],
 * mode='bilinear', align_corners=False)).  Index math follows
 * ATen/native/UpSample.h:289-313 (area_pixel_compute_source_index) and
 * :455-475 (compute_source_index_and_lambda): scale = in/out (REAL),
 * src = max(scale*(dst+0.5)-0.5, 0), i0 = floor(src) clamped to in-1,
 * i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1.
 * ---------------------------------------------------------------------- */
static inline void FN(oracle_src_index)(int dst, int in_size, int out_size,
                                        int *i0, int *i1, REAL *l0, REAL *l1)
{
    if (in_size == out_size) { *i0 = *i1 = dst; *l0 = (REAL)1; *l1 = (REAL)0; return; }
    const REAL scale = (REAL)in_size / (REAL)out_size;
    REAL src = scale * ((REAL)dst + (REAL)0.5) - (REAL)0.5;
    if (src < (REAL)0) src = (REAL)0;
    int k = (int)src;
    if (k > in_size - 1) k = in_size - 1;
    REAL lam = src - (REAL)k;
    if (lam < (REAL)0) lam = (REAL)0;
    if (lam > (REAL)1) lam = (REAL)1;
    *i0 = k;
    *i1 = k + ((k < in_size - 1) ? 1 : 0);
    *l1 = lam;
    *l0 = (REAL)1 - lam;
}

void FN(oracle_resize_bilinear)(const REAL *in, REAL *out,
                                int B, int C, int Hi, int Wi, int Ho, int Wo)
{
#pragma omp parallel for schedule(static)
    for (long pc = 0; pc < (long)B * C; ++pc) {
        const REAL *p = in + pc * Hi * Wi;
        REAL *q = out + pc * Ho * Wo;
        for (int y = 0; y < Ho; ++y) {
            int y0, y1; REAL h0, h1;
            FN(oracle_src_index)(y, Hi, Ho, &y0, &y1, &h0, &h1);
            for (int x = 0; x < Wo; ++x) {
                int x0, x1; REAL w0, w1;
                FN(oracle_src_index)(x, Wi, Wo, &x0, &x1, &w0, &w1);
                q[(long)y * Wo + x] =
                    h0 * (w0 * p[(long)y0 * Wi + x0] + w1 * p[(long)y0 * Wi + x1]) +
                    h1 * (w0 * p[(long)y1 * Wi + x0] + w1 * p[(long)y1 * Wi + x1]);
            }
        }
    }
}

/* Backward of the resize: scatter g_out * (hk * wk) into the 4 source taps. */
void FN(oracle_resize_bilinear_bwd)(const REAL *g_out, REAL *g_in,
                                    int B, int C, int Hi, int Wi, int Ho, int Wo)
{
#pragma omp parallel for schedule(static)
    for (long pc = 0; pc < (long)B * C; ++pc) {
        REAL *p = g_in + pc * Hi * Wi;
        const REAL *q = g_out + pc * Ho * Wo;
        for (long i = 0; i < (long)Hi * Wi; ++i) p[i] = (REAL)0;
        for (int y = 0; y < Ho; ++y) {
            int y0, y1; REAL h0, h1;
            FN(oracle_src_index)(y, Hi, Ho, &y0, &y1, &h0, &h1);
            for (int x = 0; x < Wo; ++x) {
                int x0, x1; REAL w0, w1;
                FN(oracle_src_index)(x, Wi, Wo, &x0, &x1, &w0, &w1);
                const REAL go = q[(long)y * Wo + x];
                p[(long)y0 * Wi + x0] += h0 * w0 * go;
                p[(long)y0 * Wi + x1] += h0 * w1 * go;
                p[(long)y1 * Wi + x0] += h1 * w0 * go;
                p[(long)y1 * Wi + x1] += h1 * w1 * go;
            }
        }
    }
}

/* ------------------------------------------------------------------------
 * LRwarp_error, utils/disparity_warper.py:109-115 (== stereo_op.py:67-73):
 *   if imgL wider than disp: resize imgL to disp's size; same for imgR;
 *   warped,_ = disp_warp(imgL, disp)  (padding 'border');  return imgR - warped
 * scratch must hold 3 * B*C*H*W elements (resized L, resized R, mask).
 * Returns the negative-disparity flag of the inner disp_warp.
 * ---------------------------------------------------------------------- */
int FN(oracle_lrwarp_error)(const REAL *imgL, const REAL *disp, const REAL *imgR,
                            REAL *out, REAL *scratch,
                            int B, int C, int H0, int W0, int H, int W)
{
    const long n = (long)B * C * H * W;
    REAL *l = scratch, *r = scratch + n, *m = scratch + 2 * n;
    const REAL *lp = imgL, *rp = imgR;
    if (W0 > W) {
        FN(oracle_resize_bilinear)(imgL, l, B, C, H0, W0, H, W);
        FN(oracle_resize_bilinear)(imgR, r, B, C, H0, W0, H, W);
        lp = l; rp = r;
    }
    int bad = FN(oracle_warp_fwd)(lp, disp, out, m, B, C, H, W, 1);
    for (long i = 0; i < n; ++i) out[i] = rp[i] - out[i];
    return bad;
}


/* ------------------------------------------------------------------------
 * Haze synthesis (input side of the training step).
 * Restates, operation for operation and with upstream's rounding points,
 *   convert_disp_to_depth  DeBug/inference.py:57-64    depth = (focal * baseline) / disp
 *   depth2trans            DeBug/inference.py:126-134  trans = exp((-depth) * beta)
 *   recover_haze_images    DeBug/inference.py:95-112   haze  = (clear * trans) + A * (1 - trans)
 * as called per batch by trainfiles/dssmd_traner_new.py:250-256.  src is the
 * disparity (src_is_disp != 0) or the depth itself.  The per-sample scalars
 * are REAL; the image-side element type is given by in_is_f32 (upstream's
 * type promotion: float32 images with the float64 scalars of the data loaders
 * compute in float64).  Any of haze / trans / depth may be NULL.
 * ---------------------------------------------------------------------- */
void FN(oracle_haze_fwd)(const void *clear, const void *src, int src_is_disp, int in_is_f32,
                         const REAL *baseline, const REAL *focal, const REAL *beta, const REAL *airlight,
                         REAL *haze, REAL *trans, REAL *depth, int B, int C, long HW)
{
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const REAL fb = src_is_disp ? focal[b] * baseline[b] : (REAL)0;
        const REAL be = (haze || trans) ? beta[b] : (REAL)0;
        const REAL A = haze ? airlight[b] : (REAL)0;
        for (long i = 0; i < HW; ++i) {
            const long p = (long)b * HW + i;
            const REAL s = in_is_f32 ? (REAL)((const float *)src)[p] : (REAL)((const double *)src)[p];
            const REAL dep = src_is_disp ? fb / s : s;
            if (depth) depth[p] = dep;
            if (!(haze || trans)) continue;
            const REAL nd = -dep;
            const REAL t = EXP(nd * be);
            if (trans) trans[p] = t;
            if (haze) {
                const REAL a = A * ((REAL)1 - t);
                for (int c = 0; c < C; ++c) {
                    const long pc = ((long)b * C + c) * HW + i;
                    const REAL J = in_is_f32 ? (REAL)((const float *)clear)[pc] : (REAL)((const double *)clear)[pc];
                    const REAL jt = J * t;
                    haze[pc] = jt + a;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------
 * ASM reconstruction loss, value and gradients.
 * Restates RecoveredCleanImagesLoss.forward, losses/DSSMDLoss.py:87-111 (the
 * `transmission_map.min==0` test compares a method with 0: always the else
 * branch), and the backward autograd derives from it, node by node:
 *   om = 1 - t;  am = A * om;  num = haze_s - am;  rec = num / t
 *   rc = clamp(rec, 0, 1);  diff = rc - clean_s;  loss = mean |diff|
 *   g_diff = sign(diff) / n;  g_rec = g_diff on 0 <= rec <= 1 else 0
 *   g_num = g_rec / t;  g_t += -g_rec * (rec / t)          (div backward)
 *   g_am = -g_num;  g_A += g_am * om;  g_om = g_am * A;  g_t += -g_om
 * haze_s / clean_s: the images at the size of t (resized by the caller with
 * oracle_resize_bilinear when larger).  Sums are accumulated in double.
 * Returns the loss; g_t: [B,1,h,w], g_A: [B] (either may be NULL).
 * ---------------------------------------------------------------------- */
double FN(oracle_recovered_l1)(const REAL *trans, const REAL *airlight, const REAL *haze_s, const REAL *clean_s,
                               REAL *g_t, REAL *g_A, int B, int C, long hw)
{
    const double n = (double)B * C * (double)hw;
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        const REAL A = airlight[b];
        double ga = 0.0;
        for (long i = 0; i < hw; ++i) {
            const REAL t = trans[(long)b * hw + i];
            const REAL om = (REAL)1 - t;
            const REAL am = A * om;
            double gt = 0.0;
            for (int c = 0; c < C; ++c) {
                const long pc = ((long)b * C + c) * hw + i;
                const REAL num = haze_s[pc] - am;
                const REAL rec = num / t;
                const REAL rc = FMIN(FMAX(rec, (REAL)0), (REAL)1);
                const REAL diff = rc - clean_s[pc];
                total += diff < 0 ? -(double)diff : (double)diff;
                const REAL sg = diff > 0 ? (REAL)1 : (diff < 0 ? (REAL)-1 : (REAL)0);
                const REAL g_rec = (rec >= 0 && rec <= 1) ? (REAL)(sg / n) : (REAL)0;
                const REAL g_num = g_rec / t;
                const REAL q = rec / t;
                gt += (double)(-g_rec * q);
                const REAL g_am = -g_num;
                ga += (double)(g_am * om);
                const REAL g_om = g_am * A;
                gt += (double)(-g_om);
            }
            if (g_t) g_t[(long)b * hw + i] = (REAL)gt;
        }
        if (g_A) g_A[b] = (REAL)ga;
    }
    return total / n;
}

/* ------------------------------------------------------------------------
 * Attention along the epipolar line, forward and backward.
 * Restates the core of STTM_Single.forward, models/BidNet/STTM_Simple.py:43-57:
 *   per image row (b, y):  dots[x][x'] = scale * sum_c q[c][x] k[c][x']
 *   attn = softmax(dots, over x');  out[c][x] = sum_x' attn[x][x'] v[c][x']
 * on NCHW tensors (upstream permutes to [B*H, W, C] first; same numbers), and
 * the backward autograd derives from it (matmul, softmax, matmul):
 *   dattn = g^T v;  ddots = attn o (dattn - rowsum(attn o dattn));  g_q = scale * ddots k;
 *   g_k = scale * ddots^T q;  g_v = attn^T g.
 * g == NULL: forward only.  scratch: 3 * W * threads REALs are taken from the heap.
 * ---------------------------------------------------------------------- */
#include <stdlib.h>
void FN(oracle_epipolar_attn)(const REAL *q, const REAL *k, const REAL *v, REAL *out, const REAL *g,
                              REAL *g_q, REAL *g_k, REAL *g_v, int B, int C, int H, int W, REAL scale)
{
    const long HW = (long)H * W;
#pragma omp parallel
    {
        REAL *attn = (REAL *)malloc(sizeof(REAL) * (size_t)W * 2);
        REAL *dat = attn + W;
#pragma omp for collapse(2) schedule(static)
        for (int b = 0; b < B; ++b)
            for (int y = 0; y < H; ++y) {
                const long ro = (long)b * C * HW + (long)y * W;
                if (g)
                    for (int c = 0; c < C; ++c)
                        for (int x = 0; x < W; ++x) { g_k[ro + c * HW + x] = 0; g_v[ro + c * HW + x] = 0; }
                for (int x = 0; x < W; ++x) {
                    REAL mx = 0;
                    for (int j = 0; j < W; ++j) {
                        REAL d = 0;
                        for (int c = 0; c < C; ++c) d += q[ro + c * HW + x] * k[ro + c * HW + j];
                        d *= scale;
                        attn[j] = d;
                        if (j == 0 || d > mx) mx = d;
                    }
                    REAL sum = 0;
                    for (int j = 0; j < W; ++j) { attn[j] = EXP(attn[j] - mx); sum += attn[j]; }
                    for (int j = 0; j < W; ++j) attn[j] /= sum;
                    for (int c = 0; c < C; ++c) {
                        REAL o = 0;
                        for (int j = 0; j < W; ++j) o += attn[j] * v[ro + c * HW + j];
                        out[ro + c * HW + x] = o;
                    }
                    if (!g) continue;
                    REAL rs = 0;
                    for (int j = 0; j < W; ++j) {
                        REAL d = 0;
                        for (int c = 0; c < C; ++c) d += g[ro + c * HW + x] * v[ro + c * HW + j];
                        dat[j] = d;
                        rs += attn[j] * d;
                    }
                    for (int j = 0; j < W; ++j) dat[j] = attn[j] * (dat[j] - rs);   /* ddots */
                    for (int c = 0; c < C; ++c) {
                        REAL a = 0;
                        const REAL qc = q[ro + c * HW + x], gc = g[ro + c * HW + x];
                        for (int j = 0; j < W; ++j) {
                            a += dat[j] * k[ro + c * HW + j];
                            g_k[ro + c * HW + j] += scale * dat[j] * qc;
                            g_v[ro + c * HW + j] += attn[j] * gc;
                        }
                        g_q[ro + c * HW + x] = scale * a;
                    }
                }
            }
        free(attn);
    }
}

#undef FN
#undef CAT
#undef CAT_
