/*
 * mgvs_oracle.c -- CPU ORACLE for the MGNet view-synthesis loss.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithm for the hot path
 *   mgnet/modeling/loss.py:111-294  (MultiViewPhotometricLoss)
 *   mgnet/geometry/camera.py:72-81,107-182, camera_utils.py:24-54, pose.py:41-47,77-82,
 *   pose_utils.py:9-59, depth.py:11-51, image.py:42-69
 * plus the ATen kernels those call (grid_sampler_2d bilinear/zeros/align_corners=True,
 * reflection_pad2d + avg_pool2d(3,1), bmm, min).  It is the checker the CUDA kernels are tested
 * against; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build or call it.  The product (mgnet_b200/) never does.
 *
 * Parity status: PINNED against outputs of the reference itself (imported from /root/reference
 * in the build container, torch 2.11 CPU) -- tests/golden/ (npz files) made by tests/golden/make_golden.py.
 * The reference ships no tests or golden vectors of its own.
 *
 * Forward arithmetic follows the fp32 op order of the reference on CPU exactly,
 * so normalised coordinates, warped images, per-pixel photometric maps and the argmin selection are
 * bit-identical to the reference.  The one deliberate deviation: sin/cos of the Euler angles are the
 * correctly rounded values (computed in double) whereas torch-CPU calls MKL VML, which differs from
 * correct rounding by 1 ulp for ~5% of arguments; fixtures use angles where the two agree.
 * Backward is the closed-form adjoint evaluated in double at the fp32 forward
 * values, with every discrete decision (argmin, floor cell, clamps, sign) taken from the fp32 forward.
 *
 * Build:  gcc -O2 -fopenmp -mfma -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off is REQUIRED: every fused multiply-add below is an explicit fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_SCALES 8
#define ORC_S 2 /* source frames: prev, next (loss.py:116) */

typedef struct {
    int B, H, W, n;
    const float *target;            /* [B,3,H,W] */
    const float *source[ORC_S];     /* [B,3,H,W] */
    const float *inv[ORC_MAX_SCALES]; /* [B,1,H,W] each */
    const float *cam;               /* camera matrices, element (b,r,c) at cam[b*cam_bs + r*cam_rs + c] */
    long cam_bs, cam_rs;
    const float *poses;             /* [B,S,6] (tx,ty,tz,rx,ry,rz) */
    const uint8_t *mask;            /* [B,1,H,W] bool or NULL */
    float ssim_w, photo_w, smooth_w;
    int automask;
    int padding_mode;               /* grid_sample padding_mode (camera_utils.py:52-54): 0 zeros, 1 border, 2 reflection */
    const float *pose_mats;         /* optional [B,S,3,4] row-major (R|t) = rows 0..2 of Pose.mat (pose.py:41-47); when non-NULL it
                                       replaces the Euler evaluation of `poses` (the caller built R, e.g. with torch's pose_vec2mat) */
} OrcIn;

typedef struct {
    /* optional dumps (NULL to skip) */
    float *coords;   /* [n,S,B,H,W,2] normalised sample coordinates (camera.py:171-182) */
    float *warped;   /* [n,S,B,3,H,W] */
    float *photo;    /* [n,S,B,H,W]   per-pixel photometric maps of the warped sources */
    float *identity; /* [S,B,H,W]     identity-reprojection maps */
    float *minmap;   /* [n,B,H,W] */
    uint8_t *sel;    /* [n,B,H,W] argmin in list order (loss.py:136-144) */
    float *posemat;  /* [B,S,12] row-major 3x4 (R|t) */
    float *kinv;     /* [B,9] */
    double *sums;    /* [3n+3]: photo_i, N, smx_i, smy_i, Nx, Ny  (rank-level sums, fp64) */
    float loss_photo, loss_smooth;
} OrcOut;

/* ------------------------------------------------------------------------------------------ */
/* exact fp32 helpers                                                                          */

static inline float dot3(float a0, float a1, float a2, float b0, float b1, float b2)
{   /* MKL sgemm K=3 == ascending FMA chain */
    float acc = a0 * b0;
    acc = fmaf(a1, b1, acc);
    acc = fmaf(a2, b2, acc);
    return acc;
}

static inline int reflect_idx(int j, int n)
{   /* F.pad(.., "reflect"): -1 -> 1, n -> n-2 */
    if (j < 0) j = -j;
    if (j >= n) j = 2 * n - 2 - j;
    if (j < 0) j = 0;
    if (j >= n) j = n - 1;
    return j;
}

/* Camera parameters per image: K, Kinv (camera.py:72-81), and per source R|t (pose_utils.py:9-51). */
static void orc_prep_cam(const OrcIn *in, int b, float K[9], float Kinv[9], float Rt[ORC_S][12],
                         int euler_fma)
{
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) K[r * 3 + c] = in->cam[b * in->cam_bs + r * in->cam_rs + c];
    memcpy(Kinv, K, 9 * sizeof(float));
    float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    Kinv[0] = 1.0f / fx;
    Kinv[4] = 1.0f / fy;
    Kinv[2] = (-1.0f * cx) / fx;
    Kinv[5] = (-1.0f * cy) / fy;
    for (int s = 0; s < ORC_S; s++) {
        if (in->pose_mats) {
            memcpy(Rt[s], in->pose_mats + ((long)b * ORC_S + s) * 12, 12 * sizeof(float));
            continue;
        }
        const float *v = in->poses + ((long)b * ORC_S + s) * 6;
        float cxr = (float)cos((double)v[3]), sxr = (float)sin((double)v[3]);
        float cyr = (float)cos((double)v[4]), syr = (float)sin((double)v[4]);
        float czr = (float)cos((double)v[5]), szr = (float)sin((double)v[5]);
        float z0 = v[5] * 0.0f, o1 = z0 + 1.0f; /* zeros = z*0, ones = zeros+1 (pose_utils.py:17-18) */
        float zm[9] = {czr, -szr, z0, szr, czr, z0, z0, z0, o1};
        float ym[9] = {cyr, z0, syr, z0, o1, z0, -syr, z0, cyr};
        float xm[9] = {o1, z0, z0, z0, cxr, -sxr, z0, sxr, cxr};
        float xy[9], R[9];
        for (int pass = 0; pass < 2; pass++) {
            const float *A = pass ? xy : xm, *Bm = pass ? zm : ym;
            float *C = pass ? R : xy;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    if (euler_fma) {
                        C[r * 3 + c] = dot3(A[r * 3], A[r * 3 + 1], A[r * 3 + 2], Bm[c], Bm[3 + c], Bm[6 + c]);
                    } else { /* small-matrix bmm path: acc = 0; acc += a*b with separate roundings */
                        float acc = 0.0f;
                        for (int k = 0; k < 3; k++) { float p = A[r * 3 + k] * Bm[k * 3 + c]; acc = acc + p; }
                        C[r * 3 + c] = acc;
                    }
                }
        }
        for (int r = 0; r < 3; r++) {
            Rt[s][r * 4 + 0] = R[r * 3 + 0];
            Rt[s][r * 4 + 1] = R[r * 3 + 1];
            Rt[s][r * 4 + 2] = R[r * 3 + 2];
            Rt[s][r * 4 + 3] = v[r];
        }
    }
}

typedef struct { /* everything the projection of one pixel produces (fp32, reference op order) */
    float r[3];   /* Kinv (u,v,1) */
    float d;      /* 1/clamp(inv) */
    float Xc[3];  /* r*d */
    float Xs[3];  /* R Xc + t */
    float P[3];   /* K Xs */
    float Z, ax, ay; /* clamp(Pz), Px/Z, Py/Z */
    float xn, yn; /* normalised coords */
    float ix, iy; /* sample position in source pixels after the padding-mode mapping (ATen GridSampler.h:27-36, 60-140) */
    float mx, my; /* d(ix)/d(un-normalised x), d(iy)/d(un-normalised y): 1 for zeros; 0 / +-1 for border / reflection */
} OrcProj;

/* grid_sampler_compute_source_index for align_corners=True (ATen GridSampler.h:60-140; CPU vectorised kernel
 * GridSamplerKernel.cpp ComputeLocation): border clips to [0, size-1]; reflection folds around 0 and size-1
 * (|c| - trunc(|c| / 2span) * 2span, then min(extra, 2span - extra)) and clips.  Probed bit for bit against torch 2.11.
 * *mult receives the derivative the backward applies (clip_coordinates_set_grad / reflect_coordinates_set_grad). */
static inline float orc_pad_coord(float c, float sm1, int mode, float *mult)
{
    *mult = 1.0f;
    if (mode == 0) return c;
    if (mode == 2) {
        float ts = sm1 + sm1, a = fabsf(c);
        float df = truncf(a / ts);
        float extra = a - df * ts;
        float other = ts - extra;
        float sgn = c < 0.0f ? -1.0f : 1.0f;
        *mult = extra <= other ? sgn : -sgn;
        c = extra < other ? extra : other;
    }
    if (!(c > 0.0f && c < sm1)) *mult = 0.0f;
    float lo = c > 0.0f ? c : 0.0f;
    return lo < sm1 ? lo : sm1;
}

static inline void orc_project(const float K[9], const float Kinv[9], const float Rt[12], int u, int v,
                               float inv, int H, int W, int pad, OrcProj *o)
{
    float gu = (float)u, gv = (float)v;
    for (int j = 0; j < 3; j++) o->r[j] = dot3(Kinv[j * 3], Kinv[j * 3 + 1], Kinv[j * 3 + 2], gu, gv, 1.0f);
    float ci = inv < 1e-6f ? 1e-6f : inv;            /* depth.py:15 */
    o->d = 1.0f / ci;
    for (int j = 0; j < 3; j++) o->Xc[j] = o->r[j] * o->d;   /* camera.py:131; Twc = I is an exact no-op */
    for (int j = 0; j < 3; j++) {
        float acc = dot3(Rt[j * 4], Rt[j * 4 + 1], Rt[j * 4 + 2], o->Xc[0], o->Xc[1], o->Xc[2]);
        o->Xs[j] = acc + Rt[j * 4 + 3];              /* pose.py:81: separate rounded add */
    }
    for (int j = 0; j < 3; j++) o->P[j] = dot3(K[j * 3], K[j * 3 + 1], K[j * 3 + 2], o->Xs[0], o->Xs[1], o->Xs[2]);
    o->Z = o->P[2] < 1e-5f ? 1e-5f : o->P[2];        /* camera.py:171 */
    o->ax = o->P[0] / o->Z;
    o->ay = o->P[1] / o->Z;
    float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    o->xn = (2.0f * o->ax) / wm1 - 1.0f;             /* camera.py:172-173 */
    o->yn = (2.0f * o->ay) / hm1 - 1.0f;
    o->ix = (o->xn + 1.0f) * (wm1 * 0.5f);           /* ((c+1)/2)*(size-1) == (c+1)*((size-1)/2) */
    o->iy = (o->yn + 1.0f) * (hm1 * 0.5f);
    o->ix = orc_pad_coord(o->ix, wm1, pad, &o->mx);
    o->iy = orc_pad_coord(o->iy, hm1, pad, &o->my);
}

typedef struct {
    int x0, y0;               /* floor cell (may be far outside) as clamped ints */
    int in_nw, in_ne, in_sw, in_se;
    float wE, wW, wS, wN;     /* w = ix-x0, e = 1-w, n = iy-y0, s = 1-n in ATen naming: wE==w etc. */
} OrcCell;

static inline void orc_cell(float ix, float iy, int H, int W, OrcCell *c)
{
    float xw = floorf(ix), yn = floorf(iy);
    float w = ix - xw, e = 1.0f - w, n = iy - yn, s = 1.0f - n;
    c->wE = w; c->wW = e; c->wS = n; c->wN = s;
    int wm = (xw > -1.0f) && (xw < (float)W);
    int em = (xw + 1.0f > -1.0f) && (xw + 1.0f < (float)W);
    int nm = (yn > -1.0f) && (yn < (float)H);
    int sm = (yn + 1.0f > -1.0f) && (yn + 1.0f < (float)H);
    c->in_nw = wm && nm; c->in_ne = em && nm; c->in_sw = wm && sm; c->in_se = em && sm;
    /* safe ints: only used when the matching in_* flag is set */
    float cx = xw < -2.0f ? -2.0f : (xw > (float)W ? (float)W : xw);
    float cy = yn < -2.0f ? -2.0f : (yn > (float)H ? (float)H : yn);
    c->x0 = (int)cx; c->y0 = (int)cy;
}

static inline float orc_bilinear(const float *img /*[H,W]*/, int W, const OrcCell *c, float v[4])
{
    v[0] = c->in_nw ? img[(long)c->y0 * W + c->x0] : 0.0f;
    v[1] = c->in_ne ? img[(long)c->y0 * W + c->x0 + 1] : 0.0f;
    v[2] = c->in_sw ? img[(long)(c->y0 + 1) * W + c->x0] : 0.0f;
    v[3] = c->in_se ? img[(long)(c->y0 + 1) * W + c->x0 + 1] : 0.0f;
    float nw = c->wN * c->wW, ne = c->wN * c->wE, sw = c->wS * c->wW, se = c->wS * c->wE;
    float acc = v[0] * nw;                 /* ATen CPU blend == this FMA chain (App. A) */
    acc = fmaf(v[1], ne, acc);
    acc = fmaf(v[2], sw, acc);
    acc = fmaf(v[3], se, acc);
    return acc;
}

typedef struct { float mu_x, mu_y, mu_xy, mu_xx, mu_yy, sig_x, sig_y, sig_xy, n1, n2, d1, d2, ssim, loss; } OrcSsim;

/* 3x3 reflect-padded window sum, row-major order, then IEEE /9 (App. A "3x3 mean"). */
static inline void orc_ssim_px(const float *x, const float *y, int H, int W, int v, int u, OrcSsim *o)
{
    float sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
    int first = 1;
    for (int dy = -1; dy <= 1; dy++) {
        int rr = reflect_idx(v + dy, H);
        for (int dx = -1; dx <= 1; dx++) {
            int cc = reflect_idx(u + dx, W);
            float xv = x[(long)rr * W + cc], yv = y[(long)rr * W + cc];
            float xx = xv * xv, yy = yv * yv, xy = xv * yv;
            if (first) { sx = xv; sy = yv; sxx = xx; syy = yy; sxy = xy; first = 0; }
            else { sx = sx + xv; sy = sy + yv; sxx = sxx + xx; syy = syy + yy; sxy = sxy + xy; }
        }
    }
    o->mu_x = sx / 9.0f; o->mu_y = sy / 9.0f;
    o->mu_xx = sxx / 9.0f; o->mu_yy = syy / 9.0f; o->mu_xy = sxy / 9.0f;
    float mxy = o->mu_x * o->mu_y, mxs = o->mu_x * o->mu_x, mys = o->mu_y * o->mu_y;
    o->sig_x = o->mu_xx - mxs; o->sig_y = o->mu_yy - mys; o->sig_xy = o->mu_xy - mxy;
    const float c1 = 1e-4f, c2 = 9e-4f;
    o->n1 = 2.0f * mxy + c1;        /* 2*a exact => same bits as separately rounded mul+add */
    o->n2 = 2.0f * o->sig_xy + c2;
    o->d1 = (mxs + mys) + c1;
    o->d2 = (o->sig_x + o->sig_y) + c2;
    o->ssim = (o->n1 * o->n2) / (o->d1 * o->d2);
    float l = (1.0f - o->ssim) / 2.0f;
    o->loss = l < 0.0f ? 0.0f : (l > 1.0f ? 1.0f : l);
}

/* photometric map of (est, tgt): alpha*mean_c(ssim) + (1-alpha)*mean_c(|est-tgt|) (loss.py:186-194) */
static void orc_photometric_map(const float *est /*[3,H,W]*/, const float *tgt, int H, int W, float alpha_f,
                                float one_minus_alpha_f, float *out)
{
    long HW = (long)H * W;
#pragma omp parallel for schedule(static)
    for (int v = 0; v < H; v++)
        for (int u = 0; u < W; u++) {
            float ss[3], l1[3];
            for (int c = 0; c < 3; c++) {
                OrcSsim s;
                orc_ssim_px(est + c * HW, tgt + c * HW, H, W, v, u, &s);
                ss[c] = s.loss;
                l1[c] = fabsf(est[c * HW + (long)v * W + u] - tgt[c * HW + (long)v * W + u]);
            }
            float sm = ((ss[0] + ss[1]) + ss[2]) / 3.0f;
            float lm = ((l1[0] + l1[1]) + l1[2]) / 3.0f;
            float a = alpha_f * sm, b = one_minus_alpha_f * lm;
            out[(long)v * W + u] = a + b;
        }
}

static void orc_warp(const OrcIn *in, int b, int i, int s, const float K[9], const float Kinv[9],
                     const float Rt[12], float *warped /*[3,H,W]*/, float *coords /*[H,W,2] or NULL*/)
{
    int H = in->H, W = in->W;
    long HW = (long)H * W;
    const float *inv = in->inv[i] + (long)b * HW;
    const float *src = in->source[s] + (long)b * 3 * HW;
#pragma omp parallel for schedule(static)
    for (int v = 0; v < H; v++)
        for (int u = 0; u < W; u++) {
            OrcProj p; OrcCell c; float vals[4];
            orc_project(K, Kinv, Rt, u, v, inv[(long)v * W + u], H, W, in->padding_mode, &p);
            orc_cell(p.ix, p.iy, H, W, &c);
            for (int ch = 0; ch < 3; ch++) warped[ch * HW + (long)v * W + u] = orc_bilinear(src + ch * HW, W, &c, vals);
            if (coords) { coords[((long)v * W + u) * 2] = p.xn; coords[((long)v * W + u) * 2 + 1] = p.yn; }
        }
}

/* float32(1 - 0.85) in Python double arithmetic, then rounded to fp32 by the tensor*scalar op */
static inline float one_minus_alpha(float ssim_w_as_given_double_rounded, double ssim_w_double)
{
    (void)ssim_w_as_given_double_rounded;
    return (float)(1.0 - ssim_w_double);
}

/* ------------------------------------------------------------------------------------------ */
/* forward                                                                                     */

int orc_forward(const OrcIn *in, double ssim_w_double, int euler_fma, OrcOut *out)
{
    const int B = in->B, H = in->H, W = in->W, n = in->n, S = ORC_S;
    const long HW = (long)H * W;
    if (n < 1 || n > ORC_MAX_SCALES || H < 2 || W < 2) return -1;
    /* ssim_loss_weight == 0: calc_photometric_loss returns the raw 3-channel |x - y| (loss.py:195-196), so the min of
       reduce_photometric_loss runs over 3 channels per list entry and sel indexes [entry][channel] */
    const int l1only = !(ssim_w_double > 0.0);
    const int cpm = l1only ? 3 : 1;
    const int nch = (in->automask ? 2 * S : S) * cpm;
    const float alpha_f = (float)ssim_w_double;
    const float oma_f = one_minus_alpha(alpha_f, ssim_w_double);

    double *sums = out->sums;
    double local_sums[3 * ORC_MAX_SCALES + 3];
    if (!sums) sums = local_sums;
    for (int k = 0; k < 3 * n + 3; k++) sums[k] = 0.0;
    double *ph = sums, *Ncnt = sums + n, *smx = sums + n + 1, *smy = sums + 2 * n + 1, *Nx = sums + 3 * n + 1,
           *Ny = sums + 3 * n + 2;

    float *warped = (float *)malloc(sizeof(float) * 3 * HW);
    float *maps = (float *)malloc(sizeof(float) * 12 * HW);  /* list-ordered loss maps of one scale */
    float *idm = (float *)malloc(sizeof(float) * S * 3 * HW);
    float *wx = (float *)malloc(sizeof(float) * HW), *wy = (float *)malloc(sizeof(float) * HW);
    if (!warped || !maps || !idm || !wx || !wy) return -2;

    for (int b = 0; b < B; b++) {
        float K[9], Kinv[9], Rt[ORC_S][12];
        orc_prep_cam(in, b, K, Kinv, Rt, euler_fma);
        if (out->posemat) memcpy(out->posemat + (long)b * S * 12, Rt, sizeof(float) * S * 12);
        if (out->kinv) memcpy(out->kinv + (long)b * 9, Kinv, sizeof(float) * 9);
        const float *tgt = in->target + (long)b * 3 * HW;
        const uint8_t *mk = in->mask ? in->mask + (long)b * HW : NULL;

        if (in->automask)
            for (int s = 0; s < S; s++) {
                if (l1only) {
                    const float *sp = in->source[s] + (long)b * 3 * HW;
                    for (long p = 0; p < 3 * HW; p++) idm[s * 3 * HW + p] = fabsf(sp[p] - tgt[p]);
                    continue;
                }
                orc_photometric_map(in->source[s] + (long)b * 3 * HW, tgt, H, W, alpha_f, oma_f, idm + s * HW);
                if (out->identity) memcpy(out->identity + ((long)s * B + b) * HW, idm + s * HW, sizeof(float) * HW);
            }
        /* edge-aware weights from the full-res target (depth.py:23-24) */
#pragma omp parallel for schedule(static)
        for (int v = 0; v < H; v++)
            for (int u = 0; u < W; u++) {
                long p = (long)v * W + u;
                float gx = 0, gy = 0;
                if (u + 1 < W) {
                    float a0 = fabsf(tgt[p] - tgt[p + 1]), a1 = fabsf(tgt[HW + p] - tgt[HW + p + 1]),
                          a2 = fabsf(tgt[2 * HW + p] - tgt[2 * HW + p + 1]);
                    gx = expf(-(((a0 + a1) + a2) / 3.0f));
                }
                if (v + 1 < H) {
                    float a0 = fabsf(tgt[p] - tgt[p + W]), a1 = fabsf(tgt[HW + p] - tgt[HW + p + W]),
                          a2 = fabsf(tgt[2 * HW + p] - tgt[2 * HW + p + W]);
                    gy = expf(-(((a0 + a1) + a2) / 3.0f));
                }
                wx[p] = gx; wy[p] = gy;
            }
        /* mask counts (scale independent) */
        double cN = 0, cNx = 0, cNy = 0;
        for (int v = 0; v < H; v++)
            for (int u = 0; u < W; u++) {
                int m = mk ? (mk[(long)v * W + u] != 0) : 1;
                cN += m; if (u + 1 < W) cNx += m; if (v + 1 < H) cNy += m;
            }
        *Ncnt += cN; *Nx += cNx; *Ny += cNy;

        for (int i = 0; i < n; i++) {
            for (int s = 0; s < S; s++) {
                float *cd = out->coords ? out->coords + ((((long)i * S + s) * B + b) * HW) * 2 : NULL;
                orc_warp(in, b, i, s, K, Kinv, Rt[s], warped, cd);
                if (out->warped) memcpy(out->warped + (((long)i * S + s) * B + b) * 3 * HW, warped, sizeof(float) * 3 * HW);
                int slot = in->automask ? 2 * s : s;
                if (l1only) {
                    for (long p = 0; p < 3 * HW; p++) maps[slot * 3 * HW + p] = fabsf(warped[p] - tgt[p]);
                    if (in->automask) memcpy(maps + (2 * s + 1) * 3 * HW, idm + s * 3 * HW, sizeof(float) * 3 * HW);
                    continue;
                }
                orc_photometric_map(warped, tgt, H, W, alpha_f, oma_f, maps + slot * HW);
                if (out->photo) memcpy(out->photo + (((long)i * S + s) * B + b) * HW, maps + slot * HW, sizeof(float) * HW);
                if (in->automask) memcpy(maps + (2 * s + 1) * HW, idm + s * HW, sizeof(float) * HW);
            }
            double acc = 0.0;
            for (long p = 0; p < HW; p++) {
                float best = maps[p]; int bi = 0;
                for (int k = 1; k < nch; k++) if (maps[k * HW + p] < best) { best = maps[k * HW + p]; bi = k; }
                if (out->sel) out->sel[((long)i * B + b) * HW + p] = (uint8_t)bi;
                if (out->minmap) out->minmap[((long)i * B + b) * HW + p] = best;
                if (!mk || mk[p]) acc += (double)best;
            }
            ph[i] += acc;
            /* smoothness, factorised (depth.py:18-51, loss.py:274-294) */
            const float *inv = in->inv[i] + (long)b * HW;
            double msum = 0;
            for (long p = 0; p < HW; p++) msum += inv[p];
            double mean = msum / (double)HW;
            double c = mean < 1e-6 ? 1e-6 : mean;
            double ax = 0, ay = 0;
            for (int v = 0; v < H; v++)
                for (int u = 0; u < W; u++) {
                    long p = (long)v * W + u;
                    int m = mk ? (mk[p] != 0) : 1;
                    if (!m) continue;
                    if (u + 1 < W) ax += fabs((double)inv[p] - (double)inv[p + 1]) * wx[p];
                    if (v + 1 < H) ay += fabs((double)inv[p] - (double)inv[p + W]) * wy[p];
                }
            smx[i] += ax / c; smy[i] += ay / c;
        }
    }
    double lp = 0, ls = 0;
    for (int i = 0; i < n; i++) {
        lp += ph[i] / *Ncnt;
        ls += (smx[i] / *Nx + smy[i] / *Ny) / (double)(1 << i);
    }
    out->loss_photo = (float)(lp / n * (double)in->photo_w);
    out->loss_smooth = (float)(ls / n * (double)in->smooth_w);
    free(warped); free(maps); free(idm); free(wx); free(wy);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* backward: closed-form adjoint, double arithmetic at fp32 forward values     */

int orc_backward(const OrcIn *in, double ssim_w_double, int euler_fma, const uint8_t *sel /*[n,B,H,W]*/,
                 const double *sums /*[3n+3] (global)*/, double g_photo, double g_smooth,
                 float *grad_inv[ORC_MAX_SCALES] /*[B,1,H,W] each*/, float *grad_poses /*[B,S,6]*/,
                 double *grad_Rt /*[B,S,12] or NULL*/)
{
    const int B = in->B, H = in->H, W = in->W, n = in->n, S = ORC_S;
    const long HW = (long)H * W;
    const double alpha = (double)(float)ssim_w_double, oma = (double)(float)(1.0 - ssim_w_double);
    const double N = sums[n], Nx = sums[3 * n + 1], Ny = sums[3 * n + 2];
    const int l1only = !(ssim_w_double > 0.0);   /* raw 3-channel L1 (loss.py:195-196): sel = entry * 3 + channel */

    float *xw[ORC_S];
    double *coef[ORC_S];
    for (int s = 0; s < S; s++) {
        xw[s] = (float *)malloc(sizeof(float) * 3 * HW);
        coef[s] = (double *)malloc(sizeof(double) * 9 * HW); /* [3 ch][a,b,c][HW] */
        if (!xw[s] || !coef[s]) return -2;
    }
    float *wx = (float *)malloc(sizeof(float) * HW), *wy = (float *)malloc(sizeof(float) * HW);

    for (int b = 0; b < B; b++) {
        float K[9], Kinv[9], Rt[ORC_S][12];
        orc_prep_cam(in, b, K, Kinv, Rt, euler_fma);
        const float *tgt = in->target + (long)b * 3 * HW;
        const uint8_t *mk = in->mask ? in->mask + (long)b * HW : NULL;
        double gRt[ORC_S][12];
        memset(gRt, 0, sizeof(gRt));
        for (long p = 0; p < HW; p++) {
            int v = (int)(p / W), u = (int)(p % W);
            float gx = 0, gy = 0;
            if (u + 1 < W) {
                float a0 = fabsf(tgt[p] - tgt[p + 1]), a1 = fabsf(tgt[HW + p] - tgt[HW + p + 1]),
                      a2 = fabsf(tgt[2 * HW + p] - tgt[2 * HW + p + 1]);
                gx = expf(-(((a0 + a1) + a2) / 3.0f));
            }
            if (v + 1 < H) {
                float a0 = fabsf(tgt[p] - tgt[p + W]), a1 = fabsf(tgt[HW + p] - tgt[HW + p + W]),
                      a2 = fabsf(tgt[2 * HW + p] - tgt[2 * HW + p + W]);
                gy = expf(-(((a0 + a1) + a2) / 3.0f));
            }
            wx[p] = gx; wy[p] = gy;
        }
        for (int i = 0; i < n; i++) {
            const float *inv = in->inv[i] + (long)b * HW;
            float *ginv = grad_inv[i] + (long)b * HW;
            const uint8_t *sl = sel + ((long)i * B + b) * HW;
            const double Wp = g_photo * (double)in->photo_w / ((double)n * N);
            /* ---- smoothness part ---- */
            {
                double msum = 0;
                for (long p = 0; p < HW; p++) msum += inv[p];
                double mean = msum / (double)HW;
                int active = mean >= 1e-6;
                double c = active ? mean : 1e-6;
                double Ws = g_smooth * (double)in->smooth_w / ((double)n * (double)(1 << i));
                double A = 0;
                for (int v = 0; v < H; v++)
                    for (int u = 0; u < W; u++) {
                        long p = (long)v * W + u;
                        int m = mk ? (mk[p] != 0) : 1;
                        if (!m) continue;
                        if (u + 1 < W) A += fabs((double)inv[p] - (double)inv[p + 1]) * wx[p] / Nx;
                        if (v + 1 < H) A += fabs((double)inv[p] - (double)inv[p + W]) * wy[p] / Ny;
                    }
                double mean_term = active ? -Ws * A / (c * c * (double)HW) : 0.0;
#pragma omp parallel for schedule(static)
                for (int v = 0; v < H; v++)
                    for (int u = 0; u < W; u++) {
                        long p = (long)v * W + u;
                        double g = mean_term;
                        /* NOTE the reference differentiates |d_hat[p]-d_hat[p+1]| with d_hat in fp32:
                           sign taken from the fp32 normalised difference */
                        float cf = (float)c;
                        int mp = mk ? (mk[p] != 0) : 1;
                        if (u + 1 < W && mp) {
                            float df = inv[p] / cf - inv[p + 1] / cf;
                            double sg = (df > 0) - (df < 0);
                            g += Ws * sg * wx[p] / (Nx * c);
                        }
                        if (u > 0 && (mk ? (mk[p - 1] != 0) : 1)) {
                            float df = inv[p - 1] / cf - inv[p] / cf;
                            double sg = (df > 0) - (df < 0);
                            g -= Ws * sg * wx[p - 1] / (Nx * c);
                        }
                        if (v + 1 < H && mp) {
                            float df = inv[p] / cf - inv[p + W] / cf;
                            double sg = (df > 0) - (df < 0);
                            g += Ws * sg * wy[p] / (Ny * c);
                        }
                        if (v > 0 && (mk ? (mk[p - W] != 0) : 1)) {
                            float df = inv[p - W] / cf - inv[p] / cf;
                            double sg = (df > 0) - (df < 0);
                            g -= Ws * sg * wy[p - W] / (Ny * c);
                        }
                        ginv[p] = (float)g;   /* photometric part added below */
                    }
            }
            /* ---- photometric part ---- */
            for (int s = 0; s < S; s++) {
                orc_warp(in, b, i, s, K, Kinv, Rt[s], xw[s], NULL);
                memset(coef[s], 0, sizeof(double) * 9 * HW);
            }
            /* coefficient maps at every output pixel p whose argmin is a warped source */
#pragma omp parallel for schedule(static)
            for (int v = 0; v < H; v++)
                for (int u = 0; u < W; u++) {
                    long p = (long)v * W + u;
                    if (mk && !mk[p]) continue;
                    int k = sl[p];
                    int s;
                    if (l1only) continue;   /* no SSIM term */
                    if (in->automask) { if (k & 1) continue; s = k >> 1; } else s = k;
                    for (int ch = 0; ch < 3; ch++) {
                        OrcSsim q;
                        orc_ssim_px(xw[s] + ch * HW, tgt + ch * HW, H, W, v, u, &q);
                        float lraw = (1.0f - q.ssim) / 2.0f;
                        if (!(lraw >= 0.0f && lraw <= 1.0f)) continue;   /* clamp gradient is inclusive */
                        double mux = q.mu_x, muy = q.mu_y, n1 = q.n1, n2 = q.n2, d1 = q.d1, d2 = q.d2, ss = q.ssim;
                        double dd = d1 * d2;
                        double ds_dmux = (2.0 * muy * n2 - 2.0 * muy * n1) / dd - ss * (2.0 * mux / d1 - 2.0 * mux / d2);
                        double ds_dexx = -ss / d2;
                        double ds_dexy = 2.0 * n1 / dd;
                        double uu = Wp * alpha / 3.0 * (-0.5);
                        coef[s][(ch * 3 + 0) * HW + p] = uu * ds_dmux / 9.0;
                        coef[s][(ch * 3 + 1) * HW + p] = uu * ds_dexx * 2.0 / 9.0;
                        coef[s][(ch * 3 + 2) * HW + p] = uu * ds_dexy / 9.0;
                    }
                }
            /* gather the box adjoint at q, chain through bilinear sampling and projection */
            for (int s = 0; s < S; s++) {
                const float *src = in->source[s] + (long)b * 3 * HW;
                double acc12[12];
                memset(acc12, 0, sizeof(acc12));
#pragma omp parallel
                {
                    double loc[12];
                    memset(loc, 0, sizeof(loc));
#pragma omp for schedule(static)
                    for (int v = 0; v < H; v++)
                        for (int u = 0; u < W; u++) {
                            long q = (long)v * W + u;
                            double G[3] = {0, 0, 0};
                            int any = 0;
                            /* adjoint of reflect-pad + 3x3 box: window p in N(q) with multiplicity 2 where
                               the padded tap of a border row/column folds back onto q (row -1 -> 1, H -> H-2) */
                            double rw[3], cw[3];
                            rw[0] = (v - 1 >= 0) ? 1.0 + (v == 1) : 0.0;
                            rw[1] = 1.0;
                            rw[2] = (v + 1 <= H - 1) ? 1.0 + (v == H - 2) : 0.0;
                            cw[0] = (u - 1 >= 0) ? 1.0 + (u == 1) : 0.0;
                            cw[1] = 1.0;
                            cw[2] = (u + 1 <= W - 1) ? 1.0 + (u == W - 2) : 0.0;
                            for (int ch = 0; ch < 3; ch++) {
                                double sa = 0, sb = 0, sc = 0;
                                for (int dy = -1; dy <= 1; dy++) {
                                    if (rw[dy + 1] == 0.0) continue;
                                    for (int dx = -1; dx <= 1; dx++) {
                                        if (cw[dx + 1] == 0.0) continue;
                                        long p = (long)(v + dy) * W + (u + dx);
                                        double wgt = rw[dy + 1] * cw[dx + 1];
                                        sa += wgt * coef[s][(ch * 3 + 0) * HW + p];
                                        sb += wgt * coef[s][(ch * 3 + 1) * HW + p];
                                        sc += wgt * coef[s][(ch * 3 + 2) * HW + p];
                                    }
                                }
                                double xq = xw[s][ch * HW + q], yq = tgt[ch * HW + q];
                                G[ch] = sa + xq * sb + yq * sc;
                                /* L1 term at q itself */
                                int m = mk ? (mk[q] != 0) : 1;
                                int k = sl[q];
                                int selq = in->automask ? (k == 2 * s) : (k == s);
                                if (l1only) {   /* the min picked one (entry, channel): only that channel of that source gets gradient */
                                    int entry = k / 3;
                                    selq = (in->automask ? (entry == 2 * s) : (entry == s)) && (k % 3 == ch);
                                    if (m && selq) {
                                        float df = xw[s][ch * HW + q] - tgt[ch * HW + q];
                                        G[ch] += Wp * (double)((df > 0) - (df < 0));
                                    }
                                } else
                                if (m && selq) {
                                    float df = xw[s][ch * HW + q] - tgt[ch * HW + q];
                                    G[ch] += Wp * oma / 3.0 * (double)((df > 0) - (df < 0));
                                }
                                if (G[ch] != 0.0) any = 1;
                            }
                            if (!any) continue;
                            OrcProj pr; OrcCell c; float vals[4];
                            orc_project(K, Kinv, Rt[s], u, v, inv[q], H, W, in->padding_mode, &pr);
                            orc_cell(pr.ix, pr.iy, H, W, &c);
                            double gix = 0, giy = 0;
                            for (int ch = 0; ch < 3; ch++) {
                                (void)orc_bilinear(src + ch * HW, W, &c, vals);
                                double nw = vals[0], ne = vals[1], sw = vals[2], se = vals[3];
                                gix += G[ch] * ((ne - nw) * (double)c.wN + (se - sw) * (double)c.wS);
                                giy += G[ch] * ((sw - nw) * (double)c.wW + (se - ne) * (double)c.wE);
                            }
                            gix *= (double)pr.mx;   /* padding-mode derivative (1 for zeros) */
                            giy *= (double)pr.my;
                            double Z = pr.Z;
                            double gP[3] = {gix / Z, giy / Z, 0.0};
                            if (pr.P[2] >= 1e-5f) gP[2] = -(gix * (double)pr.ax + giy * (double)pr.ay) / Z;
                            double gXs[3];
                            for (int j = 0; j < 3; j++) gXs[j] = (double)K[j] * gP[0] + (double)K[3 + j] * gP[1] + (double)K[6 + j] * gP[2];
                            for (int j = 0; j < 3; j++) {
                                loc[j * 4 + 0] += gXs[j] * (double)pr.Xc[0];
                                loc[j * 4 + 1] += gXs[j] * (double)pr.Xc[1];
                                loc[j * 4 + 2] += gXs[j] * (double)pr.Xc[2];
                                loc[j * 4 + 3] += gXs[j];
                            }
                            double gd = 0;
                            for (int j = 0; j < 3; j++) {
                                double gXc = (double)Rt[s][0 * 4 + j] * gXs[0] + (double)Rt[s][1 * 4 + j] * gXs[1] + (double)Rt[s][2 * 4 + j] * gXs[2];
                                gd += gXc * (double)pr.r[j];
                            }
                            if (inv[q] >= 1e-6f) ginv[q] += (float)(-(double)pr.d * (double)pr.d * gd);
                        }
#pragma omp critical
                    for (int k = 0; k < 12; k++) acc12[k] += loc[k];
                }
                for (int k = 0; k < 12; k++) gRt[s][k] += acc12[k];
            }
        }
        /* Euler chain: R = Rx Ry Rz (pose_utils.py:14-38), vec = (tx,ty,tz,rx,ry,rz) */
        for (int s = 0; s < S; s++) {
            if (in->pose_mats) {     /* matrix input: the gradient is dL/d(R|t) itself (grad_Rt); no Euler chain */
                if (grad_poses) for (int k = 0; k < 6; k++) grad_poses[((long)b * S + s) * 6 + k] = 0.0f;
                if (grad_Rt) memcpy(grad_Rt + ((long)b * S + s) * 12, gRt[s], sizeof(double) * 12);
                continue;
            }
            const float *vv = in->poses + ((long)b * S + s) * 6;
            double cx = cos((double)vv[3]), sx = sin((double)vv[3]);
            double cy = cos((double)vv[4]), sy = sin((double)vv[4]);
            double cz = cos((double)vv[5]), sz = sin((double)vv[5]);
            double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, dRx[9] = {0, 0, 0, 0, -sx, -cx, 0, cx, -sx};
            double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, dRy[9] = {-sy, 0, cy, 0, 0, 0, -cy, 0, -sy};
            double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1}, dRz[9] = {-sz, -cz, 0, cz, -sz, 0, 0, 0, 0};
            const double *Ms[3][3] = {{dRx, Ry, Rz}, {Rx, dRy, Rz}, {Rx, Ry, dRz}};
            float *gp = grad_poses + ((long)b * S + s) * 6;
            for (int a = 0; a < 3; a++) {
                double T[9], D[9];
                for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
                    double t = 0; for (int k = 0; k < 3; k++) t += Ms[a][0][r * 3 + k] * Ms[a][1][k * 3 + c]; T[r * 3 + c] = t; }
                for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
                    double t = 0; for (int k = 0; k < 3; k++) t += T[r * 3 + k] * Ms[a][2][k * 3 + c]; D[r * 3 + c] = t; }
                double g = 0;
                for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) g += gRt[s][r * 4 + c] * D[r * 3 + c];
                gp[3 + a] = (float)g;
            }
            for (int r = 0; r < 3; r++) gp[r] = (float)gRt[s][r * 4 + 3];
            if (grad_Rt) memcpy(grad_Rt + ((long)b * S + s) * 12, gRt[s], sizeof(double) * 12);
        }
    }
    for (int s = 0; s < S; s++) { free(xw[s]); free(coef[s]); }
    free(wx); free(wy);
    return 0;
}

int orc_abi_version(void) { return 2; }
