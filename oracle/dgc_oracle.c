/*
 * dgc_oracle.c -- CPU ORACLE for MGNet's DGC depth rescaling.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of
 *   mgnet/postprocessing/depth_post_proc.py:11-71   get_depth_prediction
 *   mgnet/postprocessing/depth_post_proc.py:74-104  _get_scale_recovery
 *   mgnet/postprocessing/depth_post_proc.py:107-151 _get_surface_normal
 *   mgnet/postprocessing/depth_post_proc.py:154-185 _get_ground_mask
 *   mgnet/geometry/camera.py:72-81,107-136          Camera.Kinv, Camera.reconstruct(frame="c")
 *   mgnet/postprocessing/exportable_post_proc.py:52-79  (same arithmetic, inverse camera matrix given)
 * plus the ATen CPU kernels those call (bmm, cross, normalize, replication_pad2d, cosine_similarity,
 * masked_select, median).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build or
 * call it; the product (mgnet_b200/) never does.
 *
 * Parity status: PINNED against outputs of the reference itself (imported from /root/reference in the build
 * container, torch 2.11 CPU) -- tests/golden/dgc_*.npz made by tests/golden/make_golden_dgc.py; the fp32 op
 * order below reproduces the reference's points, normals, camera heights, ground mask, median and scale
 * factor BIT FOR BIT (probed: torch.cross == fma(a1,b2,-(a2*b1)); F.normalize == v / max(sqrt((x*x+y*y)+z*z),
 * 1e-12); the 4-normal mean == (((n0+n1)+n2)+n3)/4; (P*N).sum(1) == (p0n0+p1n1)+p2n2 without contraction).
 *
 * Build: gcc -O2 -fopenmp -mfma -ffp-contract=off -shared -fPIC (oracle/oracle.py:build_dgc()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline void dgc_cross(const float a[3], const float b[3], float o[3])
{   /* ATen cross kernel: a1*b2 - a2*b1 compiled with contraction -> fma(a1, b2, -(a2*b1)) */
    o[0] = fmaf(a[1], b[2], -(a[2] * b[1]));
    o[1] = fmaf(a[2], b[0], -(a[0] * b[2]));
    o[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}

static inline void dgc_normalize(float v[3], float eps)
{   /* F.normalize(dim=1): v / max(||v||_2, eps) */
    float nn = sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    float d = nn > eps ? nn : eps;   /* clamp_min; NaN propagates like torch (NaN > eps is false -> eps: see note) */
    if (nn != nn) d = nn;
    v[0] = v[0] / d;
    v[1] = v[1] / d;
    v[2] = v[2] / d;
}

static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* Kinv of Camera (camera.py:72-81) or a caller-given inverse (exportable_post_proc.py:52-56). */
static void dgc_kinv(const float *cam, long cam_rs, int cam_is_inverse, float Kinv[9])
{
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Kinv[r * 3 + c] = cam[r * cam_rs + c];
    if (!cam_is_inverse) {
        float fx = Kinv[0], fy = Kinv[4], cx = Kinv[2], cy = Kinv[5];
        Kinv[0] = 1.0f / fx;
        Kinv[4] = 1.0f / fy;
        Kinv[2] = (-1.0f * cx) / fx;
        Kinv[5] = (-1.0f * cy) / fy;
    }
}

/*
 * One image.  Returns 0, or 1 when the ground mask is empty (torch.median raises on an empty tensor).
 *   depth      [H,W] in;   depth_out [H,W];  points_out [3,H,W] (NULL allowed)
 *   panoptic   [H,W] int64 or NULL (-> ground mask from the surface normals, depth_post_proc.py:154-185)
 *   filter_ids class ids whose depth is zeroed and whose points become NaN (depth_post_proc.py:61-69)
 *   optional diagnostics: normals_out [3,H,W], heights_out [H,W], ground_out [H,W] uint8
 */
int orc_dgc(int H, int W, const float *depth, const float *cam, long cam_rs, int cam_is_inverse, float real_height,
            const int64_t *panoptic, long long road_id, const long long *filter_ids, int n_filter, float *depth_out,
            float *points_out, float *normals_out, float *heights_out, uint8_t *ground_out, float *scale_out,
            long long *count_out)
{
    const long HW = (long)H * W;
    float Kinv[9];
    dgc_kinv(cam, cam_rs, cam_is_inverse, Kinv);
    float *P = (float *)malloc(sizeof(float) * 3 * HW);
    float *N = (float *)malloc(sizeof(float) * 3 * HW);
    uint32_t *keys = (uint32_t *)malloc(sizeof(uint32_t) * HW);
    /* Camera.reconstruct(depth, "c"): xnorm = Kinv.bmm(grid) (ascending FMA chain), Xc = xnorm * depth */
#pragma omp parallel for
    for (int v = 0; v < H; v++)
        for (int u = 0; u < W; u++) {
            float gu = (float)u, gv = (float)v, d = depth[(long)v * W + u];
            for (int j = 0; j < 3; j++) {
                float acc = Kinv[3 * j] * gu;
                acc = fmaf(Kinv[3 * j + 1], gv, acc);
                acc = fmaf(Kinv[3 * j + 2], 1.0f, acc);
                P[j * HW + (long)v * W + u] = acc * d;
            }
        }
    /* _get_surface_normal, nei = 1, interior pixels */
#pragma omp parallel for
    for (int v = 1; v < H - 1; v++)
        for (int u = 1; u < W - 1; u++) {
            float c[3], d[8][3];
            static const int off[8][2] = {{0, -1}, {-1, 0}, {0, 1}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}}; /* (dy,dx): x0,y0,x1,y1,x0y0,x0y1,x1y0,x1y1 */
            for (int j = 0; j < 3; j++) c[j] = P[j * HW + (long)v * W + u];
            for (int k = 0; k < 8; k++)
                for (int j = 0; j < 3; j++) d[k][j] = P[j * HW + (long)(v + off[k][0]) * W + (u + off[k][1])] - c[j];
            float n[4][3], m[3];
            dgc_cross(d[0], d[1], n[0]);
            dgc_cross(d[2], d[3], n[1]);
            dgc_cross(d[4], d[5], n[2]);
            dgc_cross(d[6], d[7], n[3]);
            for (int k = 0; k < 4; k++) dgc_normalize(n[k], 1e-12f);
            for (int j = 0; j < 3; j++) m[j] = (((n[0][j] + n[1][j]) + n[2][j]) + n[3][j]) / 4.0f;
            dgc_normalize(m, 1e-12f);
            for (int j = 0; j < 3; j++) N[j * HW + (long)v * W + u] = m[j];
        }
    /* F.pad(..., "replicate") */
    for (int v = 0; v < H; v++)
        for (int u = 0; u < W; u++) {
            int vc = v < 1 ? 1 : (v > H - 2 ? H - 2 : v), uc = u < 1 ? 1 : (u > W - 2 ? W - 2 : u);
            if (vc != v || uc != u)
                for (int j = 0; j < 3; j++) N[j * HW + (long)v * W + u] = N[j * HW + (long)vc * W + uc];
        }
    const float thr = (float)cos(5.0 * 3.14159265358979323846 / 180.0);
    long long cnt = 0;
    int has_nan = 0;
    for (long p = 0; p < HW; p++) {
        float n0 = N[p], n1 = N[HW + p], n2 = N[2 * HW + p];
        float h = fabsf((P[p] * n0 + P[HW + p] * n1) + P[2 * HW + p] * n2);
        int g;
        if (panoptic) {
            g = panoptic[p] == road_id;
        } else { /* cosine similarity with (0,1,0), eps 1e-6; |cos| > cos(5 deg); y > 0 */
            float nn = sqrtf((n0 * n0 + n1 * n1) + n2 * n2);
            float dn = nn > 1e-6f ? nn : 1e-6f;
            float x0 = n0 / dn, x1 = n1 / dn, x2 = n2 / dn;
            float cs = (x0 * 0.0f + x1 * 1.0f) + x2 * 0.0f;
            g = ((cs > thr) || (cs < -thr)) && !(P[HW + p] <= 0.0f);
        }
        if (heights_out) heights_out[p] = h;
        if (ground_out) ground_out[p] = (uint8_t)g;
        if (g) {
            if (h != h) has_nan = 1;
            memcpy(&keys[cnt], &h, 4);
            cnt++;
        }
    }
    if (normals_out) memcpy(normals_out, N, sizeof(float) * 3 * HW);
    if (count_out) *count_out = cnt;
    int rc = 0;
    float scale = NAN;
    if (cnt == 0) {
        rc = 1;
    } else if (!has_nan) {
        /* torch.median: the lower of the two middle elements; heights are >= 0 so the bit pattern orders them */
        qsort(keys, (size_t)cnt, sizeof(uint32_t), cmp_u32);
        float med;
        memcpy(&med, &keys[(cnt - 1) / 2], 4);
        scale = (1.0f / med) * real_height; /* torch.reciprocal(cam_height).mul_(real_cam_height) */
    }
    if (scale_out) *scale_out = scale;
#pragma omp parallel for
    for (long p = 0; p < HW; p++) {
        float d = depth[p] * scale;
        float x = P[p] * scale, y = P[HW + p] * scale, z = P[2 * HW + p] * scale;
        if (panoptic)
            for (int k = 0; k < n_filter; k++)
                if (panoptic[p] == filter_ids[k]) {
                    d = 0.0f;
                    x = y = z = NAN;
                }
        depth_out[p] = d;
        if (points_out) {
            points_out[p] = x;
            points_out[HW + p] = y;
            points_out[2 * HW + p] = z;
        }
    }
    free(P);
    free(N);
    free(keys);
    return rc;
}

int orc_dgc_abi_version(void) { return 1; }
