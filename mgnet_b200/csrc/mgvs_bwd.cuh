// mgvs_bwd.cuh -- fused backward kernel of the view-synthesis loss.
//
// Recomputes the forward from the same tiles (nothing but the uint8 selection and a handful of
// per-image scalars is carried over from the forward pass) and emits
//   * per-pixel inverse-depth gradients for every scale (each pixel written exactly once), and
//   * per-tile partial sums of dL/d(R|t) for both sources, reduced later in fixed order
//     (warp shuffle -> smem -> per-tile store -> pose_reduce_kernel; no float atomics).
//
// Per scale, one CTA (64x16 outputs) runs
//   A  warp both sources on tile+2 (exact forward arithmetic)                              -> smem
//   B  for every p on tile+1 whose argmin is a warped source: SSIM statistics of THAT source
//      (bit-identical to the forward, so the clamp decisions agree) -> the three coefficient maps
//      of the closed-form SSIM adjoint, one channel at a time               -> smem
//   C  4 outputs per thread: weighted 3x3 box adjoint (reflect-pad multiplicities) of the maps
//   D  L1 term, bilinear-sample adjoint, projection adjoint, depth gradient, pose partial sums,
//      plus the smoothness gradient
#pragma once
#include "mgvs_device.cuh"

namespace mgvs {

struct BwdParams {
    int B, H, W, n, automask;
    int pad;                      // padding_mode of the PAD kernels (1 border, 2 reflection)
    const float* tgt;
    const float* src[S];
    const float* inv[MAXN];
    const unsigned char* mask;
    const Cam* cams;
    const float4* psrc[S];        // packed RGBA + zero border copies of the sources (written by the forward)
    const unsigned char* sel;     // [n,B,H,W]
    const double* sums;           // [3n+3] global sums (N at [n], Nx at [3n+1], Ny at [3n+2])
    const double* imgsums;        // [B][4n+3] per-image sums from the forward (photo|smx|smy|invsum|N|Nx|Ny)
    const float* g_losses;        // [2]
    float* grad_inv[MAXN];
    float* pose_partials;         // [tiles][S*12]
    float alpha, oma, photo_w, smooth_w;
    int tiles_x, tiles_y;
};
struct BwdMaps {
    TmaDesc tgt, inv[MAXN];
};

constexpr int BWD_ROWS = TH + 4;                 // tile+2 halo rows
constexpr int BWD_CH = BWD_ROWS * PITCH;
constexpr int BWD_W2 = TW + 4;                   // tile+2 halo width; smem col j <-> image col x0-XOFF+j
constexpr int BWD_PROWS = TH + 2;                // tile+1 rows (coefficient maps)
constexpr int BWD_PW = TW + 2;
constexpr int BWD_MAP = BWD_PROWS * PITCH;       // coefficient texels per channel pass; col j <-> image col x0-XOFF+j (like fwd)
constexpr int BWD_PIT = (BWD_PROWS * BWD_PW + NT - 1) / NT;   // stage-B iterations per thread
constexpr int BWD_TILE3_FLOATS = (3 * BWD_CH + 31) / 32 * 32;
constexpr int BWD_INV_FLOATS = (BWD_CH + 31) / 32 * 32;
constexpr int BWD_NW = 13;                       // per-thread edge-aware weights kept across scales (see prologue)
constexpr int BWD_SMEM_FLOATS = BWD_TILE3_FLOATS + S * 3 * BWD_CH + 4 * BWD_MAP + 2 * BWD_INV_FLOATS + BWD_NW * NT + 8 * 24 + 48 + 4 * MAXN + 8;
constexpr int BWD_SMEM_BYTES = BWD_SMEM_FLOATS * 4;

__device__ __forceinline__ void bwd_load_tile(const float* __restrict__ img, float* __restrict__ dst, int x0, int y0,
                                              int H, int W, int tid)
{
    const int HW = H * W;
    for (int idx = tid; idx < 3 * BWD_ROWS * BWD_W2; idx += NT) {
        int ch = idx / (BWD_ROWS * BWD_W2);
        int r = idx - ch * (BWD_ROWS * BWD_W2);
        int hr = r / BWD_W2, hc = r - hr * BWD_W2;
        int v = reflect_idx(y0 - 2 + hr, H), u = reflect_idx(x0 - 2 + hc, W);
        dst[ch * BWD_CH + hr * PITCH + XOFF - 2 + hc] = __ldg(img + ch * HW + v * W + u);
    }
}

// 3x3 box adjoint (reflect-pad multiplicities on border tiles) of the three coefficient maps of one channel
// for 4 adjacent outputs, split by the source each tap belongs to.  One 128-bit smem load per tap brings
// (a, b, c, [source==0]); out[s][m][k] = sum over the window of [source_p == s] * w_pq * coef_m(p).
template <bool WEIGHTED>
__device__ __forceinline__ void box_adjoint4(const float4* __restrict__ mp, const float rwgt[3], const float (*cwgt)[3],
                                             float out[S][3][4])
{
    float col[S][3][6];
#pragma unroll
    for (int dy = 0; dy < 3; dy++) {
        const float4* rr = mp + dy * PITCH;              // mp -> [map row of image row v-1][my output 0]
#pragma unroll
        for (int j = 0; j < 6; j++) {
            float4 t = rr[j - 1];
            float m0 = WEIGHTED ? t.w * rwgt[dy] : t.w;
            float m1 = WEIGHTED ? rwgt[dy] - m0 : 1.0f - t.w;
            const float c3[3] = {t.x, t.y, t.z};
#pragma unroll
            for (int m = 0; m < 3; m++) {
                col[0][m][j] = dy == 0 ? m0 * c3[m] : fmaf(m0, c3[m], col[0][m][j]);
                col[1][m][j] = dy == 0 ? m1 * c3[m] : fmaf(m1, c3[m], col[1][m][j]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < S; s++)
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (WEIGHTED) out[s][m][k] = cwgt[k][0] * col[s][m][k] + cwgt[k][1] * col[s][m][k + 1] + cwgt[k][2] * col[s][m][k + 2];
                else out[s][m][k] = (col[s][m][k] + col[s][m][k + 1]) + col[s][m][k + 2];
            }
}

template <bool USE_TMA, bool PAD = false, bool L1ONLY = false>
__global__ void __launch_bounds__(NT, MIN_CTAS) bwd_kernel(const BwdParams p, const __grid_constant__ BwdMaps maps)
{
    extern __shared__ __align__(128) float smem[];
    float* sY = smem;                               // [3][BWD_ROWS][PITCH]
    float* sX = sY + BWD_TILE3_FLOATS;              // [S][3][BWD_ROWS][PITCH]
    float4* sCo = reinterpret_cast<float4*>(sX + S * 3 * BWD_CH);   // [BWD_PROWS][PITCH] texels (a, b, c, [source==0]) of
                                                                    // the current channel: one 128-bit load per tap
    float* sInv = sX + S * 3 * BWD_CH + 4 * BWD_MAP; // [2][BWD_ROWS][PITCH] inverse-depth ring (TMA path)
    float* sW = sInv + 2 * BWD_INV_FLOATS;          // [13][NT] masked smoothness weights of my 4 outputs
    float* sRed = sW + BWD_NW * NT;                 // [8 warps][24]
    float* sCam = sRed + 8 * 24;                    // 48
    float* sSm = sCam + 48;                         // [n][4]: Ws/(Nx c), Ws/(Ny c), mean term, unused
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sSm + 4 * MAXN);   // [0] target tile, [1],[2] inverse-depth ring

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tpi = p.tiles_x * p.tiles_y;
    const int b = tile / tpi;
    const int trem = tile - b * tpi;
    const int tyi = trem / p.tiles_x, txi = trem - tyi * p.tiles_x;
    const int x0 = txi * TW, y0 = tyi * TH;
    const int H = p.H, W = p.W, HW = H * W;
    const int tx = tid % CG, ty = tid / CG;
    const int u0 = x0 + 4 * tx, v = y0 + ty;

    const bool border = (x0 == 0) || (y0 == 0) || (x0 + TW + 1 >= W) || (y0 + TH + 1 >= H);
    if (USE_TMA && tid == 0) {
        tma::mbar_init(sBar + 0, 1); tma::mbar_init(sBar + 1, 1); tma::mbar_init(sBar + 2, 1);
        tma::fence_barrier_init();
        tma::mbar_expect_tx(sBar + 0, 3 * BWD_CH * 4);
        tma::load_3d(sY, &maps.tgt, x0 - XOFF, y0 - 2, 3 * b, sBar + 0);
        tma::mbar_expect_tx(sBar + 1, BWD_CH * 4);
        tma::load_3d(sInv, &maps.inv[0], x0 - XOFF, y0 - 2, b, sBar + 1);
    }
    if (tid < 48) sCam[tid] = reinterpret_cast<const float*>(p.cams + b)[tid];
    const int nq = 4 * p.n + 3;
    const double Ntot = p.sums[p.n], Nx = p.sums[3 * p.n + 1], Ny = p.sums[3 * p.n + 2];
    const float g_photo = __ldg(p.g_losses), g_smooth = __ldg(p.g_losses + 1);
    if (tid < p.n) {
        // smoothness constants of (image b, scale tid): SURVEY App. B-6
        const double* is = p.imgsums + (size_t)b * nq;
        double mean = is[3 * p.n + tid] / (double)HW;
        bool active = mean >= 1e-6;
        double c = active ? mean : 1e-6;
        double Ws = (double)g_smooth * (double)p.smooth_w / ((double)p.n * (double)(1 << tid));
        double A = is[p.n + tid] / Nx + is[2 * p.n + tid] / Ny;     // un-normalised sums / counts
        sSm[tid * 4 + 0] = (float)(Ws / (Nx * c));
        sSm[tid * 4 + 1] = (float)(Ws / (Ny * c));
        sSm[tid * 4 + 2] = active ? (float)(-Ws * A / (c * c * (double)HW)) : 0.f;
    }
    if (USE_TMA) {
        __syncthreads();                 // barrier init, camera table and smoothness constants visible
        tma::mbar_wait(sBar + 0, 0);
        if (border) {
            patch_reflect<2, BWD_ROWS>(sY, 3, BWD_CH, x0, y0, H, W, tid, NT);
            __syncthreads();
        }
    } else {
        bwd_load_tile(p.tgt + (size_t)b * 3 * HW, sY, x0, y0, H, W, tid);
        __syncthreads();
    }

    const float* K = sCam;
    const float* Kinv = sCam + 9;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const float rw = exact::rcp_refined(wm1), rh = exact::rcp_refined(hm1);
    const float Wp = (float)((double)g_photo * (double)p.photo_w / ((double)p.n * Ntot));
    const float cf_ssim = Wp * p.alpha * (1.0f / 3.0f) * (-0.5f) * (1.0f / 9.0f);   // u/9 of App. B-3
    // L1ONLY (ssim_loss_weight == 0, loss.py:195-196): raw per-channel |x - y|, no channel mean, no (1 - alpha)
    const float cf_l1 = L1ONLY ? Wp : Wp * p.oma * (1.0f / 3.0f);

    const size_t pimg = (size_t)(H + 2 * PACK_BORDER) * (W + 2 * PACK_BORDER);
    const float4* src0 = p.psrc[0] + (size_t)b * pimg;
    const float4* src1 = p.psrc[1] + (size_t)b * pimg;
    const int Wpk = W + 2 * PACK_BORDER;

    bool valid[4], msk[4];
    {
        unsigned mw = 0x01010101u;
        if (p.mask != nullptr && v < H) {
            const unsigned char* mp = p.mask + (size_t)b * HW + (size_t)v * W + u0;
            if (u0 + 3 < W && ((W & 3) == 0)) mw = *reinterpret_cast<const unsigned*>(mp);
            else {
                mw = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) if (u0 + k < W) mw |= (unsigned)(mp[k] != 0) << (8 * k);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            valid[k] = (v < H) && (u0 + k < W);
            msk[k] = valid[k] && ((mw >> (8 * k)) & 0xffu) != 0;
        }
    }
    // reflect-pad multiplicities of the box adjoint: tap (q+d) counts twice when its padded twin folds onto q
    float rwgt[3], cwgt[4][3];
    rwgt[0] = (v - 1 >= 0) ? ((v == 1) ? 2.f : 1.f) : 0.f;
    rwgt[1] = 1.f;
    rwgt[2] = (v + 1 <= H - 1) ? ((v == H - 2) ? 2.f : 1.f) : 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int u = u0 + k;
        cwgt[k][0] = (u - 1 >= 0) ? ((u == 1) ? 2.f : 1.f) : 0.f;
        cwgt[k][1] = 1.f;
        cwgt[k][2] = (u + 1 <= W - 1) ? ((u == W - 2) ? 2.f : 1.f) : 0.f;
    }

    // Edge-aware weights exp(-mean_c|dI|) (depth.py:23-24) of the four pixel pairs each output takes part in,
    // already multiplied by the mask of the pair's owner (left / upper pixel, loss.py:285-286) and zero where
    // the pair does not exist.  Scale independent: computed once, parked in smem (slot j of thread tid).
    //   [0..3] pair (q, q+1)   [4] pair (q0-1, q0)   [5..8] pair (q, q+W)   [9..12] pair (q-W, q)
    {
        auto wgt = [&](const float* a, const float* b) {
            float d = (fabsf(a[0] - b[0]) + fabsf(a[BWD_CH] - b[BWD_CH])) + fabsf(a[2 * BWD_CH] - b[2 * BWD_CH]);
            return expf(-exact::div3(d));
        };
        auto mask_at = [&](int vv, int uu) {
            return vv >= 0 && vv < H && uu >= 0 && uu < W && (p.mask == nullptr || p.mask[(size_t)b * HW + (size_t)vv * W + uu] != 0);
        };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float* yc = sY + (ty + 2) * PITCH + XOFF + 4 * tx + k;
            int u = u0 + k;
            sW[(0 + k) * NT + tid] = (msk[k] && u + 1 < W) ? wgt(yc, yc + 1) : 0.f;
            sW[(5 + k) * NT + tid] = (msk[k] && v + 1 < H) ? wgt(yc, yc + PITCH) : 0.f;
            sW[(9 + k) * NT + tid] = (valid[k] && mask_at(v - 1, u)) ? wgt(yc - PITCH, yc) : 0.f;
            if (k == 0) sW[4 * NT + tid] = (valid[0] && mask_at(v, u - 1)) ? wgt(yc - 1, yc) : 0.f;
        }
    }

    // bit it of pmask: my it-th stage-B pixel lies inside the image and its mask is true (scale independent)
    unsigned pmask = 0;
#pragma unroll
    for (int it = 0; it < BWD_PIT; it++) {
        int h = tid + it * NT;
        if (h < BWD_PROWS * BWD_PW) {
            int pr = h / BWD_PW, pc = h - pr * BWD_PW;
            int pv = y0 - 1 + pr, pu = x0 - 1 + pc;
            if (pv >= 0 && pv < H && pu >= 0 && pu < W && (p.mask == nullptr || p.mask[(size_t)b * HW + (size_t)pv * W + pu] != 0))
                pmask |= 1u << it;
        }
    }

    float pacc[S][12];
#pragma unroll
    for (int s = 0; s < S; s++)
#pragma unroll
        for (int j = 0; j < 12; j++) pacc[s][j] = 0.f;

    for (int i = 0; i < p.n; i++) {
        const float* inv = p.inv[i] + (size_t)b * HW;
        const unsigned char* sel = p.sel + ((size_t)i * p.B + b) * HW;

        const float* sI = sInv + (i & 1) * BWD_INV_FLOATS;
        if (USE_TMA) {
            if (tid == 0 && i + 1 < p.n) {
                tma::fence_proxy_async();
                tma::mbar_expect_tx(sBar + 1 + ((i + 1) & 1), BWD_CH * 4);
                tma::load_3d(sInv + ((i + 1) & 1) * BWD_INV_FLOATS, &maps.inv[i + 1], x0 - XOFF, y0 - 2, b, sBar + 1 + ((i + 1) & 1));
            }
            tma::mbar_wait(sBar + 1 + (i & 1), (i >> 1) & 1);
        }
        // which source (0/1) is selected at each of my stage-B pixels; 2 = none.  Loaded BEFORE stage A so
        // the global-load latency hides behind the warp stage (the mask bits were fetched once per tile).
        unsigned char psel[BWD_PIT];
#pragma unroll
        for (int it = 0; it < BWD_PIT; it++) {
            unsigned char code = 2;
            if ((pmask >> it) & 1u) {
                int h = tid + it * NT;
                int pr = h / BWD_PW, pc = h - pr * BWD_PW;
                unsigned k = sel[(size_t)(y0 - 1 + pr) * W + (x0 - 1 + pc)];
                if (p.automask) { if ((k & 1u) == 0) code = (unsigned char)(k >> 1); }
                else code = (unsigned char)k;
            }
            psel[it] = code;
        }
        unsigned selq = 0;      // argmin codes of my 4 outputs (stage D)
        if (v < H) {
            const unsigned char* sq = sel + (size_t)v * W + u0;
            if (u0 + 3 < W && ((W & 3) == 0)) selq = *reinterpret_cast<const unsigned*>(sq);
            else {
#pragma unroll
                for (int k = 0; k < 4; k++) if (u0 + k < W) selq |= (unsigned)sq[k] << (8 * k);
            }
        }
        // ---- stage A: warp both sources on tile+2 (software pipelined, mgvs_device.cuh) ----
#ifndef MGVS_SKIP_A
        warp_tile<2, BWD_ROWS, BWD_CH, USE_TMA, PAD>(sX, sX + 3 * BWD_CH, sI, inv, src0, src1, sCam, x0, y0, H, W, border,
                                                      wm1, hm1, rw, rh, tid, p.pad);
#endif
        __syncthreads();

        float G[S][3][4];
        if constexpr (L1ONLY) {      // no SSIM term: the box adjoint is zero
#pragma unroll
            for (int s = 0; s < S; s++)
#pragma unroll
                for (int ch = 0; ch < 3; ch++)
#pragma unroll
                    for (int k = 0; k < 4; k++) G[s][ch][k] = 0.f;
        } else {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            // ---- stage B: coefficient maps of channel ch ----
#pragma unroll
            for (int it = 0; it < BWD_PIT; it++) {
                int h = tid + it * NT;
                if (h < BWD_PROWS * BWD_PW) {
                    int pr = h / BWD_PW, pc = h - pr * BWD_PW;
                    float ca = 0.f, cb = 0.f, cc = 0.f;
                    unsigned s = psel[it];
#ifdef MGVS_SKIP_B
                    if (s < 2) { ca = sX[pr * PITCH + pc]; cb = ca; cc = ca; }
                    if (false) {
#else
                    if (s < 2) {
#endif
                        // window rows pr..pr+2, cols pc..pc+2 in tile+2 coordinates
                        const float* xw = sX + s * 3 * BWD_CH + ch * BWD_CH + pr * PITCH + XOFF - 2 + pc;
                        const float* yw = sY + ch * BWD_CH + pr * PITCH + XOFF - 2 + pc;
                        float sx, sxx, sxy, sy, syy;
#pragma unroll
                        for (int dy = 0; dy < 3; dy++)
#pragma unroll
                            for (int dx = 0; dx < 3; dx++) {
                                float xv = xw[dy * PITCH + dx], yv = yw[dy * PITCH + dx];
                                float xx = __fmul_rn(xv, xv), xy = __fmul_rn(xv, yv), yy = __fmul_rn(yv, yv);
                                if (dy == 0 && dx == 0) { sx = xv; sxx = xx; sxy = xy; sy = yv; syy = yy; }
                                else {
                                    sx = __fadd_rn(sx, xv); sxx = __fadd_rn(sxx, xx); sxy = __fadd_rn(sxy, xy);
                                    sy = __fadd_rn(sy, yv); syy = __fadd_rn(syy, yy);
                                }
                            }
                        float mu_y = exact::div9(sy);
                        float mys = __fmul_rn(mu_y, mu_y);
                        float sgy = __fadd_rn(exact::div9(syy), -mys);
                        exact::Ssim q;
                        (void)exact::ssim_from_sums(sx, sxx, sxy, mu_y, mys, sgy, &q);
                        if (q.loss_raw >= 0.f && q.loss_raw <= 1.f) {      // clamp passes gradient inclusively
                            float id1 = exact::rcp_refined(q.d1), id2 = exact::rcp_refined(q.d2);
                            float idd = id1 * id2;
                            float ds_dmux = 2.f * mu_y * (q.n2 - q.n1) * idd - q.ssim * 2.f * q.mu_x * (id1 - id2);
                            float ds_dexx = -q.ssim * id2;
                            float ds_dexy = 2.f * q.n1 * idd;
                            ca = cf_ssim * ds_dmux;
                            cb = cf_ssim * 2.f * ds_dexx;
                            cc = cf_ssim * ds_dexy;
                        }
                    }
                    sCo[pr * PITCH + XOFF - 1 + pc] = make_float4(ca, cb, cc, s == 0 ? 1.f : 0.f);   // zeros when unselected
                }
            }
            __syncthreads();
            // ---- stage C: 3x3 box adjoint (with reflect-pad multiplicities on border tiles) for my 4 outputs ----
            {
                float box[S][3][4];
                const float4* mp = sCo + ty * PITCH + XOFF + 4 * tx;
#ifdef MGVS_SKIP_C
                for (int s = 0; s < S; s++) for (int m = 0; m < 3; m++) for (int k = 0; k < 4; k++) box[s][m][k] = mp[k].x;
#else
                if (border) box_adjoint4<true>(mp, rwgt, cwgt, box);
                else box_adjoint4<false>(mp, rwgt, cwgt, box);
#endif
#pragma unroll
                for (int s = 0; s < S; s++)
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float xq = sX[s * 3 * BWD_CH + ch * BWD_CH + (ty + 2) * PITCH + XOFF + 4 * tx + k];
                        float yq = sY[ch * BWD_CH + (ty + 2) * PITCH + XOFF + 4 * tx + k];
                        G[s][ch][k] = box[s][0][k] + xq * box[s][1][k] + yq * box[s][2][k];
                    }
            }
            __syncthreads();   // maps free for the next channel
        }

        }   // !L1ONLY

        // ---- stage D: per-output chain ----
        float ginv[4];
        // smoothness gradient (App. B-6): d/dinv of sum m*w*|inv_p - inv_q| / (N*c) plus the mean term
        {
            const float kx = sSm[i * 4 + 0], ky = sSm[i * 4 + 1], mt = sSm[i * 4 + 2];
            float ic[6], iu[4], id[4];      // inverse depth: row v cols -1..4, row v-1, row v+1
            if (USE_TMA) {
                const float* iq = sI + (ty + 2) * PITCH + XOFF + 4 * tx;
                float4 c4 = *reinterpret_cast<const float4*>(iq);
                float4 u4 = *reinterpret_cast<const float4*>(iq - PITCH);
                float4 d4 = *reinterpret_cast<const float4*>(iq + PITCH);
                ic[0] = iq[-1]; ic[1] = c4.x; ic[2] = c4.y; ic[3] = c4.z; ic[4] = c4.w; ic[5] = iq[4];
                iu[0] = u4.x; iu[1] = u4.y; iu[2] = u4.z; iu[3] = u4.w;
                id[0] = d4.x; id[1] = d4.y; id[2] = d4.z; id[3] = d4.w;
            } else {
#pragma unroll
                for (int k = -1; k < 5; k++) { int u = u0 + k; ic[k + 1] = (v < H && u >= 0 && u < W) ? __ldg(inv + (size_t)v * W + u) : 0.f; }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    int u = u0 + k;
                    iu[k] = (v >= 1 && v - 1 < H && u < W) ? __ldg(inv + (size_t)(v - 1) * W + u) : 0.f;
                    id[k] = (v + 1 < H && u < W) ? __ldg(inv + (size_t)(v + 1) * W + u) : 0.f;
                }
            }
            auto sgn = [](float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); };
            float wl = sW[4 * NT + tid];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float wr = sW[(0 + k) * NT + tid], wd = sW[(5 + k) * NT + tid], wu = sW[(9 + k) * NT + tid];
                float c = ic[k + 1];
                float g = kx * (wr * sgn(c - ic[k + 2]) - wl * sgn(ic[k] - c)) + ky * (wd * sgn(c - id[k]) - wu * sgn(iu[k] - c));
                ginv[k] = valid[k] ? mt + g : 0.f;
                wl = wr;      // pair (q_k, q_k+1) is the left pair of output k+1 (same owner mask)
            }
        }
#pragma unroll
        for (int s = 0; s < S; s++) {
            const float* Rt = sCam + 18 + 12 * s;
            const float4* sp = s == 0 ? src0 : src1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (!valid[k]) continue;
                unsigned code = (selq >> (8 * k)) & 0xffu;
                bool selme = msk[k] && (p.automask ? (code == 2u * s) : (code == (unsigned)s));
                unsigned chsel = 3u;         // L1ONLY: the one channel the min picked (code = list entry * 3 + channel)
                if constexpr (L1ONLY) {
                    const unsigned entry = code / 3u;
                    chsel = code - 3u * entry;
                    selme = msk[k] && (p.automask ? (entry == 2u * s) : (entry == (unsigned)s));
                }
                float g0 = G[s][0][k], g1 = G[s][1][k], g2 = G[s][2][k];
                if (selme) {
                    const float* xq = sX + s * 3 * BWD_CH + (ty + 2) * PITCH + XOFF + 4 * tx + k;
                    const float* yq = sY + (ty + 2) * PITCH + XOFF + 4 * tx + k;
                    float d0 = xq[0] - yq[0], d1 = xq[BWD_CH] - yq[BWD_CH], d2 = xq[2 * BWD_CH] - yq[2 * BWD_CH];
                    if (!L1ONLY || chsel == 0u) g0 += cf_l1 * (d0 > 0.f ? 1.f : (d0 < 0.f ? -1.f : 0.f));
                    if (!L1ONLY || chsel == 1u) g1 += cf_l1 * (d1 > 0.f ? 1.f : (d1 < 0.f ? -1.f : 0.f));
                    if (!L1ONLY || chsel == 2u) g2 += cf_l1 * (d2 > 0.f ? 1.f : (d2 < 0.f ? -1.f : 0.f));
                }
                if (g0 == 0.f && g1 == 0.f && g2 == 0.f) continue;
#ifdef MGVS_SKIP_D
                ginv[k] += g0 + g1 + g2; continue;
#endif
                int u = u0 + k;
                float r[3], Xc[3];
                exact::ray(Kinv, u, v, r);
                float invq = USE_TMA ? sI[(ty + 2) * PITCH + XOFF + 4 * tx + k] : __ldg(inv + (size_t)v * W + u);
                float d = exact::rcp_refined(fmaxf(invq, 1e-6f));
#pragma unroll
                for (int j = 0; j < 3; j++) Xc[j] = __fmul_rn(r[j], d);
                exact::Proj pr;
                exact::project<PAD>(K, Rt, Xc, wm1, hm1, rw, rh, pr, p.pad);
                // bilinear adjoint (GridSampler backward w.r.t. the grid): out-of-image corners are zeros of the border
                float xw = floorf(pr.ix), yn = floorf(pr.iy);
                float wE = pr.ix - xw, wW = 1.0f - wE, wS = pr.iy - yn, wN = 1.0f - wS;
                int cx0 = (int)fminf(fmaxf(xw, -2.0f), (float)W), cy0 = (int)fminf(fmaxf(yn, -2.0f), (float)H);
                const float4* pc = sp + (cy0 + PACK_BORDER) * Wpk + (cx0 + PACK_BORDER);
                float4 nw = __ldg(pc), ne = __ldg(pc + 1), sw = __ldg(pc + Wpk), se = __ldg(pc + Wpk + 1);
                float gix = g0 * ((ne.x - nw.x) * wN + (se.x - sw.x) * wS) + g1 * ((ne.y - nw.y) * wN + (se.y - sw.y) * wS) +
                            g2 * ((ne.z - nw.z) * wN + (se.z - sw.z) * wS);
                float giy = g0 * ((sw.x - nw.x) * wW + (se.x - ne.x) * wE) + g1 * ((sw.y - nw.y) * wW + (se.y - ne.y) * wE) +
                            g2 * ((sw.z - nw.z) * wW + (se.z - ne.z) * wE);
                if constexpr (PAD) { gix *= pr.mx; giy *= pr.my; }   // padding-mode derivative (clip / reflect)
                // projection adjoint (App. B-5)
                float iz = exact::rcp_refined(pr.Z);
                float gP0 = gix * iz, gP1 = giy * iz;
                float gP2 = (pr.Pz >= 1e-5f) ? -(gix * pr.ax + giy * pr.ay) * iz : 0.f;
                float gX0 = K[0] * gP0 + K[3] * gP1 + K[6] * gP2;
                float gX1 = K[1] * gP0 + K[4] * gP1 + K[7] * gP2;
                float gX2 = K[2] * gP0 + K[5] * gP1 + K[8] * gP2;
                pacc[s][0] += gX0 * pr.Xc0; pacc[s][1] += gX0 * pr.Xc1; pacc[s][2] += gX0 * pr.Xc2; pacc[s][3] += gX0;
                pacc[s][4] += gX1 * pr.Xc0; pacc[s][5] += gX1 * pr.Xc1; pacc[s][6] += gX1 * pr.Xc2; pacc[s][7] += gX1;
                pacc[s][8] += gX2 * pr.Xc0; pacc[s][9] += gX2 * pr.Xc1; pacc[s][10] += gX2 * pr.Xc2; pacc[s][11] += gX2;
                float gd = 0.f;
#pragma unroll
                for (int j = 0; j < 3; j++) gd += (Rt[j] * gX0 + Rt[4 + j] * gX1 + Rt[8 + j] * gX2) * r[j];
                if (invq >= 1e-6f) ginv[k] -= d * d * gd;
            }
        }
        if (v < H) {
            float* go = p.grad_inv[i] + (size_t)b * HW + (size_t)v * W + u0;
            if (u0 + 3 < W && ((W & 3) == 0)) *reinterpret_cast<float4*>(go) = make_float4(ginv[0], ginv[1], ginv[2], ginv[3]);
            else {
#pragma unroll
                for (int k = 0; k < 4; k++) if (u0 + k < W) go[k] = ginv[k];
            }
        }
        // (the next scale's stage A overwrites sX: every reader of sX in stage D is done only after
        //  this barrier)
        __syncthreads();
    }

    // deterministic pose partials: shuffle tree -> smem -> fixed-order sum -> one store per tile
#pragma unroll
    for (int s = 0; s < S; s++)
#pragma unroll
        for (int j = 0; j < 12; j++) {
            float vsum = warp_sum(pacc[s][j]);
            if (lane == 0) sRed[warp * 24 + s * 12 + j] = vsum;
        }
    __syncthreads();
    if (tid < 24) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) acc += sRed[w * 24 + tid];
        p.pose_partials[(size_t)tile * 24 + tid] = acc;
    }
}

}  // namespace mgvs
