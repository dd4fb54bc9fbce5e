// mgvs_bwd_stash.cuh -- backward kernel of the view-synthesis loss that consumes the coefficient stash.
//
// The recompute backward (mgvs_bwd.cuh) spends ~60 % of its instructions rebuilding what the forward already
// had in registers: both warped sources on tile+2 and the SSIM statistics of the selected source on tile+1.
// On B200 the path is issue-bound and HBM is >90 % idle, so the forward can afford to write the three
// coefficients of the closed-form SSIM adjoint of the selected source per (scale, channel,
// pixel) -- 48 B/px/scale as float4 texels (a, b, c, [source == 0]) -- and this kernel only runs
//   C  3x3 box adjoint of the coefficient maps (reflect-pad multiplicities on border tiles), per source
//   D  per output: exact re-projection, bilinear gather (re-blend of x_q), L1 term, bilinear adjoint,
//      projection adjoint, depth gradient, pose partial sums; plus the smoothness gradient.
// Texels arrive by TMA (5-D map over the phase-major stash [plane][row][u & 3][u >> 2][4]) straight into the
// layout stage C reads: for a fixed tap the 16 threads of a tile row load 16 consecutive float4 -> no bank
// conflicts (the recompute kernel's interleaved layout costs 4 wavefronts per quarter-warp here).
//
// Same determinism rules as everywhere else: no float atomics, fixed-order pose partials.
#pragma once
#include <type_traits>

#include "mgvs_device.cuh"

namespace mgvs {

struct BwdSParams {
    int B, H, W, n, automask;
    const float* tgt;
    const float* inv[MAXN];
    const unsigned char* mask;
    const Cam* cams;
    const float4* psrc[S];        // packed RGBA + zero border copies of the sources (written by the forward)
    const unsigned char* sel;     // [n,B,H,W]
    const double* sums;           // [3n+3] global sums
    const double* imgsums;        // [B][4n+3] per-image sums from the forward
    const float* g_losses;        // [2]
    float* grad_inv[MAXN];
    float* pose_partials;         // [tiles][S*12]
    float alpha, oma, photo_w, smooth_w;
    int tiles_x, tiles_y;
};
struct BwdSMaps {
    TmaDesc tgt, inv[MAXN], coef;
};

constexpr int BS_ROWS = TH + 2;                       // tile+1 rows
constexpr int BS_CH = BS_ROWS * PITCH;                // floats per plane (target / inverse depth), col j <-> image col x0-XOFF+j
constexpr int BS_GROUPS = PITCH / 4;                  // 18 column groups per texel row (image cols x0-4 .. x0+67)
constexpr int BS_TEX = BS_ROWS * PITCH;               // texels per channel map: [row][phase][group]
constexpr int BS_TILE3_FLOATS = (3 * BS_CH + 31) / 32 * 32;
constexpr int BS_INV_FLOATS = (BS_CH + 31) / 32 * 32;
constexpr int BS_NW = 13;
constexpr int BS_SMEM_FLOATS = 3 * BS_TEX * 4 + BS_TILE3_FLOATS + 2 * BS_INV_FLOATS + BS_NW * NT + 8 * 24 + 48 + 4 * MAXN + 8;
constexpr int BS_SMEM_BYTES = BS_SMEM_FLOATS * 4;
static_assert((BS_TEX * 16) % 128 == 0, "channel maps must stay 128-byte aligned for TMA");

namespace tma {
// 5-D tiled load: box origin (c0..c4) in elements
__device__ __forceinline__ void load_5d(void* dst, const void* tmap, int c0, int c1, int c2, int c3, int c4, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
}
}  // namespace tma

// 3x3 box adjoint of the three coefficient maps of one channel for 4 adjacent outputs, restricted to the taps
// whose argmin is source SRC.  row0 -> texel row of image row v-1; the strip's taps are columns -1..4 relative
// to output 0 = (phase 3, group tx), (phases 0..3, group tx+1), (phase 0, group tx+2).
template <int SRC, bool WEIGHTED>
__device__ __forceinline__ void box_adjoint_src(const float4* __restrict__ row0, int tx, const float rwgt[3],
                                                const float (*cwgt)[3], const float* __restrict__ yq, float out[2][4])
{   // out[0][k] = A_k + y_q * C_k, out[1][k] = B_k   (G = out[0] + x_q * out[1]; y_q is known without the gather)
    float col[3][6];
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
        for (int j = 0; j < 6; j++) col[m][j] = 0.f;
    // a real loop over the three tap rows: fully unrolled, ptxas hoists all 18 128-bit loads of a channel (and of the
    // next one) above the arithmetic and spills them
#pragma unroll 1
    for (int dy = 0; dy < 3; dy++) {
        const float4* rr = row0 + dy * PITCH + tx;
        const float rwt = WEIGHTED ? (dy == 0 ? rwgt[0] : (dy == 1 ? rwgt[1] : rwgt[2])) : 1.0f;
#pragma unroll
        for (int j = 0; j < 6; j++) {
            // j = 0: phase 3 of group tx; j = 1..4: phases 0..3 of group tx+1; j = 5: phase 0 of group tx+2
            const int off = j == 0 ? 3 * BS_GROUPS : (j == 5 ? 2 : (j - 1) * BS_GROUPS + 1);
            float4 t = rr[off];
            float m = SRC == 0 ? t.w : 1.0f - t.w;
            if (WEIGHTED) m *= rwt;
            col[0][j] = fmaf(m, t.x, col[0][j]);
            col[1][j] = fmaf(m, t.y, col[1][j]);
            col[2][j] = fmaf(m, t.z, col[2][j]);
        }
    }
    const float4 y4 = *reinterpret_cast<const float4*>(yq);
    const float ys[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float o[3];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            if (WEIGHTED) o[m] = cwgt[k][0] * col[m][k] + cwgt[k][1] * col[m][k + 1] + cwgt[k][2] * col[m][k + 2];
            else o[m] = (col[m][k] + col[m][k + 1]) + col[m][k + 2];
        }
        out[0][k] = fmaf(ys[k], o[2], o[0]);
        out[1][k] = o[1];
    }
}

template <bool USE_TMA>
__global__ void __launch_bounds__(NT, MIN_CTAS) bwd_stash_kernel(const __grid_constant__ BwdSParams p, const __grid_constant__ BwdSMaps maps)
{
    extern __shared__ __align__(128) float smem[];
    float4* sCo = reinterpret_cast<float4*>(smem);        // [3 ch][BS_ROWS][4 phases][BS_GROUPS] texels of the current scale
    float* sY = smem + 3 * BS_TEX * 4;                     // [3][BS_ROWS][PITCH]
    float* sInv = sY + BS_TILE3_FLOATS;                    // [2][BS_ROWS][PITCH] inverse-depth ring
    float* sW = sInv + 2 * BS_INV_FLOATS;                  // [13][NT] masked smoothness weights of my 4 outputs
    float* sRed = sW + BS_NW * NT;                         // [8 warps][24]
    float* sCam = sRed + 8 * 24;                           // 48
    float* sSm = sCam + 48;                                // [n][4]
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sSm + 4 * MAXN);   // [0] target, [1],[2] inverse-depth ring, [3] coefficients

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

    // reflect-pad multiplicities are only needed where a 3x3 window of the tile can touch the image border
    const bool border = (x0 == 0) || (y0 == 0) || (x0 + TW + 1 >= W) || (y0 + TH + 1 >= H);
    constexpr uint32_t COEF_BYTES = 3 * BS_TEX * 16;
    auto load_coef = [&](int i) {      // thread 0 only: the three channel maps of scale i
        tma::mbar_expect_tx(sBar + 3, COEF_BYTES);
#pragma unroll
        for (int ch = 0; ch < 3; ch++)
            tma::load_5d(sCo + ch * BS_TEX, &maps.coef, 0, (x0 >> 2) - 1, 0, y0 - 1, (i * p.B + b) * 3 + ch, sBar + 3);
    };
    auto load_inv_manual = [&](int i, float* dst) {
        const float* inv = p.inv[i] + (size_t)b * HW;
        for (int idx = tid; idx < BS_CH; idx += NT) {
            int r = idx / PITCH, j = idx - r * PITCH;
            int vv = y0 - 1 + r, uu = x0 - XOFF + j;
            dst[idx] = (vv >= 0 && vv < H && uu >= 0 && uu < W) ? __ldg(inv + (size_t)vv * W + uu) : 0.f;
        }
    };
    if (tid == 0) {
        tma::mbar_init(sBar + 0, 1); tma::mbar_init(sBar + 1, 1); tma::mbar_init(sBar + 2, 1); tma::mbar_init(sBar + 3, 1);
        tma::fence_barrier_init();
        if (USE_TMA) {
            tma::mbar_expect_tx(sBar + 0, 3 * BS_CH * 4);
            tma::load_3d(sY, &maps.tgt, x0 - XOFF, y0 - 1, 3 * b, sBar + 0);
            tma::mbar_expect_tx(sBar + 1, BS_CH * 4);
            tma::load_3d(sInv, &maps.inv[0], x0 - XOFF, y0 - 1, b, sBar + 1);
        }
        load_coef(0);
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
        double A = is[p.n + tid] / Nx + is[2 * p.n + tid] / Ny;
        sSm[tid * 4 + 0] = (float)(Ws / (Nx * c));
        sSm[tid * 4 + 1] = (float)(Ws / (Ny * c));
        sSm[tid * 4 + 2] = active ? (float)(-Ws * A / (c * c * (double)HW)) : 0.f;
    }
    if (!USE_TMA) {
        const float* img = p.tgt + (size_t)b * 3 * HW;
        for (int idx = tid; idx < 3 * BS_CH; idx += NT) {
            int ch = idx / BS_CH, r = (idx - ch * BS_CH) / PITCH, j = idx - ch * BS_CH - r * PITCH;
            int vv = y0 - 1 + r, uu = x0 - XOFF + j;
            sY[idx] = (vv >= 0 && vv < H && uu >= 0 && uu < W) ? __ldg(img + (size_t)ch * HW + (size_t)vv * W + uu) : 0.f;
        }
        load_inv_manual(0, sInv);
    }
    __syncthreads();                 // barrier init, camera table, smoothness constants (and manual tiles) visible
    if (USE_TMA) tma::mbar_wait(sBar + 0, 0);

    const float* K = sCam;
    const float* Kinv = sCam + 9;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const float rw = exact::rcp_refined(wm1), rh = exact::rcp_refined(hm1);
    const float Wp = (float)((double)g_photo * (double)p.photo_w / ((double)p.n * Ntot));
    const float cf_ssim = Wp * p.alpha * (1.0f / 3.0f) * (-0.5f) * (1.0f / 9.0f);   // u/9 of App. B-3
    const float cf_l1 = Wp * p.oma * (1.0f / 3.0f);

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

    // Edge-aware weights exp(-mean_c|dI|) (depth.py:23-24) of the four pixel pairs each output takes part in,
    // multiplied by the mask of the pair's owner (left / upper pixel, loss.py:285-286), zero where the pair
    // does not exist.  Scale independent: computed once, parked in smem (slot j of thread tid).
    //   [0..3] pair (q, q+1)   [4] pair (q0-1, q0)   [5..8] pair (q, q+W)   [9..12] pair (q-W, q)
    {
        auto wgt = [&](const float* a, const float* c) {
            float d = (fabsf(a[0] - c[0]) + fabsf(a[BS_CH] - c[BS_CH])) + fabsf(a[2 * BS_CH] - c[2 * BS_CH]);
            return expf(-exact::div3(d));
        };
        auto mask_at = [&](int vv, int uu) {
            return vv >= 0 && vv < H && uu >= 0 && uu < W && (p.mask == nullptr || p.mask[(size_t)b * HW + (size_t)vv * W + uu] != 0);
        };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float* yc = sY + (ty + 1) * PITCH + XOFF + 4 * tx + k;
            int u = u0 + k;
            sW[(0 + k) * NT + tid] = (msk[k] && u + 1 < W) ? wgt(yc, yc + 1) : 0.f;
            sW[(5 + k) * NT + tid] = (msk[k] && v + 1 < H) ? wgt(yc, yc + PITCH) : 0.f;
            sW[(9 + k) * NT + tid] = (valid[k] && mask_at(v - 1, u)) ? wgt(yc - PITCH, yc) : 0.f;
            if (k == 0) sW[4 * NT + tid] = (valid[0] && mask_at(v, u - 1)) ? wgt(yc - 1, yc) : 0.f;
        }
    }

    // Pose-gradient partials dL/d(R|t) of this thread, both sources, all scales.  Deliberately kept in (L1-resident)
    // local memory: they are touched once per (scale, source) pass -- each pass accumulates its 12 sums in registers
    // and folds them in at the end -- and 24 more live registers would push stage D over the 128-register budget.
    volatile float pacc[S][12];
#pragma unroll
    for (int s = 0; s < S; s++)
#pragma unroll
        for (int j = 0; j < 12; j++) pacc[s][j] = 0.f;

#pragma unroll 1
    for (int i = 0; i < p.n; i++) {
        const unsigned char* sel = p.sel + ((size_t)i * p.B + b) * HW;
        const float* sI = sInv + (i & 1) * BS_INV_FLOATS;
        unsigned selq = 0;      // argmin codes of my 4 outputs
        if (v < H) {
            const unsigned char* sq = sel + (size_t)v * W + u0;
            if (u0 + 3 < W && ((W & 3) == 0)) selq = *reinterpret_cast<const unsigned*>(sq);
            else {
#pragma unroll
                for (int k = 0; k < 4; k++) if (u0 + k < W) selq |= (unsigned)sq[k] << (8 * k);
            }
        }
        if (USE_TMA) tma::mbar_wait(sBar + 1 + (i & 1), (i >> 1) & 1);

        // ---- smoothness gradient (App. B-6): d/dinv of sum m*w*|inv_p - inv_q| / (N*c) plus the mean term ----
        float ginv[4];
        {
            const float kx = sSm[i * 4 + 0], ky = sSm[i * 4 + 1], mt = sSm[i * 4 + 2];
            const float* iq = sI + (ty + 1) * PITCH + XOFF + 4 * tx;
            float4 c4 = *reinterpret_cast<const float4*>(iq);
            float4 u4 = *reinterpret_cast<const float4*>(iq - PITCH);
            float4 d4 = *reinterpret_cast<const float4*>(iq + PITCH);
            const float ic[6] = {iq[-1], c4.x, c4.y, c4.z, c4.w, iq[4]};
            const float iu[4] = {u4.x, u4.y, u4.z, u4.w}, id[4] = {d4.x, d4.y, d4.z, d4.w};
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

        tma::mbar_wait(sBar + 3, i & 1);     // coefficient texels of this scale have landed

        auto pass = [&](auto src_c) {
            constexpr int s = decltype(src_c)::value;
            // ---- stage C: box adjoint of the taps that selected source s, all three channels ----
            float box[3][2][4];      // [channel][A + y*C | B][output]
            {
                const float4* row0 = sCo + ty * PITCH;      // texel row of image row v-1
                const float* yrow = sY + (ty + 1) * PITCH + XOFF + 4 * tx;   // my 4 target values, channel 0
                if (border) {
                    // reflect-pad multiplicities: tap (q+d) counts twice when its padded twin folds onto q
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
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        if (s == 0) box_adjoint_src<0, true>(row0 + ch * BS_TEX, tx, rwgt, cwgt, yrow + ch * BS_CH, box[ch]);
                        else box_adjoint_src<1, true>(row0 + ch * BS_TEX, tx, rwgt, cwgt, yrow + ch * BS_CH, box[ch]);
                        asm volatile("" ::: "memory");     // keep the three channels' tap loads from being hoisted together
                    }
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        if (s == 0) box_adjoint_src<0, false>(row0 + ch * BS_TEX, tx, nullptr, nullptr, yrow + ch * BS_CH, box[ch]);
                        else box_adjoint_src<1, false>(row0 + ch * BS_TEX, tx, nullptr, nullptr, yrow + ch * BS_CH, box[ch]);
                        asm volatile("" ::: "memory");
                    }
                }
            }
            if (s == S - 1) {
                // every thread is done with this scale's texels (and with the previous scale's inverse-depth slot):
                // refill both for the next scale while stage D of the last source runs
                __syncthreads();
                if (i + 1 < p.n) {
                    if (tid == 0) {
                        tma::fence_proxy_async();
                        if (USE_TMA) {
                            tma::mbar_expect_tx(sBar + 1 + ((i + 1) & 1), BS_CH * 4);
                            tma::load_3d(sInv + ((i + 1) & 1) * BS_INV_FLOATS, &maps.inv[i + 1], x0 - XOFF, y0 - 1, b, sBar + 1 + ((i + 1) & 1));
                        }
                        load_coef(i + 1);
                    }
                    if (!USE_TMA) load_inv_manual(i + 1, sInv + ((i + 1) & 1) * BS_INV_FLOATS);
                }
            }

            // ---- stage D: per-output chain for source s ----
            const float* Rt = sCam + 18 + 12 * s;
            const float4* sp = s == 0 ? src0 : src1;
            float P[12];
#pragma unroll
            for (int j = 0; j < 12; j++) P[j] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (!valid[k]) continue;
                unsigned code = (selq >> (8 * k)) & 0xffu;
                bool selme = msk[k] && (p.automask ? (code == 2u * s) : (code == (unsigned)s));
                bool any = selme;
#pragma unroll
                for (int ch = 0; ch < 3; ch++) any = any || box[ch][0][k] != 0.f || box[ch][1][k] != 0.f;
                if (!any) continue;
                int u = u0 + k;
                float r[3], Xc[3];
                exact::ray(Kinv, u, v, r);
                const float invq = sI[(ty + 1) * PITCH + XOFF + 4 * tx + k];
                float d = exact::rcp_refined(fmaxf(invq, 1e-6f));
#pragma unroll
                for (int j = 0; j < 3; j++) Xc[j] = __fmul_rn(r[j], d);
                exact::Proj pr;
                exact::project(K, Rt, Xc, wm1, hm1, rw, rh, pr);
                // same footprint arithmetic as the forward (mgvs_device.cuh footprint/blend4): x_q comes out bit-identical
                float xw = floorf(pr.ix), yn = floorf(pr.iy);
                float wE = __fadd_rn(pr.ix, -xw), wW = __fadd_rn(1.0f, -wE), wS = __fadd_rn(pr.iy, -yn), wN = __fadd_rn(1.0f, -wS);
                int cx0 = (int)fminf(fmaxf(xw, -2.0f), (float)W), cy0 = (int)fminf(fmaxf(yn, -2.0f), (float)H);
                const float4* pc = sp + (cy0 + PACK_BORDER) * Wpk + (cx0 + PACK_BORDER);
                float4 nw = __ldg(pc), ne = __ldg(pc + 1), sw = __ldg(pc + Wpk), se = __ldg(pc + Wpk + 1);
                Foot f;
                f.off = 0;
                f.wnw = __fmul_rn(wN, wW); f.wne = __fmul_rn(wN, wE); f.wsw = __fmul_rn(wS, wW); f.wse = __fmul_rn(wS, wE);
                const float xq[3] = {blend4(nw.x, ne.x, sw.x, se.x, f), blend4(nw.y, ne.y, sw.y, se.y, f), blend4(nw.z, ne.z, sw.z, se.z, f)};
                const float* yp = sY + (ty + 1) * PITCH + XOFF + 4 * tx + k;
                float g[3];
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    float yq = yp[ch * BS_CH];
                    g[ch] = cf_ssim * fmaf(xq[ch], box[ch][1][k], box[ch][0][k]);
                    if (selme) {
                        float df = xq[ch] - yq;
                        g[ch] += cf_l1 * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
                    }
                }
                // bilinear adjoint (GridSampler backward w.r.t. the grid): out-of-image corners are zeros of the border
                float gix = g[0] * ((ne.x - nw.x) * wN + (se.x - sw.x) * wS) + g[1] * ((ne.y - nw.y) * wN + (se.y - sw.y) * wS) +
                            g[2] * ((ne.z - nw.z) * wN + (se.z - sw.z) * wS);
                float giy = g[0] * ((sw.x - nw.x) * wW + (se.x - ne.x) * wE) + g[1] * ((sw.y - nw.y) * wW + (se.y - ne.y) * wE) +
                            g[2] * ((sw.z - nw.z) * wW + (se.z - ne.z) * wE);
                // projection adjoint (App. B-5)
                float iz = exact::rcp_refined(pr.Z);
                float gP0 = gix * iz, gP1 = giy * iz;
                float gP2 = (pr.Pz >= 1e-5f) ? -(gix * pr.ax + giy * pr.ay) * iz : 0.f;
                float gX0 = K[0] * gP0 + K[3] * gP1 + K[6] * gP2;
                float gX1 = K[1] * gP0 + K[4] * gP1 + K[7] * gP2;
                float gX2 = K[2] * gP0 + K[5] * gP1 + K[8] * gP2;
                P[0] += gX0 * pr.Xc0; P[1] += gX0 * pr.Xc1; P[2] += gX0 * pr.Xc2; P[3] += gX0;
                P[4] += gX1 * pr.Xc0; P[5] += gX1 * pr.Xc1; P[6] += gX1 * pr.Xc2; P[7] += gX1;
                P[8] += gX2 * pr.Xc0; P[9] += gX2 * pr.Xc1; P[10] += gX2 * pr.Xc2; P[11] += gX2;
                float gd = 0.f;
#pragma unroll
                for (int j = 0; j < 3; j++) gd += (Rt[j] * gX0 + Rt[4 + j] * gX1 + Rt[8 + j] * gX2) * r[j];
                if (invq >= 1e-6f) ginv[k] -= d * d * gd;
            }
#pragma unroll
            for (int j = 0; j < 12; j++) pacc[s][j] = pacc[s][j] + P[j];
        };
        pass(std::integral_constant<int, 0>());
        pass(std::integral_constant<int, 1>());
        if (v < H) {
            float* go = p.grad_inv[i] + (size_t)b * HW + (size_t)v * W + u0;
            if (u0 + 3 < W && ((W & 3) == 0)) *reinterpret_cast<float4*>(go) = make_float4(ginv[0], ginv[1], ginv[2], ginv[3]);
            else {
#pragma unroll
                for (int k = 0; k < 4; k++) if (u0 + k < W) go[k] = ginv[k];
            }
        }
        if (!USE_TMA) __syncthreads();     // manually loaded inverse-depth tile of the next scale visible
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
