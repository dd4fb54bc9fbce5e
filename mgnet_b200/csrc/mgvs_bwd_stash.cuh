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
// The forward also leaves the two masked edge-aware smoothness weight planes (8 B/px) in the stash, so the 13
// expf per thread of the recompute kernel's prologue are gone as well.
// Thread layout: 64 columns x 4 strips of 4 rows -- lanes are consecutive columns, so the bilinear gathers touch
// 4x fewer cache lines per instruction than with the forward's 4-wide horizontal strips, scalar smem reads are
// conflict-free, and so are the 128-bit texel reads (a phase row is 18 texels = 288 B = 32 B mod 128 B, so the four
// phases of a quarter-warp land on disjoint bank groups).
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
    int pad;                      // padding_mode of the PAD kernels (1 border, 2 reflection)
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
    TmaDesc tgt, inv[MAXN], coef, wgt;     // wgt: the two masked edge-aware weight planes [2B][H][4*Wg] of the stash
};

constexpr int BS_ROWS = TH + 2;                       // tile+1 rows
constexpr int BS_CH = BS_ROWS * PITCH;                // floats per plane (target / inverse depth), col j <-> image col x0-XOFF+j
constexpr int BS_GROUPS = PITCH / 4;                  // 18 column groups per texel row (image cols x0-4 .. x0+67)
constexpr int BS_TEX = BS_ROWS * PITCH;               // texels per channel map: [row][phase][group]
constexpr int BS_TILE3_FLOATS = (3 * BS_CH + 31) / 32 * 32;
constexpr int BS_INV_FLOATS = (BS_CH + 31) / 32 * 32;
// Two CTAs per SM fit the 196 KB shared-memory carveout (2 x (99,328 + 1 KB reserved)), which leaves 60 instead of
// 28 KB of L1 for the bilinear gathers; the pose-partial scratch therefore aliases the target tile, which is dead by
// then.  (Measured: no effect on kernel time at C2/C3/C4 -- the L1 capacity is not what limits the gathers.)
constexpr int BS_SMEM_FLOATS = 3 * BS_TEX * 4 + BS_TILE3_FLOATS + 4 * BS_INV_FLOATS + 48 + 4 * MAXN + 8;
static_assert(BS_TILE3_FLOATS >= 8 * 24, "pose-partial scratch aliases the target tile");
constexpr int BS_SMEM_BYTES = BS_SMEM_FLOATS * 4;
#if MGVS_TW == 64 && MGVS_TH == 16     // (tile-shape experiment builds never run the stash backward: mgvs_api.cu refuses)
static_assert(2 * (BS_SMEM_BYTES + 1024) <= 196 * 1024, "stash backward must fit the 196 KB carveout twice");
#endif
static_assert((BS_TEX * 16) % 128 == 0, "channel maps must stay 128-byte aligned for TMA");

namespace tma {
// 5-D tiled load: box origin (c0..c4) in elements
// The coefficient texels are read exactly once: L2 evict-first keeps them from displacing the packed sources the
// bilinear gathers re-read.
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void load_5d(void* dst, const void* tmap, int c0, int c1, int c2, int c3, int c4, uint64_t* bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
}  // namespace tma

// The camera table re-laid out in smem so that every row is one broadcast 128-bit load:
//   [0..11] K rows padded to 4, [12..23] Kinv rows padded to 4, [24..35] R|t of source 0, [36..47] R|t of source 1
__device__ __forceinline__ void load_cam_padded(float* __restrict__ sCamP, const Cam* __restrict__ cam, int tid)
{
    if (tid < 48) {
        const float* c = reinterpret_cast<const float*>(cam);
        float v;
        if (tid < 24) { int m = tid / 12, r = (tid % 12) / 4, k = tid % 4; v = k < 3 ? c[m * 9 + r * 3 + k] : 0.f; }
        else v = c[18 + (tid - 24)];
        sCamP[tid] = v;
    }
}

// 3x3 box adjoint of the three coefficient maps of one channel for the 4 vertically adjacent outputs of a thread
// (column u, rows v0..v0+3), restricted to the taps whose argmin is source SRC.  col0 -> texel of (tile row of image
// row v0-1, phase/group of column u-1); toff[dx] = texel offset of column u-1+dx.  One real loop over the three tap
// columns (fully unrolled, ptxas hoists all 18 128-bit loads above the arithmetic and spills them).
template <int SRC, bool WEIGHTED>
__device__ __forceinline__ void box_adjoint_col(const float4* __restrict__ rows, const int toff[3], const float cw[3],
                                                int v0, int H, const float ys[4], float out[2][4])
{   // out[0][k] = A_k + y_q * C_k, out[1][k] = B_k   (G = out[0] + x_q * out[1]; y_q is known without the gather)
    float h[3][6];
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
        for (int r = 0; r < 6; r++) h[m][r] = 0.f;
#pragma unroll 1
    for (int dx = 0; dx < 3; dx++) {
        const float4* cc = rows + (dx == 0 ? toff[0] : (dx == 1 ? toff[1] : toff[2]));
        const float cwt = WEIGHTED ? (dx == 0 ? cw[0] : (dx == 1 ? cw[1] : cw[2])) : 1.0f;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            float4 t = cc[r * PITCH];
            float m = SRC == 0 ? t.w : 1.0f - t.w;
            if (WEIGHTED) m *= cwt;
            h[0][r] = fmaf(m, t.x, h[0][r]);
            h[1][r] = fmaf(m, t.y, h[1][r]);
            h[2][r] = fmaf(m, t.z, h[2][r]);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float o[3];
        if (WEIGHTED) {
            // reflect-pad multiplicities: tap (q+d) counts twice when its padded twin folds onto q
            const int v = v0 + k;
            const float r0 = (v - 1 >= 0) ? ((v == 1) ? 2.f : 1.f) : 0.f;
            const float r2 = (v + 1 <= H - 1) ? ((v == H - 2) ? 2.f : 1.f) : 0.f;
#pragma unroll
            for (int m = 0; m < 3; m++) o[m] = fmaf(r0, h[m][k], fmaf(r2, h[m][k + 2], h[m][k + 1]));
        } else {
#pragma unroll
            for (int m = 0; m < 3; m++) o[m] = (h[m][k] + h[m][k + 1]) + h[m][k + 2];
        }
        out[0][k] = fmaf(ys[k], o[2], o[0]);
        out[1][k] = o[1];
    }
}

// cf * sign(d) with sign(0) = 0
__device__ __forceinline__ float signed_const(float cf, float d)
{
    float s = __int_as_float((__float_as_int(d) & 0x80000000) ^ __float_as_int(cf));
    return d == 0.f ? 0.f : s;
}

constexpr int BS_SX = 64;                  // thread layout of the stash backward: 64 columns x 4 strips of 4 rows
// (assumes the 64x16 tile / 256 threads of the product build; mgvs_api.cu refuses the stash path otherwise)

template <bool USE_TMA, bool PAD = false>
__global__ void __launch_bounds__(NT, MIN_CTAS) bwd_stash_kernel(const __grid_constant__ BwdSParams p, const __grid_constant__ BwdSMaps maps)
{
    extern __shared__ __align__(128) float smem[];
    float4* sCo = reinterpret_cast<float4*>(smem);        // [3 ch][BS_ROWS][4 phases][BS_GROUPS] texels of the current scale
    float* sY = smem + 3 * BS_TEX * 4;                     // [3][BS_ROWS][PITCH]
    float* sInv = sY + BS_TILE3_FLOATS;                    // [2][BS_ROWS][PITCH] inverse-depth ring
    float* sWx = sInv + 2 * BS_INV_FLOATS;                 // [BS_ROWS][PITCH] masked edge-aware weight of pair (q, q+1), from the forward
    float* sWy = sWx + BS_INV_FLOATS;                      // [BS_ROWS][PITCH] same for pair (q, q+W)
    float* sRed = sY;                                      // [8 warps][24], after the scale loop only (aliases the dead target tile)
    float* sCam = sWy + BS_INV_FLOATS;                     // 48, padded rows (load_cam_padded)
    float* sSm = sCam + 48;                                // [n][4]
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sSm + 4 * MAXN);   // [0] target + weights, [1],[2] inverse-depth ring, [3] coefficients

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tpi = p.tiles_x * p.tiles_y;
    const int b = tile / tpi;
    const int trem = tile - b * tpi;
    const int tyi = trem / p.tiles_x, txi = trem - tyi * p.tiles_x;
    const int x0 = txi * TW, y0 = tyi * TH;
    const int H = p.H, W = p.W, HW = H * W;
    // vertical strips: lanes are consecutive columns (coalesced gathers and stores, conflict-free smem), each thread
    // owns rows v0..v0+3 of column u
    // Within a warp (32 columns) lane l takes column 4*(l & 7) + (l >> 3): every quarter-warp then reads one phase of
    // 8 consecutive column groups = 8 consecutive texels for ANY tap column (conflict-free 128-bit loads; with lanes on
    // consecutive columns the dx = +-1 taps of a quarter-warp straddle two groups of the same phase -> 2-way conflicts),
    // while the warp as a whole still covers 32 consecutive columns (coalesced gathers / stores, conflict-free scalars).
    const int tx = (tid & 32) + 4 * (lane & 7) + (lane >> 3), sy = tid / BS_SX;
    const int u = x0 + tx, v0 = y0 + 4 * sy;
    const int rl = 4 * sy;                       // tile row of output 0 (smem row rl+1: planes carry one halo row)

    // reflect-pad multiplicities are only needed where a 3x3 window of the tile can touch the image border
    const bool border = (x0 == 0) || (y0 == 0) || (x0 + TW + 1 >= W) || (y0 + TH + 1 >= H);
    constexpr uint32_t COEF_BYTES = 3 * BS_TEX * 16;
    auto load_coef = [&](int i) {      // thread 0 only: the three channel maps of scale i
        tma::mbar_expect_tx(sBar + 3, COEF_BYTES);
        const uint64_t pol = tma::policy_evict_first();
#pragma unroll
        for (int ch = 0; ch < 3; ch++)
            tma::load_5d(sCo + ch * BS_TEX, &maps.coef, 0, (x0 >> 2) - 1, 0, y0 - 1, (i * p.B + b) * 3 + ch, sBar + 3, pol);
    };
    auto load_plane_manual = [&](const float* __restrict__ img, float* dst) {     // [H][W] plane -> tile+1, zero outside
        for (int idx = tid; idx < BS_CH; idx += NT) {
            int r = idx / PITCH, j = idx - r * PITCH;
            int vv = y0 - 1 + r, uu = x0 - XOFF + j;
            dst[idx] = (vv >= 0 && vv < H && uu >= 0 && uu < W) ? __ldg(img + (size_t)vv * W + uu) : 0.f;
        }
    };
    // inverse-depth tiles: TMA from the full-resolution maps, or filled by hand (manual-loader build)
    constexpr bool inv_tma = USE_TMA;
    auto fill_inv = [&](int i, float* dst) { load_plane_manual(p.inv[i] + (size_t)b * HW, dst); };
    if (tid == 0) {
        tma::mbar_init(sBar + 0, 1); tma::mbar_init(sBar + 1, 1); tma::mbar_init(sBar + 2, 1); tma::mbar_init(sBar + 3, 1);
        tma::fence_barrier_init();
        // the two weight planes live in the stash (own pitch, always TMA-able)
        tma::mbar_expect_tx(sBar + 0, (USE_TMA ? 3 : 0) * BS_CH * 4 + 2 * BS_CH * 4);
        tma::load_3d(sWx, &maps.wgt, x0 - XOFF, y0 - 1, 2 * b, sBar + 0);
        tma::load_3d(sWy, &maps.wgt, x0 - XOFF, y0 - 1, 2 * b + 1, sBar + 0);
        if (USE_TMA) tma::load_3d(sY, &maps.tgt, x0 - XOFF, y0 - 1, 3 * b, sBar + 0);
        if (inv_tma) {
            tma::mbar_expect_tx(sBar + 1, BS_CH * 4);
            tma::load_3d(sInv, &maps.inv[0], x0 - XOFF, y0 - 1, b, sBar + 1);
        }
        load_coef(0);
    }
    load_cam_padded(sCam, p.cams + b, tid);
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
#pragma unroll 1
        for (int ch = 0; ch < 3; ch++) load_plane_manual(p.tgt + ((size_t)b * 3 + ch) * HW, sY + ch * BS_CH);
    }
    if (!inv_tma) fill_inv(0, sInv);
    __syncthreads();                 // barrier init, camera table, smoothness constants (and manual tiles) visible
    tma::mbar_wait(sBar + 0, 0);

    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const float rw = exact::rcp_refined(wm1), rh = exact::rcp_refined(hm1);
    const float Wp = (float)((double)g_photo * (double)p.photo_w / ((double)p.n * Ntot));
    const float cf_ssim = Wp * p.alpha * (1.0f / 3.0f) * (-0.5f) * (1.0f / 9.0f);   // u/9 of App. B-3
    const float cf_l1 = Wp * p.oma * (1.0f / 3.0f);

    const size_t pimg = (size_t)(H + 2 * PACK_BORDER) * (W + 2 * PACK_BORDER);
    const float4* src0 = p.psrc[0] + (size_t)b * pimg;
    const float4* src1 = p.psrc[1] + (size_t)b * pimg;
    const int Wpk = W + 2 * PACK_BORDER;

    // bit k: output k inside the image / inside the image and mask true
    unsigned valid = 0, msk = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        bool ok = (v0 + k < H) && (u < W);
        valid |= (unsigned)ok << k;
        if (ok && (p.mask == nullptr || p.mask[(size_t)b * HW + (size_t)(v0 + k) * W + u] != 0)) msk |= 1u << k;
    }

    // texel offsets of columns u-1, u, u+1 in a [phase][group] texel row, and the scalar-plane column of u
    int toff[3];
#pragma unroll
    for (int dx = 0; dx < 3; dx++) { int j = tx + dx - 1 + XOFF; toff[dx] = (j & 3) * BS_GROUPS + (j >> 2); }
    const int pc0 = (rl + 1) * PITCH + XOFF + tx;          // plane index of output 0
    float cw[3];
    cw[0] = (u - 1 >= 0) ? ((u == 1) ? 2.f : 1.f) : 0.f;
    cw[1] = 1.f;
    cw[2] = (u + 1 <= W - 1) ? ((u == W - 2) ? 2.f : 1.f) : 0.f;

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
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((valid >> k) & 1u) selq |= (unsigned)sel[(size_t)(v0 + k) * W + u] << (8 * k);
        if (inv_tma) tma::mbar_wait(sBar + 1 + (i & 1), (i >> 1) & 1);

        // ---- smoothness gradient (App. B-6): d/dinv of sum m*w*|inv_p - inv_q| / (N*c) plus the mean term ----
        float ginv[4];
        {
            const float kx = sSm[i * 4 + 0], ky = sSm[i * 4 + 1], mt = sSm[i * 4 + 2];
            auto sgn = [](float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); };
            float up = sI[pc0 - PITCH];
            float wu = sWy[pc0 - PITCH];         // pair (q-W, q), masked by its owner (the upper pixel)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int pc = pc0 + k * PITCH;
                float c = sI[pc], lf = sI[pc - 1], rt = sI[pc + 1], dn = sI[pc + PITCH];
                float wr = sWx[pc], wl = sWx[pc - 1], wd = sWy[pc];
                float g = kx * (wr * sgn(c - rt) - wl * sgn(lf - c)) + ky * (wd * sgn(c - dn) - wu * sgn(up - c));
                ginv[k] = ((valid >> k) & 1u) ? mt + g : 0.f;
                up = c; wu = wd;
            }
        }

        tma::mbar_wait(sBar + 3, i & 1);     // coefficient texels of this scale have landed

        auto pass = [&](auto src_c) {
            constexpr int s = decltype(src_c)::value;
            // ---- stage C: box adjoint of the taps that selected source s, all three channels ----
            float box[3][2][4];      // [channel][A + y*C | B][output]
            {
                const float4* rows = sCo + rl * PITCH;      // texel row of image row v0-1
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    const float ys[4] = {sY[ch * BS_CH + pc0], sY[ch * BS_CH + pc0 + PITCH], sY[ch * BS_CH + pc0 + 2 * PITCH],
                                         sY[ch * BS_CH + pc0 + 3 * PITCH]};
                    if (border) box_adjoint_col<s, true>(rows + ch * BS_TEX, toff, cw, v0, H, ys, box[ch]);
                    else box_adjoint_col<s, false>(rows + ch * BS_TEX, toff, cw, v0, H, ys, box[ch]);
                }
            }
            if (s == S - 1) {
                // every thread is done with this scale's texels (and with the previous scale's inverse-depth slot):
                // refill both for the next scale while stage D of the last source runs
                __syncthreads();
                if (i + 1 < p.n) {
                    if (tid == 0) {
                        tma::fence_proxy_async();
                        if (inv_tma) {
                            tma::mbar_expect_tx(sBar + 1 + ((i + 1) & 1), BS_CH * 4);
                            tma::load_3d(sInv + ((i + 1) & 1) * BS_INV_FLOATS, &maps.inv[i + 1], x0 - XOFF, y0 - 1, b, sBar + 1 + ((i + 1) & 1));
                        }
                        load_coef(i + 1);
                    }
                    if (!inv_tma) fill_inv(i + 1, sInv + ((i + 1) & 1) * BS_INV_FLOATS);
                }
            }

            // ---- stage D: per-output chain for source s ----
            const float4* sp = s == 0 ? src0 : src1;
            float P[12];
#pragma unroll
            for (int j = 0; j < 12; j++) P[j] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (!((valid >> k) & 1u)) continue;
                unsigned code = (selq >> (8 * k)) & 0xffu;
                bool selme = ((msk >> k) & 1u) && (p.automask ? (code == 2u * s) : (code == (unsigned)s));
                bool any = selme;
#pragma unroll
                for (int ch = 0; ch < 3; ch++) any = any || box[ch][0][k] != 0.f || box[ch][1][k] != 0.f;
                if (!any) continue;
                // camera rows as broadcast 128-bit loads (load_cam_padded)
                const float4 K0 = *reinterpret_cast<const float4*>(sCam + 0), K1 = *reinterpret_cast<const float4*>(sCam + 4),
                             K2 = *reinterpret_cast<const float4*>(sCam + 8);
                const float4 I0 = *reinterpret_cast<const float4*>(sCam + 12), I1 = *reinterpret_cast<const float4*>(sCam + 16),
                             I2 = *reinterpret_cast<const float4*>(sCam + 20);
                const float4 R0 = *reinterpret_cast<const float4*>(sCam + 24 + 12 * s), R1 = *reinterpret_cast<const float4*>(sCam + 28 + 12 * s),
                             R2 = *reinterpret_cast<const float4*>(sCam + 32 + 12 * s);
                const float Kf[9] = {K0.x, K0.y, K0.z, K1.x, K1.y, K1.z, K2.x, K2.y, K2.z};
                const float Kif[9] = {I0.x, I0.y, I0.z, I1.x, I1.y, I1.z, I2.x, I2.y, I2.z};
                const float Rtf[12] = {R0.x, R0.y, R0.z, R0.w, R1.x, R1.y, R1.z, R1.w, R2.x, R2.y, R2.z, R2.w};
                float r[3], Xc[3];
                exact::ray(Kif, u, v0 + k, r);
                const float invq = sI[pc0 + k * PITCH];
                float d = exact::rcp_refined(fmaxf(invq, 1e-6f));
#pragma unroll
                for (int j = 0; j < 3; j++) Xc[j] = __fmul_rn(r[j], d);
                exact::Proj pr;
                exact::project<PAD>(Kf, Rtf, Xc, wm1, hm1, rw, rh, pr, p.pad);
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
                float g[3];
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    g[ch] = cf_ssim * fmaf(xq[ch], box[ch][1][k], box[ch][0][k]);
                    if (selme) g[ch] += signed_const(cf_l1, xq[ch] - sY[ch * BS_CH + pc0 + k * PITCH]);
                }
                // bilinear adjoint (GridSampler backward w.r.t. the grid): out-of-image corners are zeros of the border
                float gix = g[0] * ((ne.x - nw.x) * wN + (se.x - sw.x) * wS) + g[1] * ((ne.y - nw.y) * wN + (se.y - sw.y) * wS) +
                            g[2] * ((ne.z - nw.z) * wN + (se.z - sw.z) * wS);
                float giy = g[0] * ((sw.x - nw.x) * wW + (se.x - ne.x) * wE) + g[1] * ((sw.y - nw.y) * wW + (se.y - ne.y) * wE) +
                            g[2] * ((sw.z - nw.z) * wW + (se.z - ne.z) * wE);
                if constexpr (PAD) { gix *= pr.mx; giy *= pr.my; }   // padding-mode derivative (clip / reflect)
                // projection adjoint (App. B-5)
                float iz = exact::rcp_refined(pr.Z);
                float gP0 = gix * iz, gP1 = giy * iz;
                float gP2 = (pr.Pz >= 1e-5f) ? -(gix * pr.ax + giy * pr.ay) * iz : 0.f;
                float gX0 = Kf[0] * gP0 + Kf[3] * gP1 + Kf[6] * gP2;
                float gX1 = Kf[1] * gP0 + Kf[4] * gP1 + Kf[7] * gP2;
                float gX2 = Kf[2] * gP0 + Kf[5] * gP1 + Kf[8] * gP2;
                P[0] += gX0 * pr.Xc0; P[1] += gX0 * pr.Xc1; P[2] += gX0 * pr.Xc2; P[3] += gX0;
                P[4] += gX1 * pr.Xc0; P[5] += gX1 * pr.Xc1; P[6] += gX1 * pr.Xc2; P[7] += gX1;
                P[8] += gX2 * pr.Xc0; P[9] += gX2 * pr.Xc1; P[10] += gX2 * pr.Xc2; P[11] += gX2;
                float gd = 0.f;
#pragma unroll
                for (int j = 0; j < 3; j++) gd += (Rtf[j] * gX0 + Rtf[4 + j] * gX1 + Rtf[8 + j] * gX2) * r[j];
                if (invq >= 1e-6f) ginv[k] -= d * d * gd;
            }
#pragma unroll
            for (int j = 0; j < 12; j++) pacc[s][j] = pacc[s][j] + P[j];
        };
        pass(std::integral_constant<int, 0>());
        pass(std::integral_constant<int, 1>());
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((valid >> k) & 1u) p.grad_inv[i][(size_t)b * HW + (size_t)(v0 + k) * W + u] = ginv[k];
        if (!inv_tma) __syncthreads();     // hand-filled inverse-depth tile of the next scale visible
    }

    // deterministic pose partials: shuffle tree -> smem -> fixed-order sum -> one store per tile
    __syncthreads();     // sRed aliases sY: every warp is done with the target tile
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
