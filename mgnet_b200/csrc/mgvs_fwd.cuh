// mgvs_fwd.cuh -- fused forward kernel of the view-synthesis loss (one launch per call).
//
// One CTA owns a 64x16 tile of target pixels of one image and loops over all n scales:
//   stage 0  target tile + 1-pixel SSIM halo (reflect indexed) and both source tiles -> smem;
//            per-pixel target statistics (mu_y, mu_y^2, sigma_y) -> smem; identity-reprojection
//            losses of both sources -> registers (computed once, reused by every scale; loss.py:139-144)
//   stage 1  (per scale) every halo pixel: K^-1 back-projection, SE(3), projection, bilinear gather of
//            both sources -> smem (the warped images never reach HBM)
//   stage 2  (per scale) 4 horizontally adjacent outputs per thread: 3x3 SSIM + L1 against the target,
//            min / argmin over [warp_prev, id_prev, warp_next, id_next], masked sum, smoothness terms
//   stage 3  deterministic CTA reduction -> per-tile partial sums (fp64) for the finalise pass
#pragma once
#include "mgvs_device.cuh"

namespace mgvs {

struct FwdParams {
    int B, H, W, n, automask;
    int pad;                     // grid_sample padding_mode for the PAD kernels: 1 border, 2 reflection (0 = zeros: PAD = false kernels)
    int early_wait;              // the image pointers are workspace copies written by the preceding kernel (uint8 ingestion)
    const float* tgt;
    const float* src[S];
    const float* inv[MAXN];
    const unsigned char* mask;   // may be null
    const Cam* cams;
    const float4* psrc[S];       // packed RGBA + zero border copies of the sources (pack_sources_kernel)
    unsigned char* sel;          // may be null
    float4* stash;               // STASH kernels only: SSIM-adjoint coefficient texels for the stash backward (see below)
    int Wg;                      // ceil(W/4): column groups per stash row
    float* wgt;                  // STASH kernels only: masked edge-aware weight planes [2B][H][4*Wg] (pairs (q,q+1) and (q,q+W))
    double* partials;            // [tiles][4n+3]
    float alpha, oma;
    int tiles_x, tiles_y;
};

struct FwdMaps {
    TmaDesc tgt, src[S], inv[MAXN];
};

constexpr int FWD_ROWS = TH + 2;                 // halo rows
constexpr int FWD_CH = FWD_ROWS * PITCH;         // floats per channel plane in smem
constexpr int FWD_HALO_W = TW + 2;
constexpr int FWD_TILE3_FLOATS = (3 * FWD_CH + 31) / 32 * 32;      // one 3-channel tile, padded to 128 bytes
constexpr int FWD_INV_FLOATS = (FWD_CH + 31) / 32 * 32;            // one inverse-depth tile
constexpr int FWD_SMEM_FLOATS = FWD_TILE3_FLOATS /*Y*/ + S * FWD_TILE3_FLOATS /*X*/ + 2 * FWD_INV_FLOATS /*inv ring*/ +
                                6 * TH * TW /*Y stats: mu_y, sigma_y per channel*/ + 8 * 8 /*red*/ + 48 /*cam*/ + 8 * NT /*smoothness weights*/ + 8 /*3 mbarriers*/;
constexpr int FWD_SMEM_BYTES = FWD_SMEM_FLOATS * 4;

// Loads a [3,H,W] image tile with 1-pixel halo (reflect-indexed) into smem planes.
__device__ __forceinline__ void fwd_load_tile(const float* __restrict__ img, float* __restrict__ dst, int x0, int y0,
                                              int H, int W, int tid)
{
    const int HW = H * W;
    for (int idx = tid; idx < 3 * FWD_ROWS * FWD_HALO_W; idx += NT) {
        int ch = idx / (FWD_ROWS * FWD_HALO_W);
        int r = idx - ch * (FWD_ROWS * FWD_HALO_W);
        int hr = r / FWD_HALO_W, hc = r - hr * FWD_HALO_W;
        int v = reflect_idx(y0 - 1 + hr, H), u = reflect_idx(x0 - 1 + hc, W);
        dst[ch * FWD_CH + hr * PITCH + XOFF - 1 + hc] = __ldg(img + ch * HW + v * W + u);
    }
}

// Photometric loss (alpha*mean_c SSIM + (1-alpha)*mean_c L1, loss.py:186-194) of 4 adjacent outputs.
// xs, ys: smem planes of the estimate and the target, pointing at [ch 0][halo row ty][my output 0]
// (16B aligned); the 3x3 windows of the 4 outputs span columns -1..4 of rows 0..2 from there.
// yst: target statistics at [0][ty][4*tx].
//
// KEEP: additionally hands the three coefficients of the closed-form SSIM adjoint of every
// (channel, output) to emit(ch, k, a, b, c):  d(ssim)/d(x_q) for q in the 3x3 window of p is (a + b*x_q + c*y_q)/9
// with a = ds/dmu_x, b = 2*ds/dE[xx], c = ds/dE[xy] -- zero where the clamp of loss.py:217 is inactive.  The
// forward has every SSIM internal in registers here, so emitting them costs ~20 instructions per channel and
// lets the stash backward skip the whole recomputation (tile+2 warps and SSIM statistics).
struct NoEmit { __device__ __forceinline__ void operator()(int, int, float, float, float) const {} };

template <bool KEEP, typename Emit>
__device__ __forceinline__ void photometric4(const float* __restrict__ xs, const float* __restrict__ ys,
                                             const float* __restrict__ yst, float alpha, float oma, float out[4], Emit emit)
{
    float ssum[4], lsum[4];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        // NOTE (measured, round 1): running the E[xx] / E[xy] sums as one packed fp32x2 chain (__fadd2_rn, bit-identical)
        // removed 384 issue slots per thread but changed neither kernel time (0.351 vs 0.353 ms at C2): this phase is
        // bound by the FP32 pipe itself, where FADD2 costs two passes.  Not kept.
        float sx[4], sxx[4], sxy[4], l1[4];
#pragma unroll
        for (int dy = 0; dy < 3; dy++) {
            const float* xr = xs + ch * FWD_CH + dy * PITCH;
            const float* yr = ys + ch * FWD_CH + dy * PITCH;
            float4 xm = *reinterpret_cast<const float4*>(xr), ym = *reinterpret_cast<const float4*>(yr);
#if MGVS_ABL & 8
            float x6[6] = {xm.y, xm.x, xm.y, xm.z, xm.w, xm.z};
            float y6[6] = {ym.y, ym.x, ym.y, ym.z, ym.w, ym.z};
#else
            float x6[6] = {xr[-1], xm.x, xm.y, xm.z, xm.w, xr[4]};
            float y6[6] = {yr[-1], ym.x, ym.y, ym.z, ym.w, yr[4]};
#endif
            float xx6[6], xy6[6];
#pragma unroll
            for (int j = 0; j < 6; j++) { xx6[j] = __fmul_rn(x6[j], x6[j]); xy6[j] = __fmul_rn(x6[j], y6[j]); }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // row-major 9-term sums, left to right (App. A "3x3 mean")
                if (dy == 0) { sx[k] = x6[k]; sxx[k] = xx6[k]; sxy[k] = xy6[k]; }
                else { sx[k] = __fadd_rn(sx[k], x6[k]); sxx[k] = __fadd_rn(sxx[k], xx6[k]); sxy[k] = __fadd_rn(sxy[k], xy6[k]); }
                sx[k] = __fadd_rn(sx[k], x6[k + 1]); sxx[k] = __fadd_rn(sxx[k], xx6[k + 1]); sxy[k] = __fadd_rn(sxy[k], xy6[k + 1]);
                sx[k] = __fadd_rn(sx[k], x6[k + 2]); sxx[k] = __fadd_rn(sxx[k], xx6[k + 2]); sxy[k] = __fadd_rn(sxy[k], xy6[k + 2]);
                if (dy == 1) l1[k] = fabsf(__fadd_rn(x6[k + 1], -y6[k + 1]));
            }
        }
        // mu_y^2 is one rounded multiply away from mu_y: recomputed (same bits) instead of stored -- the 12 KB it used to occupy
        // keep two resident CTAs inside the 196 KB shared-memory carveout, i.e. 60 KB of L1 for the gathers instead of 28 KB
        float4 muy = *reinterpret_cast<const float4*>(yst + (ch * 2 + 0) * TH * TW);
        float4 sgy = *reinterpret_cast<const float4*>(yst + (ch * 2 + 1) * TH * TW);
        const float muy4[4] = {muy.x, muy.y, muy.z, muy.w}, sgy4[4] = {sgy.x, sgy.y, sgy.z, sgy.w};
        const float mys4[4] = {__fmul_rn(muy.x, muy.x), __fmul_rn(muy.y, muy.y), __fmul_rn(muy.z, muy.z), __fmul_rn(muy.w, muy.w)};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            exact::Ssim q;
            float l = exact::ssim_from_sums(sx[k], sxx[k], sxy[k], muy4[k], mys4[k], sgy4[k], KEEP ? &q : nullptr);
            if (KEEP) {
                // With e = 2/(d1 d2) (the reciprocal the SSIM division already refined; zero where the clamp of
                // loss.py:217 is inactive):  a = e (mu_y (n2-n1) - ssim mu_x (d2-d1)),  b = -e ssim d1,  c = e n1
                const bool ok = q.loss_raw >= 0.f && q.loss_raw <= 1.f;      // clamp passes gradient inclusively
                const float e = ok ? q.idd + q.idd : 0.f;
                const float sm = q.ssim * q.mu_x;
                const float ca = e * fmaf(muy4[k], q.n2 - q.n1, -sm * (q.d2 - q.d1));
                const float cb = -e * (q.ssim * q.d1);
                const float cc = e * q.n1;
                emit(ch, k, ca, cb, cc);
            }
            if (ch == 0) { ssum[k] = l; lsum[k] = l1[k]; }
            else { ssum[k] = __fadd_rn(ssum[k], l); lsum[k] = __fadd_rn(lsum[k], l1[k]); }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
        out[k] = __fadd_rn(__fmul_rn(alpha, exact::div3(ssum[k])), __fmul_rn(oma, exact::div3(lsum[k])));
}

template <bool USE_TMA, bool STASH, bool PAD = false, bool L1ONLY = false>
__global__ void __launch_bounds__(NT, MIN_CTAS) fwd_kernel(const FwdParams p, const __grid_constant__ FwdMaps maps)
{
    extern __shared__ __align__(128) float smem[];
    float* sY = smem;
    float* sX = sY + FWD_TILE3_FLOATS;          // [S][3][rows][PITCH] (each source tile 128B aligned)
    float* sInv = sX + S * FWD_TILE3_FLOATS;    // [2][rows][PITCH] inverse-depth ring (TMA path)
    float* sYst = sInv + 2 * FWD_INV_FLOATS;    // [6][TH][TW]: (mu_y, sigma_y) per channel
    float* sRed = sYst + 6 * TH * TW;           // [8 warps][8]
    float* sCam = sRed + 64;                    // 48 floats
    float4* sWgt = reinterpret_cast<float4*>(sCam + 48);        // [2][NT]: each thread's masked edge-aware weights (x pairs, y pairs)
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sCam + 48 + 8 * NT);   // [0] image tiles, [1],[2] inverse-depth ring

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tpi = p.tiles_x * p.tiles_y;
    const int b = tile / tpi;
    const int trem = tile - b * tpi;
    const int tyi = trem / p.tiles_x, txi = trem - tyi * p.tiles_x;
    const int x0 = txi * TW, y0 = tyi * TH;
    const int H = p.H, W = p.W, HW = H * W;
    const int tx = tid % CG, ty = tid / CG;
    const int u0 = x0 + 4 * tx, v = y0 + ty;     // this thread's 4 outputs: (v, u0..u0+3)

    const bool border = (x0 == 0) || (y0 == 0) || (x0 + TW >= W) || (y0 + TH >= H);
    // NOTE (measured, round 1, profiles/r01k_stagger.txt): delaying the second CTA of every SM by 1-12 us at grid start, so that the
    // two co-resident CTAs alternate their gather-latency and FP32 stages, changed C2/C4 times by <0.5 %: they de-phase on their own.
    if (p.early_wait) pdl_wait();
    if (USE_TMA) {
        if (tid == 0) {
            tma::mbar_init(sBar + 0, 1); tma::mbar_init(sBar + 1, 1); tma::mbar_init(sBar + 2, 1);
            tma::fence_barrier_init();
            tma::mbar_expect_tx(sBar + 0, 3 * 3 * FWD_CH * 4);
            tma::load_3d(sY, &maps.tgt, x0 - XOFF, y0 - 1, 3 * b, sBar + 0);
            tma::load_3d(sX, &maps.src[0], x0 - XOFF, y0 - 1, 3 * b, sBar + 0);
            tma::load_3d(sX + FWD_TILE3_FLOATS, &maps.src[1], x0 - XOFF, y0 - 1, 3 * b, sBar + 0);
            tma::mbar_expect_tx(sBar + 1, FWD_CH * 4);
            tma::load_3d(sInv, &maps.inv[0], x0 - XOFF, y0 - 1, b, sBar + 1);
        }
        __syncthreads();                 // barrier init visible
        tma::mbar_wait(sBar + 0, 0);
        if (border) {                    // CTA-uniform
            // the three tiles are contiguous planes of FWD_TILE3_FLOATS / FWD_CH floats: patch them in one go
            patch_reflect<1, FWD_ROWS>(sY, 3, FWD_CH, x0, y0, H, W, tid, NT);
            patch_reflect<1, FWD_ROWS>(sX, 3, FWD_CH, x0, y0, H, W, tid, NT);
            patch_reflect<1, FWD_ROWS>(sX + FWD_TILE3_FLOATS, 3, FWD_CH, x0, y0, H, W, tid, NT);
            __syncthreads();
        }
    } else {
        fwd_load_tile(p.tgt + (size_t)b * 3 * HW, sY, x0, y0, H, W, tid);
        fwd_load_tile(p.src[0] + (size_t)b * 3 * HW, sX, x0, y0, H, W, tid);
        fwd_load_tile(p.src[1] + (size_t)b * 3 * HW, sX + FWD_TILE3_FLOATS, x0, y0, H, W, tid);
        __syncthreads();
    }

    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const float rw = exact::rcp_refined(wm1), rh = exact::rcp_refined(hm1);

    // validity and mask of the 4 outputs
    bool valid[4];
    bool msk[4];
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

    const float* ys_t = sY + ty * PITCH + XOFF + 4 * tx;      // [ch 0][halo row ty][my output 0], 16B aligned
    float* yst_t = sYst + ty * TW + 4 * tx;
    // target statistics for my 4 outputs (shared by all 2+2n photometric evaluations)
    if constexpr (!L1ONLY) {
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float sy[4], syy[4];
#pragma unroll
        for (int dy = 0; dy < 3; dy++) {
            const float* yr = ys_t + ch * FWD_CH + dy * PITCH;
            float4 ym = *reinterpret_cast<const float4*>(yr);
            float y6[6] = {yr[-1], ym.x, ym.y, ym.z, ym.w, yr[4]}, yy6[6];
#pragma unroll
            for (int j = 0; j < 6; j++) yy6[j] = __fmul_rn(y6[j], y6[j]);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (dy == 0) { sy[k] = y6[k]; syy[k] = yy6[k]; }
                else { sy[k] = __fadd_rn(sy[k], y6[k]); syy[k] = __fadd_rn(syy[k], yy6[k]); }
                sy[k] = __fadd_rn(sy[k], y6[k + 1]); syy[k] = __fadd_rn(syy[k], yy6[k + 1]);
                sy[k] = __fadd_rn(sy[k], y6[k + 2]); syy[k] = __fadd_rn(syy[k], yy6[k + 2]);
            }
        }
        float mu[4], ms[4], sg[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            mu[k] = exact::div9(sy[k]);
            ms[k] = __fmul_rn(mu[k], mu[k]);
            sg[k] = __fadd_rn(exact::div9(syy[k]), -ms[k]);
        }
        *reinterpret_cast<float4*>(yst_t + (ch * 2 + 0) * TH * TW) = make_float4(mu[0], mu[1], mu[2], mu[3]);
        *reinterpret_cast<float4*>(yst_t + (ch * 2 + 1) * TH * TW) = make_float4(sg[0], sg[1], sg[2], sg[3]);
    }
    }   // !L1ONLY
    // (each thread only reads back its own statistics: no barrier needed)

    // identity-reprojection losses (un-warped source vs target), once per tile
    float lid0[4] = {0, 0, 0, 0}, lid1[4] = {0, 0, 0, 0};
    unsigned idnext = 0;     // bit k: the smaller identity loss of output k is id_next (index 3), else id_prev (index 1)
    // L1ONLY (ssim_loss_weight == 0): calc_photometric_loss returns the raw 3-channel |x - y| (loss.py:195-196), every list
    // entry contributes 3 channels to the min and the selection index is entry * 3 + channel
    float lidc[L1ONLY ? 2 : 1][3][4];
    if constexpr (L1ONLY) {
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int ch = 0; ch < 3; ch++)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int o = ch * FWD_CH + (ty + 1) * PITCH + XOFF + 4 * tx + k;
                    lidc[s][ch][k] = fabsf(__fadd_rn(sX[s * FWD_TILE3_FLOATS + o], -sY[o]));
                }
    } else if (p.automask) {
        photometric4<false>(sX + ty * PITCH + XOFF + 4 * tx, ys_t, yst_t, p.alpha, p.oma, lid0, NoEmit());
        photometric4<false>(sX + FWD_TILE3_FLOATS + ty * PITCH + XOFF + 4 * tx, ys_t, yst_t, p.alpha, p.oma, lid1, NoEmit());
        // Only the smaller identity loss can ever win the strict-`<` scan over [warp_prev, id_prev, warp_next, id_next]: keep
        // min(id_prev, id_next) and which one it is (id_prev on ties: lower index) -- 4 live values instead of 8 through the scale loop.
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (lid1[k] < lid0[k]) { lid0[k] = lid1[k]; idnext |= 1u << k; }
        }
    }

    // edge-aware smoothness weights exp(-mean_c |dI|) (depth.py:23-24), premultiplied by mask/validity
    float wxm[4], wym[4];
    float cntN = 0.f, cntX = 0.f, cntY = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float* yc = sY + (ty + 1) * PITCH + XOFF + 4 * tx + k;
        float ax = 0.f, ay = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            float c0 = yc[ch * FWD_CH];
            float dx = fabsf(__fadd_rn(c0, -yc[ch * FWD_CH + 1]));
            float dyv = fabsf(__fadd_rn(c0, -yc[ch * FWD_CH + PITCH]));
            ax = ch == 0 ? dx : __fadd_rn(ax, dx);
            ay = ch == 0 ? dyv : __fadd_rn(ay, dyv);
        }
        bool hx = msk[k] && (u0 + k + 1 < W), hy = msk[k] && (v + 1 < H);
        wxm[k] = hx ? expf(-exact::div3(ax)) : 0.f;
        wym[k] = hy ? expf(-exact::div3(ay)) : 0.f;
        cntN += msk[k] ? 1.f : 0.f;
        cntX += hx ? 1.f : 0.f;
        cntY += hy ? 1.f : 0.f;
    }
    // parked in shared memory (own slot, no barrier needed) so that they do not occupy 8 registers through the SSIM stages
    sWgt[tid] = make_float4(wxm[0], wxm[1], wxm[2], wxm[3]);
    sWgt[NT + tid] = make_float4(wym[0], wym[1], wym[2], wym[3]);
    // mask counts: reduce now (slots 4..6 of sRed are not used by the per-scale sums), stored after the barrier below
    cntN = warp_sum(cntN); cntX = warp_sum(cntX); cntY = warp_sum(cntY);
    if (lane == 0) { sRed[warp * 8 + 4] = cntN; sRed[warp * 8 + 5] = cntX; sRed[warp * 8 + 6] = cntY; }
    if (STASH && v < H && (x0 >> 2) + tx < p.Wg) {
        // the stash backward needs these weights at q, q-1 and q-W: leave them in the stash instead of having it redo the expf
        float* wp = p.wgt + ((size_t)(2 * b) * H + v) * (4 * p.Wg) + u0;
        *reinterpret_cast<float4*>(wp) = make_float4(wxm[0], wxm[1], wxm[2], wxm[3]);
        *reinterpret_cast<float4*>(wp + (size_t)H * 4 * p.Wg) = make_float4(wym[0], wym[1], wym[2], wym[3]);
    }
    // Everything above reads only the caller's tensors.  The camera table and the packed source copies are written
    // by pack_sources_kernel, which may still be running (programmatic dependent launch): wait for it here.
    pdl_wait();
    pdl_trigger();     // let reduce_kernel's CTAs be scheduled; they wait for this grid to finish
    if (tid < 48) sCam[tid] = reinterpret_cast<const float*>(p.cams + b)[tid];
    __syncthreads();   // identity evaluation done: sX may be overwritten; camera table visible

    const int nq = 4 * p.n + 3;
    double* my_partials = p.partials + (size_t)tile * nq;
    if (tid < 3) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) acc += (double)sRed[w * 8 + 4 + tid];
        my_partials[4 * p.n + tid] = acc;
    }
    const size_t pimg = (size_t)(H + 2 * PACK_BORDER) * (W + 2 * PACK_BORDER);
    const float4* src0 = p.psrc[0] + (size_t)b * pimg;
    const float4* src1 = p.psrc[1] + (size_t)b * pimg;

    // NOTE (measured, round 1): a software pipeline by source -- SSIM(s0, i) || warp(s1, i), then SSIM(s1, i) || warp(s0, i+1),
    // the two source buffers alternating roles, half of the warps gathering first and half starting with their SSIM strip
    // -- was bit-identical but SLOWER (C2 forward 0.357 -> 0.379 ms, 0.291 -> 0.320 without the stash): the gather stage
    // is a latency chain, so warping one source per phase pays that chain twice per scale instead of once.  Overlapping the
    // warp of BOTH sources of scale i+1 with the SSIM of scale i would need two more source tiles (31 KB) of shared memory.
    for (int i = 0; i < p.n; i++) {
        const float* inv = p.inv[i] + (size_t)b * HW;
        const float* sI = sInv + (i & 1) * FWD_INV_FLOATS;
        if (USE_TMA) {
            // prefetch the next scale's inverse-depth tile into the other ring slot (its last readers
            // finished before the barrier that ended the previous scale), then wait for this scale's tile
            if (tid == 0 && i + 1 < p.n) {
                tma::fence_proxy_async();
                tma::mbar_expect_tx(sBar + 1 + ((i + 1) & 1), FWD_CH * 4);
                tma::load_3d(sInv + ((i + 1) & 1) * FWD_INV_FLOATS, &maps.inv[i + 1], x0 - XOFF, y0 - 1, b, sBar + 1 + ((i + 1) & 1));
            }
            tma::mbar_wait(sBar + 1 + (i & 1), (i >> 1) & 1);
        }
        // ---- stage 1: warp both sources at every halo pixel (software pipelined, mgvs_device.cuh) ----
#if !(MGVS_ABL & 4)
        warp_tile<1, FWD_ROWS, FWD_CH, USE_TMA, PAD>(sX, sX + FWD_TILE3_FLOATS, sI, inv, src0, src1, sCam, x0, y0, H, W, border,
                                                      wm1, hm1, rw, rh, tid, p.pad);
#endif
        __syncthreads();

        // ---- stage 2: photometric maps, min/argmin, smoothness ----
        float lw0[4], lw1[4];
        // STASH: coefficient texels (a, b, c, [source == 0]) of this scale, one float4 per (channel, pixel), in the
        // phase-major layout [scale][image][channel][row][u & 3][u >> 2] -- for a fixed output k the 16 threads of a
        // tile row write 256 contiguous bytes, and the backward's 3x3 taps are bank-conflict-free 128-bit loads.
        // Source 0's texels are stored as they are produced; source 1's wait in registers for the argmin.
        float4* st = nullptr;
        const int st_ch = H * 4 * p.Wg;          // texels per channel map (32-bit offsets: 3*H*4*Wg < 2^31 by check_problem)
        float c1[3][3][4];
        // rows below the image and column groups right of it do not exist in the stash (the backward's TMA zero-fills them)
        const bool row_ok = STASH && v < H && (x0 >> 2) + tx < p.Wg;
#if MGVS_ABL & 2
        for (int k = 0; k < 4; k++) { lw0[k] = sX[ty * PITCH + XOFF + 4 * tx + k]; lw1[k] = sX[FWD_TILE3_FLOATS + ty * PITCH + XOFF + 4 * tx + k]; }
        for (int ch = 0; ch < 3; ch++) for (int m = 0; m < 3; m++) for (int k = 0; k < 4; k++) c1[ch][m][k] = lw1[k];
        if (STASH) st = p.stash + ((size_t)i * p.B + b) * 3 * st_ch + (size_t)v * 4 * p.Wg + (x0 >> 2) + tx;
        if (false) {
#else
        if (STASH) {
#endif
            st = p.stash + ((size_t)i * p.B + b) * 3 * st_ch + (size_t)v * 4 * p.Wg + (x0 >> 2) + tx;
#if MGVS_ROLL_SRC
#pragma unroll 1
            for (int s = 0; s < 2; s++) {
                float lw[4];
                photometric4<true>(sX + s * FWD_TILE3_FLOATS + ty * PITCH + XOFF + 4 * tx, ys_t, yst_t, p.alpha, p.oma, lw,
                                   [&](int ch, int k, float a, float bq, float c) {
                                       if (s == 0) { if (row_ok) __stcs(st + ch * st_ch + k * p.Wg, make_float4(a, bq, c, 1.f)); }
                                       else { c1[ch][0][k] = a; c1[ch][1][k] = bq; c1[ch][2][k] = c; }
                                   });
#pragma unroll
                for (int k = 0; k < 4; k++) { if (s == 0) lw0[k] = lw[k]; else lw1[k] = lw[k]; }
            }
            if (false)
#endif
            photometric4<true>(sX + ty * PITCH + XOFF + 4 * tx, ys_t, yst_t, p.alpha, p.oma, lw0,
                               [&](int ch, int k, float a, float bq, float c) {
                                   if (row_ok) __stcs(st + ch * st_ch + k * p.Wg, make_float4(a, bq, c, 1.f));     // streaming: written once, read once
                               });
#if MGVS_ROLL_SRC
            if (false)
#endif
            photometric4<true>(sX + FWD_TILE3_FLOATS + ty * PITCH + XOFF + 4 * tx, ys_t, yst_t, p.alpha, p.oma, lw1,
                               [&](int ch, int k, float a, float bq, float c) { c1[ch][0][k] = a; c1[ch][1][k] = bq; c1[ch][2][k] = c; });
        } else if constexpr (!L1ONLY && !(MGVS_ABL & 2)) {
            photometric4<false>(sX + ty * PITCH + XOFF + 4 * tx, ys_t, yst_t, p.alpha, p.oma, lw0, NoEmit());
            photometric4<false>(sX + FWD_TILE3_FLOATS + ty * PITCH + XOFF + 4 * tx, ys_t, yst_t, p.alpha, p.oma, lw1, NoEmit());
        }
        float photo = 0.f;
        unsigned selw = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float best = L1ONLY ? 0.f : lw0[k];
            unsigned bi = 0;
            if constexpr (L1ONLY) {
                // strict `<` scan over [warp_prev c0..c2, (id_prev c0..c2,) warp_next c0..c2 (, id_next c0..c2)]
                float w[2][3];
#pragma unroll
                for (int s = 0; s < 2; s++)
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        const int o = ch * FWD_CH + (ty + 1) * PITCH + XOFF + 4 * tx + k;
                        w[s][ch] = fabsf(__fadd_rn(sX[s * FWD_TILE3_FLOATS + o], -sY[o]));
                    }
                best = w[0][0];
                unsigned idx = 0;
#pragma unroll
                for (int s = 0; s < 2; s++) {
#pragma unroll
                    for (int ch = 0; ch < 3; ch++, idx++) if (w[s][ch] < best) { best = w[s][ch]; bi = idx; }
                    if (p.automask) {
#pragma unroll
                        for (int ch = 0; ch < 3; ch++, idx++) if (lidc[s][ch][k] < best) { best = lidc[s][ch][k]; bi = idx; }
                    }
                }
            } else
            if (p.automask) {
                // same winner as the scan over all four (lid0 now holds min(id_prev, id_next), see above)
                if ((idnext >> k) & 1u) {
                    if (lw1[k] < best) { best = lw1[k]; bi = 2; }
                    if (lid0[k] < best) { best = lid0[k]; bi = 3; }
                } else {
                    if (lid0[k] < best) { best = lid0[k]; bi = 1; }
                    if (lw1[k] < best) { best = lw1[k]; bi = 2; }
                }
            } else {
                if (lw1[k] < best) { best = lw1[k]; bi = 1; }
            }
            if (msk[k]) photo += best;
            selw |= bi << (8 * k);
            if (STASH && row_ok) {
                // fix up the eagerly stored texels: keep source 0's where it won, source 1's where it won, zeros elsewhere
                // (identity reprojection selected, mask false, or outside the image)
                const unsigned warp1 = p.automask ? 2u : 1u;
                const bool w0 = msk[k] && bi == 0u, w1 = msk[k] && bi == warp1;
                if (!w0) {
#pragma unroll
                    for (int ch = 0; ch < 3; ch++)
                        __stcs(st + ch * st_ch + k * p.Wg, make_float4(w1 ? c1[ch][0][k] : 0.f, w1 ? c1[ch][1][k] : 0.f, w1 ? c1[ch][2][k] : 0.f, 0.f));
                }
            }
        }
        if (p.sel != nullptr && v < H) {
            unsigned char* sp = p.sel + ((size_t)i * p.B + b) * HW + (size_t)v * W + u0;
            if (u0 + 3 < W && ((W & 3) == 0)) *reinterpret_cast<unsigned*>(sp) = selw;
            else {
#pragma unroll
                for (int k = 0; k < 4; k++) if (u0 + k < W) sp[k] = (unsigned char)((selw >> (8 * k)) & 0xff);
            }
        }
        // smoothness: |inv[p]-inv[p+1]| * w_x * mask (the per-image 1/mean is applied in the finalise pass)
        float smx = 0.f, smy = 0.f, isum = 0.f;
        if (v < H) {
            const float* ir = inv + (size_t)v * W;
            const float* si = sI + (ty + 1) * PITCH + XOFF + 4 * tx;   // my 4 outputs in the smem tile
            float c4[5];
#pragma unroll
            for (int k = 0; k < 5; k++) c4[k] = USE_TMA ? si[k] : ((u0 + k < W) ? __ldg(ir + u0 + k) : 0.f);
            const float4 wx4 = sWgt[tid], wy4 = sWgt[NT + tid];
            const float wxs[4] = {wx4.x, wx4.y, wx4.z, wx4.w}, wys[4] = {wy4.x, wy4.y, wy4.z, wy4.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float below = USE_TMA ? si[PITCH + k] : ((v + 1 < H && u0 + k < W) ? __ldg(ir + W + u0 + k) : 0.f);
                smx += wxs[k] * fabsf(c4[k] - c4[k + 1]);
                smy += wys[k] * fabsf(c4[k] - below);
                isum += valid[k] ? c4[k] : 0.f;
            }
        }
        photo = warp_sum(photo); smx = warp_sum(smx); smy = warp_sum(smy); isum = warp_sum(isum);
        if (lane == 0) { sRed[warp * 8 + 0] = photo; sRed[warp * 8 + 1] = smx; sRed[warp * 8 + 2] = smy; sRed[warp * 8 + 3] = isum; }
        __syncthreads();   // sX free for the next scale; sRed complete
        if (tid < 4) {
            double acc = 0.0;
#pragma unroll
            for (int w = 0; w < NT / 32; w++) acc += (double)sRed[w * 8 + tid];
            my_partials[tid * p.n + i] = acc;   // [photo | smx | smy | invsum][n]
        }
    }
}

}  // namespace mgvs
