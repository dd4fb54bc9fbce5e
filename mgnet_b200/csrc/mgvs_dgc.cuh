// DGC depth rescaling: the inference-time geometry tail of MGNet, fused.
//
// Reference: mgnet/postprocessing/depth_post_proc.py:11-185 (get_depth_prediction, _get_scale_recovery,
// _get_surface_normal, _get_ground_mask), Camera.reconstruct(frame="c") (camera.py:107-136) and the same arithmetic
// in exportable_post_proc.py:52-79.  The reference materialises 13 [3,H,W] temporaries for the normals, runs
// masked_select (host sync) and a full sort-based median; here
//   dgc_heights_kernel      points on tile+1 in shared memory, the 4 cross-product normals, their mean, the camera
//                           height |P.N| and the ground decision per pixel; ground heights are COMPACTED (order does
//                           not matter for a median) and a level-1 radix histogram of their bit patterns is built
//   dgc_refine_kernel<L>    radix select, levels 2 and 3 (11 + 10 + 10 bits of the non-negative float): histogram
//                           of the keys that share the prefix found so far
//   dgc_apply_kernel        finds the median bit pattern from the three histograms, scale = (1/median)*real_height,
//                           then depth *= scale, points = (Kinv grid * depth) * scale, class filter
// Integer atomics only (histograms, compaction cursor): the result is deterministic.  Every fp32 operation follows
// the rounding sequence of the reference on CPU (probed bit for bit, see oracle/dgc_oracle.c).
#pragma once
#include "mgvs_device.cuh"

namespace mgvs {
namespace dgc {

constexpr int TW = 64, TH = 16, NT = 256;
constexpr int L1_BITS = 11, L2_BITS = 10, L3_BITS = 10;
constexpr int L1_BINS = 1 << L1_BITS, L2_BINS = 1 << L2_BITS, L3_BINS = 1 << L3_BITS;
constexpr int MAX_FILTER = 16;

// per-image selection state in the workspace
struct State {
    unsigned hist1[L1_BINS];
    unsigned hist2[L2_BINS];
    unsigned hist3[L3_BINS];
    unsigned count;      // number of ground pixels == compaction cursor
    unsigned nan_flag;   // a ground pixel's height was NaN -> torch.median returns NaN
    unsigned ticket[3];  // [2]: blocks of pass 3 that are done -- the last one runs the final bin search once for everybody
    unsigned median;     // bit pattern of the median, written by that block
    unsigned pad[2];
};

struct FilterIds {
    long long id[MAX_FILTER];
    int n;
};

__device__ __forceinline__ void kinv_of(const float* __restrict__ camera, long long crs, int is_inverse, float Kinv[9])
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) Kinv[r * 3 + k] = camera[r * crs + k];
    if (!is_inverse) {   // Camera.Kinv closed form (camera.py:72-81)
        const float fx = Kinv[0], fy = Kinv[4], cx = Kinv[2], cy = Kinv[5];
        Kinv[0] = __fdiv_rn(1.0f, fx);
        Kinv[4] = __fdiv_rn(1.0f, fy);
        Kinv[2] = __fdiv_rn(__fmul_rn(-1.0f, cx), fx);
        Kinv[5] = __fdiv_rn(__fmul_rn(-1.0f, cy), fy);
    }
}

__device__ __forceinline__ void cross3(const float a[3], const float b[3], float o[3])
{   // ATen's cross on CPU rounds as fma(a1, b2, -(a2*b1))
    o[0] = __fmaf_rn(a[1], b[2], -__fmul_rn(a[2], b[1]));
    o[1] = __fmaf_rn(a[2], b[0], -__fmul_rn(a[0], b[2]));
    o[2] = __fmaf_rn(a[0], b[1], -__fmul_rn(a[1], b[0]));
}

__device__ __forceinline__ float norm3(const float v[3])
{
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
}

// F.normalize(dim=1): v / max(||v||, eps), IEEE quotients through one shared refined reciprocal
__device__ __forceinline__ void normalize3(float v[3], float eps)
{
    const float nn = norm3(v);
    const float d = (nn != nn) ? nn : fmaxf(nn, eps);
    const float y = exact::rcp_refined(d);
#pragma unroll
    for (int j = 0; j < 3; j++) v[j] = exact::div_by(v[j], d, y);
}

template <typename PanT>
__device__ __forceinline__ long long pan_at(const void* pan, size_t i) { return (long long)((const PanT*)pan)[i]; }

// Finds the bin holding element `rank` (0-based) of a histogram: 256 threads, bins/256 bins per thread.
// Returns (bin, rank inside the bin) to every thread; found == 0 when rank >= total.
template <int BINS>
__device__ __forceinline__ void find_bin(const unsigned* __restrict__ hist, unsigned rank, unsigned* s_scan /*[16]*/,
                                         unsigned& bin, unsigned& rank_in_bin, unsigned& found)
{
    constexpr int PER = BINS / NT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned c[PER], s = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { c[k] = __ldcg(hist + tid * PER + k); s += c[k]; }
    unsigned inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) s_scan[warp] = inc;
    if (tid == 0) { s_scan[8] = 0xffffffffu; s_scan[9] = 0; s_scan[10] = 0; }
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < warp; w++) base += s_scan[w];
    unsigned excl = base + inc - s;
    if (rank >= excl && rank < excl + s) {   // exactly one thread
        unsigned r = rank - excl;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (r < c[k]) { s_scan[8] = tid * PER + k; s_scan[9] = r; s_scan[10] = 1; break; }
            r -= c[k];
        }
    }
    __syncthreads();
    bin = s_scan[8]; rank_in_bin = s_scan[9]; found = s_scan[10];
    __syncthreads();
}

// Takes a ticket for pass `pass` of image state `st`; true (for every thread of the block) in the block that arrives last,
// after which all histogram updates of the pass are visible to it.
__device__ __forceinline__ bool last_block(State* st, int pass, unsigned nblocks, unsigned* s_flag)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *s_flag = (atomicAdd(&st->ticket[pass], 1u) == nblocks - 1u) ? 1u : 0u;
    __syncthreads();
    const bool last = *s_flag != 0u;
    if (last) __threadfence();
    return last;
}

// ---------------------------------------------------------------------------------------------------------------
template <typename PanT, bool AUTO_MASK>
__global__ void __launch_bounds__(NT) dgc_heights_kernel(int H, int W, const float* __restrict__ depth,
                                                         const float* __restrict__ camera, long long cbs, long long crs,
                                                         int cam_is_inverse, const void* __restrict__ panoptic,
                                                         long long road_id, unsigned* __restrict__ keys, State* __restrict__ states,
                                                         float* __restrict__ dbg_heights, unsigned char* __restrict__ dbg_ground)
{
    __shared__ float sP[3][TH + 2][TW + 2];
    __shared__ unsigned sHist[L1_BINS];
    __shared__ unsigned sScan[16];
    const int tid = threadIdx.x, b = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const size_t HW = (size_t)H * W;
    const float* dimg = depth + (size_t)b * HW;
    State* st = states + b;
    float Kinv[9];
    kinv_of(camera + b * cbs, crs, cam_is_inverse, Kinv);
    for (int i = tid; i < L1_BINS; i += NT) sHist[i] = 0;
    // points of tile+1 (Camera.reconstruct "c": xnorm = Kinv.bmm(grid); Xc = xnorm * depth)
    for (int i = tid; i < (TH + 2) * (TW + 2); i += NT) {
        const int ly = i / (TW + 2), lx = i - ly * (TW + 2);
        const int v = min(max(y0 + ly - 1, 0), H - 1), u = min(max(x0 + lx - 1, 0), W - 1);
        float r[3];
        exact::ray(Kinv, u, v, r);
        const float d = __ldg(dimg + (size_t)v * W + u);
#pragma unroll
        for (int j = 0; j < 3; j++) sP[j][ly][lx] = __fmul_rn(r[j], d);
    }
    __syncthreads();
    const float thr = __uint_as_float(0x3f7f069eu);   // float32(cos(radians(5))) = 0.99619472 (depth_post_proc.py:174)
    const int tx = tid & (TW - 1), ty0 = tid / TW;  // 4 rows of 64 threads, each thread 4 pixels (rows ty0 + 4k)
    unsigned mykeys[TH / 4];
    unsigned nmine = 0;
#pragma unroll
    for (int k = 0; k < TH / 4; k++) {
        const int ly = ty0 + 4 * k, x = x0 + tx, y = y0 + ly;
        mykeys[k] = 0xffffffffu;
        if (x >= W || y >= H) continue;
        // F.pad(normals, "replicate"): a border pixel takes the normal of the nearest interior pixel
        const int qx = min(max(x, 1), W - 2) - x0 + 1, qy = min(max(y, 1), H - 2) - y0 + 1;
        float c[3], d[8][3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            c[j] = sP[j][qy][qx];
            d[0][j] = __fsub_rn(sP[j][qy][qx - 1], c[j]);       // x0
            d[1][j] = __fsub_rn(sP[j][qy - 1][qx], c[j]);       // y0
            d[2][j] = __fsub_rn(sP[j][qy][qx + 1], c[j]);       // x1
            d[3][j] = __fsub_rn(sP[j][qy + 1][qx], c[j]);       // y1
            d[4][j] = __fsub_rn(sP[j][qy - 1][qx - 1], c[j]);   // x0y0
            d[5][j] = __fsub_rn(sP[j][qy + 1][qx - 1], c[j]);   // x0y1
            d[6][j] = __fsub_rn(sP[j][qy - 1][qx + 1], c[j]);   // x1y0
            d[7][j] = __fsub_rn(sP[j][qy + 1][qx + 1], c[j]);   // x1y1
        }
        float n[4][3], m[3];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            cross3(d[2 * q], d[2 * q + 1], n[q]);
            normalize3(n[q], 1e-12f);
        }
#pragma unroll
        for (int j = 0; j < 3; j++)
            m[j] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(n[0][j], n[1][j]), n[2][j]), n[3][j]), 0.25f);
        normalize3(m, 1e-12f);
        const float p0 = sP[0][ly + 1][tx + 1], p1 = sP[1][ly + 1][tx + 1], p2 = sP[2][ly + 1][tx + 1];
        const float h = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(p0, m[0]), __fmul_rn(p1, m[1])), __fmul_rn(p2, m[2])));
        bool ground;
        if (AUTO_MASK) {   // _get_ground_mask: |cosine_similarity(normal, (0,1,0))| > cos(5 deg) and y > 0
            const float nn = norm3(m);
            const float dn = fmaxf(nn, 1e-6f);
            const float yy = exact::rcp_refined(dn);
            const float a0 = exact::div_by(m[0], dn, yy), a1 = exact::div_by(m[1], dn, yy), a2 = exact::div_by(m[2], dn, yy);
            const float cs = __fadd_rn(__fadd_rn(__fmul_rn(a0, 0.0f), __fmul_rn(a1, 1.0f)), __fmul_rn(a2, 0.0f));
            ground = ((cs > thr) || (cs < -thr)) && !(p1 <= 0.0f);
        } else {
            ground = pan_at<PanT>(panoptic, (size_t)b * HW + (size_t)y * W + x) == road_id;
        }
        if (dbg_heights) dbg_heights[(size_t)b * HW + (size_t)y * W + x] = h;
        if (dbg_ground) dbg_ground[(size_t)b * HW + (size_t)y * W + x] = ground ? 1 : 0;
        if (ground) {
            if (h != h) {
                atomicOr(&st->nan_flag, 1u);
            } else {
                mykeys[k] = __float_as_uint(h);
                nmine++;
                atomicAdd(&sHist[mykeys[k] >> (L2_BITS + L3_BITS)], 1u);
            }
        }
    }
    // compaction: block-exclusive scan of the per-thread counts, one cursor bump per block
    const int lane = tid & 31, warp = tid >> 5;
    unsigned inc = nmine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sScan[warp] = inc;
    __syncthreads();
    if (tid == 0) {
        unsigned tot = 0;
        for (int w = 0; w < NT / 32; w++) { unsigned t = sScan[w]; sScan[w] = tot; tot += t; }
        sScan[8] = tot ? atomicAdd(&st->count, tot) : 0u;
    }
    __syncthreads();
    unsigned pos = sScan[8] + sScan[warp] + inc - nmine;
    unsigned* kout = keys + (size_t)b * HW;
#pragma unroll
    for (int k = 0; k < TH / 4; k++)
        if (mykeys[k] != 0xffffffffu) kout[pos++] = mykeys[k];
    for (int i = tid; i < L1_BINS; i += NT) {
        const unsigned c = sHist[i];
        if (c) atomicAdd(&st->hist1[i], c);
    }
}

// LEVEL 2: histogram of bits [19:10] of the keys whose bits [30:20] equal the level-1 bin of the median.
// LEVEL 3: histogram of bits [9:0] of the keys whose bits [30:10] equal the prefix found so far.
template <int LEVEL>
__global__ void __launch_bounds__(NT) dgc_refine_kernel(size_t HW, const unsigned* __restrict__ keys, State* __restrict__ states)
{
    __shared__ unsigned sHist[L2_BINS];
    __shared__ unsigned sScan[16];
    const int tid = threadIdx.x, b = blockIdx.y;
    State* st = states + b;
    const unsigned count = __ldcg(&st->count);
    if (count == 0) return;
    // every block repeats the (cheap) bin search of the earlier passes; only the very last search -- which the apply kernel
    // would otherwise repeat in each of its ~1000 blocks -- is done once, by the last block of pass 3
    // (measured: a last-block hand-over in every pass costs more in fences than the repeated searches, profiles/r01k)
    unsigned bin1, rank, found;
    find_bin<L1_BINS>(st->hist1, (count - 1) >> 1, sScan, bin1, rank, found);
    unsigned prefix = bin1;
    constexpr unsigned shift = LEVEL == 2 ? L2_BITS + L3_BITS : L3_BITS;
    if (LEVEL == 3) {
        unsigned bin2;
        find_bin<L2_BINS>(st->hist2, rank, sScan, bin2, rank, found);
        prefix = (bin1 << L2_BITS) | bin2;
    }
    if ((size_t)blockIdx.x * NT < count) {      // block-uniform
        for (int i = tid; i < L2_BINS; i += NT) sHist[i] = 0;
        __syncthreads();
        const unsigned* kin = keys + (size_t)b * HW;
        for (size_t i = (size_t)blockIdx.x * NT + tid; i < count; i += (size_t)gridDim.x * NT) {
            const unsigned k = kin[i];
            if ((k >> shift) == prefix) atomicAdd(&sHist[(k >> (shift - 10)) & 1023u], 1u);
        }
        __syncthreads();
        unsigned* gh = LEVEL == 2 ? st->hist2 : st->hist3;
        for (int i = tid; i < L2_BINS; i += NT) {
            const unsigned c = sHist[i];
            if (c) atomicAdd(&gh[i], c);
        }
    }
    if (LEVEL == 3) {
        if (last_block(st, 2, gridDim.x, &sScan[12])) {
            unsigned bin3;
            find_bin<L3_BINS>(st->hist3, rank, sScan, bin3, rank, found);
            if (tid == 0) st->median = (prefix << L3_BITS) | bin3;
        }
    }
}

// scale[b] = reciprocal(median) * real_height (depth_post_proc.py:100-102); depth *= scale; points = Xc * scale;
// depth[panoptic == id] = 0, points[:, panoptic == id] = NaN (depth_post_proc.py:61-69).
template <typename PanT, bool VEC4>
__global__ void __launch_bounds__(NT) dgc_apply_kernel(int H, int W, float* __restrict__ depth, const float* __restrict__ camera,
                                                       long long cbs, long long crs, int cam_is_inverse,
                                                       const float* __restrict__ real_height, long long rh_stride,
                                                       const void* __restrict__ panoptic, FilterIds ids, int use_dgc,
                                                       float* __restrict__ points, float* __restrict__ scale_out,
                                                       long long* __restrict__ count_out, const State* __restrict__ states)
{
    const int tid = threadIdx.x, b = blockIdx.y;
    const size_t HW = (size_t)H * W;
    float scale = 1.0f;
    if (use_dgc) {
        const State* st = states + b;
        const unsigned count = st->count;
        scale = __uint_as_float(0x7fc00000u);
        if (count > 0 && !st->nan_flag) scale = __fmul_rn(__frcp_rn(__uint_as_float(st->median)), real_height[b * rh_stride]);
        if (blockIdx.x == 0 && tid == 0) {
            scale_out[b] = scale;
            if (count_out) count_out[b] = (long long)count + (st->nan_flag ? 1 : 0);
        }
    }
    float Kinv[9];
    if (points) kinv_of(camera + b * cbs, crs, cam_is_inverse, Kinv);
    float* dimg = depth + (size_t)b * HW;
    float* pimg = points ? points + (size_t)b * 3 * HW : nullptr;
    const float qnan = __uint_as_float(0x7fc00000u);
    constexpr int V = VEC4 ? 4 : 1;
    for (size_t i = ((size_t)blockIdx.x * NT + tid) * V; i < HW; i += (size_t)gridDim.x * NT * V) {
        const int v = (int)(i / W), u = (int)(i - (size_t)v * W);
        float d[V], o[V], px[V], py[V], pz[V];
        if constexpr (VEC4) {
            const float4 t = *reinterpret_cast<const float4*>(dimg + i);
            d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
        } else {
            d[0] = dimg[i];
        }
#pragma unroll
        for (int e = 0; e < V; e++) {
            bool drop = false;
            if (panoptic && ids.n) {
                const long long pv = pan_at<PanT>(panoptic, (size_t)b * HW + i + e);
                for (int k = 0; k < ids.n; k++) drop |= (pv == ids.id[k]);
            }
            o[e] = drop ? 0.0f : __fmul_rn(d[e], scale);
            if (points) {
                float r[3];
                exact::ray(Kinv, u + e, v, r);
                px[e] = drop ? qnan : __fmul_rn(__fmul_rn(r[0], d[e]), scale);
                py[e] = drop ? qnan : __fmul_rn(__fmul_rn(r[1], d[e]), scale);
                pz[e] = drop ? qnan : __fmul_rn(__fmul_rn(r[2], d[e]), scale);
            }
        }
        if constexpr (VEC4) {
            __stcs(reinterpret_cast<float4*>(dimg + i), make_float4(o[0], o[1], o[2], o[3]));
            if (points) {
                __stcs(reinterpret_cast<float4*>(pimg + i), make_float4(px[0], px[1], px[2], px[3]));
                __stcs(reinterpret_cast<float4*>(pimg + HW + i), make_float4(py[0], py[1], py[2], py[3]));
                __stcs(reinterpret_cast<float4*>(pimg + 2 * HW + i), make_float4(pz[0], pz[1], pz[2], pz[3]));
            }
        } else {
            dimg[i] = o[0];
            if (points) { pimg[i] = px[0]; pimg[HW + i] = py[0]; pimg[2 * HW + i] = pz[0]; }
        }
    }
}

}  // namespace dgc
}  // namespace mgvs
