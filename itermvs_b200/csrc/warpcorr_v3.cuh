// ROUND-1 "v3" plane-sweep kernels, kept for A/B timing against warpcorr.cu (IMVS_WARPCORR_V3=1).
// Fused plane-sweep kernels: hypothesis generation + homography warp + bilinear sampling of the
// source feature pyramids + group-wise correlation (+ pixel-wise view-weighted aggregation in the
// iteration kernel).  The [C, D, H, W] warped volume of the reference (module.py:118-120) and the
// same-size product tensor (itermvs.py:50, 103) are never materialised.
//
// Reference: models/module.py:68-125, models/itermvs.py:11-19, 45-69, 86-120, 289-293.
//
// Thread mapping (both kernels).  Features are channels-last, so one sampled tap is C contiguous
// floats (64/128/192 B).  A warp is split into 4 "slots" of 8 lanes; lane g of a slot owns
// correlation group g, i.e. channels [g*C/8, (g+1)*C/8) -- exactly one float2 / float4 / 3xfloat2
// per tap, so a slot reads a tap as one fully used 64/128/192-byte segment and the group
// reduction needs no shuffles.  The 4 slots are the 4 depth samples of ONE pixel (2 pixels x 2
// samples at level 3): neighbouring hypotheses of a pixel land within ~a pixel of each other along
// the epipolar line, so the four slots of a load instruction mostly hit the same 128-byte lines
// (fewer L1 wavefronts than four different pixels would cost).  The sampling position of
// (sample, view) is computed once -- by lane (slot, g = view) -- and broadcast with shuffles.
#pragma once
#include "common.cuh"
#include "sampling.cuh"

namespace imvs {

template <int CPG>
__device__ __forceinline__ void load_group(const float* __restrict__ p, float (&v)[CPG]) {
    if constexpr (CPG == 4) {
        float4 t = ldg4(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int i = 0; i < CPG / 2; ++i) {
            float2 t = ldg2(p + 2 * i);
            v[2 * i] = t.x; v[2 * i + 1] = t.y;
        }
    }
}

// Sampling parameters of one (pixel, hypothesis, view), computed by the owner lane and broadcast:
// clamped top-left tap offset, +1 steps (0 when clamped at the border) and the four bilinear weights
// with the zero-padding of grid_sample folded in (weight 0 for taps outside the map), so that all
// taps can be loaded unconditionally from in-bounds addresses -- no branches, loads issue back to back.
struct TapSet {
    int o00, o01, o10, o11;     // element offsets of the 4 (clamped) taps inside one view: (y*Wf + x) * C
    float w00, w01, w10, w11;   // bilinear weights, 0 for taps outside the map
};

__device__ __forceinline__ TapSet make_tapset(const Tap& tp, int Wf, int Hf, int C) {
    TapSet ts;
    const int x0c = min(max(tp.x0, 0), Wf - 1), x1c = min(max(tp.x0 + 1, 0), Wf - 1);
    const int y0c = min(max(tp.y0, 0), Hf - 1), y1c = min(max(tp.y0 + 1, 0), Hf - 1);
    ts.o00 = (y0c * Wf + x0c) * C; ts.o01 = (y0c * Wf + x1c) * C;
    ts.o10 = (y1c * Wf + x0c) * C; ts.o11 = (y1c * Wf + x1c) * C;
    const float gx = 1.f - tp.fx, gy = 1.f - tp.fy;
    ts.w00 = (tp.mask & 1u) ? gx * gy : 0.f;
    ts.w01 = (tp.mask & 2u) ? tp.fx * gy : 0.f;
    ts.w10 = (tp.mask & 4u) ? gx * tp.fy : 0.f;
    ts.w11 = (tp.mask & 8u) ? tp.fx * tp.fy : 0.f;
    return ts;
}

// Owner lanes publish their tap set (+ view weight) in shared memory; every lane of the slot then reads
// it back as three broadcast 16-byte loads (cheaper than seven shuffles, and no per-lane offset math).
struct __align__(16) TapRecord { int4 off; float4 w; float4 extra; };     // extra.x = view weight

__device__ __forceinline__ void publish_tapset(TapRecord* rec, const TapSet& ts, float wv) {
    rec->off = make_int4(ts.o00, ts.o01, ts.o10, ts.o11);
    rec->w = make_float4(ts.w00, ts.w01, ts.w10, ts.w11);
    rec->extra = make_float4(wv, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ TapSet read_tapset(const TapRecord* rec, float& wv) {
    const int4 o = rec->off;
    const float4 w = rec->w;
    wv = rec->extra.x;
    TapSet ts;
    ts.o00 = o.x; ts.o01 = o.y; ts.o10 = o.z; ts.o11 = o.w;
    ts.w00 = w.x; ts.w01 = w.y; ts.w10 = w.z; ts.w11 = w.w;
    return ts;
}

// the four taps of this lane's channel group (issued back to back), then interpolate and dot with
// the reference feature group: mean_c( warped_c * ref_c )  (itermvs.py:50-51)
template <int CPG>
struct TapLoads { float t00[CPG], t01[CPG], t10[CPG], t11[CPG]; };

template <int CPG>
__device__ __forceinline__ void issue_taps(TapLoads<CPG>& L, const float* __restrict__ fea_view_g, const TapSet& ts) {
    load_group<CPG>(fea_view_g + ts.o00, L.t00);
    load_group<CPG>(fea_view_g + ts.o01, L.t01);
    load_group<CPG>(fea_view_g + ts.o10, L.t10);
    load_group<CPG>(fea_view_g + ts.o11, L.t11);
}

template <int CPG>
__device__ __forceinline__ float finish_taps(const TapLoads<CPG>& L, const TapSet& ts, const float (&ref)[CPG]) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < CPG; ++i) {
        float a = L.t00[i] * ts.w00;
        a = fmaf(L.t01[i], ts.w01, a);
        a = fmaf(L.t10[i], ts.w10, a);
        a = fmaf(L.t11[i], ts.w11, a);
        dot = fmaf(a, ref[i], dot);
    }
    return dot * (1.0f / (float)CPG);     // exact for 2 and 4 channels per group, <= 1 ulp for 6
}

// ---------------------------------------------------------------------------------------------
// K2: init plane sweep at level 3 (C = 48), per-view group correlation.
//   grid (ceil(W3/TPX), ceil(H3/8), B*DSPLIT), block 256 (8 warps = 8 rows)
// ---------------------------------------------------------------------------------------------
constexpr int INIT_TPX = 4;

__global__ void __launch_bounds__(256)
warpcorr_init_v3_kernel(const float* __restrict__ fea3, const float* __restrict__ rt3,
                     const float* __restrict__ depth_min, const float* __restrict__ depth_max,
                     const float* __restrict__ samples, float* __restrict__ corr, int B, int V, int H3, int W3, int D,
                     int dsplit) {
    constexpr int CPG = 6, C = 48;
    __shared__ float sP[IMVS_MAX_VIEWS * 12];
    __shared__ TapRecord sTap[8][4][8];          // [warp][slot][view of the current chunk]
    const int S = V - 1;
    const int b = blockIdx.z / dsplit, dpart = blockIdx.z % dsplit;
    for (int i = threadIdx.x; i < S * 12; i += blockDim.x) sP[i] = rt3[(size_t)b * S * 12 + i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane >> 3, g = lane & 7;
    const int y = blockIdx.y * 8 + warp;
    if (y >= H3) return;
    const int P3 = H3 * W3;
    const float inv_min = samples ? 0.f : 1.0f / depth_min[b], inv_max = samples ? 0.f : 1.0f / depth_max[b];
    const int chunks = (D + 3) / 4;
    const int c_begin = (chunks * dpart) / dsplit, c_end = (chunks * (dpart + 1)) / dsplit;
    const int x_begin = blockIdx.x * INIT_TPX, x_end = min(x_begin + INIT_TPX, W3);
    const float* ref_view = fea3 + (size_t)(b * V) * P3 * C;
    const size_t view_stride = (size_t)P3 * C;
    const float* src_base = fea3 + (size_t)(b * V + 1) * view_stride + g * CPG;      // source view 0, this lane's group

    for (int x = x_begin; x < x_end; ++x) {
        const int p = y * W3 + x;
        float ref[CPG];
        load_group<CPG>(ref_view + (size_t)p * C + g * CPG, ref);
        for (int ch = c_begin; ch < c_end; ++ch) {
            const int d = ch * 4 + slot;
            const bool dvalid = d < D;
            // itermvs.py:13-17 (or the caller's explicit hypotheses, Evaluation.forward's depth_sample)
            const float depth = samples ? ldg(samples + ((size_t)b * D + (dvalid ? d : 0)) * P3 + p)
                                        : 1.0f / (inv_max + ((float)d / (float)(D - 1)) * (inv_min - inv_max));
            for (int v0 = 0; v0 < S; v0 += 8) {
                Tap tp;
                tp.x0 = tp.y0 = 0; tp.fx = tp.fy = 0.f; tp.mask = 0u;
                if (v0 + g < S && dvalid)
                    tp = project_tap(sP + (v0 + g) * 12, (float)x, (float)y, depth, (float)W3, (float)H3, W3, H3);
                __syncwarp();
                publish_tapset(&sTap[warp][slot][g], make_tapset(tp, W3, H3, C), 0.f);
                __syncwarp();
                const int nv = min(8, S - v0);
                for (int j0 = 0; j0 < nv; j0 += 2) {          // two views per batch: 8 x 3 float2 loads in flight
                    TapSet ts[2];
                    TapLoads<CPG> L[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        float unused;
                        ts[u] = read_tapset(&sTap[warp][slot][min(j0 + u, 7)], unused);
                        const int v = min(v0 + j0 + u, S - 1);
                        issue_taps<CPG>(L[u], src_base + (size_t)v * view_stride, ts[u]);
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int v = v0 + j0 + u;
                        const float c = finish_taps<CPG>(L[u], ts[u], ref);
                        if (dvalid && v < S) corr[((((size_t)b * S + v) * D + d) * P3 + p) * 8 + g] = c;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: iteration kernel -- three pyramid levels, R = (4,4,2) samples per pixel around the current
// normalized depth, all source views, view-weighted aggregation.  One launch covers the three
// levels (blockIdx.z = b*3 + level).
//   grid (ceil(W2/ITER_TPX), ceil(H2/8), B*3), block 256 (8 warps = 8 rows)
// ---------------------------------------------------------------------------------------------
constexpr int ITER_TPX = 16;

struct IterParams {
    const float* fea[3];   // level 1,2,3 pyramids  [B][V][Hf][Wf][C]
    const float* rt[3];    // composed projections  [B][S][12]
    const float* nd;       // [B][nd_stride]
    size_t nd_stride, nd_pstride;
    const float* vw2;      // [B][S][P2]
    const float* depth_min;
    const float* depth_max;
    const float* samples[3];   // optional explicit hypotheses [B][R_l][P2] per level (else from nd)
    float* agg;            // [B][10][P2][8]
    int B, V, H2, W2;
};

// MODE: how the reference-view feature of this level is brought to level-2 resolution
// (itermvs.py:95-98): 0 same, 1 F.interpolate(x0.5) == 2x2 mean, 2 F.interpolate(x2) bilinear.
template <int CPG, int R, int MODE>
__device__ __forceinline__ void iter_level(const IterParams& prm, const float* sP, TapRecord (*sTap)[8], int b, int y, int x_begin,
                                           int x_end, int slice_base, float o0, float o1, float o2, float o3) {
    constexpr int C = CPG * 8;
    constexpr int PPS = 4 / R;     // pixels per warp step
    const int lane = threadIdx.x & 31;
    const int slot = lane >> 3, g = lane & 7;
    const int r = slot % R, pxo = slot / R;
    const int V = prm.V, S = V - 1, H2 = prm.H2, W2 = prm.W2, P2 = H2 * W2;
    const int Hf = MODE == 1 ? H2 * 2 : (MODE == 2 ? H2 / 2 : H2);
    const int Wf = MODE == 1 ? W2 * 2 : (MODE == 2 ? W2 / 2 : W2);
    const float sx = (float)((double)Wf / (double)W2), sy = (float)((double)Hf / (double)H2);   // module.py:95-96
    const float* fea = prm.fea[MODE == 1 ? 0 : (MODE == 0 ? 1 : 2)];
    const size_t view_stride = (size_t)Hf * Wf * C;
    const float* ref_view = fea + (size_t)(b * V) * view_stride + g * CPG;
    const float* src_base = ref_view + view_stride;                                   // source view 0, this lane's group
    const float* smp = prm.samples[MODE == 1 ? 0 : (MODE == 0 ? 1 : 2)];
    const float inv_min = smp ? 0.f : 1.0f / prm.depth_min[b], inv_max = smp ? 0.f : 1.0f / prm.depth_max[b];
    const float off = (r == 0 ? o0 : r == 1 ? o1 : r == 2 ? o2 : o3) * (1.0f / 256.0f);   // itermvs.py:229,290

    for (int xs = x_begin; xs < x_end; xs += PPS) {
        const int x = xs + pxo;
        const bool pvalid = x < x_end;
        const int xc = pvalid ? x : x_end - 1;
        const int p = y * W2 + xc;
        // hypotheses: itermvs.py:290-293
        float depth;
        if (smp) {
            depth = ldg(smp + ((size_t)b * R + r) * P2 + p);
        } else {
            const float ndv = ldg(prm.nd + (size_t)b * prm.nd_stride + (size_t)p * prm.nd_pstride);
            const float s = fminf(fmaxf(ndv + off, 0.f), 1.f);
            depth = unnormalize_depth(s, inv_min, inv_max);
        }
        // reference feature of this group at level-2 resolution
        float ref[CPG];
        if constexpr (MODE == 0) {
            load_group<CPG>(ref_view + (size_t)p * C, ref);
        } else if constexpr (MODE == 1) {
            float a[CPG], bq[CPG], c[CPG], d[CPG];
            const float* q = ref_view + ((size_t)(2 * y) * Wf + 2 * xc) * C;
            load_group<CPG>(q, a);
            load_group<CPG>(q + C, bq);
            load_group<CPG>(q + (size_t)Wf * C, c);
            load_group<CPG>(q + (size_t)(Wf + 1) * C, d);
#pragma unroll
            for (int i = 0; i < CPG; ++i) ref[i] = 0.5f * (0.5f * a[i] + 0.5f * bq[i]) + 0.5f * (0.5f * c[i] + 0.5f * d[i]);
        } else {
            int h0, h1, w0, w1;
            float lh, lw;
            up_index(y, 0.5f, Hf, h0, h1, lh);
            up_index(xc, 0.5f, Wf, w0, w1, lw);
            float a[CPG], bq[CPG], c[CPG], d[CPG];
            load_group<CPG>(ref_view + ((size_t)h0 * Wf + w0) * C, a);
            load_group<CPG>(ref_view + ((size_t)h0 * Wf + w1) * C, bq);
            load_group<CPG>(ref_view + ((size_t)h1 * Wf + w0) * C, c);
            load_group<CPG>(ref_view + ((size_t)h1 * Wf + w1) * C, d);
#pragma unroll
            for (int i = 0; i < CPG; ++i)
                ref[i] = (1.f - lh) * ((1.f - lw) * a[i] + lw * bq[i]) + lh * ((1.f - lw) * c[i] + lw * d[i]);
        }
        float num = 0.f, wsum = 1e-5f;                          // itermvs.py:88-89
        constexpr int VG = CPG == 2 ? 4 : 2;                    // views per load batch (bounded by registers)
        for (int v0 = 0; v0 < S; v0 += 8) {
            Tap tp;
            tp.x0 = tp.y0 = 0; tp.fx = tp.fy = 0.f; tp.mask = 0u;
            float wv = 0.f;
            if (v0 + g < S) {
                tp = project_tap(sP + (v0 + g) * 12, (float)xc * sx, (float)y * sy, depth, (float)W2, (float)H2, Wf, Hf);
                wv = ldg(prm.vw2 + ((size_t)b * S + v0 + g) * P2 + p);
            }
            __syncwarp();
            publish_tapset(&sTap[slot][g], make_tapset(tp, Wf, Hf, C), wv);
            __syncwarp();
            const int nv = min(8, S - v0);
            for (int j0 = 0; j0 < nv; j0 += VG) {
                TapSet ts[VG];
                TapLoads<CPG> L[VG];
                float wj[VG];
#pragma unroll
                for (int u = 0; u < VG; ++u) {
                    ts[u] = read_tapset(&sTap[slot][min(j0 + u, 7)], wj[u]);          // weight 0 for views >= S
                    const int v = min(v0 + j0 + u, S - 1);
                    issue_taps<CPG>(L[u], src_base + (size_t)v * view_stride, ts[u]);
                }
#pragma unroll
                for (int u = 0; u < VG; ++u) {
                    const float c = finish_taps<CPG>(L[u], ts[u], ref);
                    if (v0 + j0 + u < S) {
                        num = fmaf(c, wj[u], num);                       // itermvs.py:114
                        wsum += wj[u];                                   // itermvs.py:115
                    }
                }
            }
        }
        if (pvalid) prm.agg[(((size_t)b * IMVS_ITER_SLICES + slice_base + r) * P2 + p) * 8 + g] = num / wsum;
    }
}

__global__ void __launch_bounds__(256, 4) warpcorr_iter_v3_kernel(const IterParams prm) {
    __shared__ float sP[IMVS_MAX_VIEWS * 12];
    __shared__ TapRecord sTapAll[8][4][8];       // [warp][slot][view of the current chunk]
    const int b = blockIdx.z / 3, lvl = blockIdx.z % 3;
    const int S = prm.V - 1;
    for (int i = threadIdx.x; i < S * 12; i += blockDim.x) sP[i] = prm.rt[lvl][(size_t)b * S * 12 + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int y = blockIdx.y * 8 + warp;
    if (y >= prm.H2) return;
    const int x_begin = blockIdx.x * ITER_TPX, x_end = min(x_begin + ITER_TPX, prm.W2);
    if (x_begin >= x_end) return;
    // itermvs.py:231-235
    TapRecord (*sTap)[8] = sTapAll[warp];
    if (lvl == 0)      iter_level<2, 4, 1>(prm, sP, sTap, b, y, x_begin, x_end, 0, -2.f, -2.0f / 3, 2.0f / 3, 2.f);
    else if (lvl == 1) iter_level<4, 4, 0>(prm, sP, sTap, b, y, x_begin, x_end, 4, -8.f, -8.0f / 3, 8.0f / 3, 8.f);
    else               iter_level<6, 2, 2>(prm, sP, sTap, b, y, x_begin, x_end, 8, -32.f, 32.f, 0.f, 0.f);
}

}  // namespace imvs
