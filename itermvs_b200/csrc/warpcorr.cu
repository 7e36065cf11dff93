// Fused plane-sweep kernels: hypothesis generation + homography warp + bilinear sampling of the
// source feature pyramids + group-wise correlation (+ pixel-wise view-weighted aggregation in the
// iteration kernel).  The [C, D, H, W] warped volume of the reference (module.py:118-120) and the
// same-size product tensor (itermvs.py:50, 103) are never materialised.
//
// Reference: models/module.py:68-125, models/itermvs.py:11-19, 45-69, 86-120, 289-293.
//
// Structure.  These kernels are bound by the L1 data pipe (one 128-byte wavefront per cycle per SM)
// and by instruction issue, not by HBM: every sample is a 4-tap gather of C contiguous floats whose
// lines are reused ~8x out of L1.  The design attacks both:
//   * two phases per warp.  Phase A: the 32 lanes compute 32 DIFFERENT (pixel, hypothesis, view)
//     sampling positions (projection, 4 IEEE divisions, clamping, bilinear weights) and leave a 24-byte
//     record per sample in shared memory.  Phase B: groups of 4/8 lanes walk the records and do
//     nothing but loads and FMAs.
//   * the taps of a record sit at FIXED offsets from one base (top-left tap clamped to
//     [0, W-2] x [0, H-2]; the bilinear weights are permuted / zeroed to match, which also folds in
//     grid_sample's zero padding), so a sample costs one address computation.
//   * every load instruction reads 64 or 128 CONTIGUOUS bytes per sample.  At level 3 (48 channels,
//     6 per correlation group) a lane therefore holds channel pairs {g, 8+g, 16+g} instead of "its"
//     group; the per-pair partial sums are accumulated over the views (the aggregation is linear) and
//     regrouped once per (pixel, hypothesis) with three shuffles.  (Loading a lane's own 6 channels
//     means 8-byte pieces at a 24-byte stride: 2 lines per instruction per sample, 3x the wavefronts.)
//   * iteration kernel: one persistent 768-thread block per SM owns a compact pixel region (L1 hit
//     rate 79 %), its warps draw (4-pixel row, level) items from a shared-memory counter, and the
//     gather loops are software-pipelined and fully unrolled per source-view count.
// Measured history at 640x512 / 4 views: profiles/README.md.
#include "common.cuh"
#include "sampling.cuh"

#include <algorithm>

namespace imvs {

int tune(const char* name, int def);   // IMVS_TUNE_<NAME> experiment switch, defined in warp.cu

constexpr int WC_WARPS = 4;     // warps per block = rows of the pixel tile
constexpr int WC_NPX = 4;       // consecutive pixels of a row handled by one warp
constexpr int WC_ITER_WARPS = 24;   // iteration kernel: one 768-thread block per SM (<= 85 registers)

// Phase-A result for one (pixel, hypothesis, view).  Taps are read at element offsets
// off, off + C, off + pitch, off + pitch + C (pitch = Wf * C) with weights w.x .. w.w.
__device__ __forceinline__ void make_record(const Tap& tp, int Wf, int Hf, int C, int view, float4& w, int& off) {
    const int xb = min(max(tp.x0, 0), Wf - 2), yb = min(max(tp.y0, 0), Hf - 2);
    const float gx = 1.f - tp.fx, gy = 1.f - tp.fy;
    // weight of the loaded column xb / xb+1: column x0 carries gx, column x0+1 carries fx, columns outside
    // the map carry nothing (grid_sample padding_mode='zeros').  Same for the rows.
    const float wxa = tp.x0 == xb ? gx : (tp.x0 + 1 == xb ? tp.fx : 0.f);
    const float wxb = tp.x0 == xb ? tp.fx : (tp.x0 == xb + 1 ? gx : 0.f);
    const float wya = tp.y0 == yb ? gy : (tp.y0 + 1 == yb ? tp.fy : 0.f);
    const float wyb = tp.y0 == yb ? tp.fy : (tp.y0 == yb + 1 ? gy : 0.f);
    w = make_float4(wxa * wya, wxb * wya, wxa * wyb, wxb * wyb);
    off = ((view * Hf + yb) * Wf + xb) * C;
}

__device__ __forceinline__ float bilerp(float t00, float t01, float t10, float t11, const float4& w) {
    float a = t00 * w.x;
    a = fmaf(t01, w.y, a);
    a = fmaf(t10, w.z, a);
    return fmaf(t11, w.w, a);
}

// Level-3 regrouping: lane g of an 8-lane slot holds the partial sums of channel pairs g, 8+g, 16+g
// (acc[k] <-> pair 8k+g); correlation group G is pairs 3G, 3G+1, 3G+2.  Three shuffles, each lane
// supplying the register its reader needs.
__device__ __forceinline__ float regroup48(const float (&acc)[3], int lane) {
    const int g = lane & 7, slot_base = lane & ~7;
    float sum = 0.f;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const int Gd = (3 * (g - m + 8)) & 7;          // the group whose pair 3*Gd+m lives in this lane
        const int ks = (3 * Gd + m) >> 3;
        const float mine = ks == 0 ? acc[0] : (ks == 1 ? acc[1] : acc[2]);
        sum += __shfl_sync(0xffffffffu, mine, slot_base | ((3 * g + m) & 7));
    }
    return sum;
}

// ---------------------------------------------------------------------------------------------
// K3: iteration kernel -- three pyramid levels, R = (4,4,2) samples per pixel around the current
// normalized depth, all source views, view-weighted aggregation.
//
// Work item = (4 consecutive pixels of a row, one pyramid level).  ONE persistent block per SM owns a
// contiguous, spatially compact range of 4x4-pixel tiles (all three levels), so that the halo of every tap
// stays in that SM's L1; its warps draw items from a shared-memory counter, level 3 first (heaviest), so the
// tail is one light level-1 item.  Inside an item the gather loop is software-pipelined by hand: the taps of
// step i+1 are in flight while step i is multiplied out (the v4 kernel spent issue time and L1 time
// strictly one after the other: 43 % + 49 % busy).
//   grid (#SMs), block 768 or 1024
// ---------------------------------------------------------------------------------------------
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

struct IterParams5 {
    IterParams p;
    unsigned int n_tiles;       // 4x4-pixel tiles of the level-2 maps of all batch items
    int tiles_x, tiles_y;
    unsigned long long strip_magic;     // fastdiv constant for 2 * tiles_x
};

struct IterSmem {
    float4* recW;      // [16*S] bilinear weights of the warp's samples of the current item
    int2* recO;        // [16*S] {element offset inside this batch item's pyramid, view weight bits}
    float* sP;         // [S][12] composed projections of this item's batch element / level
    float* nd;         // [4] normalized depth of the item's pixels
    float* vw;         // [S][4] view weights of the item's pixels
    float* park;       // [32][6] level 3: reference feature of the second pixel pair
};

// LVL 0: level 1 (C=16, feature map 2x the depth map), 1: level 2 (C=32), 2: level 3 (C=48, half size)
template <int LVL, int ST, bool PAD3 = false>
__device__ __forceinline__ void iter_build(const IterParams& prm, const IterSmem& sm, int b, int y, int x0,
                                           float inv_min, float inv_max, float o0, float o1, float o2, float o3) {
    constexpr int C = LVL == 0 ? 16 : (LVL == 1 ? 32 : (PAD3 ? 64 : 48));
    constexpr int R = LVL == 2 ? 2 : 4;
    constexpr int NPR = WC_NPX * R;
    constexpr float SC = LVL == 0 ? 2.f : (LVL == 1 ? 1.f : 0.5f);       // module.py:95-96 (Wf / W2, exact)
    const int lane = threadIdx.x & 31;
    const int S = ST ? ST : prm.V - 1, H2 = prm.H2, W2 = prm.W2;
    const int Hf = LVL == 0 ? H2 * 2 : (LVL == 2 ? H2 / 2 : H2);
    const int Wf = LVL == 0 ? W2 * 2 : (LVL == 2 ? W2 / 2 : W2);
    const float* smp = prm.samples[LVL];
    const int total = NPR * S;
#pragma unroll (ST ? 4 : 1)
    for (int t = lane; t < total; t += 32) {
        const int v = t / NPR, pr = t % NPR, px = pr / R, r = pr % R;
        const int x = min(x0 + px, W2 - 1);
        float depth;
        if (smp) {
            depth = ldg(smp + ((size_t)b * R + r) * ((size_t)H2 * W2) + (size_t)y * W2 + x);
        } else {                                                         // itermvs.py:229, 290-293
            const float off = (r == 0 ? o0 : r == 1 ? o1 : r == 2 ? o2 : o3) * (1.0f / 256.0f);
            const float s = fminf(fmaxf(sm.nd[px] + off, 0.f), 1.f);
            depth = unnormalize_depth(s, inv_min, inv_max);
        }
        const Tap tp = project_tap(sm.sP + v * 12, (float)x * SC, (float)y * SC, depth, (float)W2, (float)H2, Wf, Hf);
        float4 w;
        int off;
        make_record(tp, Wf, Hf, C, v + 1, w, off);
        sm.recW[t] = w;
        sm.recO[t] = make_int2(off, __float_as_int(sm.vw[v * WC_NPX + px]));
    }
}

struct Buf4 { float4 a, b, c, d, w; float wv; };      // the four taps of one sample (one float4 per lane) + its record

template <int C>
__device__ __forceinline__ void fetch4(Buf4& B, const IterSmem& sm, const float* __restrict__ base, int pitch, int t) {
    B.w = sm.recW[t];
    const int2 o = sm.recO[t];
    B.wv = __int_as_float(o.y);
    const float* p = base + o.x;
    B.a = ldg4(p); B.b = ldg4(p + C); B.c = ldg4(p + pitch); B.d = ldg4(p + pitch + C);
}

// level 1: 4 lanes per sample (float4 = correlation groups 2j, 2j+1), 8 samples per instruction =
// 4 hypotheses x 2 pixels.  Reference feature: F.interpolate(x0.5) == 2x2 mean (itermvs.py:95-96).
template <int ST>
__device__ __forceinline__ void iter_gather_l1(const IterParams& prm, const IterSmem& sm, int b, int y, int x0) {
    constexpr int NPR = WC_NPX * 4;
    const int lane = threadIdx.x & 31, j = lane & 3, q = lane >> 2, r = q & 3, pxpar = q >> 2;
    const int V = prm.V, S = ST ? ST : V - 1, H2 = prm.H2, W2 = prm.W2, Wf = 2 * W2, Hf = 2 * H2;
    const float* base = prm.fea[0] + (size_t)b * V * Hf * Wf * 16 + 4 * j;
    const int pitch = Wf * 16;
    float4 refs[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int xc = min(x0 + 2 * h + pxpar, W2 - 1);
        const float* qr = base + ((size_t)(2 * y) * Wf + 2 * xc) * 16;
        const float4 ra = ldg4(qr), rb = ldg4(qr + 16), rc = ldg4(qr + pitch), rd = ldg4(qr + pitch + 16);
        refs[h].x = 0.5f * (0.5f * ra.x + 0.5f * rb.x) + 0.5f * (0.5f * rc.x + 0.5f * rd.x);
        refs[h].y = 0.5f * (0.5f * ra.y + 0.5f * rb.y) + 0.5f * (0.5f * rc.y + 0.5f * rd.y);
        refs[h].z = 0.5f * (0.5f * ra.z + 0.5f * rb.z) + 0.5f * (0.5f * rc.z + 0.5f * rd.z);
        refs[h].w = 0.5f * (0.5f * ra.w + 0.5f * rb.w) + 0.5f * (0.5f * rc.w + 0.5f * rd.w);
    }
    float4 ref = refs[0];
    float n0 = 0.f, n1 = 0.f, wsum = 1e-5f;                               // itermvs.py:88-89
    auto tix = [&](int h, int v) { return v * NPR + (2 * h + pxpar) * 4 + r; };
    auto consume = [&](const Buf4& B, int h, int v) {
        if (v == 0) { ref = h == 0 ? refs[0] : refs[1]; n0 = 0.f; n1 = 0.f; wsum = 1e-5f; }
        const float lo = fmaf(bilerp(B.a.y, B.b.y, B.c.y, B.d.y, B.w), ref.y, bilerp(B.a.x, B.b.x, B.c.x, B.d.x, B.w) * ref.x) * 0.5f;
        const float hi = fmaf(bilerp(B.a.w, B.b.w, B.c.w, B.d.w, B.w), ref.w, bilerp(B.a.z, B.b.z, B.c.z, B.d.z, B.w) * ref.z) * 0.5f;
        n0 = fmaf(lo, B.wv, n0);                                          // itermvs.py:114
        n1 = fmaf(hi, B.wv, n1);
        wsum += B.wv;                                                     // itermvs.py:115
        if (v == S - 1) {
            const int x = x0 + 2 * h + pxpar;
            if (x < W2) {
                float* out = prm.agg + (((size_t)b * IMVS_ITER_SLICES + r) * ((size_t)H2 * W2) + (size_t)y * W2 + x) * 8 + 2 * j;
                *reinterpret_cast<float2*>(out) = make_float2(n0 / wsum, n1 / wsum);
            }
        }
    };
    Buf4 A, Bb;
    int h = 0, v = 0;
    fetch4<16>(A, sm, base, pitch, tix(0, 0));
    const int n = 2 * S;
#pragma unroll (ST ? 64 : 1)
    for (int i = 0; i < n; i += 2) {
        int vb = v + 1, hb = h;
        if (vb == S) { vb = 0; hb = h + 1; }
        fetch4<16>(Bb, sm, base, pitch, tix(hb, vb));
        consume(A, h, v);
        v = vb + 1; h = hb;
        if (v == S) { v = 0; h = hb + 1; }
        if (i + 2 < n) fetch4<16>(A, sm, base, pitch, tix(h, v));
        consume(Bb, hb, vb);
    }
}

// level 2: 8 lanes per sample (float4 = one correlation group), 4 samples per instruction = the 4 hypotheses
// of one pixel; one tap is exactly one 128-byte line.
template <int ST>
__device__ __forceinline__ void iter_gather_l2(const IterParams& prm, const IterSmem& sm, int b, int y, int x0) {
    constexpr int NPR = WC_NPX * 4;
    const int lane = threadIdx.x & 31, g = lane & 7, r = lane >> 3;
    const int V = prm.V, S = ST ? ST : V - 1, H2 = prm.H2, W2 = prm.W2;
    const float* base = prm.fea[1] + (size_t)b * V * H2 * W2 * 32 + 4 * g;
    const int pitch = W2 * 32;
    const float* refrow = base + (size_t)y * W2 * 32;
    float4 ref, refn = ldg4(refrow + (size_t)min(x0, W2 - 1) * 32);
    float num = 0.f, wsum = 1e-5f;
    auto tix = [&](int px, int v) { return v * NPR + px * 4 + r; };
    auto consume = [&](const Buf4& B, int px, int v) {
        if (v == 0) {
            ref = refn; num = 0.f; wsum = 1e-5f;
            if (px + 1 < WC_NPX) refn = ldg4(refrow + (size_t)min(x0 + px + 1, W2 - 1) * 32);
        }
        float dot = bilerp(B.a.x, B.b.x, B.c.x, B.d.x, B.w) * ref.x;
        dot = fmaf(bilerp(B.a.y, B.b.y, B.c.y, B.d.y, B.w), ref.y, dot);
        dot = fmaf(bilerp(B.a.z, B.b.z, B.c.z, B.d.z, B.w), ref.z, dot);
        dot = fmaf(bilerp(B.a.w, B.b.w, B.c.w, B.d.w, B.w), ref.w, dot);
        num = fmaf(dot * 0.25f, B.wv, num);
        wsum += B.wv;
        if (v == S - 1) {
            const int x = x0 + px;
            if (x < W2) prm.agg[(((size_t)b * IMVS_ITER_SLICES + 4 + r) * ((size_t)H2 * W2) + (size_t)y * W2 + x) * 8 + g] = num / wsum;
        }
    };
    Buf4 A, Bb;
    int px = 0, v = 0;
    fetch4<32>(A, sm, base, pitch, tix(0, 0));
    const int n = WC_NPX * S;
#pragma unroll (ST ? 64 : 1)
    for (int i = 0; i < n; i += 2) {
        int vb = v + 1, pxb = px;
        if (vb == S) { vb = 0; pxb = px + 1; }
        fetch4<32>(Bb, sm, base, pitch, tix(pxb, vb));
        consume(A, px, v);
        v = vb + 1; px = pxb;
        if (v == S) { v = 0; px = pxb + 1; }
        if (i + 2 < n) fetch4<32>(A, sm, base, pitch, tix(px, v));
        consume(Bb, pxb, vb);
    }
}

// one row (two horizontally adjacent taps) of a level-3 sample: 6 float2 loads, each instruction reading 64
// contiguous bytes per sample (lane g: channel pairs g, 8+g, 16+g)
struct Row48 { float2 l[3], r[3]; float wl, wr, wv; };
__device__ __forceinline__ void fetch_row48(Row48& R, const IterSmem& sm, const float* __restrict__ base, int pitch, int t, int row) {
    const float4 w = sm.recW[t];
    const int2 o = sm.recO[t];
    R.wl = row ? w.z : w.x; R.wr = row ? w.w : w.y; R.wv = __int_as_float(o.y);
    const float* p = base + o.x + (row ? pitch : 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) { R.l[k] = ldg2(p + 16 * k); R.r[k] = ldg2(p + 48 + 16 * k); }
}

// level 3: 8 lanes per sample, 4 samples per instruction = 2 hypotheses x 2 pixels.  Reference feature:
// F.interpolate(x2, bilinear, align_corners=False) of the level-3 map (itermvs.py:97-98).
template <int ST>
__device__ __forceinline__ void iter_gather_l3(const IterParams& prm, const IterSmem& sm, int b, int y, int x0) {
    constexpr int NPR = WC_NPX * 2;
    const int lane = threadIdx.x & 31, g = lane & 7, slot = lane >> 3, r = slot & 1, pxpar = slot >> 1;
    const int V = prm.V, S = ST ? ST : V - 1, H2 = prm.H2, W2 = prm.W2, Wf = W2 / 2, Hf = H2 / 2;
    const float* base = prm.fea[2] + (size_t)b * V * Hf * Wf * 48 + 2 * g;
    const int pitch = Wf * 48;
    int h0, h1;
    float lh;
    up_index(y, 0.5f, Hf, h0, h1, lh);
    float2 ref[3];
#pragma unroll
    for (int h = 1; h >= 0; --h) {          // second pixel pair first: parked in shared memory
        const int xc = min(x0 + 2 * h + pxpar, W2 - 1);
        int w0i, w1i;
        float lw;
        up_index(xc, 0.5f, Wf, w0i, w1i, lw);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float2 a = ldg2(base + ((size_t)h0 * Wf + w0i) * 48 + 16 * k), bq = ldg2(base + ((size_t)h0 * Wf + w1i) * 48 + 16 * k);
            const float2 c = ldg2(base + ((size_t)h1 * Wf + w0i) * 48 + 16 * k), d = ldg2(base + ((size_t)h1 * Wf + w1i) * 48 + 16 * k);
            ref[k].x = (1.f - lh) * ((1.f - lw) * a.x + lw * bq.x) + lh * ((1.f - lw) * c.x + lw * d.x);
            ref[k].y = (1.f - lh) * ((1.f - lw) * a.y + lw * bq.y) + lh * ((1.f - lw) * c.y + lw * d.y);
        }
        if (h == 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) *reinterpret_cast<float2*>(sm.park + (k * 32 + lane) * 2) = ref[k];
        }
    }
    float acc[3] = {0.f, 0.f, 0.f};
    float ax[3], ay[3];
    float wsum = 1e-5f;
    auto tix = [&](int h, int v) { return v * NPR + (2 * h + pxpar) * 2 + r; };
    Row48 A, Bb;
    int h = 0, v = 0;
    fetch_row48(A, sm, base, pitch, tix(0, 0), 0);
    const int n = 2 * S;
#pragma unroll (ST ? 64 : 1)
    for (int i = 0; i < n; ++i) {
        fetch_row48(Bb, sm, base, pitch, tix(h, v), 1);
        if (v == 0) {
            acc[0] = acc[1] = acc[2] = 0.f; wsum = 1e-5f;
            if (h == 1) {
#pragma unroll
                for (int k = 0; k < 3; ++k) ref[k] = *reinterpret_cast<const float2*>(sm.park + (k * 32 + lane) * 2);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ax[k] = fmaf(A.r[k].x, A.wr, A.l[k].x * A.wl);
            ay[k] = fmaf(A.r[k].y, A.wr, A.l[k].y * A.wl);
        }
        const int hc = h, vc = v;
        if (++v == S) { v = 0; ++h; }
        if (i + 1 < n) fetch_row48(A, sm, base, pitch, tix(h, v), 0);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ax[k] = fmaf(Bb.r[k].x, Bb.wr, fmaf(Bb.l[k].x, Bb.wl, ax[k]));
            ay[k] = fmaf(Bb.r[k].y, Bb.wr, fmaf(Bb.l[k].y, Bb.wl, ay[k]));
            acc[k] = fmaf(fmaf(ay[k], ref[k].y, ax[k] * ref[k].x), Bb.wv, acc[k]);
        }
        wsum += Bb.wv;
        if (vc == S - 1) {
            const float num = regroup48(acc, lane) * (1.0f / 6.0f);
            const int x = x0 + 2 * hc + pxpar;
            if (x < W2)
                prm.agg[(((size_t)b * IMVS_ITER_SLICES + 8 + r) * ((size_t)H2 * W2) + (size_t)y * W2 + x) * 8 + g] = num / wsum;
        }
    }
}

// ---- level 3 on the PADDED pyramid (imvs_pad_level3: 64 floats = 256 bytes per texel; correlation group g's six channels at
// floats 4g..4g+3 and 32+4g, 32+4g+1, the rest zero).  Lane g of a sample's 8 lanes then loads ITS OWN group: one float4 + one
// float2 per tap, every instruction one full aligned 128-byte line per sample (the 48-channel layout needs three 64-byte pieces
// per tap -- half-used wavefronts -- and three shuffles per (pixel, hypothesis) to regroup).
struct Row64 { float4 la, ra; float2 lb, rb; float wl, wr, wv; };
__device__ __forceinline__ void fetch_row64(Row64& R, const IterSmem& sm, const float* __restrict__ base, int pitch, int t, int row) {
    const float4 w = sm.recW[t];
    const int2 o = sm.recO[t];
    R.wl = row ? w.z : w.x; R.wr = row ? w.w : w.y; R.wv = __int_as_float(o.y);
    const float* p = base + o.x + (row ? pitch : 0);
    R.la = ldg4(p); R.lb = ldg2(p + 32); R.ra = ldg4(p + 64); R.rb = ldg2(p + 96);
}

template <int ST>
__device__ __forceinline__ void iter_gather_l3p(const IterParams& prm, const IterSmem& sm, int b, int y, int x0) {
    constexpr int NPR = WC_NPX * 2;
    const int lane = threadIdx.x & 31, g = lane & 7, slot = lane >> 3, r = slot & 1, pxpar = slot >> 1;
    const int V = prm.V, S = ST ? ST : V - 1, H2 = prm.H2, W2 = prm.W2, Wf = W2 / 2, Hf = H2 / 2;
    const float* base = prm.fea[2] + (size_t)b * V * Hf * Wf * 64 + 4 * g;
    const int pitch = Wf * 64;
    int h0, h1;
    float lh;
    up_index(y, 0.5f, Hf, h0, h1, lh);
    float ref[6];
#pragma unroll
    for (int h = 1; h >= 0; --h) {          // second pixel pair first: parked in shared memory
        const int xc = min(x0 + 2 * h + pxpar, W2 - 1);
        int w0i, w1i;
        float lw;
        up_index(xc, 0.5f, Wf, w0i, w1i, lw);
        const float* pa = base + ((size_t)h0 * Wf + w0i) * 64, *pb = base + ((size_t)h0 * Wf + w1i) * 64;
        const float* pc = base + ((size_t)h1 * Wf + w0i) * 64, *pd = base + ((size_t)h1 * Wf + w1i) * 64;
        const float4 a4 = ldg4(pa), b4 = ldg4(pb), c4 = ldg4(pc), d4 = ldg4(pd);
        const float2 a2 = ldg2(pa + 32), b2 = ldg2(pb + 32), c2 = ldg2(pc + 32), d2 = ldg2(pd + 32);
        ref[0] = (1.f - lh) * ((1.f - lw) * a4.x + lw * b4.x) + lh * ((1.f - lw) * c4.x + lw * d4.x);
        ref[1] = (1.f - lh) * ((1.f - lw) * a4.y + lw * b4.y) + lh * ((1.f - lw) * c4.y + lw * d4.y);
        ref[2] = (1.f - lh) * ((1.f - lw) * a4.z + lw * b4.z) + lh * ((1.f - lw) * c4.z + lw * d4.z);
        ref[3] = (1.f - lh) * ((1.f - lw) * a4.w + lw * b4.w) + lh * ((1.f - lw) * c4.w + lw * d4.w);
        ref[4] = (1.f - lh) * ((1.f - lw) * a2.x + lw * b2.x) + lh * ((1.f - lw) * c2.x + lw * d2.x);
        ref[5] = (1.f - lh) * ((1.f - lw) * a2.y + lw * b2.y) + lh * ((1.f - lw) * c2.y + lw * d2.y);
        if (h == 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) *reinterpret_cast<float2*>(sm.park + (k * 32 + lane) * 2) = make_float2(ref[2 * k], ref[2 * k + 1]);
        }
    }
    float acc = 0.f, wsum = 1e-5f;
    float a[6];
    auto tix = [&](int h, int v) { return v * NPR + (2 * h + pxpar) * 2 + r; };
    Row64 A, Bb;
    int h = 0, v = 0;
    fetch_row64(A, sm, base, pitch, tix(0, 0), 0);
    const int n = 2 * S;
#pragma unroll (ST ? 64 : 1)
    for (int i = 0; i < n; ++i) {
        fetch_row64(Bb, sm, base, pitch, tix(h, v), 1);
        if (v == 0) {
            acc = 0.f; wsum = 1e-5f;
            if (h == 1) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float2 q = *reinterpret_cast<const float2*>(sm.park + (k * 32 + lane) * 2);
                    ref[2 * k] = q.x; ref[2 * k + 1] = q.y;
                }
            }
        }
        a[0] = fmaf(A.ra.x, A.wr, A.la.x * A.wl); a[1] = fmaf(A.ra.y, A.wr, A.la.y * A.wl);
        a[2] = fmaf(A.ra.z, A.wr, A.la.z * A.wl); a[3] = fmaf(A.ra.w, A.wr, A.la.w * A.wl);
        a[4] = fmaf(A.rb.x, A.wr, A.lb.x * A.wl); a[5] = fmaf(A.rb.y, A.wr, A.lb.y * A.wl);
        const int hc = h, vc = v;
        if (++v == S) { v = 0; ++h; }
        if (i + 1 < n) fetch_row64(A, sm, base, pitch, tix(h, v), 0);
        a[0] = fmaf(Bb.ra.x, Bb.wr, fmaf(Bb.la.x, Bb.wl, a[0])); a[1] = fmaf(Bb.ra.y, Bb.wr, fmaf(Bb.la.y, Bb.wl, a[1]));
        a[2] = fmaf(Bb.ra.z, Bb.wr, fmaf(Bb.la.z, Bb.wl, a[2])); a[3] = fmaf(Bb.ra.w, Bb.wr, fmaf(Bb.la.w, Bb.wl, a[3]));
        a[4] = fmaf(Bb.rb.x, Bb.wr, fmaf(Bb.lb.x, Bb.wl, a[4])); a[5] = fmaf(Bb.rb.y, Bb.wr, fmaf(Bb.lb.y, Bb.wl, a[5]));
        float dot = a[0] * ref[0];
#pragma unroll
        for (int c = 1; c < 6; ++c) dot = fmaf(a[c], ref[c], dot);
        acc = fmaf(dot, Bb.wv, acc);
        wsum += Bb.wv;
        if (vc == S - 1) {
            const float num = acc * (1.0f / 6.0f);
            const int x = x0 + 2 * hc + pxpar;
            if (x < W2)
                prm.agg[(((size_t)b * IMVS_ITER_SLICES + 8 + r) * ((size_t)H2 * W2) + (size_t)y * W2 + x) * 8 + g] = num / wsum;
        }
    }
}

// Item header: everything a warp needs before phase A of an item, fetched into registers one item ahead
// (the loads fly while the current item is processed) and parked in shared memory when the item starts.
template <int ST>
struct ItemHeader {
    static constexpr int NRT = ST ? (12 * ST + 31) / 32 : (12 * IMVS_MAX_VIEWS + 31) / 32;      // rot|trans floats per lane
    static constexpr int NVW = ST ? (WC_NPX * ST + 31) / 32 : (WC_NPX * IMVS_MAX_VIEWS + 31) / 32;
    int b, y, x0, lvl;          // lvl 0..2 = pyramid level 1..3, -1 = no item
    float rt[NRT], vw[NVW], misc;       // misc: lanes 0..3 normalized depth, lane 4 / 5 depth_min / depth_max
};

// n / d for n < 2^28, d <= 4096, m = ceil(2^40 / d)
__device__ __forceinline__ unsigned fastdiv(unsigned n, unsigned long long m) { return (unsigned)(((unsigned long long)n * m) >> 40); }

template <int ST>
__device__ __forceinline__ void fetch_header(ItemHeader<ST>& h, const IterParams5& q, unsigned item, unsigned n_items,
                                             unsigned r0, unsigned per_level, int S, int lane) {
    const IterParams& prm = q.p;
    if (item >= n_items) { h.lvl = -1; return; }
    // items of a block: level 3 rows first (heaviest), then level 2, then level 1
    const unsigned li = (item >= per_level) + (item >= 2 * per_level);
    const unsigned row_item = r0 + (item - li * per_level);
    const unsigned tile = row_item >> 2;
    // tile order: batch item, then strips of two tile rows, column by column inside a strip -- consecutive
    // tiles (what one SM works on) form a compact region, 8 rows high
    const unsigned per_b = (unsigned)(q.tiles_x * q.tiles_y);
    const unsigned b = q.p.B == 1 ? 0u : tile / per_b;
    const unsigned k = tile - b * per_b;
    const unsigned strip = fastdiv(k, q.strip_magic), within = k - strip * 2 * q.tiles_x;
    const bool two = (int)(2 * strip + 1) < q.tiles_y;
    const int tx = two ? within >> 1 : within, ty = 2 * strip + (two ? (within & 1) : 0);
    h.b = b; h.lvl = 2 - (int)li;
    h.y = ty * 4 + (row_item & 3); h.x0 = tx * WC_NPX;
    const int H2 = prm.H2, W2 = prm.W2;
    if (h.y >= H2) { h.lvl = -1; return; }          // ragged last tile row: nothing to do
    const float* rt = prm.rt[h.lvl] + (size_t)b * S * 12;
#pragma unroll
    for (int i = 0; i < ItemHeader<ST>::NRT; ++i) h.rt[i] = (lane + 32 * i < 12 * S) ? ldg(rt + lane + 32 * i) : 0.f;
#pragma unroll
    for (int i = 0; i < ItemHeader<ST>::NVW; ++i) {
        const int e = lane + 32 * i, v = e / WC_NPX, px = e % WC_NPX;
        h.vw[i] = e < WC_NPX * S ? ldg(prm.vw2 + ((size_t)b * S + v) * ((size_t)H2 * W2) + (size_t)h.y * W2 + min(h.x0 + px, W2 - 1)) : 0.f;
    }
    h.misc = 0.f;
    if (prm.samples[0] == nullptr) {
        if (lane < WC_NPX) h.misc = ldg(prm.nd + (size_t)b * prm.nd_stride + ((size_t)h.y * W2 + min(h.x0 + lane, W2 - 1)) * prm.nd_pstride);
        else if (lane == 4) h.misc = ldg(prm.depth_min + b);
        else if (lane == 5) h.misc = ldg(prm.depth_max + b);
    }
}

// ST = number of source views when it is 1..8 (gather loops fully unrolled: the software pipeline becomes
// straight-line code with statically renamed buffers), 0 = any number (rolled loops).
// NW = warps of the (single) block an SM runs.
template <int ST, int NW, bool PAD3 = false>
__global__ void __launch_bounds__(NW * 32, NW <= 16 ? 2 : 1) warpcorr_iter_kernel(const IterParams5 q) {
    extern __shared__ float4 smem4[];
    __shared__ unsigned int next_item;
    pdl_trigger();
    const IterParams& prm = q.p;
    const int S = ST ? ST : prm.V - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per-warp carve: [16S] float4 | [16S] int2 | [32][6] float | [S][12] float | [S][4] float | [4] float
    const int per_warp4 = 16 * S + 8 * S + 48 + 3 * S + S + 1;      // in float4 units
    float4* mine = smem4 + warp * per_warp4;
    IterSmem sm;
    sm.recW = mine;
    sm.recO = reinterpret_cast<int2*>(mine + 16 * S);
    sm.park = reinterpret_cast<float*>(mine + 24 * S);
    sm.sP = sm.park + 192;
    sm.vw = sm.sP + 12 * S;
    sm.nd = sm.vw + 4 * S;
    // this block's contiguous range of 4-pixel rows (4 per tile) and its item queue
    const unsigned r0 = (unsigned)(((unsigned long long)q.n_tiles * 4 * blockIdx.x) / gridDim.x);
    const unsigned r1 = (unsigned)(((unsigned long long)q.n_tiles * 4 * (blockIdx.x + 1)) / gridDim.x);
    const unsigned per_level = r1 - r0, n_items = per_level * 3;
    if (threadIdx.x == 0) next_item = 2 * NW;
    __syncthreads();
    pdl_wait();                     // nd / view weights / pyramids come from the preceding kernels
    ItemHeader<ST> cur, nxt;
    fetch_header<ST>(cur, q, warp, n_items, r0, per_level, S, lane);
    unsigned nxt_item = NW + warp;
    while (cur.lvl >= 0 || nxt_item < n_items) {
        // the next item's header loads fly while this item is processed
        fetch_header<ST>(nxt, q, nxt_item, n_items, r0, per_level, S, lane);
        if (cur.lvl >= 0) {
#pragma unroll
            for (int i = 0; i < ItemHeader<ST>::NRT; ++i)
                if (lane + 32 * i < 12 * S) sm.sP[lane + 32 * i] = cur.rt[i];
#pragma unroll
            for (int i = 0; i < ItemHeader<ST>::NVW; ++i)
                if (lane + 32 * i < WC_NPX * S) sm.vw[lane + 32 * i] = cur.vw[i];
            if (lane < WC_NPX) sm.nd[lane] = cur.misc;
            const float dmin = __shfl_sync(0xffffffffu, cur.misc, 4), dmax = __shfl_sync(0xffffffffu, cur.misc, 5);
            const bool explicit_samples = prm.samples[0] != nullptr;
            const float inv_min = explicit_samples ? 0.f : 1.0f / dmin;
            const float inv_max = explicit_samples ? 0.f : 1.0f / dmax;
            __syncwarp();
            const int b = cur.b, y = cur.y, x0 = cur.x0;
            // itermvs.py:231-235
            if (cur.lvl == 2) {
                iter_build<2, ST, PAD3>(prm, sm, b, y, x0, inv_min, inv_max, -32.f, 32.f, 0.f, 0.f);
                __syncwarp();
                if constexpr (PAD3) iter_gather_l3p<ST>(prm, sm, b, y, x0);
                else iter_gather_l3<ST>(prm, sm, b, y, x0);
            } else if (cur.lvl == 1) {
                iter_build<1, ST>(prm, sm, b, y, x0, inv_min, inv_max, -8.f, -8.0f / 3, 8.0f / 3, 8.f);
                __syncwarp();
                iter_gather_l2<ST>(prm, sm, b, y, x0);
            } else {
                iter_build<0, ST>(prm, sm, b, y, x0, inv_min, inv_max, -2.f, -2.0f / 3, 2.0f / 3, 2.f);
                __syncwarp();
                iter_gather_l1<ST>(prm, sm, b, y, x0);
            }
            __syncwarp();
        }
        cur = nxt;
        unsigned drawn = 0;
        if (lane == 0 && nxt_item < n_items) drawn = atomicAdd(&next_item, 1u);
        nxt_item = nxt_item < n_items ? __shfl_sync(0xffffffffu, drawn, 0) : n_items;
    }
}

static size_t iter_smem_bytes(int S, int nw) { return (size_t)nw * (16 * S + 8 * S + 48 + 3 * S + S + 1) * sizeof(float4); }

static void (*iter_kernel_padded_for(int S))(const IterParams5) {
    switch (S) {
        case 1: return warpcorr_iter_kernel<1, WC_ITER_WARPS, true>;
        case 2: return warpcorr_iter_kernel<2, WC_ITER_WARPS, true>;
        case 3: return warpcorr_iter_kernel<3, WC_ITER_WARPS, true>;
        case 4: return warpcorr_iter_kernel<4, WC_ITER_WARPS, true>;
        case 5: return warpcorr_iter_kernel<5, WC_ITER_WARPS, true>;
        case 6: return warpcorr_iter_kernel<6, WC_ITER_WARPS, true>;
        case 7: return warpcorr_iter_kernel<7, WC_ITER_WARPS, true>;
        case 8: return warpcorr_iter_kernel<8, WC_ITER_WARPS, true>;
        default: return warpcorr_iter_kernel<0, WC_ITER_WARPS, true>;
    }
}

template <int NW>
static void (*iter_kernel_for(int S))(const IterParams5) {
    switch (S) {
        case 1: return warpcorr_iter_kernel<1, NW>;
        case 2: return warpcorr_iter_kernel<2, NW>;
        case 3: return warpcorr_iter_kernel<3, NW>;
        case 4: return warpcorr_iter_kernel<4, NW>;
        case 5: return warpcorr_iter_kernel<5, NW>;
        case 6: return warpcorr_iter_kernel<6, NW>;
        case 7: return warpcorr_iter_kernel<7, NW>;
        case 8: return warpcorr_iter_kernel<8, NW>;
        default: return warpcorr_iter_kernel<0, NW>;
    }
}

// 12 contiguous-per-instruction float2 loads of one level-3 sample (4 taps x 3 chunks of 64 bytes) and the
// partial correlation of this lane's three channel pairs with the reference feature
struct Taps48 { float2 t[4][3]; };
__device__ __forceinline__ void load48(Taps48& T, const float* __restrict__ p, int pitch) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        T.t[0][k] = ldg2(p + 16 * k);
        T.t[1][k] = ldg2(p + 48 + 16 * k);
        T.t[2][k] = ldg2(p + pitch + 16 * k);
        T.t[3][k] = ldg2(p + pitch + 48 + 16 * k);
    }
}
__device__ __forceinline__ float pair_dot(const Taps48& T, int k, const float4& w, const float2& ref) {
    const float ax = bilerp(T.t[0][k].x, T.t[1][k].x, T.t[2][k].x, T.t[3][k].x, w);
    const float ay = bilerp(T.t[0][k].y, T.t[1][k].y, T.t[2][k].y, T.t[3][k].y, w);
    return fmaf(ay, ref.y, ax * ref.x);
}

// one sample of the padded level-3 pyramid: this lane's own group, 4 taps x (float4 + float2)
struct Taps64 { float4 a[4]; float2 b[4]; };
__device__ __forceinline__ void load64(Taps64& T, const float* __restrict__ p, int pitch) {
    T.a[0] = ldg4(p); T.b[0] = ldg2(p + 32);
    T.a[1] = ldg4(p + 64); T.b[1] = ldg2(p + 96);
    T.a[2] = ldg4(p + pitch); T.b[2] = ldg2(p + pitch + 32);
    T.a[3] = ldg4(p + pitch + 64); T.b[3] = ldg2(p + pitch + 96);
}
__device__ __forceinline__ float group_dot(const Taps64& T, const float4& w, const float (&ref)[6]) {
    float dot = bilerp(T.a[0].x, T.a[1].x, T.a[2].x, T.a[3].x, w) * ref[0];
    dot = fmaf(bilerp(T.a[0].y, T.a[1].y, T.a[2].y, T.a[3].y, w), ref[1], dot);
    dot = fmaf(bilerp(T.a[0].z, T.a[1].z, T.a[2].z, T.a[3].z, w), ref[2], dot);
    dot = fmaf(bilerp(T.a[0].w, T.a[1].w, T.a[2].w, T.a[3].w, w), ref[3], dot);
    dot = fmaf(bilerp(T.b[0].x, T.b[1].x, T.b[2].x, T.b[3].x, w), ref[4], dot);
    return fmaf(bilerp(T.b[0].y, T.b[1].y, T.b[2].y, T.b[3].y, w), ref[5], dot);
}

// ---------------------------------------------------------------------------------------------
// K2: init plane sweep at level 3 (C = 48), per-view group correlation.  One warp = one pixel; its D hypotheses in
// chunks of 32: phase A, lane = hypothesis (projection, clamped 2x2 tap base, permuted bilinear weights with the
// zero padding folded in) -> a 20-byte record per hypothesis in shared memory; phase B, 8 lanes per hypothesis
// (lane g: channel pairs g, 8+g, 16+g -> every load instruction reads 64 contiguous bytes per hypothesis), 4
// consecutive hypotheses per instruction (they lie within a few texels of each other on the epipolar line and share
// cache lines), regrouped to the 8 correlation groups with three shuffles.
// Measured alternative (round 2, profiles/ps_experiments_r02.md): taking the dot products first, once per texel of the
// epipolar segment's bounding box, and interpolating 8 group scalars per tap ("band") executes MORE instructions than
// this gather at the benchmark's hypothesis spacing (30 M vs 25 M warp instructions, 52 us vs 46 us) and 1.4x more at
// the sparser spacing of 1920x1056 -- it was removed.
//   grid (ceil(W3/2), ceil(H3/2), B), block 128 (4 warps = 2 x 2 pixels)
// ---------------------------------------------------------------------------------------------
template <bool PAD3>
__global__ void __launch_bounds__(WC_WARPS * 32, 6)
warpcorr_init_kernel(const float* __restrict__ fea3, const float* __restrict__ rt3,
                     const float* __restrict__ depth_min, const float* __restrict__ depth_max,
                     const float* __restrict__ samples, float* __restrict__ corr, int B, int V, int H3, int W3, int D, int vper) {
    __shared__ float4 s_recW[WC_WARPS][32];
    __shared__ int s_recO[WC_WARPS][32];
    __shared__ float sP[IMVS_MAX_VIEWS * 12];
    const int S = V - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // blockIdx.z = (batch item, group of vper source views): finer blocks fill the last wave (1 280 four-warp blocks at 6 per SM
    // are 1.44 waves at 640x512)
    const int nsplit = (S + vper - 1) / vper;
    const int b = blockIdx.z / nsplit, v_lo = (blockIdx.z % nsplit) * vper, v_hi = min(S, v_lo + vper);
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < S * 12; i += blockDim.x) sP[i] = rt3[(size_t)b * S * 12 + i];
    __syncthreads();
    const int x = blockIdx.x * 2 + (warp & 1), y = blockIdx.y * 2 + (warp >> 1);
    if (x >= W3 || y >= H3) return;
    const int g = lane & 7, slot = lane >> 3;
    const int P3 = H3 * W3, p = y * W3 + x;
    const float inv_min = samples ? 0.f : 1.0f / depth_min[b], inv_max = samples ? 0.f : 1.0f / depth_max[b];
    constexpr int CS = PAD3 ? 64 : 48;          // floats per texel
    const float* fb = fea3 + (size_t)b * V * P3 * CS;
    const int pitch = W3 * CS;
    float4* recW = s_recW[warp];
    int* recO = s_recO[warp];
    float2 ref[3];
    float refp[6];
    if constexpr (PAD3) {
        const float4 r4 = ldg4(fb + (size_t)p * 64 + 4 * g);
        const float2 r2 = ldg2(fb + (size_t)p * 64 + 32 + 4 * g);
        refp[0] = r4.x; refp[1] = r4.y; refp[2] = r4.z; refp[3] = r4.w; refp[4] = r2.x; refp[5] = r2.y;
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) ref[k] = ldg2(fb + (size_t)p * 48 + 2 * g + 16 * k);
    }

    for (int d0 = 0; d0 < D; d0 += 32) {
        const int dc = min(d0 + lane, D - 1);
        // itermvs.py:13-17 (or the caller's explicit hypotheses, Evaluation.forward's depth_sample)
        const float depth = samples ? ldg(samples + ((size_t)b * D + dc) * P3 + p)
                                    : 1.0f / (inv_max + ((float)dc / (float)(D - 1)) * (inv_min - inv_max));
        const int nd = min(32, D - d0);
        for (int v = v_lo; v < v_hi; ++v) {
            const Tap tp = project_tap(sP + v * 12, (float)x, (float)y, depth, (float)W3, (float)H3, W3, H3);
            float4 w;
            int off;
            make_record(tp, W3, H3, CS, v + 1, w, off);
            __syncwarp();                               // the previous view's records have been consumed
            recW[lane] = w;
            recO[lane] = off;
            __syncwarp();
            float* out = corr + ((((size_t)b * S + v) * D + d0) * P3 + p) * 8 + g;
            if constexpr (PAD3) {
                const float* src = fb + 4 * g;
#pragma unroll 2
                for (int dd = 0; dd < nd; dd += 4) {
                    const int t = dd + slot;
                    const float4 wt = recW[t];
                    Taps64 T;
                    load64(T, src + recO[t], pitch);
                    const float c = group_dot(T, wt, refp) * (1.0f / 6.0f);
                    if (t < nd) out[(size_t)t * P3 * 8] = c;
                }
            } else {
            const float* src = fb + 2 * g;
#pragma unroll 2
            for (int dd = 0; dd < nd; dd += 4) {
                const int t = dd + slot;
                const float4 wt = recW[t];
                Taps48 T;
                load48(T, src + recO[t], pitch);
                float acc[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) acc[k] = pair_dot(T, k, wt, ref[k]);
                const float c = regroup48(acc, lane) * (1.0f / 6.0f);
                if (t < nd) out[(size_t)t * P3 * 8] = c;
            }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fea3 [B][V][P3][48] -> padded [B][V][P3][64]: thread = (texel, correlation group g): channels 6g..6g+5 go to floats
// 4g..4g+3 and 32+4g, 32+4g+1 of the texel, floats 32+4g+2, +3 are zero.
// ---------------------------------------------------------------------------------------------
__global__ void pad_level3_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t texels) {
    pdl_trigger();
    pdl_wait();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < texels * 8) pad_level3_item(src, dst, t);
}

// ---------------------------------------------------------------------------------------------
// view-weighted aggregation of the init volume (itermvs.py:59-69): one thread per float4 of the
// [B][D][P3][8] output.
// ---------------------------------------------------------------------------------------------
__global__ void aggregate_init_kernel(const float* __restrict__ corr, const float* __restrict__ vw3,
                                      float* __restrict__ agg, int B, int S, int D, int P3) {
    pdl_trigger();
    pdl_wait();
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)B * D * P3 * 2;
    if (t >= total) return;
    int half = (int)(t & 1);
    size_t q = t >> 1;                 // (b*D + d)*P3 + p
    int p = (int)(q % P3);
    size_t bd = q / P3;
    int d = (int)(bd % D), b = (int)(bd / D);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float wsum = 1e-5f;                // itermvs.py:38
    for (int v = 0; v < S; ++v) {
        float w = ldg(vw3 + ((size_t)b * S + v) * P3 + p);
        float4 c = ldg4(corr + ((((size_t)b * S + v) * D + d) * P3 + p) * 8 + half * 4);
        acc.x = fmaf(c.x, w, acc.x); acc.y = fmaf(c.y, w, acc.y);
        acc.z = fmaf(c.z, w, acc.z); acc.w = fmaf(c.w, w, acc.w);
        wsum += w;
    }
    acc.x /= wsum; acc.y /= wsum; acc.z /= wsum; acc.w /= wsum;
    reinterpret_cast<float4*>(agg)[t] = acc;
}

}  // namespace imvs

using namespace imvs;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int warpcorr_init_impl(bool pad3, const float* fea3, const float* rt3, const float* depth_min, const float* depth_max,
                              const float* depth_samples, float* corr, int B, int V, int H3, int W3, int D, void* stream) {
    IMVS_REQUIRE(fea3 && rt3 && corr && (depth_samples || (depth_min && depth_max)), "warpcorr_init: null pointer");
    IMVS_REQUIRE(B >= 1 && V >= 2 && V - 1 <= IMVS_MAX_VIEWS, "warpcorr_init: need 1..%d source views (V=%d)", IMVS_MAX_VIEWS, V);
    IMVS_REQUIRE(H3 >= 2 && W3 >= 2 && D >= 2, "warpcorr_init: bad shape H3=%d W3=%d D=%d", H3, W3, D);
    IMVS_REQUIRE((double)V * H3 * W3 * 64 < 2147483647.0, "warpcorr_init: one batch item's pyramid exceeds 2^31 elements");
    IMVS_REQUIRE(aligned16(fea3) && aligned16(corr), "warpcorr_init: feature/corr pointers must be 16-byte aligned");
    // source views per block (IMVS_TUNE_WCI_VPER, 0 = all of them in one block)
    const int S = V - 1, vt = tune("WCI_VPER", 0), vper = vt >= 1 && vt < S ? vt : S;
    dim3 grid(cdiv(W3, 2), cdiv(H3, 2), B * cdiv(S, vper));
    IMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "warpcorr_init: grid too large");
    ApiScope api_;
    IMVS_CUDA(launch_k(pad3 ? warpcorr_init_kernel<true> : warpcorr_init_kernel<false>, grid, dim3(WC_WARPS * 32), 0, (cudaStream_t)stream,
                       fea3, rt3, depth_min, depth_max, depth_samples, corr, B, V, H3, W3, D, vper));
    return 0;
}

extern "C" int imvs_warpcorr_init(const float* fea3, const float* rt3, const float* depth_min, const float* depth_max,
                                  const float* depth_samples, float* corr, int B, int V, int H3, int W3, int D, void* stream) {
    return warpcorr_init_impl(false, fea3, rt3, depth_min, depth_max, depth_samples, corr, B, V, H3, W3, D, stream);
}

extern "C" int imvs_warpcorr_init_padded(const float* fea3p, const float* rt3, const float* depth_min, const float* depth_max,
                                         const float* depth_samples, float* corr, int B, int V, int H3, int W3, int D, void* stream) {
    return warpcorr_init_impl(true, fea3p, rt3, depth_min, depth_max, depth_samples, corr, B, V, H3, W3, D, stream);
}

extern "C" int imvs_pad_level3(const float* fea3, float* fea3p, int B, int V, int H3, int W3, void* stream) {
    IMVS_REQUIRE(fea3 && fea3p && B >= 1 && V >= 1 && H3 >= 1 && W3 >= 1, "pad_level3: bad argument");
    IMVS_REQUIRE(aligned16(fea3) && aligned16(fea3p), "pad_level3: pointers must be 16-byte aligned");
    const size_t texels = (size_t)B * V * H3 * W3;
    IMVS_REQUIRE(texels * 8 < (size_t)2147483647 * 256, "pad_level3: too many texels");
    ApiScope api_;
    IMVS_CUDA(launch_k(pad_level3_kernel, dim3((unsigned)((texels * 8 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, fea3, fea3p, texels));
    return 0;
}

static int warpcorr_iter_impl(bool pad3, const float* fea1, const float* fea2, const float* fea3,
                              const float* rt1, const float* rt2, const float* rt3,
                              const float* nd, size_t nd_batch_stride, size_t nd_pixel_stride, const float* vw2,
                              const float* depth_min, const float* depth_max,
                              const float* samples1, const float* samples2, const float* samples3, float* agg,
                              int B, int V, int H2, int W2, void* stream) {
    const bool explicit_samples = samples1 && samples2 && samples3;
    IMVS_REQUIRE(fea1 && fea2 && fea3 && rt1 && rt2 && rt3 && vw2 && agg, "warpcorr_iter: null pointer");
    IMVS_REQUIRE(explicit_samples || (!samples1 && !samples2 && !samples3 && nd && depth_min && depth_max),
                 "warpcorr_iter: pass either all three sample tensors or nd + depth range");
    IMVS_REQUIRE(B >= 1 && V >= 2 && V - 1 <= IMVS_MAX_VIEWS, "warpcorr_iter: need 1..%d source views (V=%d)", IMVS_MAX_VIEWS, V);
    IMVS_REQUIRE(H2 >= 4 && W2 >= 4 && H2 % 2 == 0 && W2 % 2 == 0, "warpcorr_iter: H2, W2 must be even and >= 4 (H2=%d W2=%d)", H2, W2);
    IMVS_REQUIRE((double)V * H2 * W2 * 64 < 2147483647.0, "warpcorr_iter: one batch item's level-1 pyramid exceeds 2^31 elements");
    IMVS_REQUIRE(aligned16(fea1) && aligned16(fea2) && aligned16(fea3) && aligned16(agg),
                 "warpcorr_iter: feature/agg pointers must be 16-byte aligned");
    IterParams prm;
    prm.fea[0] = fea1; prm.fea[1] = fea2; prm.fea[2] = fea3;
    prm.rt[0] = rt1; prm.rt[1] = rt2; prm.rt[2] = rt3;
    prm.nd = nd; prm.nd_stride = nd_batch_stride; prm.nd_pstride = nd_pixel_stride; prm.vw2 = vw2;
    prm.depth_min = depth_min; prm.depth_max = depth_max; prm.agg = agg;
    prm.samples[0] = samples1; prm.samples[1] = samples2; prm.samples[2] = samples3;
    prm.B = B; prm.V = V; prm.H2 = H2; prm.W2 = W2;
    // one persistent block per SM, each owning a contiguous, compact range of 4-pixel rows
    const int S = V - 1;
    // occupancy experiment (IMVS_TUNE_WC_WARPS): 24 warps x 1 block per SM (default, <= 85 registers), 32 warps x 1 block or
    // 16 warps x 2 blocks (both <= 64 registers: 32 resident warps per SM) -- profiles/ps_experiments_r02.md section 5
    // 26 / 28 warps (72 registers): 104 items per SM at 640x512 are 4 full rounds instead of 4.33
    const int nwt = tune("WC_WARPS", WC_ITER_WARPS);
    const int nw0 = nwt == 32 || nwt == 16 || nwt == 26 || nwt == 28 ? nwt : WC_ITER_WARPS;
    const int nw = pad3 ? WC_ITER_WARPS : nw0;       // the padded variant is built for the default block shape only
    auto kern = pad3 ? iter_kernel_padded_for(S)
              : nw == 32 ? iter_kernel_for<32>(S) : nw == 16 ? iter_kernel_for<16>(S) : nw == 26 ? iter_kernel_for<26>(S)
              : nw == 28 ? iter_kernel_for<28>(S) : iter_kernel_for<WC_ITER_WARPS>(S);
    const size_t smem = iter_smem_bytes(S, nw);
    int dev = 0, sms = 0;
    IMVS_CUDA(cudaGetDevice(&dev));
    IMVS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (smem > 48 * 1024) IMVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IterParams5 q;
    q.p = prm;
    q.tiles_x = cdiv(W2, WC_NPX); q.tiles_y = cdiv(H2, 4);
    const long long tiles = (long long)B * q.tiles_x * q.tiles_y;
    IMVS_REQUIRE(tiles < (1LL << 26), "warpcorr_iter: too many tiles");
    IMVS_REQUIRE(2 * q.tiles_x <= 4096, "warpcorr_iter: W2=%d too wide", W2);
    q.n_tiles = (unsigned)tiles;
    q.strip_magic = ((1ULL << 40) + 2 * q.tiles_x - 1) / (2 * q.tiles_x);
    const int blocks = (int)std::min<long long>(tiles, nw == 16 ? 2 * sms : sms);
    ApiScope api_;
    IMVS_CUDA(launch_k(kern, dim3(blocks), dim3(nw * 32), smem, (cudaStream_t)stream, q));
    return 0;
}

extern "C" int imvs_warpcorr_iter(const float* fea1, const float* fea2, const float* fea3,
                                  const float* rt1, const float* rt2, const float* rt3,
                                  const float* nd, size_t nd_batch_stride, size_t nd_pixel_stride, const float* vw2,
                                  const float* depth_min, const float* depth_max,
                                  const float* samples1, const float* samples2, const float* samples3, float* agg,
                                  int B, int V, int H2, int W2, void* stream) {
    return warpcorr_iter_impl(false, fea1, fea2, fea3, rt1, rt2, rt3, nd, nd_batch_stride, nd_pixel_stride, vw2, depth_min, depth_max,
                              samples1, samples2, samples3, agg, B, V, H2, W2, stream);
}

extern "C" int imvs_warpcorr_iter_padded(const float* fea1, const float* fea2, const float* fea3p,
                                         const float* rt1, const float* rt2, const float* rt3,
                                         const float* nd, size_t nd_batch_stride, size_t nd_pixel_stride, const float* vw2,
                                         const float* depth_min, const float* depth_max,
                                         const float* samples1, const float* samples2, const float* samples3, float* agg,
                                         int B, int V, int H2, int W2, void* stream) {
    return warpcorr_iter_impl(true, fea1, fea2, fea3p, rt1, rt2, rt3, nd, nd_batch_stride, nd_pixel_stride, vw2, depth_min, depth_max,
                              samples1, samples2, samples3, agg, B, V, H2, W2, stream);
}

extern "C" int imvs_aggregate_init(const float* corr, const float* vw3, float* agg, int B, int S, int D, int P3, void* stream) {
    IMVS_REQUIRE(corr && vw3 && agg && B >= 1 && S >= 1 && D >= 1 && P3 >= 1, "aggregate_init: bad argument");
    size_t total = (size_t)B * D * P3 * 2;
    ApiScope api_;
    IMVS_CUDA(launch_k(aggregate_init_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, corr, vw3, agg, B, S, D, P3));
    return 0;
}
