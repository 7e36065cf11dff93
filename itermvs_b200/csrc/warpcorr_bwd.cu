// Backward of the fused plane-sweep kernels of warpcorr.cu: gradients of the group-wise correlation volumes with
// respect to the feature pyramids (reference view and source views).  The training path of the reference keeps, for
// every one of its 4 + 12 * iterations differentiable_warping calls, the [C, R, H, W] warped volume and the same-size
// product alive for autograd (module.py:118-120, itermvs.py:50, 103); here nothing is saved: the backward recomputes
// the sampling positions from the (detached) hypotheses and re-gathers the taps.
//
// What carries a gradient: the source features (grid_sample's input gradient) and the reference
// feature (through the product and, at levels 1 / 3, through its resampling, itermvs.py:95-98).  The sampling grid
// does not (module.py:77: built under no_grad), nor do the per-iteration view weights (itermvs.py:295: detached) or
// the hypotheses (itermvs.py:282-283: normalized_depth detached).
//
// Thread mapping: one thread = (pixel, 4 consecutive channels); it loops over views and hypotheses, keeps the
// gradient of "its" reference-feature channels in registers and scatters the source-feature gradient with 16-byte
// vector atomics (red.global.add.v4.f32, sm_90+).  Consecutive threads hold consecutive channel chunks of one pixel,
// so every tap is one contiguous 64 / 128 / 192-byte read and one contiguous vector-atomic burst per sample.
#include "common.cuh"
#include "sampling.cuh"

namespace imvs {

__device__ __forceinline__ void red_add4(float* p, const float4& v) {
    atomicAdd(reinterpret_cast<float4*>(p), v);
}

// One (pixel, hypothesis, view) sample, channels c0 .. c0+3 (src / gsrc already point at channel c0 of the view):
//   warped = sum_t w_t * src[tap_t]                    grid_sample(bilinear, zeros, align_corners=True)
//   gref  += gc * warped                               d corr / d ref
//   gsrc[tap_t] += w_t * gc * ref                      d corr / d src
__device__ __forceinline__ void bwd_sample(const Tap& tp, const float* __restrict__ src, float* __restrict__ gsrc, int Wf, int C,
                                           const float4& gc, const float4& ref, float4& gref) {
    if (tp.mask == 0) return;
    const float gx = 1.f - tp.fx, gy = 1.f - tp.fy;
    const float w[4] = {gx * gy, tp.fx * gy, gx * tp.fy, tp.fx * tp.fy};
    const int base = (tp.y0 * Wf + tp.x0) * C;
    const int offs[4] = {0, C, Wf * C, Wf * C + C};
    const float4 gr = make_float4(gc.x * ref.x, gc.y * ref.y, gc.z * ref.z, gc.w * ref.w);
    float4 warped = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (!((tp.mask >> t) & 1u)) continue;
        const float4 s = ldg4(src + base + offs[t]);
        warped.x = fmaf(s.x, w[t], warped.x); warped.y = fmaf(s.y, w[t], warped.y);
        warped.z = fmaf(s.z, w[t], warped.z); warped.w = fmaf(s.w, w[t], warped.w);
        red_add4(gsrc + base + offs[t], make_float4(gr.x * w[t], gr.y * w[t], gr.z * w[t], gr.w * w[t]));
    }
    gref.x = fmaf(gc.x, warped.x, gref.x); gref.y = fmaf(gc.y, warped.y, gref.y);
    gref.z = fmaf(gc.z, warped.z, gref.z); gref.w = fmaf(gc.w, warped.w, gref.w);
}

// ---------------------------------------------------------------------------------------------
// Init plane sweep (itermvs.py:45-51): corr[b][v][d][p][g] = mean_{c in group g} warped_v[c][d][p] * ref[c][p].
//   grid (ceil(B * P3 * 12 / 128)), block 128
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
warpcorr_init_bwd_kernel(const float* __restrict__ fea3, const float* __restrict__ rt3, const float* __restrict__ depth_min,
                         const float* __restrict__ depth_max, const float* __restrict__ samples,
                         const float* __restrict__ gcorr, float* __restrict__ gfea3, int B, int V, int H3, int W3, int D) {
    pdl_trigger();
    pdl_wait();
    const int P3 = H3 * W3, S = V - 1;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * P3 * 12) return;
    const int j = (int)(t % 12);
    const long long q = t / 12;
    const int p = (int)(q % P3), b = (int)(q / P3);
    const int x = p % W3, y = p / W3, c0 = 4 * j;
    const float* fb = fea3 + (size_t)b * V * P3 * 48 + c0;
    float* gb = gfea3 + (size_t)b * V * P3 * 48 + c0;
    const float4 ref = ldg4(fb + (size_t)p * 48);
    float4 gref = make_float4(0.f, 0.f, 0.f, 0.f);
    const float inv_min = samples ? 0.f : 1.0f / depth_min[b], inv_max = samples ? 0.f : 1.0f / depth_max[b];
    const int g0 = c0 / 6, g1 = (c0 + 1) / 6, g2 = (c0 + 2) / 6, g3 = (c0 + 3) / 6;
    for (int v = 0; v < S; ++v) {
        const float* Pm = rt3 + ((size_t)b * S + v) * 12;
        const float* src = fb + (size_t)(v + 1) * P3 * 48;
        float* gsrc = gb + (size_t)(v + 1) * P3 * 48;
        for (int d = 0; d < D; ++d) {
            // the forward's hypotheses (warpcorr_init_kernel; itermvs.py:13-17)
            const float depth = samples ? ldg(samples + ((size_t)b * D + d) * P3 + p)
                                        : 1.0f / (inv_max + ((float)d / (float)(D - 1)) * (inv_min - inv_max));
            const Tap tp = project_tap(Pm, (float)x, (float)y, depth, (float)W3, (float)H3, W3, H3);
            const float* gp = gcorr + ((((size_t)b * S + v) * D + d) * P3 + p) * 8;
            const float k = 1.0f / 6.0f;
            const float4 gc = make_float4(ldg(gp + g0) * k, ldg(gp + g1) * k, ldg(gp + g2) * k, ldg(gp + g3) * k);
            bwd_sample(tp, src, gsrc, W3, 48, gc, ref, gref);
        }
    }
    // the reference view's map is written by exactly one thread per (pixel, chunk)
    *reinterpret_cast<float4*>(gb + (size_t)p * 48) = gref;
}

// ---------------------------------------------------------------------------------------------
// Iteration kernel (itermvs.py:86-120), one launch per pyramid level:
//   agg[b][slice0 + r][p][g] = sum_v w_v * corr_v[r][p][g] / (1e-5 + sum_v w_v),   w_v detached.
//   LVL 0: level 1 (C = 16, map 2x the depth map, reference = 2x2 mean), 1: level 2 (C = 32), 2: level 3 (C = 48, half
//   size, reference = bilinear x2, align_corners=False).
// ---------------------------------------------------------------------------------------------
struct IterBwdParams {
    const float* fea;      // this level's pyramid [B][V][Hf][Wf][C]
    const float* rt;       // [B][S][12]
    const float* nd;       // [B][nd_stride] or null when explicit samples are given
    size_t nd_stride, nd_pstride;
    const float* vw2;      // [B][S][P2]
    const float* depth_min;
    const float* depth_max;
    const float* samples;  // optional explicit hypotheses [B][R][P2]
    const float* gagg;     // [B][10][P2][8]
    float* gfea;           // [B][V][Hf][Wf][C], zeroed by the caller
    int B, V, H2, W2;
};

template <int LVL>
__global__ void __launch_bounds__(128) warpcorr_iter_bwd_kernel(const IterBwdParams prm) {
    constexpr int C = LVL == 0 ? 16 : (LVL == 1 ? 32 : 48);
    constexpr int R = LVL == 2 ? 2 : 4;
    constexpr int CPG = C / 8, NCH = C / 4;
    constexpr int SLICE0 = LVL == 0 ? 0 : (LVL == 1 ? 4 : 8);
    constexpr float SC = LVL == 0 ? 2.f : (LVL == 1 ? 1.f : 0.5f);
    pdl_trigger();
    pdl_wait();
    const int H2 = prm.H2, W2 = prm.W2, P2 = H2 * W2, V = prm.V, S = V - 1;
    const int Hf = LVL == 0 ? H2 * 2 : (LVL == 2 ? H2 / 2 : H2);
    const int Wf = LVL == 0 ? W2 * 2 : (LVL == 2 ? W2 / 2 : W2);
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)prm.B * P2 * NCH) return;
    const int j = (int)(t % NCH);
    const long long q = t / NCH;
    const int p = (int)(q % P2), b = (int)(q / P2);
    const int x = p % W2, y = p / W2, c0 = 4 * j;
    const size_t view_elems = (size_t)Hf * Wf * C;
    const float* fb = prm.fea + (size_t)b * V * view_elems + c0;
    float* gb = prm.gfea + (size_t)b * V * view_elems + c0;

    // the reference feature of this level-2 pixel and the (<= 4) texels of the reference map it is made of
    constexpr int NREF = LVL == 1 ? 1 : 4;
    int ro[4];
    float rw[4];
    if (LVL == 1) {
        ro[0] = (y * Wf + x) * C; rw[0] = 1.f;
    } else if (LVL == 0) {                     // F.interpolate(scale 0.5, bilinear) == 2x2 mean
        ro[0] = ((2 * y) * Wf + 2 * x) * C; ro[1] = ro[0] + C; ro[2] = ro[0] + Wf * C; ro[3] = ro[2] + C;
        rw[0] = rw[1] = rw[2] = rw[3] = 0.25f;
    } else {                                   // F.interpolate(scale 2, bilinear, align_corners=False)
        int h0, h1, w0, w1;
        float lh, lw;
        up_index(y, 0.5f, Hf, h0, h1, lh);
        up_index(x, 0.5f, Wf, w0, w1, lw);
        ro[0] = (h0 * Wf + w0) * C; ro[1] = (h0 * Wf + w1) * C; ro[2] = (h1 * Wf + w0) * C; ro[3] = (h1 * Wf + w1) * C;
        rw[0] = (1.f - lh) * (1.f - lw); rw[1] = (1.f - lh) * lw; rw[2] = lh * (1.f - lw); rw[3] = lh * lw;
    }
    float4 ref = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NREF; ++i) {
        const float4 a = ldg4(fb + ro[i]);
        ref.x = fmaf(a.x, rw[i], ref.x); ref.y = fmaf(a.y, rw[i], ref.y);
        ref.z = fmaf(a.z, rw[i], ref.z); ref.w = fmaf(a.w, rw[i], ref.w);
    }
    float wsum = 1e-5f;                                                   // itermvs.py:88-89
    for (int v = 0; v < S; ++v) wsum += ldg(prm.vw2 + ((size_t)b * S + v) * P2 + p);
    float ndv = 0.f, inv_min = 0.f, inv_max = 0.f;
    if (!prm.samples) {
        ndv = ldg(prm.nd + (size_t)b * prm.nd_stride + (size_t)p * prm.nd_pstride);
        inv_min = 1.0f / prm.depth_min[b];
        inv_max = 1.0f / prm.depth_max[b];
    }
    const int g0 = c0 / CPG, g1 = (c0 + 1) / CPG, g2 = (c0 + 2) / CPG, g3 = (c0 + 3) / CPG;
    float4 gref = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int v = 0; v < S; ++v) {
        const float* Pm = prm.rt + ((size_t)b * S + v) * 12;
        const float* src = fb + (size_t)(v + 1) * view_elems;
        float* gsrc = gb + (size_t)(v + 1) * view_elems;
        const float k = ldg(prm.vw2 + ((size_t)b * S + v) * P2 + p) / wsum * (1.0f / (float)CPG);
        for (int r = 0; r < R; ++r) {
            float depth;
            if (prm.samples) {
                depth = ldg(prm.samples + ((size_t)b * R + r) * P2 + p);
            } else {                                                      // itermvs.py:229-235, 290-293
                const float o = LVL == 0 ? (r == 0 ? -2.f : r == 1 ? -2.0f / 3 : r == 2 ? 2.0f / 3 : 2.f)
                              : LVL == 1 ? (r == 0 ? -8.f : r == 1 ? -8.0f / 3 : r == 2 ? 8.0f / 3 : 8.f)
                                         : (r == 0 ? -32.f : 32.f);
                const float s = fminf(fmaxf(ndv + o * (1.0f / 256.0f), 0.f), 1.f);
                depth = unnormalize_depth(s, inv_min, inv_max);
            }
            const Tap tp = project_tap(Pm, (float)x * SC, (float)y * SC, depth, (float)W2, (float)H2, Wf, Hf);
            const float* gp = prm.gagg + (((size_t)b * IMVS_ITER_SLICES + SLICE0 + r) * P2 + p) * 8;
            const float4 gc = make_float4(ldg(gp + g0) * k, ldg(gp + g1) * k, ldg(gp + g2) * k, ldg(gp + g3) * k);
            bwd_sample(tp, src, gsrc, Wf, C, gc, ref, gref);
        }
    }
    // adjoint of the reference resampling: several level-2 pixels share a level-3 texel -> atomics
#pragma unroll
    for (int i = 0; i < NREF; ++i)
        red_add4(gb + ro[i], make_float4(gref.x * rw[i], gref.y * rw[i], gref.z * rw[i], gref.w * rw[i]));
}

}  // namespace imvs

using namespace imvs;

static inline bool aligned16b(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int imvs_warpcorr_init_backward(const float* fea3, const float* rt3, const float* depth_min, const float* depth_max,
                                           const float* depth_samples, const float* grad_corr, float* grad_fea3,
                                           int B, int V, int H3, int W3, int D, void* stream) {
    IMVS_REQUIRE(fea3 && rt3 && grad_corr && grad_fea3 && (depth_samples || (depth_min && depth_max)),
                 "warpcorr_init_backward: null pointer");
    IMVS_REQUIRE(B >= 1 && V >= 2 && V - 1 <= IMVS_MAX_VIEWS, "warpcorr_init_backward: need 1..%d source views (V=%d)", IMVS_MAX_VIEWS, V);
    IMVS_REQUIRE(H3 >= 2 && W3 >= 2 && D >= 2, "warpcorr_init_backward: bad shape H3=%d W3=%d D=%d", H3, W3, D);
    IMVS_REQUIRE((double)V * H3 * W3 * 48 < 2147483647.0, "warpcorr_init_backward: one batch item's pyramid exceeds 2^31 elements");
    IMVS_REQUIRE(aligned16b(fea3) && aligned16b(grad_fea3), "warpcorr_init_backward: feature pointers must be 16-byte aligned");
    const long long threads = (long long)B * H3 * W3 * 12;
    IMVS_REQUIRE(threads / 128 + 1 < 2147483647LL, "warpcorr_init_backward: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    ApiScope api_;
    IMVS_CUDA(cudaMemsetAsync(grad_fea3, 0, sizeof(float) * (size_t)B * V * H3 * W3 * 48, st));
    IMVS_CUDA(launch_k(warpcorr_init_bwd_kernel, dim3((unsigned)((threads + 127) / 128)), dim3(128), 0, st, fea3, rt3, depth_min,
                       depth_max, depth_samples, grad_corr, grad_fea3, B, V, H3, W3, D));
    return 0;
}

extern "C" int imvs_warpcorr_iter_backward(const float* fea1, const float* fea2, const float* fea3,
                                           const float* rt1, const float* rt2, const float* rt3,
                                           const float* nd, size_t nd_batch_stride, size_t nd_pixel_stride, const float* vw2,
                                           const float* depth_min, const float* depth_max,
                                           const float* samples1, const float* samples2, const float* samples3,
                                           const float* grad_agg, float* grad_fea1, float* grad_fea2, float* grad_fea3,
                                           int B, int V, int H2, int W2, void* stream) {
    const bool explicit_samples = samples1 && samples2 && samples3;
    IMVS_REQUIRE(fea1 && fea2 && fea3 && rt1 && rt2 && rt3 && vw2 && grad_agg && grad_fea1 && grad_fea2 && grad_fea3,
                 "warpcorr_iter_backward: null pointer");
    IMVS_REQUIRE(explicit_samples || (!samples1 && !samples2 && !samples3 && nd && depth_min && depth_max),
                 "warpcorr_iter_backward: pass either all three sample tensors or nd + depth range");
    IMVS_REQUIRE(B >= 1 && V >= 2 && V - 1 <= IMVS_MAX_VIEWS, "warpcorr_iter_backward: need 1..%d source views (V=%d)", IMVS_MAX_VIEWS, V);
    IMVS_REQUIRE(H2 >= 4 && W2 >= 4 && H2 % 2 == 0 && W2 % 2 == 0, "warpcorr_iter_backward: H2, W2 must be even and >= 4 (H2=%d W2=%d)", H2, W2);
    IMVS_REQUIRE((double)V * H2 * W2 * 64 < 2147483647.0, "warpcorr_iter_backward: one batch item's level-1 pyramid exceeds 2^31 elements");
    IMVS_REQUIRE(aligned16b(fea1) && aligned16b(fea2) && aligned16b(fea3) && aligned16b(grad_fea1) && aligned16b(grad_fea2) && aligned16b(grad_fea3),
                 "warpcorr_iter_backward: feature pointers must be 16-byte aligned");
    const long long px = (long long)B * H2 * W2;
    IMVS_REQUIRE(px * 12 / 128 + 1 < 2147483647LL, "warpcorr_iter_backward: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    IterBwdParams prm;
    prm.nd = nd; prm.nd_stride = nd_batch_stride; prm.nd_pstride = nd_pixel_stride; prm.vw2 = vw2;
    prm.depth_min = depth_min; prm.depth_max = depth_max; prm.gagg = grad_agg;
    prm.B = B; prm.V = V; prm.H2 = H2; prm.W2 = W2;
    ApiScope api_;
    const size_t vpx = (size_t)B * V * H2 * W2;
    IMVS_CUDA(cudaMemsetAsync(grad_fea1, 0, sizeof(float) * vpx * 4 * 16, st));
    IMVS_CUDA(cudaMemsetAsync(grad_fea2, 0, sizeof(float) * vpx * 32, st));
    IMVS_CUDA(cudaMemsetAsync(grad_fea3, 0, sizeof(float) * vpx / 4 * 48, st));
    prm.fea = fea1; prm.rt = rt1; prm.samples = samples1; prm.gfea = grad_fea1;
    IMVS_CUDA(launch_k(warpcorr_iter_bwd_kernel<0>, dim3((unsigned)((px * 4 + 127) / 128)), dim3(128), 0, st, prm));
    prm.fea = fea2; prm.rt = rt2; prm.samples = samples2; prm.gfea = grad_fea2;
    IMVS_CUDA(launch_k(warpcorr_iter_bwd_kernel<1>, dim3((unsigned)((px * 8 + 127) / 128)), dim3(128), 0, st, prm));
    prm.fea = fea3; prm.rt = rt3; prm.samples = samples3; prm.gfea = grad_fea3;
    IMVS_CUDA(launch_k(warpcorr_iter_bwd_kernel<2>, dim3((unsigned)((px * 12 + 127) / 128)), dim3(128), 0, st, prm));
    return 0;
}
