// Output stage: upsampling-weight net + softmax over the 9 taps + convex x4 upsampling of the
// normalized depth + depth_unnormalization, and the bilinear x4 of the confidence.
// Reference: models/itermvs.py:246-250, 262-264, 321-324; models/module.py:127-152.
// The [B,144,H2,W2] weight tensor of the reference (11.8 MB at 640x512) is never materialised.
#include <algorithm>

#include "common.cuh"
#include "mmaconv.cuh"

namespace imvs {

constexpr int UPS_THREADS = 256;

struct UpsParams {
    const float* t;        // [B][P2][64] relu'd conv output, channels-last
    const float* fc;       // [64][144]
    const float* nd;
    size_t nd_bstride, nd_pstride;
    const float* depth_min;
    const float* depth_max;
    float* depth_up;       // [B][4H2][4W2]
    int B, H2, W2;
};

// warp step = 8 consecutive quarter-res pixels; lane = 16*pp + s: sub-pixel s = 4*i + j of pixels
// pp*4 .. pp*4+3; 9 tap logits per (pixel, sub-pixel) live in registers, softmax needs no shuffles.
__global__ void __launch_bounds__(UPS_THREADS) convex_upsample_kernel(const UpsParams prm) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                       // [64][144]
    float* sWarp = sW + 64 * 144;           // per warp [64][8]
    pdl_trigger();                          // the 1x1 weights are constants: staged while the predecessor drains
    for (int i = threadIdx.x; i < 64 * 144 / 4; i += UPS_THREADS) reinterpret_cast<float4*>(sW)[i] = ldg4(prm.fc + 4 * i);
    pdl_wait();
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* st = sWarp + warp * 64 * 8;
    const int H2 = prm.H2, W2 = prm.W2, P = H2 * W2;
    const int items = (prm.B * P) / 8;
    const int nwarps = gridDim.x * (UPS_THREADS / 32);
    const int s = lane & 15, pp = lane >> 4;
    for (int item = blockIdx.x * (UPS_THREADS / 32) + warp; item < items; item += nwarps) {
        const int gp = item * 8;
        const int b = gp / P, p0 = gp % P;
        {
            const float* tb = prm.t + ((size_t)b * P + p0) * 64;
            for (int i = lane; i < 8 * 16; i += 32) {
                const int px = i / 16, k4 = i % 16;
                const float4 v = ldg4(tb + (size_t)px * 64 + 4 * k4);
                st[(4 * k4 + 0) * 8 + px] = v.x;
                st[(4 * k4 + 1) * 8 + px] = v.y;
                st[(4 * k4 + 2) * 8 + px] = v.z;
                st[(4 * k4 + 3) * 8 + px] = v.w;
            }
        }
        __syncwarp();
        float acc[4][9];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[i][k] = 0.f;
#pragma unroll 2
        for (int kk = 0; kk < 64; ++kk) {
            const float4 tv4 = reinterpret_cast<const float4*>(st + kk * 8)[pp];
            const float tv[4] = {tv4.x, tv4.y, tv4.z, tv4.w};
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float w = sW[kk * 144 + k * 16 + s];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][k] = fmaf(tv[i], w, acc[i][k]);
            }
        }
        const float inv_min = 1.0f / prm.depth_min[b], inv_max = 1.0f / prm.depth_max[b];
        const float* ndb = prm.nd + (size_t)b * prm.nd_bstride;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int p = p0 + pp * 4 + i;
            const int y = p / W2, x = p % W2;
            float m = acc[i][0];
#pragma unroll
            for (int k = 1; k < 9; ++k) m = fmaxf(m, acc[i][k]);
            float e[9], sum = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) { e[k] = expf(acc[i][k] - m); sum += e[k]; }
            float up = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int yy = min(max(y + k / 3 - 1, 0), H2 - 1), xx = min(max(x + k % 3 - 1, 0), W2 - 1);   // ReplicationPad2d(1)
                up = fmaf(ldg(ndb + (size_t)(yy * W2 + xx) * prm.nd_pstride), e[k] / sum, up);
            }
            const int oy = 4 * y + (s >> 2), ox = 4 * x + (s & 3);
            prm.depth_up[((size_t)b * 4 * H2 + oy) * (4 * W2) + ox] = unnormalize_depth(up, inv_min, inv_max);
        }
        __syncwarp();
    }
}

// F.interpolate(scale_factor=4, mode='bilinear') of a [N][H][W] map (itermvs.py:323-324)
__global__ void upsample4x_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W) {
    pdl_trigger();
    pdl_wait();
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int Ho = 4 * H, Wo = 4 * W;
    if (t >= (size_t)N * Ho * Wo) return;
    const int ox = (int)(t % Wo), oy = (int)((t / Wo) % Ho);
    const size_t n = t / ((size_t)Wo * Ho);
    int h0, h1, w0, w1;
    float lh, lw;
    up_index(oy, 0.25f, H, h0, h1, lh);
    up_index(ox, 0.25f, W, w0, w1, lw);
    const float* q = in + n * H * W;
    out[t] = (1.f - lh) * ((1.f - lw) * ldg(q + h0 * W + w0) + lw * ldg(q + h0 * W + w1)) +
             lh * ((1.f - lw) * ldg(q + h1 * W + w0) + lw * ldg(q + h1 * W + w1));
}

}  // namespace imvs

using namespace imvs;

extern "C" int imvs_upsample_outputs(const imvs_weights* w, const float* ref_fea2, size_t ref_batch_stride, const float* nd,
                                     size_t nd_batch_stride, size_t nd_pixel_stride, const float* conf, const float* depth_min,
                                     const float* depth_max, float* depth_up, float* conf_up, float* scratch,
                                     int B, int H2, int W2, void* stream) {
    IMVS_REQUIRE(w && ref_fea2 && nd && depth_min && depth_max && depth_up && scratch, "upsample_outputs: null pointer");
    IMVS_REQUIRE(B >= 1 && H2 >= 1 && W2 >= 1 && (H2 * W2) % 8 == 0, "upsample_outputs: H2*W2 must be a multiple of 8");
    IMVS_REQUIRE(!conf || conf_up, "upsample_outputs: conf given without conf_up");
    IMVS_REQUIRE(nd_pixel_stride >= 1, "upsample_outputs: nd_pixel_stride must be >= 1");
    ApiScope api_;
    cudaStream_t st = (cudaStream_t)stream;
    // IMVS_TUNE_UPS_TILE: 0 = 8-row tiles x 64 couts (160 CTAs at 640x512, 864 dependent MMAs per warp), 1 = 4-row tiles (320 CTAs),
    // 2 = 4-row tiles x two 32-cout blocks (640 CTAs)
    const EpiNHWC e0{scratch, nullptr, nullptr, H2, W2, 64, 64, 1};
    const int ut = tune("UPS_TILE", 2);          // default 2: stage 50.7 -> 46.6 us (gpurun call r2c58)
    if (ut == 1)
        IMVS_TRY((mma_conv<32, 64, 1, 4, 1, false>("upsample.conv0", in_nhwc(ref_fea2, H2, W2, 32, ref_batch_stride), e0, WSets::single(w->ups_conv0),
                                                   conv_tables(3, 1, 1, 4), B, 64, H2, W2, 1, st)));
    else if (ut == 2)
        IMVS_TRY((mma_conv<32, 32, 1, 4, 1, false>("upsample.conv0", in_nhwc(ref_fea2, H2, W2, 32, ref_batch_stride), e0, WSets::single(w->ups_conv0),
                                                   conv_tables(3, 1, 1, 4), B, 64, H2, W2, 2, st)));
    else
        IMVS_TRY((mma_conv<32, 64, 2, 4, 1, false>("upsample.conv0", in_nhwc(ref_fea2, H2, W2, 32, ref_batch_stride), e0, WSets::single(w->ups_conv0),
                                                   conv_tables(3, 1, 1, 8), B, 64, H2, W2, 1, st)));
    UpsParams prm;
    prm.t = scratch; prm.fc = w->ups_fc; prm.nd = nd; prm.nd_bstride = nd_batch_stride; prm.nd_pstride = nd_pixel_stride;
    prm.depth_min = depth_min; prm.depth_max = depth_max; prm.depth_up = depth_up;
    prm.B = B; prm.H2 = H2; prm.W2 = W2;
    const size_t smem = (size_t)(64 * 144 + (UPS_THREADS / 32) * 64 * 8) * sizeof(float);
    static int smem_ok = 0;
    IMVS_TRY(ensure_dynamic_smem(convex_upsample_kernel, smem, &smem_ok));
    int dev = 0, sms = 148;
    IMVS_CUDA(cudaGetDevice(&dev));
    IMVS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int items = (B * H2 * W2) / 8;
    const int blocks = std::min(cdiv(items, UPS_THREADS / 32), 3 * sms);
    IMVS_CUDA(launch_k(convex_upsample_kernel, dim3(blocks), dim3(UPS_THREADS), smem, st, prm));
    if (conf) {
        size_t total = (size_t)B * H2 * W2 * 16;
        IMVS_CUDA(launch_k(upsample4x_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, conf, conf_up, B, H2, W2));
    }
    return 0;
}
