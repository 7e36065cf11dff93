// f-4 on the device: the loaders' image path (reference datasets/dtu_yao_eval.py:61-76, `read_img`) after the JPEG / PNG
// decode -- normalisation 2 x / 255 - 1, cv2.resize(INTER_LINEAR) to the network size, and the three coarser pyramid levels
// cv2.resize'd from level 0 -- on the raw 8-bit image, so that a view costs H0 * W0 * 3 bytes of upload instead of the
// 4-level float pyramid (5.3 x as many bytes at 1600x1200 -> 1600x1152).
//
// cv2.resize, INTER_LINEAR, float32, no anti-aliasing: source coordinate f = (d + 0.5) * scale - 0.5 (double), i = floor(f),
// clamped to [0, src - 1] with the fraction zeroed where it clamps below and the right / lower neighbour clamped to
// src - 1; separable: rows are interpolated horizontally first (S[i] * (1 - fx) + S[i + 1] * fx), then the two row results
// vertically.  The arithmetic here follows that order with explicit roundings (no fused multiply-add); OpenCV's SIMD paths may
// fuse one of the products, so results agree with cv2 to 1 ulp of the interpolated value (tests/test_gpu_parity.py).
#include "common.cuh"

namespace imvs {

struct LinCoord { int i0, i1; float w0, w1; };

// scale = 1. / ((double)dst / src) as cv2 forms it; coordinate and fraction in double (what OpenCV's resize -- its IPP path on
// x86 -- was measured to do: with a float coordinate 27 % of the values differ by up to 2.5e-6), weights rounded to float once
__device__ __forceinline__ LinCoord lin_coord(int d, double scale, int n) {
    double f = ((double)d + 0.5) * scale - 0.5;
    int i = (int)floor(f);
    f -= (double)i;
    if (i < 0) { i = 0; f = 0.0; }
    if (i >= n - 1) { i = n - 1; f = 0.0; }         // cv2: sx clamped, the neighbour index then equals sx
    LinCoord c;
    c.i0 = i; c.i1 = min(i + 1, n - 1);
    c.w0 = (float)(1.0 - f); c.w1 = (float)f;
    return c;
}

__device__ __forceinline__ float lerp2(float a, float b, float w0, float w1) { return __fadd_rn(__fmul_rn(a, w0), __fmul_rn(b, w1)); }

// level 0: raw [H0][W0][3] uint8 -> planar [3][H][W] float in [-1, 1]
__global__ void __launch_bounds__(256) image_level0_kernel(const unsigned char* __restrict__ img, float* __restrict__ out, int H0, int W0,
                                                           int H, int W) {
    pdl_trigger();
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const bool same = H == H0 && W == W0;
    const LinCoord cx = lin_coord(x, 1.0 / ((double)W / W0), W0), cy = lin_coord(y, 1.0 / ((double)H / H0), H0);
    auto px = [&](int yy, int xx, int c) {
        return __fsub_rn(__fdiv_rn(2.0f * (float)__ldg(img + ((size_t)yy * W0 + xx) * 3 + c), 255.0f), 1.0f);     // 2 * x / 255. - 1
    };
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v;
        if (same) {
            v = px(y, x, c);
        } else {
            const float r0 = lerp2(px(cy.i0, cx.i0, c), px(cy.i0, cx.i1, c), cx.w0, cx.w1);
            const float r1 = lerp2(px(cy.i1, cx.i0, c), px(cy.i1, cx.i1, c), cx.w0, cx.w1);
            v = lerp2(r0, r1, cy.w0, cy.w1);
        }
        out[((size_t)c * H + y) * W + x] = v;
    }
}

// coarser level: planar [3][H][W] float -> [3][Hd][Wd]
__global__ void __launch_bounds__(256) image_down_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int Hd, int Wd) {
    pdl_trigger();
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, c = blockIdx.z;
    if (x >= Wd) return;
    const LinCoord cx = lin_coord(x, 1.0 / ((double)Wd / W), W), cy = lin_coord(y, 1.0 / ((double)Hd / H), H);
    const float* p = src + (size_t)c * H * W;
    const float r0 = lerp2(ldg(p + (size_t)cy.i0 * W + cx.i0), ldg(p + (size_t)cy.i0 * W + cx.i1), cx.w0, cx.w1);
    const float r1 = lerp2(ldg(p + (size_t)cy.i1 * W + cx.i0), ldg(p + (size_t)cy.i1 * W + cx.i1), cx.w0, cx.w1);
    dst[((size_t)c * Hd + y) * Wd + x] = lerp2(r0, r1, cy.w0, cy.w1);
}

}  // namespace imvs

using namespace imvs;

extern "C" int imvs_image_pyramid_u8(const unsigned char* img, int H0, int W0, float* level0, float* level1, float* level2, float* level3,
                                     int H, int W, void* stream) {
    IMVS_REQUIRE(img && level0, "image_pyramid_u8: null pointer");
    IMVS_REQUIRE(H0 >= 1 && W0 >= 1 && H >= 1 && W >= 1 && H <= 65535, "image_pyramid_u8: bad shape");
    ApiScope api_;
    cudaStream_t st = (cudaStream_t)stream;
    IMVS_CUDA(launch_k(image_level0_kernel, dim3(cdiv(W, 256), H), dim3(256), 0, st, img, level0, H0, W0, H, W));
    float* lv[3] = {level1, level2, level3};
    for (int k = 1; k <= 3; ++k) {
        if (!lv[k - 1]) continue;
        const int Hd = H >> k, Wd = W >> k;            // (w // 2^k, h // 2^k), dtu_yao_eval.py:69-72
        IMVS_REQUIRE(Hd >= 1 && Wd >= 1, "image_pyramid_u8: level %d is empty", k);
        IMVS_CUDA(launch_k(image_down_kernel, dim3(cdiv(Wd, 256), Hd, 3), dim3(256), 0, st, (const float*)level0, lv[k - 1], H, W, Hd, Wd));
    }
    return 0;
}
