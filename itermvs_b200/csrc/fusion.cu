// Depth-map filtering for fusion: the per-pixel geometric consistency check between a reference
// depth map and a source depth map, and the accumulation over the source views.
// Reference: eval.py:154-215 (reproject_with_depth, check_geometric_consistency) and eval.py:243-262 (the loop
// of filter_depth).  The reference runs this in numpy on the CPU, one pass per (reference, source) pair over
// full-resolution maps; here one thread owns one reference pixel and does the whole chain
//   back-project -> source camera -> project -> bilinear sample of the source depth (cv2.remap) ->
//   back-project with the sampled depth -> reference camera -> project -> distance / relative depth tests
// in registers.  Arithmetic types follow numpy's promotions in the reference: the camera matrices are float32
// values (inverses and products are formed in float32 on the host, as np.linalg.inv / np.matmul do there), the
// per-pixel geometry is float64, the sampled / reprojected depths and the tests on them float32.
#include "common.cuh"

namespace imvs {

struct PairCams {          // float32 values, row-major
    float Kinv_ref[9];     // inv(intrinsics_ref)
    float Mrs[16];         // extrinsics_src @ inv(extrinsics_ref)
    float K_src[9];
    float Kinv_src[9];
    float Msr[16];         // extrinsics_ref @ inv(extrinsics_src)
    float K_ref[9];
};

__device__ __forceinline__ void mat3(const float* M, double a, double b, double c, double& x, double& y, double& z) {
    x = fma((double)M[2], c, fma((double)M[1], b, (double)M[0] * a));
    y = fma((double)M[5], c, fma((double)M[4], b, (double)M[3] * a));
    z = fma((double)M[8], c, fma((double)M[7], b, (double)M[6] * a));
}
__device__ __forceinline__ void mat34(const float* M, double a, double b, double c, double& x, double& y, double& z) {
    x = fma((double)M[2], c, fma((double)M[1], b, (double)M[0] * a)) + (double)M[3];
    y = fma((double)M[6], c, fma((double)M[5], b, (double)M[4] * a)) + (double)M[7];
    z = fma((double)M[10], c, fma((double)M[9], b, (double)M[8] * a)) + (double)M[11];
}

// cv2.remap(src, map_x, map_y, INTER_LINEAR) for a CV_32FC1 image, BORDER_CONSTANT 0: the float map is
// converted to fixed point with 5 fractional bits (round half to even), the integer part saturates to int16,
// the four tap weights come from a 32 x 32 table of float products (1 - fy) * (1 - fx) ..., taps outside the image
// read the border value 0.
__device__ __forceinline__ float remap_linear(const float* __restrict__ img, int H, int W, float mx, float my) {
    // cvRound of x * 32 (lrint; out-of-range / NaN conversions give INT_MIN on x86: far outside either way)
    const float fx32 = mx * 32.0f, fy32 = my * 32.0f;
    int ix = (fx32 >= -2147483648.0f && fx32 < 2147483648.0f) ? __float2int_rn(fx32) : (int)0x80000000;
    int iy = (fy32 >= -2147483648.0f && fy32 < 2147483648.0f) ? __float2int_rn(fy32) : (int)0x80000000;
    int sx = ix >> 5, sy = iy >> 5;
    sx = min(max(sx, -32768), 32767);
    sy = min(max(sy, -32768), 32767);
    const float ax = (float)(ix & 31) * (1.0f / 32.0f), ay = (float)(iy & 31) * (1.0f / 32.0f);
    const float w00 = (1.0f - ay) * (1.0f - ax), w01 = (1.0f - ay) * ax, w10 = ay * (1.0f - ax), w11 = ay * ax;
    auto px = [&](int yy, int xx) { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? ldg(img + (size_t)yy * W + xx) : 0.0f; };
    // no contraction: products and sums are rounded one by one, left to right, as the scalar loop of remapBilinear does
    float r = __fmul_rn(px(sy, sx), w00);
    r = __fadd_rn(r, __fmul_rn(px(sy, sx + 1), w01));
    r = __fadd_rn(r, __fmul_rn(px(sy + 1, sx), w10));
    r = __fadd_rn(r, __fmul_rn(px(sy + 1, sx + 1), w11));
    return r;
}

// One (reference, source) pair.  Optional outputs may be null.  With `acc_sum` / `acc_cnt` the masked reprojected
// depth and the mask are also accumulated (eval.py:259-260), in source order across successive launches.
__global__ void __launch_bounds__(256)
geo_consistency_kernel(const float* __restrict__ depth_ref, const float* __restrict__ depth_src, const PairCams cam,
                       float pix_thres, float depth_thres, unsigned char* __restrict__ mask_out,
                       float* __restrict__ depth_rep_out, float* __restrict__ x_src_out, float* __restrict__ y_src_out,
                       float* __restrict__ acc_sum, int* __restrict__ acc_cnt, int H, int W) {
    pdl_trigger();
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t p = (size_t)y * W + x;
    const float dref = ldg(depth_ref + p);
    // eval.py:162-170
    double X, Y, Z, sxw, syw, szw, kx, ky, kz;
    mat3(cam.Kinv_ref, (double)x * (double)dref, (double)y * (double)dref, (double)dref, X, Y, Z);
    mat34(cam.Mrs, X, Y, Z, sxw, syw, szw);
    mat3(cam.K_src, sxw, syw, szw, kx, ky, kz);
    const double xs = kx / kz, ys = ky / kz;
    // eval.py:174-176
    const float xsf = (float)xs, ysf = (float)ys;
    const float sampled = remap_linear(depth_src, H, W, xsf, ysf);
    // eval.py:181-190
    mat3(cam.Kinv_src, xs * (double)sampled, ys * (double)sampled, (double)sampled, X, Y, Z);
    double rx, ry, rz;
    mat34(cam.Msr, X, Y, Z, rx, ry, rz);
    float drep = (float)rz;
    mat3(cam.K_ref, rx, ry, rz, kx, ky, kz);
    const float xr = (float)(kx / (kz + 1e-6)), yr = (float)(ky / (kz + 1e-6));
    // eval.py:205-210: float32 maps minus int64 grids -> float64; depths stay float32
    const double dx = (double)xr - (double)x, dy = (double)yr - (double)y;
    const double dist = sqrt(dx * dx + dy * dy);
    const float rel = __fdiv_rn(fabsf(__fsub_rn(drep, dref)), dref);
    const bool ok = dist < (double)pix_thres && rel < depth_thres;
    if (!ok) drep = 0.0f;                                               // eval.py:213
    if (mask_out) mask_out[p] = ok ? 1 : 0;
    if (depth_rep_out) depth_rep_out[p] = drep;
    if (x_src_out) x_src_out[p] = xsf;
    if (y_src_out) y_src_out[p] = ysf;
    if (acc_sum) acc_sum[p] = __fadd_rn(acc_sum[p], drep);
    if (acc_cnt) acc_cnt[p] += ok ? 1 : 0;
}

// eval.py:262-265: averaged depth (float32 sum / integer count -> float64) and the three masks
__global__ void fuse_finalize_kernel(const float* __restrict__ depth_ref, const float* __restrict__ conf,
                                     const float* __restrict__ acc_sum, const int* __restrict__ acc_cnt, float photo_thres,
                                     int geo_mask_thres, double* __restrict__ depth_avg, unsigned char* __restrict__ photo_mask,
                                     unsigned char* __restrict__ geo_mask, unsigned char* __restrict__ final_mask, size_t n) {
    pdl_trigger();
    pdl_wait();
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int cnt = acc_cnt[p];
    const float s = __fadd_rn(acc_sum[p], ldg(depth_ref + p));
    depth_avg[p] = (double)s / (double)(cnt + 1);
    const bool ph = ldg(conf + p) > photo_thres, ge = cnt >= geo_mask_thres;
    if (photo_mask) photo_mask[p] = ph;
    if (geo_mask) geo_mask[p] = ge;
    final_mask[p] = ph && ge;
}

static void fill_cams(PairCams& c, const float* m) {      // 68 floats in PairCams order
    memcpy(&c, m, sizeof(PairCams));
}

}  // namespace imvs

using namespace imvs;

extern "C" int imvs_check_geometric_consistency(const float* depth_ref, const float* depth_src, const float* cams68_host,
                                                float geo_pixel_thres, float geo_depth_thres, unsigned char* mask,
                                                float* depth_reprojected, float* x2d_src, float* y2d_src,
                                                float* acc_sum, int* acc_count, int H, int W, void* stream) {
    IMVS_REQUIRE(depth_ref && depth_src && cams68_host, "check_geometric_consistency: null pointer");
    IMVS_REQUIRE(mask || depth_reprojected || acc_sum || acc_count, "check_geometric_consistency: no output requested");
    IMVS_REQUIRE(H >= 1 && W >= 1 && H <= 65535, "check_geometric_consistency: bad shape H=%d W=%d", H, W);
    static_assert(sizeof(PairCams) == 68 * sizeof(float), "PairCams layout");
    PairCams cam;
    fill_cams(cam, cams68_host);
    ApiScope api_;
    IMVS_CUDA(launch_k(geo_consistency_kernel, dim3(cdiv(W, 256), H), dim3(256), 0, (cudaStream_t)stream, depth_ref, depth_src, cam,
                       geo_pixel_thres, geo_depth_thres, mask, depth_reprojected, x2d_src, y2d_src, acc_sum, acc_count, H, W));
    return 0;
}

extern "C" int imvs_filter_depth_view(const float* depth_ref, const float* confidence, const float* depth_srcs,
                                      const float* cams68_host, int S, float geo_pixel_thres, float geo_depth_thres,
                                      float photo_thres, int geo_mask_thres, float* acc_sum, int* acc_count,
                                      double* depth_averaged, unsigned char* photo_mask, unsigned char* geo_mask,
                                      unsigned char* final_mask, int H, int W, void* stream) {
    IMVS_REQUIRE(depth_ref && confidence && depth_srcs && cams68_host && acc_sum && acc_count && depth_averaged && final_mask,
                 "filter_depth_view: null pointer");
    IMVS_REQUIRE(S >= 1 && H >= 1 && W >= 1 && H <= 65535, "filter_depth_view: bad shape S=%d H=%d W=%d", S, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)H * W;
    ApiScope api_;
    IMVS_CUDA(cudaMemsetAsync(acc_sum, 0, sizeof(float) * n, st));
    IMVS_CUDA(cudaMemsetAsync(acc_count, 0, sizeof(int) * n, st));
    for (int s = 0; s < S; ++s) {
        PairCams cam;
        fill_cams(cam, cams68_host + (size_t)s * 68);
        IMVS_CUDA(launch_k(geo_consistency_kernel, dim3(cdiv(W, 256), H), dim3(256), 0, st, depth_ref, depth_srcs + (size_t)s * n, cam,
                           geo_pixel_thres, geo_depth_thres, (unsigned char*)nullptr, (float*)nullptr, (float*)nullptr,
                           (float*)nullptr, acc_sum, acc_count, H, W));
    }
    IMVS_CUDA(launch_k(fuse_finalize_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, depth_ref, confidence,
                       (const float*)acc_sum, (const int*)acc_count, photo_thres, geo_mask_thres, depth_averaged, photo_mask, geo_mask,
                       final_mask, n));
    return 0;
}
