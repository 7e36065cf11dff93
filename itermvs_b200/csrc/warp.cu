// Projection composition, the stand-alone differentiable_warping operator and layout helpers.
// Reference: models/module.py:68-125 (FangjinhuaWang/IterMVS).
#include <cstdlib>
#include "common.cuh"
#include "sampling.cuh"

namespace imvs {

thread_local char g_err[512] = {0};
char* err_buf() { return g_err; }
int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
long long launches_total() { return g_launches; }
static int g_conv_passes = 4;
int conv_passes() { return g_conv_passes; }
// experiment switch for tile-shape A/B runs: IMVS_TUNE_<NAME>=<int> in the environment (read on every call; cheap)
int tune(const char* name, int def) {
    char key[64];
    snprintf(key, sizeof key, "IMVS_TUNE_%s", name);
    const char* e = getenv(key);
    return e ? atoi(e) : def;
}
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("IMVS_PDL"); return e ? atoi(e) != 0 : true; }();
    return on;
}
static thread_local bool g_pdl_armed = false;
static thread_local int g_api_depth = 0;
bool* pdl_armed() { return &g_pdl_armed; }
int* api_depth() { return &g_api_depth; }
static int g_tc5 = 1;
int tc5_enabled() { return g_tc5; }
__device__ int g_tc5_err = 0;
__device__ long long g_tc5_clk[16];
static int g_tc5_clk_on = 0;
long long* tc5_clock_buffer() {
    if (!g_tc5_clk_on) return nullptr;
    static long long* p = nullptr;
    if (!p) {
        void* q = nullptr;
        if (cudaGetSymbolAddress(&q, g_tc5_clk) == cudaSuccess) p = static_cast<long long*>(q);
    }
    return p;
}
int* tc5_error_flag() {
    static int* p = nullptr;
    if (!p) {
        void* q = nullptr;
        if (cudaGetSymbolAddress(&q, g_tc5_err) == cudaSuccess) p = static_cast<int*>(q);
    }
    return p;
}

// ---- stage timing taps -----------------------------------------------------------------------
struct ProfileState {
    bool active = false;
    int cap = 0, n = 0;
    cudaEvent_t* ev = nullptr;     // 2 per record: start, stop
    int* tags = nullptr;
};
static ProfileState g_prof;
bool profile_active() { return g_prof.active; }
void profile_mark(int tag, cudaStream_t st, bool is_start) {
    if (!g_prof.active) return;
    if (is_start) {
        if (g_prof.n >= g_prof.cap) return;
        g_prof.tags[g_prof.n] = tag;
        cudaEventRecord(g_prof.ev[2 * g_prof.n], st);
    } else {
        if (g_prof.n >= g_prof.cap) return;
        cudaEventRecord(g_prof.ev[2 * g_prof.n + 1], st);
        g_prof.n++;
    }
}

// ---------------------------------------------------------------------------------------------
// K1: proj = src_proj @ inverse(ref_proj)  (module.py:78-90).  One thread per (b, source view).
// Done in fp64 from the fp32 inputs (the reference uses an fp32 LU; both round to the same fp32
// value up to ~1 ulp of the result) -- 4x4, so cost is irrelevant.
// ---------------------------------------------------------------------------------------------
__device__ bool invert4x4(const double* a, double* inv) {
    double m[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            m[i][j] = a[i * 4 + j];
            m[i][4 + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        double best = fabs(m[c][c]);
        for (int r = c + 1; r < 4; ++r)
            if (fabs(m[r][c]) > best) { best = fabs(m[r][c]); piv = r; }
        if (piv != c)
            for (int j = 0; j < 8; ++j) { double t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t; }
        double d = 1.0 / m[c][c];          // singular -> inf/NaN, reported through nan_flag
        for (int j = 0; j < 8; ++j) m[c][j] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            double f = m[r][c];
            for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) inv[i * 4 + j] = m[i][4 + j];
    return true;
}

__device__ void compose_one(const float* ref, const float* src, float* out12, int* nan_flag) {
    double r[16], s[16], inv[16];
    for (int i = 0; i < 16; ++i) { r[i] = (double)ref[i]; s[i] = (double)src[i]; }
    invert4x4(r, inv);
    bool bad = false;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += s[i * 4 + k] * inv[k * 4 + j];
            float f = (float)acc;
            bad |= isnan(f);
            if (j < 3) out12[i * 3 + j] = f; else out12[9 + i] = f;
        }
    }
    if (bad && nan_flag) atomicExch(nan_flag, 1);
}

__global__ void compose_kernel(const float* __restrict__ proj, int B, int V, float* __restrict__ out, int* nan_flag) {
    pdl_trigger();
    pdl_wait();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int S = V - 1;
    if (t >= B * S) return;
    int b = t / S, s = t % S;
    compose_one(proj + (size_t)(b * V) * 16, proj + (size_t)(b * V + 1 + s) * 16, out + (size_t)t * 12, nan_flag);
}

// First launch of imvs_itermvs_forward (forward.cu): the composed projections of the three levels (K1, module.py:78-90) and the
// padded copy of the level-3 pyramid in ONE launch -- blocks [0, pad_blocks) pad 256 (texel, group) items each, the blocks
// after them compose the 3 * B * S projection pairs.
__global__ void __launch_bounds__(256)
forward_prologue_kernel(const float* __restrict__ proj1, const float* __restrict__ proj2, const float* __restrict__ proj3, int B, int V,
                        float* __restrict__ rt1, float* __restrict__ rt2, float* __restrict__ rt3, int* nan_flag,
                        const float* __restrict__ fea3, float* __restrict__ fea3p, size_t texels, unsigned pad_blocks) {
    pdl_trigger();
    pdl_wait();
    if (blockIdx.x < pad_blocks) {
        const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (t < texels * 8) pad_level3_item(fea3, fea3p, t);
        return;
    }
    const int S = V - 1, n = B * S;
    const int t = (int)(blockIdx.x - pad_blocks) * (int)blockDim.x + (int)threadIdx.x;
    if (t >= 3 * n) return;
    const int lvl = t / n, r = t - lvl * n, b = r / S, s = r - b * S;
    const float* proj = lvl == 0 ? proj1 : (lvl == 1 ? proj2 : proj3);
    float* out = lvl == 0 ? rt1 : (lvl == 1 ? rt2 : rt3);
    compose_one(proj + (size_t)(b * V) * 16, proj + (size_t)(b * V + 1 + s) * 16, out + (size_t)r * 12, nan_flag);
}

int forward_prologue(const float* proj1, const float* proj2, const float* proj3, int B, int V, float* rt1, float* rt2, float* rt3,
                     int* nan_flag, const float* fea3, float* fea3p, int H3, int W3, cudaStream_t st) {
    const size_t texels = (size_t)B * V * H3 * W3;
    const size_t pad_blocks = (texels * 8 + 255) / 256;
    const int comp_blocks = cdiv(3 * B * (V - 1), 256);
    IMVS_REQUIRE(pad_blocks + comp_blocks < 2147483647ull, "forward prologue: grid too large");
    IMVS_CUDA(launch_k(forward_prologue_kernel, dim3((unsigned)(pad_blocks + comp_blocks)), dim3(256), 0, st, proj1, proj2, proj3, B, V,
                       rt1, rt2, rt3, nan_flag, fea3, fea3p, texels, (unsigned)pad_blocks));
    return 0;
}

__global__ void compose_pair_kernel(const float* __restrict__ src_proj, const float* __restrict__ ref_proj, int B,
                                    float* __restrict__ out, int* nan_flag) {
    pdl_trigger();
    pdl_wait();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    compose_one(ref_proj + (size_t)b * 16, src_proj + (size_t)b * 16, out + (size_t)b * 12, nan_flag);
}

// ---------------------------------------------------------------------------------------------
// a1: differentiable_warping in the reference's own layouts (NCHW in, [B,C,D,H,W] out).
// One thread per (b, d, y, x): sampling position once, then a loop over channels; stores are
// coalesced along x.  This is the compatibility operator -- the estimator itself uses the fused
// kernels of warpcorr.cu and never materialises this volume.
// ---------------------------------------------------------------------------------------------
__global__ void warp_nchw_kernel(const float* __restrict__ fea, const float* __restrict__ rt,
                                 const float* __restrict__ depth, float* __restrict__ out,
                                 int B, int C, int H1, int W1, int D, int H, int W) {
    pdl_trigger();
    pdl_wait();
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int bd = blockIdx.z;
    if (x >= W) return;
    int b = bd / D, d = bd % D;
    const float* P = rt + (size_t)b * 12;
    float dep = depth[((size_t)bd * H + y) * W + x];
    float sx = (float)((double)W1 / (double)W), sy = (float)((double)H1 / (double)H);
    Tap tp = project_tap(P, (float)x * sx, (float)y * sy, dep, (float)W, (float)H, W1, H1);
    float w00 = (1.f - tp.fx) * (1.f - tp.fy), w01 = tp.fx * (1.f - tp.fy);
    float w10 = (1.f - tp.fx) * tp.fy, w11 = tp.fx * tp.fy;
    const size_t plane = (size_t)H1 * W1;
    const float* base = fea + (size_t)b * C * plane;
    size_t o = (((size_t)b * C) * D + d) * (size_t)H * W + (size_t)y * W + x;
    const size_t ostride = (size_t)D * H * W;
    int i00 = tp.y0 * W1 + tp.x0;
    for (int c = 0; c < C; ++c) {
        const float* p = base + (size_t)c * plane;
        float acc = 0.f;
        if (tp.mask & 1) acc = ldg(p + i00) * w00;
        if (tp.mask & 2) acc = fmaf(ldg(p + i00 + 1), w01, acc);
        if (tp.mask & 4) acc = fmaf(ldg(p + i00 + W1), w10, acc);
        if (tp.mask & 8) acc = fmaf(ldg(p + i00 + W1 + 1), w11, acc);
        out[o + (size_t)c * ostride] = acc;
    }
}

// Backward of the operator above with respect to src_fea (grid_sample's d/d input; the grid itself carries no
// gradient, module.py:77).  Same thread mapping as the forward: the sampling position of (b, d, y, x) is computed
// once, then one scatter of four weighted atomics per channel.
__global__ void warp_nchw_backward_kernel(const float* __restrict__ gout, const float* __restrict__ rt,
                                          const float* __restrict__ depth, float* __restrict__ gfea,
                                          int B, int C, int H1, int W1, int D, int H, int W) {
    pdl_trigger();
    pdl_wait();
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int bd = blockIdx.z;
    if (x >= W) return;
    int b = bd / D, d = bd % D;
    const float* P = rt + (size_t)b * 12;
    float dep = depth[((size_t)bd * H + y) * W + x];
    float sx = (float)((double)W1 / (double)W), sy = (float)((double)H1 / (double)H);
    Tap tp = project_tap(P, (float)x * sx, (float)y * sy, dep, (float)W, (float)H, W1, H1);
    if (tp.mask == 0) return;
    float w00 = (1.f - tp.fx) * (1.f - tp.fy), w01 = tp.fx * (1.f - tp.fy);
    float w10 = (1.f - tp.fx) * tp.fy, w11 = tp.fx * tp.fy;
    const size_t plane = (size_t)H1 * W1;
    float* base = gfea + (size_t)b * C * plane;
    size_t o = (((size_t)b * C) * D + d) * (size_t)H * W + (size_t)y * W + x;
    const size_t ostride = (size_t)D * H * W;
    int i00 = tp.y0 * W1 + tp.x0;
    for (int c = 0; c < C; ++c) {
        float* p = base + (size_t)c * plane;
        const float g = ldg(gout + o + (size_t)c * ostride);
        if (tp.mask & 1) atomicAdd(p + i00, g * w00);
        if (tp.mask & 2) atomicAdd(p + i00 + 1, g * w01);
        if (tp.mask & 4) atomicAdd(p + i00 + W1, g * w10);
        if (tp.mask & 8) atomicAdd(p + i00 + W1 + 1, g * w11);
    }
}

// ---------------------------------------------------------------------------------------------
// layout helpers (tiled transpose through shared memory, 32x32 tiles over (C, H*W))
// ---------------------------------------------------------------------------------------------
__global__ void transpose_cp_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
    // in: [N][R][Cc] -> out: [N][Cc][R]
    __shared__ float tile[32][33];
    pdl_trigger();
    pdl_wait();
    int n = blockIdx.z;
    const float* src = in + (size_t)n * R * Cc;
    float* dst = out + (size_t)n * R * Cc;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < Cc) tile[i][threadIdx.x] = src[(size_t)r * Cc + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < Cc) dst[(size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

static int transpose_launch(const float* in, float* out, int N, int R, int Cc, cudaStream_t st) {
    dim3 grid(cdiv(Cc, 32), cdiv(R, 32), N), block(32, 8);
    IMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "transpose: grid too large (R=%d N=%d)", R, N);
    IMVS_CUDA(launch_k(transpose_cp_kernel, grid, block, 0, st, in, out, R, Cc));
    return 0;
}

}  // namespace imvs

using namespace imvs;

extern "C" int imvs_abi_version(void) { return IMVS_ABI_VERSION; }
extern "C" const char* imvs_last_error(void) { return err_buf(); }
extern "C" long long imvs_launches_total(void) { return launches_total(); }
extern "C" int imvs_set_conv_passes(int passes) {
    IMVS_REQUIRE(passes == 1 || passes == 3 || passes == 4,
                 "set_conv_passes: mode must be 1 (TF32), 3 (3xTF32, fp32-grade) or 4 (3xFP16 split, fp32-grade), got %d", passes);
    g_conv_passes = passes;
    return 0;
}
extern "C" int imvs_get_conv_passes(void) { return g_conv_passes; }
extern "C" int imvs_set_tcgen05(int enabled) { g_tc5 = enabled ? 1 : 0; return 0; }
extern "C" int imvs_tc5_debug_clocks(int on, long long* out16) {   // debug: phase stamps of CTA 0 of the last tcgen05 launch
    g_tc5_clk_on = on;
    if (out16) {
        if (cudaDeviceSynchronize() != cudaSuccess) return 2;
        if (cudaMemcpyFromSymbol(out16, g_tc5_clk, sizeof(long long) * 16) != cudaSuccess) return 2;
    }
    return 0;
}
extern "C" int imvs_tcgen05_status(void) {
    int v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return 4;
    if (cudaMemcpyFromSymbol(&v, g_tc5_err, sizeof(int)) != cudaSuccess) return 4;
    return v & 1;
}
extern "C" int imvs_device_status(int clear) {
    int v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return 4;
    if (cudaMemcpyFromSymbol(&v, g_tc5_err, sizeof(int)) != cudaSuccess) return 4;
    if (clear && v) {
        const int zero = 0;
        if (cudaMemcpyToSymbol(g_tc5_err, &zero, sizeof(int)) != cudaSuccess) return 4;
    }
    return v;
}

// Profiling facility (NOT graph-capturable, synchronises in _end): records one CUDA-event pair
// around every stage of imvs_itermvs_forward / imvs_featurenet_forward issued between begin and end.
extern "C" int imvs_profile_begin(int capacity) {
    IMVS_REQUIRE(!g_prof.active, "profile_begin: already active");
    IMVS_REQUIRE(capacity > 0 && capacity <= 65536, "profile_begin: bad capacity");
    g_prof.ev = new cudaEvent_t[2 * capacity];
    g_prof.tags = new int[capacity];
    for (int i = 0; i < 2 * capacity; ++i) IMVS_CUDA(cudaEventCreate(&g_prof.ev[i]));
    g_prof.cap = capacity; g_prof.n = 0; g_prof.active = true;
    return 0;
}
extern "C" int imvs_profile_end(float* ms_out, int* tags_out, int capacity) {
    IMVS_REQUIRE(g_prof.active, "profile_end: not active");
    g_prof.active = false;
    int n = g_prof.n < capacity ? g_prof.n : capacity;
    if (g_prof.n > 0) IMVS_CUDA(cudaEventSynchronize(g_prof.ev[2 * g_prof.n - 1]));
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        IMVS_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
        ms_out[i] = ms; tags_out[i] = g_prof.tags[i];
    }
    for (int i = 0; i < 2 * g_prof.cap; ++i) cudaEventDestroy(g_prof.ev[i]);
    delete[] g_prof.ev; delete[] g_prof.tags;
    g_prof.ev = nullptr; g_prof.tags = nullptr; g_prof.cap = 0;
    int total = g_prof.n; g_prof.n = 0;
    return -total;   // negative count = success with `total` records (0 -> none); positive = error
}

extern "C" int imvs_compose_projections(const float* proj, int B, int V, float* out, int* nan_flag, void* stream) {
    IMVS_REQUIRE(proj && out, "compose_projections: null pointer");
    IMVS_REQUIRE(B >= 1 && V >= 2, "compose_projections: need B>=1 and at least one source view (B=%d V=%d)", B, V);
    int n = B * (V - 1);
    ApiScope api_;
    IMVS_CUDA(launch_k(compose_kernel, dim3(cdiv(n, 64)), dim3(64), 0, (cudaStream_t)stream, proj, B, V, out, nan_flag));
    return 0;
}

extern "C" int imvs_differentiable_warping(const float* src_fea, const float* src_proj, const float* ref_proj,
                                           const float* depth_samples, float* out, int B, int C, int H1, int W1,
                                           int D, int H, int W, float* rt_scratch, int* nan_flag, void* stream) {
    IMVS_REQUIRE(src_fea && src_proj && ref_proj && depth_samples && out && rt_scratch,
                 "differentiable_warping: null pointer");
    IMVS_REQUIRE(B >= 1 && C >= 1 && H1 >= 1 && W1 >= 1 && D >= 1 && H >= 1 && W >= 1,
                 "differentiable_warping: bad shape");
    IMVS_REQUIRE(H <= 65535 && (long long)B * D <= 65535, "differentiable_warping: H or B*D exceeds grid limits");
    cudaStream_t st = (cudaStream_t)stream;
    float* rt = rt_scratch;
    ApiScope api_;
    IMVS_CUDA(launch_k(compose_pair_kernel, dim3(cdiv(B, 32)), dim3(32), 0, st, src_proj, ref_proj, B, rt, nan_flag));
    dim3 grid(cdiv(W, 128), H, B * D);
    IMVS_CUDA(launch_k(warp_nchw_kernel, grid, dim3(128), 0, st, src_fea, (const float*)rt, depth_samples, out, B, C, H1, W1, D, H, W));
    return 0;
}

extern "C" int imvs_differentiable_warping_backward(const float* grad_out, const float* src_proj, const float* ref_proj,
                                                    const float* depth_samples, float* grad_src_fea, int B, int C, int H1,
                                                    int W1, int D, int H, int W, float* rt_scratch, int* nan_flag, void* stream) {
    IMVS_REQUIRE(grad_out && src_proj && ref_proj && depth_samples && grad_src_fea && rt_scratch,
                 "differentiable_warping_backward: null pointer");
    IMVS_REQUIRE(B >= 1 && C >= 1 && H1 >= 1 && W1 >= 1 && D >= 1 && H >= 1 && W >= 1,
                 "differentiable_warping_backward: bad shape");
    IMVS_REQUIRE(H <= 65535 && (long long)B * D <= 65535, "differentiable_warping_backward: H or B*D exceeds grid limits");
    cudaStream_t st = (cudaStream_t)stream;
    ApiScope api_;
    IMVS_CUDA(cudaMemsetAsync(grad_src_fea, 0, sizeof(float) * (size_t)B * C * H1 * W1, st));
    IMVS_CUDA(launch_k(compose_pair_kernel, dim3(cdiv(B, 32)), dim3(32), 0, st, src_proj, ref_proj, B, rt_scratch, nan_flag));
    dim3 grid(cdiv(W, 128), H, B * D);
    IMVS_CUDA(launch_k(warp_nchw_backward_kernel, grid, dim3(128), 0, st, grad_out, (const float*)rt_scratch, depth_samples,
                       grad_src_fea, B, C, H1, W1, D, H, W));
    return 0;
}

extern "C" int imvs_nchw_to_nhwc(const float* in, float* out, int N, int C, int H, int W, void* stream) {
    IMVS_REQUIRE(in && out && N >= 1 && C >= 1 && H >= 1 && W >= 1, "nchw_to_nhwc: bad argument");
    ApiScope api_;
    return transpose_launch(in, out, N, C, H * W, (cudaStream_t)stream);   // [C][P] -> [P][C]
}

extern "C" int imvs_nhwc_to_nchw(const float* in, float* out, int N, int C, int H, int W, void* stream) {
    IMVS_REQUIRE(in && out && N >= 1 && C >= 1 && H >= 1 && W >= 1, "nhwc_to_nchw: bad argument");
    ApiScope api_;
    return transpose_launch(in, out, N, H * W, C, (cudaStream_t)stream);   // [P][C] -> [C][P]
}
