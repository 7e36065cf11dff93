// Shared helpers for the sm_100a kernels of itermvs_b200 (host-side error plumbing + device utils).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>

#include "../../include/itermvs_b200.h"

namespace imvs {

// ---- thread-local error text -------------------------------------------------------------
char* err_buf();
int fail(const char* fmt, ...);

#define IMVS_REQUIRE(cond, ...)                                 \
    do {                                                        \
        if (!(cond)) return ::imvs::fail(__VA_ARGS__);          \
    } while (0)

#define IMVS_CUDA(expr)                                                                       \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return ::imvs::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define IMVS_LAUNCH_CHECK(name)                                                               \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess)                                                               \
            return ::imvs::fail("launch of %s failed: %s", name, cudaGetErrorString(e__));     \
    } while (0)

#define IMVS_TRY(expr)              \
    do {                            \
        int rc__ = (expr);          \
        if (rc__ != 0) return rc__; \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// launch counter (bench.py reports gpu_launches); bumped by every host-side launch helper
void count_launch(int n = 1);
long long launches_total();

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// A forward pass is ~95 short kernels (20 us on average) in one stream / one CUDA graph.  Every kernel
// starts with pdl_trigger() (lets the NEXT kernel's CTAs become resident as soon as this grid's last CTAs
// have started) and executes pdl_wait() before its first access to memory another kernel produces (blocks
// until the PREVIOUS grid has completed and flushed), so launch latency, CTA ramp-up and the
// producer-independent prologue (weight staging, index arithmetic) overlap the predecessor's tail.
// Host side: launch_k() adds cudaLaunchAttributeProgrammaticStreamSerialization except for the first
// launch of an outermost C-ABI call (its predecessor in the stream may be a memcpy / memset / foreign
// kernel).  IMVS_PDL=0 in the environment (read once) disables the attribute; the device-side
// instructions are then no-ops.
bool pdl_enabled();
bool* pdl_armed();                 // thread-local: a kernel of this API call has already been launched
int* api_depth();                  // thread-local nesting depth of extern "C" entry points
struct ApiScope {
    ApiScope() { if ((*api_depth())++ == 0) *pdl_armed() = false; }
    ~ApiScope() { --*api_depth(); }
};

template <class... KArgs, class... Args>
cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (*pdl_armed() && pdl_enabled()) ? 1 : 0;
    *pdl_armed() = true;
    count_launch();
    return cudaLaunchKernelEx(&cfg, kern, KArgs(static_cast<Args&&>(args))...);
}

// ---- optional per-stage CUDA-event timing (bench.py's roofline measurement; off by default) ----
enum StageTag { ST_COMPOSE = 0, ST_WARPCORR_INIT, ST_PVW, ST_AGG_INIT, ST_CORRNET, ST_HIDDEN_INIT, ST_HEAD,
                ST_WARPCORR_ITER, ST_GRU, ST_UPSAMPLE, ST_FEATURENET, ST_COUNT };
bool profile_active();
void profile_mark(int tag, cudaStream_t st, bool is_start);
struct StageTimer {
    int tag; cudaStream_t st; bool on;
    StageTimer(int t, void* s) : tag(t), st((cudaStream_t)s), on(profile_active()) { if (on) profile_mark(tag, st, true); }
    ~StageTimer() { if (on) profile_mark(tag, st, false); }
};

// opt in to > 48 KB dynamic shared memory once per (kernel, device); safe to call every launch
template <class K>
int ensure_dynamic_smem(K kern, size_t smem, int* done_mask) {
    if (smem <= 48 * 1024) return 0;
    int dev = 0;
    IMVS_CUDA(cudaGetDevice(&dev));
    if ((*done_mask >> (dev & 31)) & 1) return 0;
    IMVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    *done_mask |= 1 << (dev & 31);
    return 0;
}

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float ldg(const float* p) { return __ldg(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// one (texel, correlation group g) of the padded level-3 pyramid (include/itermvs_b200.h: imvs_pad_level3): channels 6g..6g+5 of the
// 48-float texel go to floats 4g..4g+3 and 32+4g, 32+4g+1 of the 64-float texel; floats 32+4g+2, +3 are zero
__device__ __forceinline__ void pad_level3_item(const float* __restrict__ src, float* __restrict__ dst, size_t t) {
    const size_t tx = t >> 3;
    const int g = (int)(t & 7);
    const float* s = src + tx * 48 + 6 * g;
    const float2 a = ldg2(s), b = ldg2(s + 2), c = ldg2(s + 4);
    float* d = dst + tx * 64 + 4 * g;
    *reinterpret_cast<float4*>(d) = make_float4(a.x, a.y, b.x, b.y);
    *reinterpret_cast<float4*>(d + 32) = make_float4(c.x, c.y, 0.f, 0.f);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// F.interpolate(scale_factor=f, mode='bilinear', align_corners=False) source index (ATen
// area_pixel_compute_source_index): src = (dst + 0.5) / f - 0.5, clamped below at 0.
__device__ __forceinline__ void up_index(int dst, float inv_scale, int n, int& i0, int& i1, float& lam) {
    float src = inv_scale * (dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    i1 = i0 + (i0 < n - 1 ? 1 : 0);
    lam = src - (float)i0;
}

// inverse-depth parametrisation, module.py:148-152
__device__ __forceinline__ float unnormalize_depth(float nd, float inv_min, float inv_max) {
    return 1.0f / (inv_max + nd * (inv_min - inv_max));
}

}  // namespace imvs
