// cusim -- a minimal CPU emulation of the CUDA execution model, for TESTS ONLY.
//
// Purpose: compile the *.cu sources of itermvs_b200/csrc that contain no tensor-core / PTX data-path
// instructions (warp.cu, warpcorr.cu, fusion.cu) with g++ and run their kernels thread by thread on the CPU,
// so that the `-m "not gpu"` suite exercises the real kernel source (index arithmetic, shared-memory
// protocols, warp shuffles, persistent-block work queues) against the oracle without a GPU.
//
// This is test infrastructure: it lives under tests/, is built into tests/cusim/_build/ only by
// tests/cusim/cusim_build.py, and nothing in the itermvs_b200 package ever loads it -- the product path has no CPU
// fallback (itermvs_b200/_lib.py raises LibraryMissing).  It is slow (one ucontext fiber per CUDA thread) and
// proves nothing about performance, memory-model races across warps, or alignment faults beyond the
// explicit checks in __ldg.
//
// Model: blocks run one after the other; the threads of a block are fibers scheduled round-robin; a fiber
// runs until it reaches __syncthreads / __syncwarp / a warp shuffle (or exits).  That is a legal schedule
// of a data-race-free CUDA program that only uses full-mask warp primitives convergently -- which is what
// the simulated sources do.
//
// This header shadows <cuda_runtime.h> (tests/cusim is first on the include path of the simulated build).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <utility>

#define CUSIM 1

// ---- qualifiers ------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static               // block-static storage; `extern __shared__` is rewritten by build.py
#define __constant__ static
#define CUSIM_ASM(...) ((void)0)       // build.py rewrites `asm volatile(` / `asm(` to this

// ---- vector types ----------------------------------------------------------------------------------
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct double2 { double x, y; };
struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

// ---- the runtime (tests/cusim/cusim.cpp) -----------------------------------------------------------
namespace cusim {
struct Thread {
    uint3 tid;
    int warp, lane;
};
extern Thread* cur;
void block_barrier();
void warp_barrier();
uint32_t shfl(uint32_t v, int src_lane);
void* dyn_smem();
uint32_t (*warp_xchg())[12];           // the current warp's [32][12]-word exchange area (tensor-core emulation)
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
[[noreturn]] void die(const char* what);
// optional access-pattern tracer (CUSIM_TRACE=<json path>): groups the __ldg calls of a warp into warp-level requests
// and counts requests / 32-byte sectors / 128-byte lines per launch -- tools/wavefront_model.py
extern bool g_trace;
void trace_load(const void* p, int bytes, const void* site);
}  // namespace cusim

// the built-in variables: plain globals, rewritten by the scheduler at every fiber switch
extern uint3 threadIdx;
extern dim3 blockIdx, blockDim, gridDim;
#define warpSize 32

static inline void __syncthreads() { cusim::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cusim::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T>
static inline T cusim_shfl_(T v, int src) {
    static_assert(sizeof(T) == 4, "cusim: 32-bit shuffles only");
    uint32_t u;
    std::memcpy(&u, &v, 4);
    u = cusim::shfl(u, src);
    std::memcpy(&v, &u, 4);
    return v;
}
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    if (mask != 0xffffffffu) cusim::die("__shfl_sync with a partial mask");
    const int lane = cusim::cur->lane;
    return cusim_shfl_(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
    if (mask != 0xffffffffu) cusim::die("__shfl_xor_sync with a partial mask");
    (void)width;
    return cusim_shfl_(v, cusim::cur->lane ^ lane_mask);
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    if (mask != 0xffffffffu) cusim::die("__shfl_down_sync with a partial mask");
    const int lane = cusim::cur->lane;
    const int src = lane + (int)delta;
    return cusim_shfl_(v, (src & ~(width - 1)) == (lane & ~(width - 1)) ? src : lane);
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    if (mask != 0xffffffffu) cusim::die("__shfl_up_sync with a partial mask");
    const int lane = cusim::cur->lane;
    const int src = lane - (int)delta;
    return cusim_shfl_(v, src >= (lane & ~(width - 1)) ? src : lane);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    if (mask != 0xffffffffu) cusim::die("__ballot_sync with a partial mask");
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (cusim_shfl_<unsigned>(pred ? 1u : 0u, l) & 1u) << l;
    return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __reduce_min_sync(unsigned mask, int v) {
    if (mask != 0xffffffffu) cusim::die("__reduce_min_sync with a partial mask");
    int r = v;
    for (int l = 0; l < 32; ++l) { const int o = cusim_shfl_<int>(v, l); r = o < r ? o : r; }
    return r;
}
static inline int __reduce_max_sync(unsigned mask, int v) {
    if (mask != 0xffffffffu) cusim::die("__reduce_max_sync with a partial mask");
    int r = v;
    for (int l = 0; l < 32; ++l) { const int o = cusim_shfl_<int>(v, l); r = o > r ? o : r; }
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }

// ---- memory access ---------------------------------------------------------------------------------
template <class T>
static __attribute__((noinline)) T __ldg(const T* p) {       // noinline: the return address identifies the static load site
    if (reinterpret_cast<uintptr_t>(p) % alignof(T) != 0 || reinterpret_cast<uintptr_t>(p) % sizeof(T) != 0)
        cusim::die("misaligned __ldg");
    if (cusim::g_trace) cusim::trace_load(p, (int)sizeof(T), __builtin_extract_return_addr(__builtin_return_address(0)));
    return *p;
}
// fibers never pre-empt each other between sync points: plain read-modify-write is atomic here
template <class T>
static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline float4 atomicAdd(float4* p, float4 v) {          // red.global.add.v4.f32 (sm_90+): 16-byte aligned
    if (reinterpret_cast<uintptr_t>(p) % 16 != 0) cusim::die("misaligned vector atomicAdd");
    float4 o = *p;
    p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w;
    return o;
}
static inline unsigned atomicAdd(unsigned* p, int v) { unsigned o = *p; *p = o + (unsigned)v; return o; }
template <class T>
static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T>
static inline T atomicMax(T* p, T v) { T o = *p; *p = o > v ? o : v; return o; }
template <class T>
static inline T atomicMin(T* p, T v) { T o = *p; *p = o < v ? o : v; return o; }
template <class T>
static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }

// ---- math ------------------------------------------------------------------------------------------
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float2int_rn(float f) { return (int)lrintf(f); }          // round half to even (default mode)
static inline int __float2int_rd(float f) { return (int)floorf(f); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float cusim_expf(float a) { return expf(a); }
#define __expf cusim_expf
static inline float __saturatef(float a) { return fminf(fmaxf(a, 0.f), 1.f); }
using std::isnan;
using std::isinf;

// ---- tensor-core / async-copy emulation used by the CUSIM branches of csrc/mmaconv.cuh -----------------------------
// Fragment layouts as in the PTX ISA (mma.sync.aligned.m16n8k8 / m16n8k16, .row.col): g = lane >> 2, t = lane & 3;
// C/D: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
namespace cusim {
static inline float tf32_value(uint32_t u) { u &= 0xffffe000u; float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t cvt_rna_tf32(float x) {          // round to nearest, ties away from zero, 10-bit mantissa
    uint32_t u; std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return u;
    return (u + 0x1000u) & 0xffffe000u;
}
static inline float half_bits_to_float(uint16_t h) { _Float16 v; std::memcpy(&v, &h, 2); return (float)v; }
static inline uint16_t float_to_half_bits(float f) {     // cvt.rn.satfinite.f16.f32
    f = f > 65504.f ? 65504.f : (f < -65504.f ? -65504.f : f);
    _Float16 v = (_Float16)f; uint16_t h; std::memcpy(&h, &v, 2); return h;
}
static inline uint32_t cvt_f16x2(float upper, float lower) { return ((uint32_t)float_to_half_bits(upper) << 16) | float_to_half_bits(lower); }
static inline float2 unpack_f16x2(uint32_t u) { return float2{half_bits_to_float((uint16_t)(u & 0xffffu)), half_bits_to_float((uint16_t)(u >> 16))}; }
static inline float half_of(uint32_t reg, int k) { return half_bits_to_float((uint16_t)((k & 1) ? (reg >> 16) : (reg & 0xffffu))); }

// A: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B: b0 (k = t, n = g) b1 (k = t+4, n = g)
static inline void mma_m16n8k8_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    uint32_t (*x)[12] = warp_xchg();
    const int lane = cur->lane, g = lane >> 2, t = lane & 3;
    x[lane][0] = a0; x[lane][1] = a1; x[lane][2] = a2; x[lane][3] = a3; x[lane][4] = b0; x[lane][5] = b1;
    warp_barrier();
    for (int i = 0; i < 4; ++i) {
        const int row = g + (i >= 2 ? 8 : 0), col = 2 * t + (i & 1);
        float acc = c[i];
        for (int k = 0; k < 8; ++k)
            acc = fmaf(tf32_value(x[(row & 7) * 4 + (k & 3)][(row >= 8) + 2 * (k >= 4)]), tf32_value(x[col * 4 + (k & 3)][4 + (k >= 4)]), acc);
        c[i] = acc;
    }
    warp_barrier();
}
// A: a0 (g; k = 2t, 2t+1) a1 (g+8; same) a2 (g; 2t+8, 2t+9) a3 (g+8; same);  B: b0 (k = 2t, 2t+1; n = g) b1 (k = 2t+8, 2t+9)
// Every lane first unpacks its own fragments to floats: slot [lane][2 * reg + half].
static inline void mma_m16n8k16_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    float (*x)[12] = reinterpret_cast<float (*)[12]>(warp_xchg());
    const int lane = cur->lane, g = lane >> 2, t = lane & 3;
    const uint32_t regs[6] = {a0, a1, a2, a3, b0, b1};
    for (int r = 0; r < 6; ++r) { x[lane][2 * r] = half_of(regs[r], 0); x[lane][2 * r + 1] = half_of(regs[r], 1); }
    warp_barrier();
    for (int i = 0; i < 4; ++i) {
        const int row = g + (i >= 2 ? 8 : 0), col = 2 * t + (i & 1);
        float acc = c[i];
        for (int k = 0; k < 16; ++k)
            acc = fmaf(x[(row & 7) * 4 + ((k & 7) >> 1)][2 * ((row >= 8) + 2 * (k >= 8)) + (k & 1)],
                       x[col * 4 + ((k & 7) >> 1)][2 * (4 + (k >= 8)) + (k & 1)], acc);
        c[i] = acc;
    }
    warp_barrier();
}
// A: a0 (g; k = 2t, 2t+1) a1 (g+8; same);  B: b0 (k = 2t, 2t+1; n = g)
static inline void mma_m16n8k8_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    float (*x)[12] = reinterpret_cast<float (*)[12]>(warp_xchg());
    const int lane = cur->lane, g = lane >> 2, t = lane & 3;
    x[lane][0] = half_of(a0, 0); x[lane][1] = half_of(a0, 1); x[lane][2] = half_of(a1, 0); x[lane][3] = half_of(a1, 1);
    x[lane][8] = half_of(b0, 0); x[lane][9] = half_of(b0, 1);
    warp_barrier();
    for (int i = 0; i < 4; ++i) {
        const int row = g + (i >= 2 ? 8 : 0), col = 2 * t + (i & 1);
        float acc = c[i];
        for (int k = 0; k < 8; ++k)
            acc = fmaf(x[(row & 7) * 4 + (k >> 1)][2 * (row >= 8) + (k & 1)], x[col * 4 + (k >> 1)][8 + (k & 1)], acc);
        c[i] = acc;
    }
    warp_barrier();
}
static inline void cp_async16(void* smem, const void* gmem, bool valid) {
    if (reinterpret_cast<uintptr_t>(smem) % 16 != 0 || (valid && reinterpret_cast<uintptr_t>(gmem) % 16 != 0)) die("misaligned cp.async 16");
    if (valid) std::memcpy(smem, gmem, 16); else std::memset(smem, 0, 16);
}
}  // namespace cusim
static inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) {}

// ---- host API stubs --------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t {
    dim3 gridDim, blockDim;
    size_t dynamicSmemBytes;
    cudaStream_t stream;
    cudaLaunchAttribute* attrs;
    unsigned numAttrs;
};
static inline const char* cudaGetErrorString(cudaError_t) { return "cusim error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
// CUSIM_SMS in the environment = number of "SMs" the persistent kernels see (default 3: a ragged multi-block grid)
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
    const char* e = getenv("CUSIM_SMS");
    *v = e ? atoi(e) : 3;
    return cudaSuccess;
}
template <class K>
static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
template <class T>
static inline cudaError_t cudaGetSymbolAddress(void** p, T& sym) { *p = (void*)&sym; return cudaSuccess; }
template <class T>
static inline cudaError_t cudaMemcpyFromSymbol(void* dst, const T& sym, size_t n) { std::memcpy(dst, &sym, n); return cudaSuccess; }
template <class T>
static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) { std::memcpy((void*)&sym, src, n); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

template <class... KArgs, class... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kern)(KArgs...), Args&&... args) {
    std::tuple<KArgs...> held(static_cast<Args&&>(args)...);
    cusim::run_grid(cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, [&] { std::apply(kern, held); });
    return cudaSuccess;
}
