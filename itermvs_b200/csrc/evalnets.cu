// The small convolutional nets of the evaluation stage and hidden-state initialisation:
//   CorrNet (itermvs.py:352-381), PixelViewWeight (itermvs.py:333-350, 53-57), hidden_init
//   (itermvs.py:153-164) -- all on the tensor-core implicit-GEMM convolution of mmaconv.cuh,
//   activations channels-last.
#include "common.cuh"
#include "mmaconv.cuh"
#include "corrnet_tile.cuh"
#ifndef CUSIM
#include "tc5pconv.cuh"
#endif

namespace imvs {

// ------------------------------------------------------------------------------- CorrNet ----
struct EpiCorrOut {          // conv5 (1 valid cout): + bias, scatter to out[(n/period)*bstride + p*pstride + n%period]
    float* out;
    const float* bias[3];
    int period, split1, split2;
    size_t bstride, pstride;
    int H, W;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int, int t, const float (&v)[2 * NT], int) const {
        if (t != 0 || oy >= H || ox >= W) return;
        const int r = n % period;
        const float* b = r < split1 ? bias[0] : (r < split2 ? bias[1] : bias[2]);
        out[(size_t)(n / period) * bstride + ((size_t)oy * W + ox) * pstride + r] = v[0] + ldg(b);
    }
};

#ifndef CUSIM     // (the CPU emulation of the test-suite has no TMA / tcgen05 model)
// ---- CorrNet on the persistent TMA + tcgen05 kernel (tc5pconv.cuh) ------------------------------------------------------
// transposed convolution + U-Net skip (itermvs.py:374-377): thread = (input pixel, output parity); H, W = INPUT grid
struct EpiTconvP {
    static constexpr int kAhead = 1;
    __device__ __forceinline__ bool wants_prefetch() const { return true; }
    tc5p::Split out;         // [N][kco][2H][2W][8]
    tc5p::Split skip;        // same shape
    int H, W, kco;
    template <int NCH> struct Pre { uint4 h[NCH / 8], l[NCH / 8]; };
    __device__ __forceinline__ size_t index(int n, int iy, int ix, int c0, int NB, int j) const {
        const int par = c0 / NB, oy = 2 * iy + (par >> 1), ox = 2 * ix + (par & 1);
        return (((size_t)n * kco + j) * (2 * H) + oy) * (size_t)(2 * W) + ox;
    }
    template <int NB, int NCH>
    __device__ __forceinline__ void prefetch(int n, int iy, int ix, int c0, Pre<NCH>& p) const {
#pragma unroll
        for (int j = 0; j < NCH / 8; ++j) {
            if (j >= kco) break;
            const size_t idx = index(n, iy, ix, c0, NB, j);
            p.h[j] = __ldcg(reinterpret_cast<const uint4*>(skip.hi) + idx);     // .cg: in the fused kernel the skip tensor was written by other
            p.l[j] = __ldcg(reinterpret_cast<const uint4*>(skip.lo) + idx);     // CTAs earlier in the same launch (never through the read-only path)
        }
    }
    template <int NB, int NCH>
    __device__ __forceinline__ void store(int n, int iy, int ix, int c0, float (&v)[NCH], const Pre<NCH>& p, int* status) const {
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < NCH / 8; ++j) {
            if (j >= kco) break;
            float* x = v + 8 * j;
            const uint32_t hh[4] = {p.h[j].x, p.h[j].y, p.h[j].z, p.h[j].w}, ll[4] = {p.l[j].x, p.l[j].y, p.l[j].z, p.l[j].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hh[q]));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&ll[q]));
                x[2 * q] += a.x + b.x;
                x[2 * q + 1] += a.y + b.y;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) amax = fmaxf(amax, fabsf(x[q]));
            uint4 h, l;
            split_f16(make_float2(x[0], x[1]), h.x, l.x);
            split_f16(make_float2(x[2], x[3]), h.y, l.y);
            split_f16(make_float2(x[4], x[5]), h.z, l.z);
            split_f16(make_float2(x[6], x[7]), h.w, l.w);
            const size_t idx = index(n, iy, ix, c0, NB, j);
            reinterpret_cast<uint4*>(out.hi)[idx] = h;
            reinterpret_cast<uint4*>(out.lo)[idx] = l;
        }
        if (!(amax <= 65504.f) && status) atomicOr(status, 2);
    }
};

struct EpiCorrOutP {         // EpiCorrOut for the tcgen05 kernel: channel 0 + bias -> out[(n/period)*bstride + p*pstride + n%period]
    static constexpr int kAhead = 1;
    __device__ __forceinline__ bool wants_prefetch() const { return false; }
    float* out;
    const float* bias[3];
    int period, split1, split2;
    size_t bstride, pstride;
    int H, W;
    template <int NCH> struct Pre {};
    template <int NB, int NCH>
    __device__ __forceinline__ void prefetch(int, int, int, int, Pre<NCH>&) const {}
    template <int NB, int NCH>
    __device__ __forceinline__ void store(int n, int oy, int ox, int c0, float (&v)[NCH], const Pre<NCH>&, int*) const {
        if (c0 != 0) return;
        const int r = n % period;
        const float* b = r < split1 ? bias[0] : (r < split2 ? bias[1] : bias[2]);
        out[(size_t)(n / period) * bstride + ((size_t)oy * W + ox) * pstride + r] = v[0] + ldg(b);
    }
};

static tc5p::WSel wsel_of(const imvs_corrnet_weights* sets, int period, int split1, int split2, int which) {
    tc5p::WSel s;
    for (int i = 0; i < 3; ++i) {
        const imvs_corrnet_weights& c = sets[i];
        const imvs_wpair& p = which == 0 ? c.conv0 : which == 1 ? c.conv1 : which == 2 ? c.conv2 : which == 3 ? c.conv3 : which == 4 ? c.conv4 : c.conv5;
        s.w[i] = p.f16ummai;
    }
    s.nsets = period == 1 ? 1 : 3;
    s.period = period; s.split1 = split1; s.split2 = split2;
    return s;
}

// ---- the whole CorrNet pass as ONE persistent cooperative launch ---------------------------------------------------------
// Seven phases -- operand split of the input volume, conv0, conv1, conv2, conv3^T, conv4^T, conv5 -- each the layer body of
// tc5pconv.cuh over all CTAs, separated by grid-wide barriers.  What it removes is not work but the six launch gaps, TMEM
// allocations and pipeline fills / drains of the per-layer launches (~7 us each against 1..5 us of work on these small maps).
// Launched with cudaLaunchAttributeCooperative: all CTAs are co-resident, so the barriers cannot deadlock against other
// streams' kernels.  Data written in one phase and read in a later one travels through L2 only (st.global -> membar.gl ->
// barrier -> TMA loads / ld.global.cg); `counter` must be zero at launch (a memset node precedes the kernel).
struct CorrFusedParams {
    CUtensorMap m[12];           // layer L: m[2L] hi plane, m[2L + 1] lo plane of its input
    tc5p::Epi e0, e1, e2;
    EpiTconvP e3, e4;
    EpiCorrOutP e5;
    tc5p::WSel w[6];
    tc5p::Geo g[6];
    const float* vol;            // [N][HW][8] fp32
    __half* vs_hi;
    __half* vs_lo;
    unsigned long long npix;
    unsigned int* counter;
    int* err_flag;
};

constexpr int CF_THREADS = 576;          // the widest layer: 2 + 16 warps
static_assert(sizeof(CorrFusedParams) <= 4000, "kernel parameter space");

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target, int* err_flag) {
    __threadfence();                     // this thread's global writes are visible device-wide before the arrival below
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned int seen = 0;
        int spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            if (seen >= target) break;
            __nanosleep(32);
        } while (++spins < (1 << 22));
        if (seen < target && err_flag) atomicOr(err_flag, 1);
    }
    __syncthreads();
    asm volatile("fence.proxy.async.global;" ::: "memory");      // later TMA (async proxy) reads are ordered after the barrier
}

__global__ void __launch_bounds__(CF_THREADS, 1) corrnet_fused_kernel(const __grid_constant__ CorrFusedParams P) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem_raw = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem_raw + tc5p::SMEM_HEAD_BYTES - 16);
    const int tid = threadIdx.x;
    if (tid < 32) tc5p::tmem_alloc(tc5p::smem_u32(sTmem), 256);
    if (tid == 32) {                      // the layer bodies invalidate and re-initialise these between phases
        for (int b = 0; b < 2 * tc5p::MAX_STAGES + 5; ++b) tc5p::mbar_init(tc5p::smem_u32(smem_raw) + 8u * b, 1);
        tc5p::fence_mbar_init();
    }
    tc5p::fence_before_sync();
    __syncthreads();
    tc5p::fence_after_sync();
    const uint32_t tmem_d = *sTmem;
    // phase 0: the aggregated correlation volume -> fp16 hi / lo split planes (one 8-channel chunk per pixel)
    for (unsigned long long i = (unsigned long long)blockIdx.x * CF_THREADS + tid; i < P.npix; i += (unsigned long long)gridDim.x * CF_THREADS) {
        const float4 a = ldg4(P.vol + i * 8), b = ldg4(P.vol + i * 8 + 4);
        uint4 h, l;
        split_f16(make_float2(a.x, a.y), h.x, l.x);
        split_f16(make_float2(a.z, a.w), h.y, l.y);
        split_f16(make_float2(b.x, b.y), h.z, l.z);
        split_f16(make_float2(b.z, b.w), h.w, l.w);
        reinterpret_cast<uint4*>(P.vs_hi)[i] = h;
        reinterpret_cast<uint4*>(P.vs_lo)[i] = l;
    }
    unsigned int target = gridDim.x;
    grid_barrier(P.counter, target, P.err_flag); target += gridDim.x;
    tc5p::layer_body<8, 16, 2, 1, 3, 1, false>(P.m[0], P.m[1], P.e0, P.w[0], P.g[0], P.err_flag, smem_raw, tmem_d);
    grid_barrier(P.counter, target, P.err_flag); target += gridDim.x;
    tc5p::layer_body<8, 16, 1, 1, 3, 2, false>(P.m[2], P.m[3], P.e1, P.w[1], P.g[1], P.err_flag, smem_raw, tmem_d);
    grid_barrier(P.counter, target, P.err_flag); target += gridDim.x;
    tc5p::layer_body<16, 32, 1, 1, 3, 2, false>(P.m[4], P.m[5], P.e2, P.w[2], P.g[2], P.err_flag, smem_raw, tmem_d);
    grid_barrier(P.counter, target, P.err_flag); target += gridDim.x;
    tc5p::layer_body<32, 16, 1, 1, 3, 0, false>(P.m[6], P.m[7], P.e3, P.w[3], P.g[3], P.err_flag, smem_raw, tmem_d);
    grid_barrier(P.counter, target, P.err_flag); target += gridDim.x;
    tc5p::layer_body<16, 16, 1, 1, 3, 0, false>(P.m[8], P.m[9], P.e4, P.w[4], P.g[4], P.err_flag, smem_raw, tmem_d);
    grid_barrier(P.counter, target, P.err_flag);
    tc5p::layer_body<8, 16, 2, 1, 3, 1, false>(P.m[10], P.m[11], P.e5, P.w[5], P.g[5], P.err_flag, smem_raw, tmem_d);
    if (tid < 32) tc5p::tmem_dealloc(tmem_d, 256);
}

static bool corrnet_tc5p_ready(const imvs_corrnet_weights* sets) {
    // OFF by default: measured on B200 (gpurun call r2c22) the seven tcgen05 launches of a pass take 75 us against 66 us for the
    // six mma.sync launches -- every persistent launch has a ~7 us floor (TMEM allocation, barrier and weight staging, first
    // TMA round trip, drain) that these 160..960-tile layers cannot amortise.  IMVS_TUNE_TC5P_CORR=1 selects it (parity-tested).
    if (conv_passes() != 4 || !(tune("TC5P_CORR", 0) || tune("CORR_FUSED", 0)) || !tc5p::encode_tiled_fn()) return false;
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 6; ++k)
            if (!wsel_of(sets, 1, 1, 1, k).w[i]) return false;
    return true;
}

#endif  // !CUSIM

static WSets sel_of(const imvs_corrnet_weights* sets, int period, int split1, int split2, int which) {
    WSets s;
    for (int i = 0; i < 3; ++i) {
        const imvs_corrnet_weights& c = sets[i];
        s.w[i] = which == 0 ? c.conv0 : which == 1 ? c.conv1 : which == 2 ? c.conv2
               : which == 3 ? c.conv3 : which == 4 ? c.conv4 : c.conv5;
    }
    s.period = period; s.split1 = split1; s.split2 = split2;
    return s;
}

// ------------------------------------------------------------------------ PixelViewWeight ----
struct EpiPvw {              // relu(16 channels) . w1 + b1 -> logits[n][y][x]; the 16 channels of a pixel
    float* logits;           // are spread over the 4 lanes of a quad -> two xor-shuffles
    const float* w1;
    const float* b1;
    int H, W;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT], int) const {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = co0 + 8 * j + 2 * t;
            s = fmaf(fmaxf(v[2 * j], 0.f), ldg(w1 + co), s);
            s = fmaf(fmaxf(v[2 * j + 1], 0.f), ldg(w1 + co + 1), s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (t == 0 && oy < H && ox < W) logits[((size_t)n * H + oy) * W + ox] = s + ldg(b1);
    }
};

// softmax over D then max over D == 1 / sum_d exp(l_d - max_d l)   (itermvs.py:347-348)
__global__ void pvw_reduce_kernel(const float* __restrict__ logits, float* __restrict__ vw3, int BS, int D, int P3) {
    pdl_trigger();
    pdl_wait();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= BS * P3) return;
    int bs = t / P3, p = t % P3;
    const float* l = logits + (size_t)bs * D * P3 + p;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, ldg(l + (size_t)d * P3));
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += expf(ldg(l + (size_t)d * P3) - m);
    vw3[t] = 1.0f / s;
}

// F.interpolate(scale_factor=2, mode='bilinear') of channels-last maps [N][H][W][C] -> [N][2H][2W][C]
// (C = 1: itermvs.py:56-57; C = 32 with tanh: itermvs.py:161-163)
__global__ void upsample2x_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C,
                                       bool apply_tanh) {
    pdl_trigger();
    pdl_wait();
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int Ho = 2 * H, Wo = 2 * W;
    if (t >= (size_t)N * Ho * Wo * C) return;
    const int c = (int)(t % C);
    size_t q = t / C;
    const int ox = (int)(q % Wo), oy = (int)((q / Wo) % Ho);
    const size_t n = q / ((size_t)Wo * Ho);
    int h0, h1, w0, w1;
    float lh, lw;
    up_index(oy, 0.5f, H, h0, h1, lh);
    up_index(ox, 0.5f, W, w0, w1, lw);
    const float* b = in + n * H * W * C + c;
    float v = (1.f - lh) * ((1.f - lw) * ldg(b + ((size_t)h0 * W + w0) * C) + lw * ldg(b + ((size_t)h0 * W + w1) * C)) +
              lh * ((1.f - lw) * ldg(b + ((size_t)h1 * W + w0) * C) + lw * ldg(b + ((size_t)h1 * W + w1) * C));
    out[t] = apply_tanh ? tanhf(v) : v;
}

int launch_upsample2x_nhwc(const float* in, float* out, int N, int H, int W, int C, bool apply_tanh, cudaStream_t st) {
    size_t total = (size_t)N * H * W * C * 4;
    IMVS_CUDA(launch_k(upsample2x_nhwc_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, in, out, N, H, W, C, apply_tanh));
    return 0;
}

// a7, models/itermvs.py:74-81 (the init branch's own depth estimate; consumed by the training loss only): softmax over the D
// hypotheses of the CorrNet output, expectation of the hypothesis index, / (D - 1), depth_unnormalization (module.py:148-152).
// corr element (b, d, p) at b * bstride + d * dstride + p * pstride (planar [B][D][P] or channels-last [B][P][D]).
__global__ void init_expectation_kernel(const float* __restrict__ corr, size_t bstride, size_t dstride, size_t pstride,
                                        const float* __restrict__ depth_min, const float* __restrict__ depth_max,
                                        float* __restrict__ depth3, int B, int D, int P) {
    pdl_trigger();
    pdl_wait();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * P) return;
    const int b = (int)(t / P), p = (int)(t % P);
    const float* c = corr + (size_t)b * bstride + (size_t)p * pstride;
    float m = ldg(c);
    for (int d = 1; d < D; ++d) m = fmaxf(m, ldg(c + (size_t)d * dstride));
    float s = 0.f, num = 0.f;
    for (int d = 0; d < D; ++d) {
        const float e = expf(ldg(c + (size_t)d * dstride) - m);
        s += e;
        num = fmaf((float)d, e, num);
    }
    // sum_d d * (e_d / s), as torch.sum(index * softmax) does, up to the rounding of the common division
    const float nd = (num / s) / (float)(D - 1);
    const float inv_min = 1.0f / depth_min[b], inv_max = 1.0f / depth_max[b];
    depth3[t] = unnormalize_depth(nd, inv_min, inv_max);
}

template <int D>
static int hinit_conv0(const imvs_weights* w, const float* corr, float* t, int B, int H3, int W3, cudaStream_t st) {
    // the 1/8-resolution map is small (80 x 64 at 640x512: 40 CTAs of 8 rows x 16 columns x 64 couts, each with 864 dependent MMAs per
    // warp): IMVS_TUNE_HINIT_TILE=1: 4-row tiles (80 CTAs), 2: 4-row tiles x two 32-cout blocks (160 CTAs)
    const EpiNHWC e{t, nullptr, nullptr, H3, W3, 64, 64, 1};
    const int ht = tune("HINIT_TILE", 2);      // default 2: 29.6 -> 23.7 us (gpurun call r2c58)
    if (ht == 1) return mma_conv<D, 64, 1, 4, 1, false>("hidden_init.conv0", in_nhwc(corr, H3, W3, D), e, WSets::single(w->hinit_conv0), conv_tables(3, 1, 1, 4), B, 64, H3, W3, 1, st);
    if (ht == 2) return mma_conv<D, 32, 1, 4, 1, false>("hidden_init.conv0", in_nhwc(corr, H3, W3, D), e, WSets::single(w->hinit_conv0), conv_tables(3, 1, 1, 4), B, 64, H3, W3, 2, st);
    return mma_conv<D, 64, 2, 4, 1, false>("hidden_init.conv0", in_nhwc(corr, H3, W3, D), e, WSets::single(w->hinit_conv0), conv_tables(3, 1, 1, 8), B, 64, H3, W3, 1, st);
}

}  // namespace imvs

using namespace imvs;

extern "C" size_t imvs_corrnet_scratch_floats(int N, int H, int W) { return (size_t)N * 48 * H * W; }

extern "C" int imvs_corrnet(const imvs_corrnet_weights* sets, int period, int split1, int split2, const float* vol,
                            float* out, size_t out_batch_stride, size_t out_pixel_stride, float* scratch,
                            int N, int H, int W, void* stream) {
    IMVS_REQUIRE(sets && vol && out && scratch, "corrnet: null pointer");
    IMVS_REQUIRE(N >= 1 && H >= 4 && W >= 4 && H % 4 == 0 && W % 4 == 0, "corrnet: H, W must be multiples of 4 (H=%d W=%d)", H, W);
    IMVS_REQUIRE(period >= 1 && N % period == 0, "corrnet: N=%d not a multiple of period=%d", N, period);
    ApiScope api_;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t HW = (size_t)H * W;
    float* c0 = scratch;                    // [N][H][W][8]
    float* c1 = c0 + (size_t)N * 8 * HW;    // [N][H/2][W/2][16]
    float* c2 = c1 + (size_t)N * 4 * HW;    // [N][H/4][W/4][32]
    float* x3 = c2 + (size_t)N * 2 * HW;    // [N][H/2][W/2][16]
    float* x4 = x3 + (size_t)N * 4 * HW;    // [N][H][W][8]
    const int H1 = H / 2, W1 = W / 2, H2 = H / 4, W2 = W / 4;
    if (tune("CORR_TILE", 0) && sets[0].conv0.fp32 && (size_t)cdiv(H, ctile::T) <= 65535 && N <= 65535) {
        // option (IMVS_TUNE_CORR_TILE=1; measured 76 us per pass against 66 us for the six mma.sync launches, profiles/README.md):
        // one launch, a 32 x 32 tile resident in shared memory through all six layers (corrnet_tile.cuh), exact fp32
        ctile::Params P{};
        P.vol = vol;
        for (int i = 0; i < 3; ++i) {
            const imvs_corrnet_weights& c = sets[i];
            const imvs_wpair* wp[6] = {&c.conv0, &c.conv1, &c.conv2, &c.conv3, &c.conv4, &c.conv5};
            for (int l = 0; l < 6; ++l) {
                IMVS_REQUIRE(wp[l]->fp32, "corrnet: fp32 weights missing");
                P.w[i][l] = wp[l]->fp32;
            }
            P.b5[i] = c.conv5_b;
        }
        P.period = period; P.split1 = split1; P.split2 = split2;
        P.out = out; P.bstride = out_batch_stride; P.pstride = out_pixel_stride;
        P.N = N; P.H = H; P.W = W; P.tiles_x = cdiv(W, ctile::T); P.tiles_y = cdiv(H, ctile::T);
        static int smem_ok = 0;
        IMVS_TRY(ensure_dynamic_smem(ctile::corrnet_tile_kernel, ctile::SMEM_BYTES, &smem_ok));
        IMVS_CUDA(launch_k(ctile::corrnet_tile_kernel, dim3(P.tiles_x, P.tiles_y, N), dim3(ctile::THREADS), ctile::SMEM_BYTES, st, P));
        return 0;
    }
#ifndef CUSIM
    if (corrnet_tc5p_ready(sets)) {
        // optional (IMVS_TUNE_TC5P_CORR=1): all six layers on the TMA + tcgen05 kernel; activations as fp16 hi / lo split planes, the inputs of the two
        // stride-2 layers also as parity planes, the transposed layers with four parity accumulators (tc5pconv.cuh)
        float* p = scratch;
        auto take = [&](size_t floats) { float* q = p; p += floats; return q; };
        const size_t n8 = (size_t)N * 8 * HW, n4 = (size_t)N * 4 * HW, n2 = (size_t)N * 2 * HW;
        const tc5p::Split vs = tc5p::split_at(take(n8), n8), c0s = tc5p::split_at(take(n8), n8), c0p = tc5p::split_at(take(n8), n8),
                          c1s = tc5p::split_at(take(n4), n4), c1p = tc5p::split_at(take(n4), n4), c2s = tc5p::split_at(take(n2), n2),
                          x3s = tc5p::split_at(take(n4), n4), x4s = tc5p::split_at(take(n8), n8), none{nullptr, nullptr};
        int* flag = tc5_error_flag();
        auto ws = [&](int which) { return wsel_of(sets, period, split1, split2, which); };
        EpiCorrOutP e5;
        e5.out = out;
        for (int i = 0; i < 3; ++i) e5.bias[i] = sets[i].conv5_b;
        e5.period = period; e5.split1 = split1; e5.split2 = split2;
        e5.bstride = out_batch_stride; e5.pstride = out_pixel_stride; e5.H = H; e5.W = W;
        if (tune("CORR_FUSED", 0)) {
            // one cooperative launch for the whole pass
            CorrFusedParams P{};
            const size_t budget = 200 * 1024;
            const int ctas = tc5p::sm_count();
            size_t smem = 0, need = 0;
            for (int k = 0; k < 6; ++k) P.w[k] = ws(k);
            IMVS_TRY((tc5p::plan_layer<8, 16, 2, 1, 3, 1>("corrnet.conv0", P.w[0], N, H, W, ctas, budget, P.g[0], need))); smem = std::max(smem, need);
            IMVS_TRY((tc5p::plan_layer<8, 16, 1, 1, 3, 2>("corrnet.conv1", P.w[1], N, H1, W1, ctas, budget, P.g[1], need))); smem = std::max(smem, need);
            IMVS_TRY((tc5p::plan_layer<16, 32, 1, 1, 3, 2>("corrnet.conv2", P.w[2], N, H2, W2, ctas, budget, P.g[2], need))); smem = std::max(smem, need);
            IMVS_TRY((tc5p::plan_layer<32, 16, 1, 1, 3, 0>("corrnet.conv3", P.w[3], N, H2, W2, ctas, budget, P.g[3], need))); smem = std::max(smem, need);
            IMVS_TRY((tc5p::plan_layer<16, 16, 1, 1, 3, 0>("corrnet.conv4", P.w[4], N, H1, W1, ctas, budget, P.g[4], need))); smem = std::max(smem, need);
            IMVS_TRY((tc5p::plan_layer<8, 16, 2, 1, 3, 1>("corrnet.conv5", P.w[5], N, H, W, ctas, budget, P.g[5], need))); smem = std::max(smem, need);
            struct In { const tc5p::Split* t; int n, kc, h, w, rows; };
            const In ins[6] = {{&vs, N, 1, H, W, P.g[0].rows}, {&c0p, 4 * N, 1, H1, W1, P.g[1].rows}, {&c1p, 4 * N, 2, H2, W2, P.g[2].rows},
                               {&c2s, N, 4, H2, W2, P.g[3].rows}, {&x3s, N, 2, H1, W1, P.g[4].rows}, {&x4s, N, 1, H, W, P.g[5].rows}};
            for (int k = 0; k < 6; ++k) {
                IMVS_TRY(tc5p::make_plane_map(&P.m[2 * k], ins[k].t->hi, ins[k].n, ins[k].kc, ins[k].h, ins[k].w, ins[k].rows));
                IMVS_TRY(tc5p::make_plane_map(&P.m[2 * k + 1], ins[k].t->lo, ins[k].n, ins[k].kc, ins[k].h, ins[k].w, ins[k].rows));
            }
            P.e0 = tc5p::Epi{c0s, nullptr, none, nullptr, H, W, 1, c0p, 1};
            P.e1 = tc5p::Epi{c1s, nullptr, none, nullptr, H1, W1, 1, c1p, 0};
            P.e2 = tc5p::Epi{c2s, nullptr, none, nullptr, H2, W2, 1};
            P.e3 = EpiTconvP{x3s, c1s, H2, W2, 2};
            P.e4 = EpiTconvP{x4s, c0s, H1, W1, 1};
            P.e5 = e5;
            P.vol = vol; P.vs_hi = vs.hi; P.vs_lo = vs.lo; P.npix = (unsigned long long)N * HW;
            P.counter = reinterpret_cast<unsigned int*>(take(64));
            P.err_flag = flag;
            static int smem_ok = 0;
            IMVS_TRY(ensure_dynamic_smem(corrnet_fused_kernel, smem, &smem_ok));
            IMVS_CUDA(cudaMemsetAsync(P.counter, 0, sizeof(unsigned int), st));
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(CF_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeCooperative;
            attr[0].val.cooperative = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            *pdl_armed() = false;            // the next launch of this call must not be a programmatic dependent of the memset / this kernel
            count_launch();
            IMVS_CUDA(cudaLaunchKernelEx(&cfg, corrnet_fused_kernel, P));
            return 0;
        }
        IMVS_TRY(tc5p::launch_nhwc_to_split(vol, vs, (size_t)N * HW, (int)HW, 8, st));
        IMVS_TRY((tc5p::launch<8, 16>("corrnet.conv0", vs, tc5p::Epi{c0s, nullptr, none, nullptr, H, W, 1, c0p, 1}, ws(0), N, H, W, flag, st)));
        IMVS_TRY((tc5p::launch<8, 16, 1, true, 3, 2>("corrnet.conv1", c0p, tc5p::Epi{c1s, nullptr, none, nullptr, H1, W1, 1, c1p, 0}, ws(1), N, H1, W1, flag, st)));
        IMVS_TRY((tc5p::launch<16, 32, 1, true, 3, 2>("corrnet.conv2", c1p, tc5p::Epi{c2s, nullptr, none, nullptr, H2, W2, 1}, ws(2), N, H2, W2, flag, st)));
        IMVS_TRY((tc5p::launch<32, 16, 1, false, 3, 0>("corrnet.conv3", c2s, EpiTconvP{x3s, c1s, H2, W2, 2}, ws(3), N, H2, W2, flag, st)));
        IMVS_TRY((tc5p::launch<16, 16, 1, false, 3, 0>("corrnet.conv4", x3s, EpiTconvP{x4s, c0s, H1, W1, 1}, ws(4), N, H1, W1, flag, st)));
        IMVS_TRY((tc5p::launch<8, 16>("corrnet.conv5", x4s, e5, ws(5), N, H, W, flag, st)));
        return 0;
    }
#endif
    auto sel = [&](int which) { return sel_of(sets, period, split1, split2, which); };
    // tile-shape experiment for the two full-resolution layers (IMVS_TUNE_CORR_TILE05: rows per CTA = 8 (default), 16 as 4 warps x 4
    // row-tiles, 16 as 8 warps x 2, 4 as 4 warps x 1)
    const int t05 = tune("CORR_TILE05", 0);
    const EpiNHWC e0{c0, nullptr, nullptr, H, W, 8, 8, 1};
    if (t05 == 1) IMVS_TRY((mma_conv<8, 8, 4, 4, 1, true>("corrnet.conv0", in_nhwc(vol, H, W, 8), e0, sel(0), conv_tables(3, 1, 1, 16), N, 8, H, W, 1, st)));
    else if (t05 == 2) IMVS_TRY((mma_conv<8, 8, 2, 8, 1, true>("corrnet.conv0", in_nhwc(vol, H, W, 8), e0, sel(0), conv_tables(3, 1, 1, 16), N, 8, H, W, 1, st)));
    else if (t05 == 3) IMVS_TRY((mma_conv<8, 8, 1, 4, 1, true>("corrnet.conv0", in_nhwc(vol, H, W, 8), e0, sel(0), conv_tables(3, 1, 1, 4), N, 8, H, W, 1, st)));
    else IMVS_TRY((mma_conv<8, 8, 2, 4, 1, true>("corrnet.conv0", in_nhwc(vol, H, W, 8), e0, sel(0), conv_tables(3, 1, 1, 8), N, 8, H, W, 1, st)));
    // conv3 / conv4: transposed convolutions, the four output parities as variants of one launch,
    // + U-Net skips c1 / c0 (itermvs.py:374-377).  IMVS_TUNE_CORR_TILEMID=1: 4-row instead of 8-row tiles for the coarse layers
    if (tune("CORR_TILEMID", 0) == 1) {
        IMVS_TRY((mma_conv<8, 16, 1, 4, 2, true>("corrnet.conv1", in_nhwc(c0, H, W, 8), EpiNHWC{c1, nullptr, nullptr, H1, W1, 16, 16, 1},
                                                 sel(1), conv_tables(3, 2, 1, 4), N, 16, H1, W1, 1, st)));
        IMVS_TRY((mma_conv<16, 32, 1, 4, 2, true>("corrnet.conv2", in_nhwc(c1, H1, W1, 16), EpiNHWC{c2, nullptr, nullptr, H2, W2, 32, 32, 1},
                                                  sel(2), conv_tables(3, 2, 1, 4), N, 32, H2, W2, 1, st)));
        IMVS_TRY((mma_conv<32, 16, 1, 4, 1, true>("corrnet.conv3", in_nhwc(c2, H2, W2, 32), EpiTconvNHWC{x3, c1, H2, W2, 16},
                                                  sel(3), tconv_tables(4), N, 16, H2, W2, 1, st)));
        IMVS_TRY((mma_conv<16, 8, 1, 4, 1, true>("corrnet.conv4", in_nhwc(x3, H1, W1, 16), EpiTconvNHWC{x4, c0, H1, W1, 8},
                                                 sel(4), tconv_tables(4), N, 8, H1, W1, 1, st)));
    } else {
        IMVS_TRY((mma_conv<8, 16, 2, 4, 2, true>("corrnet.conv1", in_nhwc(c0, H, W, 8), EpiNHWC{c1, nullptr, nullptr, H1, W1, 16, 16, 1},
                                                 sel(1), conv_tables(3, 2, 1, 8), N, 16, H1, W1, 1, st)));
        IMVS_TRY((mma_conv<16, 32, 2, 4, 2, true>("corrnet.conv2", in_nhwc(c1, H1, W1, 16), EpiNHWC{c2, nullptr, nullptr, H2, W2, 32, 32, 1},
                                                  sel(2), conv_tables(3, 2, 1, 8), N, 32, H2, W2, 1, st)));
        IMVS_TRY((mma_conv<32, 16, 2, 4, 1, true>("corrnet.conv3", in_nhwc(c2, H2, W2, 32), EpiTconvNHWC{x3, c1, H2, W2, 16},
                                                  sel(3), tconv_tables(8), N, 16, H2, W2, 1, st)));
        IMVS_TRY((mma_conv<16, 8, 2, 4, 1, true>("corrnet.conv4", in_nhwc(x3, H1, W1, 16), EpiTconvNHWC{x4, c0, H1, W1, 8},
                                                 sel(4), tconv_tables(8), N, 8, H1, W1, 1, st)));
    }
    EpiCorrOut e5;
    e5.out = out;
    for (int i = 0; i < 3; ++i) e5.bias[i] = sets[i].conv5_b;
    e5.period = period; e5.split1 = split1; e5.split2 = split2;
    e5.bstride = out_batch_stride; e5.pstride = out_pixel_stride; e5.H = H; e5.W = W;
    if (t05 == 1) IMVS_TRY((mma_conv<8, 8, 4, 4, 1, true>("corrnet.conv5", in_nhwc(x4, H, W, 8), e5, sel(5), conv_tables(3, 1, 1, 16), N, 8, H, W, 1, st)));
    else if (t05 == 2) IMVS_TRY((mma_conv<8, 8, 2, 8, 1, true>("corrnet.conv5", in_nhwc(x4, H, W, 8), e5, sel(5), conv_tables(3, 1, 1, 16), N, 8, H, W, 1, st)));
    else if (t05 == 3) IMVS_TRY((mma_conv<8, 8, 1, 4, 1, true>("corrnet.conv5", in_nhwc(x4, H, W, 8), e5, sel(5), conv_tables(3, 1, 1, 4), N, 8, H, W, 1, st)));
    else IMVS_TRY((mma_conv<8, 8, 2, 4, 1, true>("corrnet.conv5", in_nhwc(x4, H, W, 8), e5, sel(5), conv_tables(3, 1, 1, 8), N, 8, H, W, 1, st)));
    return 0;
}

extern "C" int imvs_pixel_view_weight(const imvs_weights* w, const float* corr, float* logits, float* vw3, float* vw2,
                                      int B, int S, int D, int H3, int W3, void* stream) {
    IMVS_REQUIRE(w && corr && logits && vw3 && vw2, "pixel_view_weight: null pointer");
    IMVS_REQUIRE(B >= 1 && S >= 1 && D >= 1 && H3 >= 1 && W3 >= 1, "pixel_view_weight: bad shape");
    ApiScope api_;
    cudaStream_t st = (cudaStream_t)stream;
    const int N = B * S * D, P3 = H3 * W3;
    // IMVS_TUNE_PVW_TILE: rows per CTA = 8 (0: 4 warps x 2 row-tiles, 5 120 CTAs at 640x512 / 4 src / D = 32), 16 as 4 x 4 (1, default:
    // 47.6 -> 44.3 us, gpurun call r2c56), 16 as 8 x 2 (2: 48.0 us), 32 as 8 x 4 (3: 44.4-45.1 us)
    const int pt = tune("PVW_TILE", 1);
    const EpiPvw ep{logits, w->pvw_conv1, w->pvw_conv1_b, H3, W3};
    if (pt == 1) IMVS_TRY((mma_conv<8, 16, 4, 4, 1, true>("pvw.conv", in_nhwc(corr, H3, W3, 8), ep, WSets::single(w->pvw_conv0), conv_tables(3, 1, 1, 16), N, 16, H3, W3, 1, st)));
    else if (pt == 2) IMVS_TRY((mma_conv<8, 16, 2, 8, 1, true>("pvw.conv", in_nhwc(corr, H3, W3, 8), ep, WSets::single(w->pvw_conv0), conv_tables(3, 1, 1, 16), N, 16, H3, W3, 1, st)));
    else if (pt == 3) IMVS_TRY((mma_conv<8, 16, 4, 8, 1, true>("pvw.conv", in_nhwc(corr, H3, W3, 8), ep, WSets::single(w->pvw_conv0), conv_tables(3, 1, 1, 32), N, 16, H3, W3, 1, st)));
    else IMVS_TRY((mma_conv<8, 16, 2, 4, 1, true>("pvw.conv", in_nhwc(corr, H3, W3, 8), ep, WSets::single(w->pvw_conv0), conv_tables(3, 1, 1, 8), N, 16, H3, W3, 1, st)));
    IMVS_CUDA(launch_k(pvw_reduce_kernel, dim3(cdiv(B * S * P3, 128)), dim3(128), 0, st, (const float*)logits, vw3, B * S, D, P3));
    return launch_upsample2x_nhwc(vw3, vw2, B * S, H3, W3, 1, false, st);
}

extern "C" int imvs_hidden_init(const imvs_weights* w, const float* corr, float* hidden, float* scratch,
                                int B, int D, int H3, int W3, void* stream) {
    IMVS_REQUIRE(w && corr && hidden && scratch, "hidden_init: null pointer");
    IMVS_REQUIRE(B >= 1 && H3 >= 1 && W3 >= 1, "hidden_init: bad shape");
    ApiScope api_;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t P3 = (size_t)H3 * W3;
    float* t = scratch;                       // [B][P3][64]
    float* u = scratch + (size_t)B * 64 * P3; // [B][P3][32]
    switch (D) {
        case 8: IMVS_TRY(hinit_conv0<8>(w, corr, t, B, H3, W3, st)); break;
        case 16: IMVS_TRY(hinit_conv0<16>(w, corr, t, B, H3, W3, st)); break;
        case 32: IMVS_TRY(hinit_conv0<32>(w, corr, t, B, H3, W3, st)); break;
        case 48: IMVS_TRY(hinit_conv0<48>(w, corr, t, B, H3, W3, st)); break;
        case 64: IMVS_TRY(hinit_conv0<64>(w, corr, t, B, H3, W3, st)); break;
        default: return fail("hidden_init: D=%d not supported (8, 16, 32, 48 or 64 hypotheses)", D);
    }
    const EpiNHWC efc{u, w->hinit_fc_b, nullptr, H3, W3, 32, 32, 0};
    if (tune("HINIT_TILE", 2) >= 1)
        IMVS_TRY((mma_conv<64, 32, 1, 4, 1, true>("hidden_init.fc", in_nhwc(t, H3, W3, 64), efc, WSets::single(w->hinit_fc), conv_tables(1, 1, 1, 4), B, 32, H3, W3, 1, st)));
    else
        IMVS_TRY((mma_conv<64, 32, 2, 4, 1, true>("hidden_init.fc", in_nhwc(t, H3, W3, 64), efc, WSets::single(w->hinit_fc), conv_tables(1, 1, 1, 8), B, 32, H3, W3, 1, st)));
    return launch_upsample2x_nhwc(u, hidden, B, H3, W3, 32, true, st);
}

extern "C" int imvs_init_depth(const float* corr, size_t batch_stride, size_t slice_stride, size_t pixel_stride,
                               const float* depth_min, const float* depth_max, float* scratch, float* depth_out,
                               int B, int D, int H3, int W3, void* stream) {
    IMVS_REQUIRE(corr && depth_min && depth_max && scratch && depth_out, "init_depth: null pointer");
    IMVS_REQUIRE(B >= 1 && D >= 2 && H3 >= 1 && W3 >= 1, "init_depth: bad shape");
    ApiScope api_;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * H3 * W3;
    IMVS_CUDA(launch_k(init_expectation_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, st, corr, batch_stride, slice_stride,
                       pixel_stride, depth_min, depth_max, scratch, B, D, H3 * W3));
    // F.interpolate(depth, scale_factor=2, mode="bilinear") (itermvs.py:80)
    return launch_upsample2x_nhwc(scratch, depth_out, B, H3, W3, 1, false, st);
}
