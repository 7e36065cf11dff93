// Update stage: ConvGRU (module.py:52-66) and depth / confidence heads with softmax, arg-max and
// clamped-window regression (itermvs.py:139-151, 171-190, 192-220).  Activations channels-last.
#include <algorithm>

#include "common.cuh"
#include "mmaconv.cuh"
#include "tc5conv.cuh"
#ifndef CUSIM
#include "tc5pconv.cuh"
#endif
#include "headfused.cuh"

namespace imvs {

// --------------------------------------------------------------------------------- ConvGRU ----
// z|r as ONE 48 -> 64 dilated implicit GEMM (convz and convr share their input hx = [h, x]); the
// epilogue applies the sigmoids and writes z and r*h.  q = tanh(convq([r*h, x])) is a second GEMM
// whose epilogue performs the gate  h <- (1-z) h + z q  in place (h is only read point-wise there).
struct EpiGruZR {
    const float* bias;   // [64]
    const float* h;      // [N][H][W][32]
    float* z;            // [N][H][W][32]
    float* rh;           // [N][H][W][32]
    int H, W;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT], int) const {
        if (oy >= H || ox >= W) return;
        const size_t base = (((size_t)n * H + oy) * W + ox) * 32;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = co0 + 8 * j + 2 * t;
            const float s0 = sigmoidf_(v[2 * j] + ldg(bias + co)), s1 = sigmoidf_(v[2 * j + 1] + ldg(bias + co + 1));
            if (co < 32) {
                *reinterpret_cast<float2*>(z + base + co) = make_float2(s0, s1);
            } else {
                const float2 hh = ldg2(h + base + co - 32);
                *reinterpret_cast<float2*>(rh + base + co - 32) = make_float2(s0 * hh.x, s1 * hh.y);
            }
        }
    }
};

struct EpiGruQ {
    const float* bias;   // [32]
    const float* z;
    float* h;            // updated in place
    int H, W;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT], int) const {
        if (oy >= H || ox >= W) return;
        const size_t base = (((size_t)n * H + oy) * W + ox) * 32;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = co0 + 8 * j + 2 * t;
            const float q0 = tanhf(v[2 * j] + ldg(bias + co)), q1 = tanhf(v[2 * j + 1] + ldg(bias + co + 1));
            const float2 zz = ldg2(z + base + co);
            float2 hh = *reinterpret_cast<const float2*>(h + base + co);
            hh.x = (1.f - zz.x) * hh.x + zz.x * q0;
            hh.y = (1.f - zz.y) * hh.y + zz.y * q1;
            *reinterpret_cast<float2*>(h + base + co) = hh;
        }
    }
};

#ifndef CUSIM     // (the CPU emulation of the test-suite has no TMA / tcgen05 model)
// ---- the same two GEMMs on the persistent TMA + tcgen05 kernel (tc5pconv.cuh), the default in the fp32-grade mode -------
// hx = [h (32) | x (16)] and rhx = [r*h (32) | x (16)] live as 48-channel split-plane tensors [B][6][H][W][8 halves]: one
// conversion kernel writes h and x into hx and x into rhx, the z|r kernel's epilogue writes r*h into rhx.
__global__ void gru_split_inputs_kernel(const float* __restrict__ h, const float* __restrict__ x, __half* __restrict__ a_hi,
                                        __half* __restrict__ a_lo, __half* __restrict__ q_hi, __half* __restrict__ q_lo, int B, int P) {
    pdl_trigger();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // (b, chunk, pixel)
    if (i >= (size_t)B * 6 * P) return;
    const int p = (int)(i % P), kc = (int)((i / P) % 6), b = (int)(i / ((size_t)6 * P));
    const float* src = kc < 4 ? h + ((size_t)b * P + p) * 32 + kc * 8 : x + ((size_t)b * P + p) * IMVS_XCH + (kc - 4) * 8;
    const float4 u = ldg4(src), v = ldg4(src + 4);
    uint4 hi, lo;
    split_f16(make_float2(u.x, u.y), hi.x, lo.x);
    split_f16(make_float2(u.z, u.w), hi.y, lo.y);
    split_f16(make_float2(v.x, v.y), hi.z, lo.z);
    split_f16(make_float2(v.z, v.w), hi.w, lo.w);
    reinterpret_cast<uint4*>(a_hi)[i] = hi;
    reinterpret_cast<uint4*>(a_lo)[i] = lo;
    if (kc >= 4) {
        reinterpret_cast<uint4*>(q_hi)[i] = hi;
        reinterpret_cast<uint4*>(q_lo)[i] = lo;
    }
}

struct EpiGruZRp {           // 64 stacked channels: a thread's slice is part of z (c0 < 32) or of r (c0 >= 32)   (module.py:61-62)
    static constexpr int kAhead = 1;
    __device__ __forceinline__ bool wants_prefetch() const { return true; }
    const float* bias;       // [64]
    const float* h;          // [B][P][32]
    float* z;                // [B][P][32]
    tc5p::Split rhx;         // [B][6][P][8]: chunks 0..3 <- r * h
    int H, W;
    template <int NCH> struct Pre { float4 hh[NCH / 4]; };
    template <int NB, int NCH>
    __device__ __forceinline__ void prefetch(int n, int oy, int ox, int c0, Pre<NCH>& p) const {
        static_assert(NB == 64 && 32 % NCH == 0, "z | r halves");
        if (c0 < 32) return;
        const float* src = h + (((size_t)n * H + oy) * W + ox) * 32 + (c0 - 32);
#pragma unroll
        for (int c = 0; c < NCH / 4; ++c) p.hh[c] = ldg4(src + 4 * c);
    }
    template <int NB, int NCH>
    __device__ __forceinline__ void store(int n, int oy, int ox, int c0, float (&v)[NCH], const Pre<NCH>& p, int*) const {
        const size_t plane = (size_t)H * W, pix = (size_t)oy * W + ox;
        if (c0 < 32) {
            float* dst = z + ((size_t)n * plane + pix) * 32 + c0;
#pragma unroll
            for (int c = 0; c < NCH; c += 4) {
                const float4 b = ldg4(bias + c0 + c);
                *reinterpret_cast<float4*>(dst + c) =
                    make_float4(sigmoidf_(v[c] + b.x), sigmoidf_(v[c + 1] + b.y), sigmoidf_(v[c + 2] + b.z), sigmoidf_(v[c + 3] + b.w));
            }
        } else {
#pragma unroll
            for (int j = 0; j < NCH / 8; ++j) {
                const float4 b0 = ldg4(bias + c0 + 8 * j), b1 = ldg4(bias + c0 + 4 + 8 * j), h0 = p.hh[2 * j], h1 = p.hh[2 * j + 1];
                const float* a = v + 8 * j;
                uint4 hi, lo;
                split_f16(make_float2(sigmoidf_(a[0] + b0.x) * h0.x, sigmoidf_(a[1] + b0.y) * h0.y), hi.x, lo.x);
                split_f16(make_float2(sigmoidf_(a[2] + b0.z) * h0.z, sigmoidf_(a[3] + b0.w) * h0.w), hi.y, lo.y);
                split_f16(make_float2(sigmoidf_(a[4] + b1.x) * h1.x, sigmoidf_(a[5] + b1.y) * h1.y), hi.z, lo.z);
                split_f16(make_float2(sigmoidf_(a[6] + b1.z) * h1.z, sigmoidf_(a[7] + b1.w) * h1.w), hi.w, lo.w);
                const size_t idx = ((size_t)n * 6 + ((c0 - 32) / 8 + j)) * plane + pix;
                reinterpret_cast<uint4*>(rhx.hi)[idx] = hi;
                reinterpret_cast<uint4*>(rhx.lo)[idx] = lo;
            }
        }
    }
};

struct EpiGruQp {            // q = tanh, h <- (1 - z) h + z q in place   (module.py:63-64)
    static constexpr int kAhead = 1;
    __device__ __forceinline__ bool wants_prefetch() const { return true; }
    const float* bias;       // [32]
    const float* z;          // [B][P][32]
    float* h;                // [B][P][32]
    int H, W;
    template <int NCH> struct Pre { float4 zz[NCH / 4], hh[NCH / 4]; };
    template <int NB, int NCH>
    __device__ __forceinline__ void prefetch(int n, int oy, int ox, int c0, Pre<NCH>& p) const {
        static_assert(NB == 32, "convq");
        const size_t base = (((size_t)n * H + oy) * W + ox) * 32 + c0;
#pragma unroll
        for (int c = 0; c < NCH / 4; ++c) {
            p.zz[c] = ldg4(z + base + 4 * c);
            p.hh[c] = *reinterpret_cast<const float4*>(h + base + 4 * c);
        }
    }
    template <int NB, int NCH>
    __device__ __forceinline__ void store(int n, int oy, int ox, int c0, float (&v)[NCH], const Pre<NCH>& p, int*) const {
        const size_t base = (((size_t)n * H + oy) * W + ox) * 32 + c0;
#pragma unroll
        for (int c = 0; c < NCH / 4; ++c) {
            const float4 b = ldg4(bias + c0 + 4 * c), zz = p.zz[c];
            float4 hh = p.hh[c];
            hh.x = (1.f - zz.x) * hh.x + zz.x * tanhf(v[4 * c] + b.x);
            hh.y = (1.f - zz.y) * hh.y + zz.y * tanhf(v[4 * c + 1] + b.y);
            hh.z = (1.f - zz.z) * hh.z + zz.z * tanhf(v[4 * c + 2] + b.z);
            hh.w = (1.f - zz.w) * hh.w + zz.w * tanhf(v[4 * c + 3] + b.w);
            *reinterpret_cast<float4*>(h + base + 4 * c) = hh;
        }
    }
};

#endif  // !CUSIM

// ------------------------------------------------------------------------------------ heads ----
// depth_head: 3x3 dilated conv (tensor cores) -> fc1 32->64 relu (1x1, tensor cores) -> fc2 64->256 + b
// (1x1, tensor cores) -> this kernel: softmax over the 256 bins, arg-max, clamped +-4 window regression
// (itermvs.py:173-190 / 203-219) and the confidence head's 1x1 + sigmoid (itermvs.py:147-151, 197-199).
// One warp per pixel: lane j owns logits {4j..4j+3, 128+4j..128+4j+3} (two coalesced float4 loads),
// reductions by shuffles.  The logits are written once and read once from L2 (the reference
// re-reads that 21 MB tensor ~6x through softmax / argmax / 9 gathers).
struct RegressParams {
    const float* logits;     // [B][P][256]
    const float* t;          // [B][P][64] conv0 output; channels 32..63 feed the confidence head
    const float* conf_w;     // [32]
    const float* conf_b;     // [1]
    float* nd_out;
    size_t nd_bstride, nd_pstride;
    float* prob;             // [B][256][P] or null
    float* conf;             // [B][P] or null
    float* conf_logit;       // [B][P] or null
    float* depth_out;        // [B][P] or null
    const float* depth_min;
    const float* depth_max;
    int B, P;
};

__global__ void __launch_bounds__(256) softmax_regress_kernel(const RegressParams prm) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= prm.B * prm.P) return;
    const int b = gw / prm.P, p = gw % prm.P;
    const float* lg = prm.logits + (size_t)gw * IMVS_OUT_BINS;
    const float4 la = ldg4(lg + 4 * lane), lb = ldg4(lg + 128 + 4 * lane);
    const float l[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
    float m = l[0];
#pragma unroll
    for (int a = 1; a < 8; ++a) m = fmaxf(m, l[a]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float e[8], s = 0.f;
#pragma unroll
    for (int a = 0; a < 8; ++a) { e[a] = expf(l[a] - m); s += e[a]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    // probabilities e / s are needed (a) to find the arg-max and (b) inside the +-4 window around it.  An IEEE
    // division whose numerator is denormal / tiny takes the slow path, and after a few iterations most of the
    // 256 bins are that small (the kernel went from 14 us to 25 us over the iterations).  (a) only bins with
    // e > 0.5 can reach the maximum e_max / s = 1 / s (e <= 0.5 gives exactly half of it or less), so only
    // those are divided; (b) divides the <= 9 window bins.  Same values as dividing everything.  The guards
    // are on the NUMERATOR (the compiler hoists a division out of a conditional and selects afterwards).
    float pr[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) pr[a] = e[a] > 0.5f ? fmaxf(e[a], 0.5f) / s : 0.f;
    // arg-max over probabilities, first index on ties (torch.argmax)
    float bv = -1.f;
    int bi = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int ch = a < 4 ? 4 * lane + a : 128 + 4 * lane + (a - 4);
        if (pr[a] > bv) { bv = pr[a]; bi = ch; }          // channels visited in increasing order
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    // window: indices clamp(bi-4 .. bi+4, 0, 255); clamped duplicates are counted repeatedly
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int ch = a < 4 ? 4 * lane + a : 128 + 4 * lane + (a - 4);
        int mult = (ch >= bi - IMVS_RADIUS && ch <= bi + IMVS_RADIUS) ? 1 : 0;
        if (ch == 0) mult = max(0, IMVS_RADIUS + 1 - bi);
        if (ch == IMVS_OUT_BINS - 1) mult = max(0, bi - (IMVS_OUT_BINS - 2 - IMVS_RADIUS));
        const float pw = (mult ? e[a] : 1.0f) / s;          // bins outside the window: harmless fast-path division
        num = fmaf((float)(mult * ch), pw, num);
        den = fmaf((float)mult, pw, den);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    const float ndv = (num / (1e-6f + den)) / (float)(IMVS_OUT_BINS - 1);
    if (prm.prob) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int ch = a < 4 ? 4 * lane + a : 128 + 4 * lane + (a - 4);
            prm.prob[((size_t)b * IMVS_OUT_BINS + ch) * prm.P + p] = e[a] / s;
        }
    }
    const bool want_conf = (prm.conf != nullptr) || (prm.conf_logit != nullptr);
    float cs = 0.f;
    if (want_conf) {
        cs = ldg(prm.t + (size_t)gw * 64 + 32 + lane) * ldg(prm.conf_w + lane);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o);
        cs += ldg(prm.conf_b);
    }
    if (lane == 0) {
        prm.nd_out[(size_t)b * prm.nd_bstride + (size_t)p * prm.nd_pstride] = ndv;
        if (prm.depth_out) {
            const float inv_min = 1.0f / prm.depth_min[b], inv_max = 1.0f / prm.depth_max[b];
            prm.depth_out[gw] = unnormalize_depth(ndv, inv_min, inv_max);
        }
        if (prm.conf_logit) prm.conf_logit[gw] = cs;
        if (prm.conf) prm.conf[gw] = sigmoidf_(cs);
    }
}

}  // namespace imvs

using namespace imvs;

extern "C" int imvs_conv_gru(const imvs_weights* w, float* h, const float* x, float* scratch, int B, int H, int W, void* stream) {
    IMVS_REQUIRE(w && h && x && scratch, "conv_gru: null pointer");
    IMVS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "conv_gru: bad shape");
    ApiScope api_;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * 32 * H * W;
    float* z = scratch;
    float* rh = scratch + n;
    if (conv_passes() == 1 && tc5_enabled() && w->gru_zr.umma && w->gru_q.umma) {
        // tcgen05 path: z|r as one M=128 x N=64 x K=432 UMMA chain per tile, q as N=32
        IMVS_TRY((tc5::launch<48, 64>("gru.zr(tcgen05)", InNHWC2{h, x, H, W, 32, IMVS_XCH}, tc5::PixGruZR{w->gru_zr_b, h, z, rh, H, W},
                                      w->gru_zr.umma, 3, 2, B, H, W, tc5_error_flag(), st)));
        IMVS_TRY((tc5::launch<48, 32>("gru.q(tcgen05)", InNHWC2{rh, x, H, W, 32, IMVS_XCH}, tc5::PixGruQ{w->gru_q_b, z, h, H, W},
                                      w->gru_q.umma, 3, 2, B, H, W, tc5_error_flag(), st)));
        return 0;
    }
#ifndef CUSIM
    if (conv_passes() == 4 && w->gru_zr.f16ummai && w->gru_q.f16ummai && tune("TC5P_GRU", 1) && tc5p::encode_tiled_fn()) {
        // default: persistent TMA + tcgen05 kernel on split-plane operands (tc5pconv.cuh)
        const size_t P = (size_t)H * W, e48 = (size_t)B * 48 * P;
        float* zbuf = scratch;                                               // [B][P][32] fp32
        const tc5p::Split hx = tc5p::split_at(scratch + n, e48), rhx = tc5p::split_at(scratch + n + e48, e48);
        const size_t items = (size_t)B * 6 * P;
        IMVS_REQUIRE(items < 2147483647ull, "conv_gru: too many pixels");
        IMVS_CUDA(launch_k(gru_split_inputs_kernel, dim3((unsigned)((items + 255) / 256)), dim3(256), 0, st, (const float*)h, x, hx.hi, hx.lo,
                           rhx.hi, rhx.lo, B, (int)P));
        IMVS_TRY((tc5p::launch<48, 64, 2>("gru.zr(tma+tcgen05)", hx, EpiGruZRp{w->gru_zr_b, h, zbuf, rhx, H, W}, w->gru_zr.f16ummai, B, H, W,
                                          tc5_error_flag(), st)));
        IMVS_TRY((tc5p::launch<48, 32, 2, false>("gru.q(tma+tcgen05)", rhx, EpiGruQp{w->gru_q_b, zbuf, h, H, W}, w->gru_q.f16ummai, B, H, W,
                                          tc5_error_flag(), st)));
        return 0;
    }
    if (conv_passes() == 4 && w->gru_zr.f16umma && w->gru_q.f16umma && tune("TC5H_GRU", 0)) {
        // fp32-grade mode on tcgen05: fp16 hi/lo 3-product chains (tc5conv.cuh:tc5h_conv_kernel), exact-grade gates
        IMVS_TRY((tc5::launch_h<48, 64>("gru.zr(tcgen05 f16x3)", InNHWC2{h, x, H, W, 32, IMVS_XCH}, tc5::PixGruZR{w->gru_zr_b, h, z, rh, H, W, 1},
                                        w->gru_zr.f16umma, 3, 2, B, H, W, tc5_error_flag(), st)));
        IMVS_TRY((tc5::launch_h<48, 32>("gru.q(tcgen05 f16x3)", InNHWC2{rh, x, H, W, 32, IMVS_XCH}, tc5::PixGruQ{w->gru_q_b, z, h, H, W, 1},
                                        w->gru_q.f16umma, 3, 2, B, H, W, tc5_error_flag(), st)));
        return 0;
    }
#endif
    const InNHWC2 in_zr{h, x, H, W, 32, IMVS_XCH}, in_q{rh, x, H, W, 32, IMVS_XCH};
    const EpiGruZR ezr{w->gru_zr_b, h, z, rh, H, W};
    const EpiGruQ eq{w->gru_q_b, z, h, H, W};
    const WSets wzr = WSets::single(w->gru_zr), wq = WSets::single(w->gru_q);
    switch (tune("GRU", 2)) {
        case 1:      // 16-cout blocks: 4x / 2x the CTAs
            IMVS_TRY((mma_conv<48, 16, 2, 4, 1, false>("gru.zr", in_zr, ezr, wzr, conv_tables(3, 1, 2, 8), B, 64, H, W, 4, st)));
            IMVS_TRY((mma_conv<48, 16, 2, 4, 1, false>("gru.q", in_q, eq, wq, conv_tables(3, 1, 2, 8), B, 32, H, W, 2, st)));
            break;
        case 2:      // 4-row tiles
            IMVS_TRY((mma_conv<48, 32, 1, 4, 1, false>("gru.zr", in_zr, ezr, wzr, conv_tables(3, 1, 2, 4), B, 64, H, W, 2, st)));
            IMVS_TRY((mma_conv<48, 32, 1, 4, 1, false>("gru.q", in_q, eq, wq, conv_tables(3, 1, 2, 4), B, 32, H, W, 1, st)));
            break;
        case 3:      // both
            IMVS_TRY((mma_conv<48, 16, 1, 4, 1, false>("gru.zr", in_zr, ezr, wzr, conv_tables(3, 1, 2, 4), B, 64, H, W, 4, st)));
            IMVS_TRY((mma_conv<48, 16, 1, 4, 1, false>("gru.q", in_q, eq, wq, conv_tables(3, 1, 2, 4), B, 32, H, W, 2, st)));
            break;
        default:
            IMVS_TRY((mma_conv<48, 32, 2, 4, 1, false>("gru.zr", in_zr, ezr, wzr, conv_tables(3, 1, 2, 8), B, 64, H, W, 2, st)));
            IMVS_TRY((mma_conv<48, 32, 2, 4, 1, false>("gru.q", in_q, eq, wq, conv_tables(3, 1, 2, 8), B, 32, H, W, 1, st)));
    }
    return 0;
}

extern "C" int imvs_depth_head(const imvs_weights* w, const float* hidden, float* nd_out, size_t nd_batch_stride,
                               size_t nd_pixel_stride, float* probability, float* conf, float* conf_logit, float* depth_out,
                               const float* depth_min, const float* depth_max, float* scratch,
                               int B, int H, int W, void* stream) {
    IMVS_REQUIRE(w && hidden && nd_out && scratch, "depth_head: null pointer");
    IMVS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "depth_head: bad shape");
    IMVS_REQUIRE(!depth_out || (depth_min && depth_max), "depth_head: depth_out needs depth_min/depth_max");
    IMVS_REQUIRE(nd_pixel_stride >= 1, "depth_head: nd_pixel_stride must be >= 1");
    ApiScope api_;
    cudaStream_t st = (cudaStream_t)stream;
    const bool want_conf = conf || conf_logit;
    const size_t P = (size_t)H * W;
    float* t = scratch;                          // [B][P][64]
    float* h1 = scratch + (size_t)B * P * 64;    // [B][P][64]
    float* logits = h1 + (size_t)B * P * 64;     // [B][P][256]
    // stacked weight [9][32][64]: channel block 0 = depth_head.0, block 1 = confidence_head.0; the
    // confidence block only runs when a confidence output is requested (itermvs.py:196-199)
    if (conv_passes() == 1 && tc5_enabled() && w->head_conv0.umma) {
        IMVS_TRY((tc5::launch<32, 64>("head.conv0(tcgen05)", in_nhwc(hidden, H, W, 32), tc5::PixNHWC{t, nullptr, nullptr, H, W, 64, 1},
                                      w->head_conv0.umma, 3, 2, B, H, W, tc5_error_flag(), st)));
    } else {
        const EpiNHWC e0{t, nullptr, nullptr, H, W, 64, 64, 1};
        const int nc = want_conf ? 2 : 1;
        switch (tune("HEAD", 2)) {
            case 1:
                IMVS_TRY((mma_conv<32, 16, 2, 4, 1, false>("head.conv0", in_nhwc(hidden, H, W, 32), e0, WSets::single(w->head_conv0),
                                                           conv_tables(3, 1, 2, 8), B, 64, H, W, 2 * nc, st)));
                break;
            case 2:
                IMVS_TRY((mma_conv<32, 32, 1, 4, 1, false>("head.conv0", in_nhwc(hidden, H, W, 32), e0, WSets::single(w->head_conv0),
                                                           conv_tables(3, 1, 2, 4), B, 64, H, W, nc, st)));
                break;
            case 3:
                IMVS_TRY((mma_conv<32, 16, 1, 4, 1, false>("head.conv0", in_nhwc(hidden, H, W, 32), e0, WSets::single(w->head_conv0),
                                                           conv_tables(3, 1, 2, 4), B, 64, H, W, 2 * nc, st)));
                break;
            default:
                IMVS_TRY((mma_conv<32, 32, 2, 4, 1, false>("head.conv0", in_nhwc(hidden, H, W, 32), e0, WSets::single(w->head_conv0),
                                                           conv_tables(3, 1, 2, 8), B, 64, H, W, nc, st)));
        }
    }
    // fp32-grade mode without a probability output (test mode): fc1 + fc2 + softmax / arg-max / window regression (+ the
    // confidence 1x1) as ONE tcgen05 kernel -- logits and the fc1 activation stay in TMEM / shared memory
#ifndef CUSIM        // (the CPU emulation of the test-suite has no tensor-core / TMEM model: it runs the unfused kernels)
    if (conv_passes() == 4 && !probability && w->head_fused && tune("HEADFUSED", 1)) {
        hf::Params hp;
        hp.t = t; hp.blob = w->head_fused; hp.conf_w = w->conf_fc; hp.conf_b = w->conf_fc_b;
        hp.nd_out = nd_out; hp.nd_bstride = nd_batch_stride; hp.nd_pstride = nd_pixel_stride;
        hp.conf = conf; hp.conf_logit = conf_logit; hp.depth_out = depth_out;
        hp.depth_min = depth_min; hp.depth_max = depth_max;
        hp.n_px = (int)((size_t)B * P); hp.P = (int)P;
        hp.err_flag = tc5_error_flag();
        IMVS_REQUIRE((size_t)B * P < 2147483647ull, "depth_head: too many pixels");
        IMVS_REQUIRE((reinterpret_cast<uintptr_t>(w->head_fused) & 15u) == 0, "depth_head: head_fused blob must be 16-byte aligned");
        return hf::launch(hp, st);
    }
#endif
    {
        const EpiNHWC e1{h1, nullptr, nullptr, H, W, 64, 64, 1};
        const EpiNHWC e2{logits, w->head_fc2_b, nullptr, H, W, 256, 256, 0};
        switch (tune("FC", 2)) {
            case 1:      // fc1 in 16-cout blocks, fc2 in 32-cout blocks
                IMVS_TRY((mma_conv<32, 16, 2, 4, 1, true>("head.fc1", in_nhwc(t, H, W, 64), e1, WSets::single(w->head_fc1), conv_tables(1, 1, 1, 8), B, 64, H, W, 4, st)));
                IMVS_TRY((mma_conv<64, 32, 2, 4, 1, true>("head.fc2", in_nhwc(h1, H, W, 64), e2, WSets::single(w->head_fc2), conv_tables(1, 1, 1, 8), B, 256, H, W, 8, st)));
                break;
            case 2:      // 4-row tiles
                IMVS_TRY((mma_conv<32, 64, 1, 4, 1, true>("head.fc1", in_nhwc(t, H, W, 64), e1, WSets::single(w->head_fc1), conv_tables(1, 1, 1, 4), B, 64, H, W, 1, st)));
                IMVS_TRY((mma_conv<64, 64, 1, 4, 1, true>("head.fc2", in_nhwc(h1, H, W, 64), e2, WSets::single(w->head_fc2), conv_tables(1, 1, 1, 4), B, 256, H, W, 4, st)));
                break;
            default:
                IMVS_TRY((mma_conv<32, 64, 2, 4, 1, true>("head.fc1", in_nhwc(t, H, W, 64), e1, WSets::single(w->head_fc1), conv_tables(1, 1, 1, 8), B, 64, H, W, 1, st)));
                IMVS_TRY((mma_conv<64, 64, 2, 4, 1, true>("head.fc2", in_nhwc(h1, H, W, 64), e2, WSets::single(w->head_fc2), conv_tables(1, 1, 1, 8), B, 256, H, W, 4, st)));
        }
    }
    RegressParams prm;
    prm.logits = logits; prm.t = t; prm.conf_w = w->conf_fc; prm.conf_b = w->conf_fc_b;
    prm.nd_out = nd_out; prm.nd_bstride = nd_batch_stride; prm.nd_pstride = nd_pixel_stride;
    prm.prob = probability; prm.conf = conf; prm.conf_logit = conf_logit; prm.depth_out = depth_out;
    prm.depth_min = depth_min; prm.depth_max = depth_max;
    prm.B = B; prm.P = (int)P;
    const size_t warps = (size_t)B * P;
    IMVS_CUDA(launch_k(softmax_regress_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, st, prm));
    return 0;
}
