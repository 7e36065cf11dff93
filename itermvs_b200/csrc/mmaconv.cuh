// Tensor-core implicit-GEMM convolution for the estimator's and FeatureNet's convolutions.
//
//   D[pixel, cout] = sum_{tap} sum_{cin} A[pixel shifted by tap, cin] * W[tap][cin][cout]
//
// * activations are channels-last (NHWC, channel count padded to a multiple of 8);
// * a CTA owns a TH x 16 output tile (TH = WARPS*MT rows, one m16 MMA row-tile = 16 consecutive x)
//   and NB output channels; it copies the input tile WITH HALO once into shared memory with
//   cp.async (zero-filled outside the image), so every tap of the 3x3 / dilated / strided /
//   transposed stencil is the same smem tile read at a shifted address -- no im2col copy, each
//   input element is fetched from L2 once per CTA instead of once per tap;
// * weights stream through a 3-stage cp.async ring, one tap ([Cin][NB]) per stage, loaded two taps
//   ahead (small layers keep all taps resident: WALL);
// * math: mma.sync.m16n8k8 TF32 with fp32 accumulation.
//     PASSES = 1: A rounded to TF32 (round-to-nearest) at fragment load, weights pre-rounded on
//                 the host -- the precision of the reference's own cuDNN path on Ampere+
//                 (torch.backends.cudnn.allow_tf32 defaults to True);
//     PASSES = 3: error-compensated split computed in registers from the fp32 operands
//                 (x_hi = x & ~0x1fff, x_lo = x - x_hi):  lo*hi + hi*lo + hi*hi  -> fp32-grade
//                 (relative error ~2^-21), no extra shared memory.
//     PASSES = 4: the same 3-product compensation on the FP16 tensor-core path (mma.sync.m16n8k16,
//                 twice the TF32 rate): x_hi = fp16(x), x_lo = fp16(x - x_hi) -> 22 significant bits
//                 for |x| < 65504 (saturating conversion above), absolute floor 2^-25 below 2^-14.
//                 Activations are split in registers from the fp32 tile (splitting the tile once in
//                 shared memory was measured slower: the extra pass costs more than the ALU it saves);
//                 weights come pre-split from the host ([slice][CinK/2][Cout] uint2).
// * the stencil is a runtime tap table (up to 4 variants per launch), so regular, dilated, strided
//   and the four output parities of a transposed convolution share one kernel / one launch.
#pragma once
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace imvs {

struct TapTable {
    int n;                  // taps
    int dy[9], dx[9];       // input offset of each tap relative to (oy*STRIDE, ox*STRIDE)
    int widx[9];            // weight slice of each tap in the packed [slice][CINP][COUT] array
    int dy_min, dx_min;     // tile origin offset
    int IH, IW;             // smem input tile extent for a TH x 16 output tile
    unsigned iw_magic;      // floor(2^22 / IW) + 1: slot / IW == (slot * iw_magic) >> 22 for every slot of a tile
};

struct TapTables {          // launch variants (blockIdx.z % count)
    TapTable t[4];
    int count;
    int max_slots;          // max over the variants of IH * IW (the shared-memory tile is carved for the largest)
};

inline TapTable make_taps_conv(int ks, int stride, int dil, int TH) {
    TapTable t{};
    const int pad = dil * (ks - 1) / 2;
    t.n = ks * ks;
    for (int ky = 0; ky < ks; ++ky)
        for (int kx = 0; kx < ks; ++kx) {
            int i = ky * ks + kx;
            t.dy[i] = ky * dil - pad; t.dx[i] = kx * dil - pad; t.widx[i] = i;
        }
    t.dy_min = -pad; t.dx_min = -pad;
    t.IH = (TH - 1) * stride + (ks - 1) * dil + 1;
    t.IW = 15 * stride + (ks - 1) * dil + 1;
    t.iw_magic = (1u << 22) / (unsigned)t.IW + 1u;
    return t;
}

inline TapTables conv_tables(int ks, int stride, int dil, int TH) {
    TapTables tt{};
    tt.t[0] = make_taps_conv(ks, stride, dil, TH);
    tt.count = 1;
    tt.max_slots = tt.t[0].IH * tt.t[0].IW;
    return tt;
}

// ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1): the outputs (2*iy + a, 2*ix + b) of
// parity (a, b) as a stride-1 "convolution" over the INPUT grid (y = 2*iy - 1 + ky):
//   a = 0: ky = 1 (dy 0)          a = 1: ky = 0 (dy +1), ky = 2 (dy 0)        likewise in x.
// Variant v = 2*a + b.
inline TapTables tconv_tables(int TH) {
    TapTables tt{};
    for (int v = 0; v < 4; ++v) {
        const int a = v >> 1, b = v & 1;
        TapTable& t = tt.t[v];
        int kys[2], dys[2], nky, kxs[2], dxs[2], nkx;
        if (a == 0) { nky = 1; kys[0] = 1; dys[0] = 0; } else { nky = 2; kys[0] = 0; dys[0] = 1; kys[1] = 2; dys[1] = 0; }
        if (b == 0) { nkx = 1; kxs[0] = 1; dxs[0] = 0; } else { nkx = 2; kxs[0] = 0; dxs[0] = 1; kxs[1] = 2; dxs[1] = 0; }
        t.n = 0;
        for (int i = 0; i < nky; ++i)
            for (int j = 0; j < nkx; ++j) {
                t.dy[t.n] = dys[i]; t.dx[t.n] = dxs[j]; t.widx[t.n] = kys[i] * 3 + kxs[j]; t.n++;
            }
        t.dy_min = 0; t.dx_min = 0;
        t.IH = TH + 1; t.IW = 17;
        t.iw_magic = (1u << 22) / 17u + 1u;
    }
    tt.count = 4;
    tt.max_slots = (TH + 1) * 17;
    return tt;
}

// Under CUSIM (tests/cusim: the CPU emulation the test-suite builds, never the product) the PTX helpers of this
// header have plain C++ bodies with the documented fragment layouts; nvcc never sees those branches.
__device__ __forceinline__ uint32_t f2tf32(float x) {
#ifdef CUSIM
    return cusim::cvt_rna_tf32(x);
#else
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
#endif
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
#ifdef CUSIM
    cusim::mma_m16n8k8_tf32(c, a0, a1, a2, a3, b0, b1);
    return;
#endif
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// x = hi + lo with hi = upper 19 bits (TF32 by truncation), lo = TF32-truncated remainder
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}

__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                        uint32_t b0, uint32_t b1) {
#ifdef CUSIM
    cusim::mma_m16n8k16_f16(c, a0, a1, a2, a3, b0, b1);
    return;
#endif
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// m16n8k8 FP16: A = 2 regs (row g / g+8, k = 2t, 2t+1), B = 1 reg (k = 2t, 2t+1; n = g) -- 8-channel inputs
__device__ __forceinline__ void mma_f16_k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
#ifdef CUSIM
    cusim::mma_m16n8k8_f16(c, a0, a1, b0);
    return;
#endif
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}

// (v.x, v.y) = channels (k, k+1) -> hi = half2(fp16(v.x), fp16(v.y)) [k in the low half], lo = the remainder
__device__ __forceinline__ void split_f16(float2 v, uint32_t& hi, uint32_t& lo) {
#ifdef IMVS_FAKE_SPLIT      // timing experiment only (wrong results): what the loop costs when the tile is already split
    hi = __float_as_uint(v.x); lo = __float_as_uint(v.y);
    return;
#endif
#ifdef CUSIM
    hi = cusim::cvt_f16x2(v.y, v.x);
    const float2 hs = cusim::unpack_f16x2(hi);
    lo = cusim::cvt_f16x2(v.y - hs.y, v.x - hs.x);
    return;
#endif
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.y), "f"(v.x));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v.y - hf.y), "f"(v.x - hf.x));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
#ifdef CUSIM
    cusim::cp_async16(smem, gmem, valid);     // completes at once: a legal outcome of an asynchronous copy
    return;
#endif
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;            // src-size 0 -> the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// WALL_ = true: all taps' weights are staged up front (small layers: no per-tap barrier);
// false: one tap per stage through a 3-deep ring (large layers).
template <int CINP_, int NB_, int MT_, int WARPS_, int STRIDE_, int PASSES_, bool WALL_ = false>
struct MmaCfg {
    static constexpr int CINP = CINP_, NB = NB_, MT = MT_, WARPS = WARPS_, STRIDE = STRIDE_, PASSES = PASSES_;
    static constexpr bool WALL = WALL_;
    static constexpr int TH = WARPS * MT;
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int NT = NB / 8;                       // n-tiles per CTA
    static constexpr bool F16 = PASSES == 4 || PASSES == 5;
    // PASSES = 5: mode 4 for 8-channel inputs on m16n8k8 (one k-step of 8 instead of a half-empty k-step of 16:
    // half the A fragments, splits and tensor work); same packed weights, only the first 4 channel-pair rows are read
    static constexpr bool K8 = PASSES == 5;
    static_assert(!K8 || CINP == 8, "the k8 variant serves 8-channel inputs");
    static constexpr int CINK = F16 ? (K8 ? 8 : (CINP + 15) / 16 * 16) : CINP;   // K extent staged in smem (zero-filled beyond CINP)
    static constexpr int WROWS = (CINP + 15) / 16 * 8;      // FP16: channel-pair rows per tap in the PACKED weights
    // smem channel pitch (floats): TF32: CP/4 odd -> conflict-free LDS.32 A loads; FP16: CP = 8 or 24 mod 32
    // -> conflict-free LDS.64 A loads (k8: 8 for stride 1, 12 for stride 2)
    static constexpr int CP = K8 ? (STRIDE == 2 ? 12 : 8) : (F16 ? CINK + 8 : CINP + 4);
    // smem cout pitch: TF32: floats, = 8 or 24 mod 32; FP16: uint2 (hi, lo) units, 2*NP = 8 or 24 mod 32
    static constexpr int NP = F16 ? NB + 4 : NB + 8 + (NB == 8 ? 8 : 0);
    static constexpr int KSTEPS = F16 ? CINK / 16 : CINP / 8;
    static constexpr int WBUF = CINK * NP;                  // floats per weight stage (FP16: CINK/2 rows of NP uint2)
    static constexpr int RING = 3;
    static_assert(CINP % 8 == 0 && NB % 8 == 0, "channel padding");
    static_assert(PASSES == 1 || PASSES == 3 || PASSES == 4 || PASSES == 5, "PASSES");
    static size_t smem_bytes(const TapTables& tt) {
        size_t tile = 0, taps = 0;
        for (int v = 0; v < tt.count; ++v) {
            tile = std::max(tile, (size_t)tt.t[v].IH * tt.t[v].IW * CP);
            taps = std::max(taps, (size_t)tt.t[v].n);
        }
        return sizeof(float) * (tile + (WALL ? taps : (size_t)RING) * WBUF);
    }
};

// ---- input functors ----------------------------------------------------------------------------
// async functors give the address of channels [4*c4, 4*c4+4) at (n, iy, ix) (valid=false outside)
struct InNHWC {
    static constexpr bool kAsync = true;
    const float* p;
    int H, W, C;            // C = stored channel count (>= CINP used by the kernel)
    size_t nstride;         // floats between consecutive images n (H*W*C when dense)
    bool fits(int) const { return (double)H * W * C < 4294967296.0; }
    // per-CTA view of image n: the 64-bit part of the address is formed once, the staging loop adds 32-bit offsets
    // (one image holds < 2^32 floats, checked by the launcher)
    struct Image {
        const float* p;
        int H, W, C;
        __device__ __forceinline__ const float* ptr4(int iy, int ix, int c4, bool& valid) const {
            valid = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
            return p + (valid ? (unsigned)((iy * W + ix) * C + 4 * c4) : 0u);
        }
    };
    __device__ __forceinline__ Image image(int n) const { return Image{p + (size_t)n * nstride, H, W, C}; }
};
inline InNHWC in_nhwc(const float* p, int H, int W, int C, size_t nstride = 0) {
    return InNHWC{p, H, W, C, nstride ? nstride : (size_t)H * W * C};
}

struct InNHWC2 {            // channels [0,CA) from a, then [CA, CA+CB) from b (both NHWC, multiples of 4)
    static constexpr bool kAsync = true;
    const float* a;
    const float* b;
    int H, W, CA, CBc;
    bool fits(int) const { return (double)H * W * (CA > CBc ? CA : CBc) < 4294967296.0; }
    struct Image {
        const float* a;
        const float* b;
        int H, W, CA, CBc;
        __device__ __forceinline__ const float* ptr4(int iy, int ix, int c4, bool& valid) const {
            valid = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
            const unsigned pix = valid ? (unsigned)(iy * W + ix) : 0u;
            const int c = 4 * c4;
            return c < CA ? a + (pix * (unsigned)CA + (unsigned)c) : b + (pix * (unsigned)CBc + (unsigned)(c - CA));
        }
    };
    __device__ __forceinline__ Image image(int n) const {
        const size_t px = (size_t)n * H * W;
        return Image{a + px * CA, b + px * CBc, H, W, CA, CBc};
    }
};

struct InNCHW3 {            // 3-channel planar image [N][3][H][W] -> channels (r,g,b,0,0,0,0,0); synchronous staging
    static constexpr bool kAsync = false;
    const float* p;
    int H, W;
    bool fits(int) const { return true; }
    struct Image {
        const float* p;
        int H, W;
        __device__ __forceinline__ float4 load4(int iy, int ix, int c4) const {
            if (c4 != 0 || iy < 0 || iy >= H || ix < 0 || ix >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
            const size_t plane = (size_t)H * W, o = (size_t)iy * W + ix;
            return make_float4(ldg(p + o), ldg(p + o + plane), ldg(p + o + 2 * plane), 0.f);
        }
    };
    __device__ __forceinline__ Image image(int n) const { return Image{p + (size_t)n * 3 * H * W, H, W}; }
};

struct MmaWeightSel {       // image n of a batched launch picks one of up to three weight sets
    const float* w[3];      // [slice][CINP][cout_total]: TF32-rounded (PASSES 1) or plain fp32 (PASSES 3);
                            // PASSES 4: [slice][CINK/2][cout_total] uint2 = (hi half2, lo half2) of a channel pair
    int period, split1, split2;
    int* status;            // device status word (imvs_device_status): bit 1 <- an accumulator left the fp16 range in mode 4
    __device__ __forceinline__ const float* pick(int n) const {
        if (period == 1) return w[0];
        const int r = n % period;
        return r < split1 ? w[0] : (r < split2 ? w[1] : w[2]);
    }
};

// grid: (ceil(Wout/16), ceil(Hout/TH), N * ncb * variants); block: THREADS; dyn smem: Cfg::smem_bytes
// Epi::row<NT>(n, oy, ox, co0, t, v, variant): v[2*j], v[2*j+1] = couts co0 + 8*j + 2*t, +1 of pixel
// (oy, ox) for this thread's quad position t = lane & 3; called for every (row-tile, half) with
// FULL-WARP uniformity (pixels outside the image included: the functor masks its stores), so it may
// shuffle within quads.
template <class Cfg, class In, class Epi>
__global__ void __launch_bounds__(Cfg::THREADS)
mma_conv_kernel(const In in, const Epi epi, const MmaWeightSel wsel, const TapTables tabs, int cout_total, int Hout, int Wout, int ncb) {
    constexpr int CP = Cfg::CP, NP = Cfg::NP, CINP = Cfg::CINP, NB = Cfg::NB, MT = Cfg::MT, NT = Cfg::NT;
    extern __shared__ __align__(16) float smem[];
    // (uniform special cases: most launches have one stencil variant and / or one cout block -- no runtime divisions)
    const int variant = tabs.count == 1 ? 0 : blockIdx.z % tabs.count;
    const int ncbz = tabs.count == 1 ? blockIdx.z : blockIdx.z / tabs.count;
    const int n = ncb == 1 ? ncbz : ncbz / ncb, cb = ncb == 1 ? 0 : ncbz % ncb;
    const TapTable& taps = tabs.t[variant];
    const int tile_floats = tabs.max_slots * CP;
    float* sA = smem;
    float* sW = smem + tile_floats;
    const int oy0 = blockIdx.y * Cfg::TH, ox0 = blockIdx.x * 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const float* __restrict__ wg = wsel.pick(n);
    const int ntaps = taps.n;

    auto issue_weights = [&](int tap, int buf) {
        float* dst = sW + buf * Cfg::WBUF;
        if constexpr (Cfg::F16) {      // rows = channel pairs, NB (hi, lo) uint2 per row
            const float* src = wg + 2 * (((size_t)taps.widx[tap] * Cfg::WROWS) * cout_total + cb * NB);
            constexpr int Q = NB / 2;
            for (int i = tid; i < (Cfg::CINK / 2) * Q; i += Cfg::THREADS) {
                const int k = i / Q, q = i % Q;
                cp_async16(dst + 2 * (k * NP) + 4 * q, src + (unsigned)(2 * (k * cout_total) + 4 * q), true);
            }
        } else {
            const float* src = wg + ((size_t)taps.widx[tap] * CINP) * cout_total + cb * NB;
            constexpr int Q = NB / 4;
            for (int i = tid; i < CINP * Q; i += Cfg::THREADS) {
                const int k = i / Q, q = i % Q;
                cp_async16(dst + k * NP + 4 * q, src + (unsigned)(k * cout_total + 4 * q), true);
            }
        }
    };

    // weights are constants of the forward pass: their first stages are requested BEFORE the grid-dependency
    // wait, so they arrive while the producer of the input is still draining (programmatic dependent launch)
    pdl_trigger();
    if constexpr (Cfg::WALL) {
        for (int tp = 0; tp < ntaps; ++tp) issue_weights(tp, tp);
    } else {
        issue_weights(0, 0);
    }
    pdl_wait();
    {   // input tile (with halo)
        const int iy0 = oy0 * Cfg::STRIDE + taps.dy_min, ix0 = ox0 * Cfg::STRIDE + taps.dx_min;
        constexpr int C4 = Cfg::CINK / 4;
        const int total = taps.IH * taps.IW * C4;
        const unsigned magic = taps.iw_magic;
        const int IW = taps.IW;
        const auto img = in.image(n);
        for (int i = tid; i < total; i += Cfg::THREADS) {
            const int slot = i / C4, c4 = i % C4;                      // C4 is a compile-time constant
            const int row = (int)(((unsigned)slot * magic) >> 22);     // slot / IW without the runtime division
            const int iy = iy0 + row, ix = ix0 + (slot - row * IW);
            if constexpr (In::kAsync) {
                bool valid;
                const float* src = img.ptr4(iy, ix, c4 < CINP / 4 ? c4 : 0, valid);
                cp_async16(sA + (size_t)slot * CP + 4 * c4, src, valid && c4 < CINP / 4);
            } else {
                *reinterpret_cast<float4*>(sA + (size_t)slot * CP + 4 * c4) = img.load4(iy, ix, c4);
            }
        }
    }
    if constexpr (Cfg::WALL) {
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    } else {
        cp_async_commit();              // group 0: tap-0 weights + tile
        if (ntaps > 1) issue_weights(1, 1);
        cp_async_commit();              // group 1: tap-1 weights
    }

    // FP16 3-product mode: the products of one output tile go to SEPARATE accumulator sets when the warp owns
    // few tiles (set 0: hi*hi, set 1: cross terms, or one set per product), summed once in the epilogue.  With
    // one set, consecutive HMMAs on a tile are MT*NT (= 4 for a 16-channel layer) instructions apart -- about the
    // HMMA latency -- and the warps sat in fixed-latency dependency stalls (ncu: 'wait' 30 %, tensor pipe 34 %).
    constexpr int NACC = !Cfg::F16 ? 1 : (MT * NT <= 4 ? 3 : (MT * NT <= 8 ? 2 : 1));
    constexpr int ACC_LH = NACC == 3 ? 1 : NACC - 1, ACC_HL = NACC == 3 ? 2 : NACC - 1;     // lo*hi, hi*lo; hi*hi -> set 0
    float accs[NACC][MT][NT][4];
    float (&acc)[MT][NT][4] = accs[0];
#pragma unroll
    for (int s = 0; s < NACC; ++s)
#pragma unroll
        for (int r = 0; r < MT; ++r)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) accs[s][r][j][q] = 0.f;

    int slot_base[MT];              // smem offset of this lane's first pixel (x = g) of row-tile r at tap offset (0, 0)
#pragma unroll
    for (int r = 0; r < MT; ++r) slot_base[r] = ((warp * MT + r) * Cfg::STRIDE * taps.IW + g * Cfg::STRIDE) * CP;
    int ring_cur = 0, ring_fill = 2 % Cfg::RING;       // ring slots of this tap / of the tap loaded two ahead
    for (int tap = 0; tap < ntaps; ++tap) {
        int wsel_buf = tap;
        if constexpr (!Cfg::WALL) {
            cp_async_wait<1>();            // everything but the newest group has landed: tile + this tap's weights
            __syncthreads();               // ... for all threads; also: everyone is done with tap-1's buffer
            if (tap + 2 < ntaps) issue_weights(tap + 2, ring_fill);
            cp_async_commit();
            wsel_buf = ring_cur;
            ring_cur = ring_cur + 1 == Cfg::RING ? 0 : ring_cur + 1;
            ring_fill = ring_fill + 1 == Cfg::RING ? 0 : ring_fill + 1;
        }
        const float* wb = sW + wsel_buf * Cfg::WBUF;
        const int ry = taps.dy[tap] - taps.dy_min, rx = taps.dx[tap] - taps.dx_min;
        const int tap_off = (ry * taps.IW + rx) * CP;              // the tap = the same tile read at a shifted address
        int slot0[MT], slot1[MT];
#pragma unroll
        for (int r = 0; r < MT; ++r) {
            slot0[r] = slot_base[r] + tap_off;
            slot1[r] = slot0[r] + 8 * Cfg::STRIDE * CP;
        }
        if constexpr (Cfg::K8) {
            const uint2* wb2 = reinterpret_cast<const uint2*>(wb);
            uint32_t a[MT][2], al[MT][2];
#pragma unroll
            for (int r = 0; r < MT; ++r) {
                split_f16(*reinterpret_cast<const float2*>(sA + slot0[r] + 2 * t), a[r][0], al[r][0]);
                split_f16(*reinterpret_cast<const float2*>(sA + slot1[r] + 2 * t), a[r][1], al[r][1]);
            }
            uint2 w0[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) w0[j] = wb2[t * NP + 8 * j + g];
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int r = 0; r < MT; ++r) mma_f16_k8(accs[ACC_LH][r][j], al[r][0], al[r][1], w0[j].x);
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int r = 0; r < MT; ++r) mma_f16_k8(accs[ACC_HL][r][j], a[r][0], a[r][1], w0[j].y);
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int r = 0; r < MT; ++r) mma_f16_k8(acc[r][j], a[r][0], a[r][1], w0[j].x);
        } else if constexpr (Cfg::F16) {
            const uint2* wb2 = reinterpret_cast<const uint2*>(wb);
#pragma unroll
            for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
                const int k0 = ks * 16 + 2 * t;
                uint32_t a[MT][4], al[MT][4];
#pragma unroll
                for (int r = 0; r < MT; ++r) {
                    split_f16(*reinterpret_cast<const float2*>(sA + slot0[r] + k0), a[r][0], al[r][0]);
                    split_f16(*reinterpret_cast<const float2*>(sA + slot1[r] + k0), a[r][1], al[r][1]);
                    split_f16(*reinterpret_cast<const float2*>(sA + slot0[r] + k0 + 8), a[r][2], al[r][2]);
                    split_f16(*reinterpret_cast<const float2*>(sA + slot1[r] + k0 + 8), a[r][3], al[r][3]);
                }
                // the three products of one accumulator tile are issued MT*NT MMAs apart: back-to-back they
                // serialise on the accumulator (HMMA latency), which left the tensor pipe 26 % busy
                uint2 w0[NT], w1[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    w0[j] = wb2[(ks * 8 + t) * NP + 8 * j + g];
                    w1[j] = wb2[(ks * 8 + t + 4) * NP + 8 * j + g];
                }
#pragma unroll
#ifndef IMVS_EXP_ONE_PRODUCT     // (timing experiment: hi*hi only -- wrong results)
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int r = 0; r < MT; ++r) mma_f16(accs[ACC_LH][r][j], al[r][0], al[r][1], al[r][2], al[r][3], w0[j].x, w1[j].x);
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int r = 0; r < MT; ++r) mma_f16(accs[ACC_HL][r][j], a[r][0], a[r][1], a[r][2], a[r][3], w0[j].y, w1[j].y);
#endif
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int r = 0; r < MT; ++r) mma_f16(acc[r][j], a[r][0], a[r][1], a[r][2], a[r][3], w0[j].x, w1[j].x);
            }
        } else {
#pragma unroll
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
            const int k0 = ks * 8;
            uint32_t a[MT][4], al[MT][4];
#pragma unroll
            for (int r = 0; r < MT; ++r) {
                const float f0 = sA[slot0[r] + k0 + t], f1 = sA[slot1[r] + k0 + t];
                const float f2 = sA[slot0[r] + k0 + t + 4], f3 = sA[slot1[r] + k0 + t + 4];
                if constexpr (Cfg::PASSES == 3) {
                    split_tf32(f0, a[r][0], al[r][0]); split_tf32(f1, a[r][1], al[r][1]);
                    split_tf32(f2, a[r][2], al[r][2]); split_tf32(f3, a[r][3], al[r][3]);
                } else {
                    a[r][0] = f2tf32(f0); a[r][1] = f2tf32(f1); a[r][2] = f2tf32(f2); a[r][3] = f2tf32(f3);
                }
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const float w0 = wb[(k0 + t) * NP + 8 * j + g], w1 = wb[(k0 + t + 4) * NP + 8 * j + g];
                if constexpr (Cfg::PASSES == 3) {
                    uint32_t b0, b1, bl0, bl1;
                    split_tf32(w0, b0, bl0);
                    split_tf32(w1, b1, bl1);
#pragma unroll
                    for (int r = 0; r < MT; ++r) {
                        mma_tf32(acc[r][j], al[r][0], al[r][1], al[r][2], al[r][3], b0, b1);
                        mma_tf32(acc[r][j], a[r][0], a[r][1], a[r][2], a[r][3], bl0, bl1);
                        mma_tf32(acc[r][j], a[r][0], a[r][1], a[r][2], a[r][3], b0, b1);
                    }
                } else {
                    const uint32_t b0 = __float_as_uint(w0), b1 = __float_as_uint(w1);     // pre-rounded on the host
#pragma unroll
                    for (int r = 0; r < MT; ++r) mma_tf32(acc[r][j], a[r][0], a[r][1], a[r][2], a[r][3], b0, b1);
                }
            }
        }
        }
    }

    if constexpr (NACC > 1) {           // small terms first, then onto the hi*hi sums
#pragma unroll
        for (int r = 0; r < MT; ++r)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float cross = accs[NACC - 1][r][j][q];
                    if constexpr (NACC == 3) cross += accs[1][r][j][q];
                    acc[r][j][q] += cross;
                }
    }
    // fp16-split modes: a value beyond +-65504 would saturate in the NEXT layer's split (cvt.rn.satfinite) -- raise the flag
    if constexpr (Cfg::PASSES >= 4) {
        float amax = 0.f;
#pragma unroll
        for (int r = 0; r < MT; ++r)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) amax = fmaxf(amax, fabsf(acc[r][j][q]));
        if (!(amax <= 65504.f) && wsel.status) atomicOr(wsel.status, 2);
    }
    // epilogue: thread holds, per row-tile r and half h, couts {8j + 2t, 8j + 2t + 1} of pixel x = g + 8h
#pragma unroll
    for (int r = 0; r < MT; ++r) {
        const int oy = oy0 + warp * MT + r;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float v[2 * NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) { v[2 * j] = acc[r][j][2 * h]; v[2 * j + 1] = acc[r][j][2 * h + 1]; }
            epi.template row<NT>(n, oy, ox0 + g + 8 * h, cb * NB, t, v, variant);
        }
    }
}

template <class Cfg, class In, class Epi>
int launch_mma_conv(const char* name, const In& in, const Epi& epi, const MmaWeightSel& wsel,
                    const TapTables& tabs, int N, int cout_total, int Hout, int Wout, int ncb, cudaStream_t st) {
    for (int i = 0; i < 3; ++i) IMVS_REQUIRE(wsel.w[i], "%s: null weights", name);
    IMVS_REQUIRE(cout_total % 4 == 0 && ncb * Cfg::NB <= cout_total, "%s: cout_total=%d must be a multiple of 4 and >= %d", name,
                 cout_total, ncb * Cfg::NB);
    // (a persistent, tile-double-buffered variant of this kernel for the small layers was measured in round 1
    //  and was 5% SLOWER end to end: the doubled tile buffer halves the resident warps and these layers are
    //  issue/latency bound, not load bound -- see profiles/README.md)
    IMVS_REQUIRE(in.fits(N), "%s: input too large for 32-bit in-image offsets", name);
    const size_t smem = Cfg::smem_bytes(tabs);
    IMVS_REQUIRE(smem <= 220 * 1024, "%s: %zu bytes of shared memory needed", name, smem);
    auto kern = mma_conv_kernel<Cfg, In, Epi>;
    static int smem_ok = 0;
    IMVS_TRY(ensure_dynamic_smem(kern, smem, &smem_ok));
    dim3 grid(cdiv(Wout, 16), cdiv(Hout, Cfg::TH), N * ncb * tabs.count);
    IMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "%s: grid too large", name);
    if (launch_k(kern, grid, dim3(Cfg::THREADS), smem, st, in, epi, wsel, tabs, cout_total, Hout, Wout, ncb) != cudaSuccess)
        return fail("launch of %s failed: %s", name, cudaGetErrorString(cudaGetLastError()));
    return 0;
}

int tune(const char* name, int def);   // IMVS_TUNE_<NAME> experiment switch, defined in warp.cu
int* tc5_error_flag();                 // the device status word (warp.cu), or nullptr
int conv_passes();       // process-wide precision switch (imvs_set_conv_passes), defined in warp.cu

struct WSets {           // host-side: up to three packed weights + the slice -> set mapping
    imvs_wpair w[3];
    int period, split1, split2;
    static WSets single(imvs_wpair p) { return WSets{{p, p, p}, 1, 1, 1}; }
};

// precision dispatch: one call site, both instantiations
template <int CINP, int NB, int MT, int WARPS, int STRIDE, bool WALL, class In, class Epi>
int mma_conv(const char* name, const In& in, const Epi& epi, const WSets& ws, const TapTables& tabs, int N,
             int cout_total, int Hout, int Wout, int ncb, cudaStream_t st) {
    MmaWeightSel sel;
    sel.period = ws.period; sel.split1 = ws.split1; sel.split2 = ws.split2;
    sel.status = tc5_error_flag();
    if (conv_passes() == 4) {
        for (int i = 0; i < 3; ++i) sel.w[i] = static_cast<const float*>(ws.w[i].f16x3);
        if constexpr (CINP == 8) {
            if (tune("K8", 1))
                return launch_mma_conv<MmaCfg<CINP, NB, MT, WARPS, STRIDE, 5, WALL>>(name, in, epi, sel, tabs, N, cout_total, Hout, Wout, ncb, st);
        }
        return launch_mma_conv<MmaCfg<CINP, NB, MT, WARPS, STRIDE, 4, WALL>>(name, in, epi, sel, tabs, N, cout_total, Hout, Wout, ncb, st);
    }
    if (conv_passes() == 3) {
        for (int i = 0; i < 3; ++i) sel.w[i] = ws.w[i].fp32;
        return launch_mma_conv<MmaCfg<CINP, NB, MT, WARPS, STRIDE, 3, WALL>>(name, in, epi, sel, tabs, N, cout_total, Hout, Wout, ncb, st);
    }
    for (int i = 0; i < 3; ++i) sel.w[i] = ws.w[i].tf32;
    return launch_mma_conv<MmaCfg<CINP, NB, MT, WARPS, STRIDE, 1, WALL>>(name, in, epi, sel, tabs, N, cout_total, Hout, Wout, ncb, st);
}

// ---- common epilogues (NHWC outputs) --------------------------------------------------------------
// out[n][oy][ox][co] (+bias) (+residual) (+relu); channels >= Cvalid are not written
struct EpiNHWC {
    float* out;
    const float* bias;       // [cout] or null
    const float* residual;   // NHWC same shape as out, or null
    int H, W, C;             // C = stored channels of out (and residual)
    int Cvalid;              // couts actually produced (<= C)
    int relu;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT], int) const {
        if (oy >= H || ox >= W) return;
        const size_t base = (((size_t)n * H + oy) * W + ox) * C;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = co0 + 8 * j + 2 * t;
            if (co >= Cvalid) continue;
            float a = v[2 * j], b = v[2 * j + 1];
            if (bias) { a += ldg(bias + co); b += (co + 1 < Cvalid) ? ldg(bias + co + 1) : 0.f; }
            if (residual) { a += ldg(residual + base + co); b += (co + 1 < Cvalid) ? ldg(residual + base + co + 1) : 0.f; }
            if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            if (co + 1 < Cvalid) *reinterpret_cast<float2*>(out + base + co) = make_float2(a, b);
            else out[base + co] = a;
        }
    }
};

// transposed-conv parity variant v = 2a + b (see tconv_tables): grid position (oy, ox) -> output
// (2*oy + a, 2*ox + b), plus the U-Net skip connection (itermvs.py:374-377)
struct EpiTconvNHWC {
    float* out;
    const float* skip;
    int Hin, Win, C;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT], int variant) const {
        if (oy >= Hin || ox >= Win) return;
        const int a = variant >> 1, b = variant & 1;
        const size_t base = (((size_t)n * 2 * Hin + 2 * oy + a) * (2 * Win) + 2 * ox + b) * C;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = co0 + 8 * j + 2 * t;
            if (co >= C) continue;
            const float2 s = ldg2(skip + base + co);
            *reinterpret_cast<float2*>(out + base + co) = make_float2(v[2 * j] + s.x, v[2 * j + 1] + s.y);
        }
    }
};

}  // namespace imvs
