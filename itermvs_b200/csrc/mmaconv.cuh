// Tensor-core implicit-GEMM convolution for the estimator's and FeatureNet's convolutions.
//
//   D[pixel, cout] = sum_{tap} sum_{cin} A[pixel shifted by tap, cin] * W[tap][cin][cout]
//
// * activations are channels-last (NHWC, channel count padded to a multiple of 8);
// * a CTA owns a TH x 16 output tile (TH = WARPS*MT rows, one m16 MMA row-tile = 16 consecutive x)
//   and NB output channels; it stages the input tile WITH HALO once in shared memory (rounded to
//   TF32, round-to-nearest), so every tap of the 3x3 / dilated / strided / transposed stencil is the
//   same smem tile read at a shifted address -- no im2col copy, each input element is fetched from
//   L2 once per CTA instead of once per tap;
// * weights stream through a double-buffered smem ring, one tap ([Cin][NB]) per stage, cp.async;
// * math: mma.sync.m16n8k8 TF32 with fp32 accumulation.  PASSES = 3 adds the error-compensated
//   split (a = a_hi + a_lo, w = w_hi + w_lo; hi*hi + hi*lo + lo*hi) for fp32-grade results.
//   The stock PyTorch/cuDNN GPU path of the reference runs its convolutions in TF32 as well
//   (torch.backends.cudnn.allow_tf32 defaults to True), so PASSES = 1 is the reference's own
//   GPU precision; the parity tests state which mode they ran.
// * the stencil is a runtime tap table, so regular, dilated, strided and (per output parity)
//   transposed convolutions share one kernel.
#pragma once
#include "common.cuh"

namespace imvs {

struct TapTable {
    int n;                  // taps
    int dy[9], dx[9];       // input offset of each tap relative to (oy*STRIDE, ox*STRIDE)
    int widx[9];            // weight slice of each tap in the packed [slice][CINP][COUT] array
    int dy_min, dx_min;     // tile origin offset
    int IH, IW;             // smem input tile extent for a TH x 16 output tile
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
    return t;
}

// ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1), output parity (a, b): the outputs
// (2*iy + a, 2*ix + b) as a stride-1 "convolution" over the INPUT grid (y = 2*iy - 1 + ky):
//   a = 0: ky = 1 (dy 0)          a = 1: ky = 0 (dy +1), ky = 2 (dy 0)        likewise in x
inline TapTable make_taps_tconv(int a, int b, int TH) {
    TapTable t{};
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
    return t;
}

__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// WALL_ = true: all taps' weights are staged up front (small layers: no per-tap barrier);
// false: one tap per stage through a double-buffered ring (large layers).
template <int CINP_, int NB_, int MT_, int WARPS_, int STRIDE_, int PASSES_, bool WALL_ = false>
struct MmaCfg {
    static constexpr int CINP = CINP_, NB = NB_, MT = MT_, WARPS = WARPS_, STRIDE = STRIDE_, PASSES = PASSES_;
    static constexpr bool WALL = WALL_;
    static constexpr int TH = WARPS * MT;
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int NT = NB / 8;                       // n-tiles per CTA
    static constexpr int CP = CINP + 4;                     // smem channel pitch (floats): CP/4 odd -> conflict-free A loads
    static constexpr int NP = NB + 8 + (NB == 8 ? 8 : 0);   // smem cout pitch: = 8 or 24 mod 32 -> conflict-free B loads
    static constexpr int KSTEPS = CINP / 8;
    static constexpr int WBUF = CINP * NP;                  // floats per weight stage
    static_assert(CINP % 8 == 0 && NB % 8 == 0, "channel padding");
    static_assert(PASSES == 1 || PASSES == 3, "PASSES");
    static size_t smem_bytes(const TapTable& t) {
        size_t tile = (size_t)t.IH * t.IW * CP;
        const size_t stages = WALL ? (size_t)t.n : 2;
        return sizeof(float) * ((PASSES == 3 ? 2 : 1) * (tile + stages * (size_t)WBUF));
    }
};

// ---- input functors: float4 of channels [4*c4, 4*c4+4) at (n, iy, ix); zeros outside the image
struct InNHWC {
    const float* p;
    int H, W, C;            // C = stored channel count (>= CINP used by the kernel)
    size_t nstride;         // floats between consecutive images n (H*W*C when dense)
    __device__ __forceinline__ float4 load4(int n, int iy, int ix, int c4) const {
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        return ldg4(p + (size_t)n * nstride + ((size_t)iy * W + ix) * C + 4 * c4);
    }
};
inline InNHWC in_nhwc(const float* p, int H, int W, int C, size_t nstride = 0) {
    return InNHWC{p, H, W, C, nstride ? nstride : (size_t)H * W * C};
}

struct InNHWC2 {            // channels [0,CA) from a, then [CA, CA+CB) from b (both NHWC, multiples of 4)
    const float* a;
    const float* b;
    int H, W, CA, CBc;
    __device__ __forceinline__ float4 load4(int n, int iy, int ix, int c4) const {
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t pix = ((size_t)n * H + iy) * W + ix;
        const int c = 4 * c4;
        return c < CA ? ldg4(a + pix * CA + c) : ldg4(b + pix * CBc + (c - CA));
    }
};

struct InNCHW3 {            // 3-channel planar image [N][3][H][W] -> channels (r,g,b,0,0,0,0,0)
    const float* p;
    int H, W;
    __device__ __forceinline__ float4 load4(int n, int iy, int ix, int c4) const {
        if (c4 != 0 || iy < 0 || iy >= H || ix < 0 || ix >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t plane = (size_t)H * W, o = (size_t)n * 3 * plane + (size_t)iy * W + ix;
        return make_float4(ldg(p + o), ldg(p + o + plane), ldg(p + o + 2 * plane), 0.f);
    }
};

// grid: (ceil(Wout/16), ceil(Hout/TH), N * ncb); block: THREADS; dyn smem: Cfg::smem_bytes(taps)
// weights: w_hi (and w_lo when PASSES == 3): [slice][CINP][cout_total], values already rounded to TF32.
// Epi::row(n, oy, ox, co0, v): v[2*j], v[2*j+1] = couts co0 + 8*j + 2*t, +1 of pixel (oy, ox) for this
// thread's quad position t = lane & 3; called for every (row-tile, half) with FULL-WARP uniformity
// (pixels outside the image included: the functor masks its stores), so it may shuffle within quads.
struct MmaWeightSel {       // image n of a batched launch picks one of up to three weight sets
    const float* hi[3];
    const float* lo[3];
    int period, split1, split2;
    __device__ __forceinline__ int pick(int n) const {
        const int r = n % period;
        return r < split1 ? 0 : (r < split2 ? 1 : 2);
    }
    static MmaWeightSel single(imvs_wpair w) {
        MmaWeightSel s;
        for (int i = 0; i < 3; ++i) { s.hi[i] = w.hi; s.lo[i] = w.lo; }
        s.period = 1; s.split1 = 1; s.split2 = 1;
        return s;
    }
};

template <class Cfg, class In, class Epi>
__global__ void __launch_bounds__(Cfg::THREADS)
mma_conv_kernel(const In in, const Epi epi, const MmaWeightSel wsel,
                const TapTable taps, int cout_total, int Hout, int Wout, int ncb) {
    constexpr int CP = Cfg::CP, NP = Cfg::NP, CINP = Cfg::CINP, NB = Cfg::NB, MT = Cfg::MT, NT = Cfg::NT;
    extern __shared__ __align__(16) float smem[];
    const int tile_floats = taps.IH * taps.IW * CP;
    float* sA = smem;
    float* sAlo = smem + tile_floats;                                        // PASSES == 3 only
    const int wstages = Cfg::WALL ? taps.n : 2;
    float* sW = smem + (Cfg::PASSES == 3 ? 2 : 1) * tile_floats;             // [stages][WBUF] (hi) then the same (lo)
    float* sWlo = sW + wstages * Cfg::WBUF;

    const int n = blockIdx.z / ncb, cb = blockIdx.z % ncb;
    const int oy0 = blockIdx.y * Cfg::TH, ox0 = blockIdx.x * 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wset = wsel.pick(n);
    const float* __restrict__ w_hi = wsel.hi[wset];
    const float* __restrict__ w_lo = wsel.lo[wset];

    auto load_weights = [&](int tap, int buf) {
        const float* src = w_hi + ((size_t)taps.widx[tap] * CINP) * cout_total + cb * NB;
        float* dst = sW + buf * Cfg::WBUF;
        constexpr int Q = NB / 4;
        for (int i = tid; i < CINP * Q; i += Cfg::THREADS) {
            const int k = i / Q, q = i % Q;
            cp_async16(dst + k * NP + 4 * q, src + (size_t)k * cout_total + 4 * q);
        }
        if constexpr (Cfg::PASSES == 3) {
            const float* srcl = w_lo + ((size_t)taps.widx[tap] * CINP) * cout_total + cb * NB;
            float* dstl = sWlo + buf * Cfg::WBUF;
            for (int i = tid; i < CINP * Q; i += Cfg::THREADS) {
                const int k = i / Q, q = i % Q;
                cp_async16(dstl + k * NP + 4 * q, srcl + (size_t)k * cout_total + 4 * q);
            }
        }
        cp_async_commit();
    };

    if constexpr (Cfg::WALL) {
        for (int tp = 0; tp < taps.n; ++tp) load_weights(tp, tp);
    } else {
        load_weights(0, 0);
    }
    {   // stage the input tile (with halo), rounded to TF32
        const int iy0 = oy0 * Cfg::STRIDE + taps.dy_min, ix0 = ox0 * Cfg::STRIDE + taps.dx_min;
        constexpr int C4 = CINP / 4;
        const int total = taps.IH * taps.IW * C4;
        for (int i = tid; i < total; i += Cfg::THREADS) {
            const int slot = i / C4, c4 = i % C4;
            const int iy = iy0 + slot / taps.IW, ix = ix0 + slot % taps.IW;
            const float4 v = in.load4(n, iy, ix, c4);
            uint4 hi = make_uint4(f2tf32(v.x), f2tf32(v.y), f2tf32(v.z), f2tf32(v.w));
            *reinterpret_cast<uint4*>(sA + (size_t)slot * CP + 4 * c4) = hi;
            if constexpr (Cfg::PASSES == 3) {
                uint4 lo = make_uint4(f2tf32(v.x - __uint_as_float(hi.x)), f2tf32(v.y - __uint_as_float(hi.y)),
                                      f2tf32(v.z - __uint_as_float(hi.z)), f2tf32(v.w - __uint_as_float(hi.w)));
                *reinterpret_cast<uint4*>(sAlo + (size_t)slot * CP + 4 * c4) = lo;
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    float acc[MT][NT][4];
#pragma unroll
    for (int r = 0; r < MT; ++r)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[r][j][q] = 0.f;

    for (int tap = 0; tap < taps.n; ++tap) {
        if constexpr (!Cfg::WALL) {
            if (tap + 1 < taps.n) load_weights(tap + 1, (tap + 1) & 1);
        }
        const int wsel = Cfg::WALL ? tap : (tap & 1);
        const float* wb = sW + wsel * Cfg::WBUF;
        const float* wbl = sWlo + wsel * Cfg::WBUF;
        const int ry = taps.dy[tap] - taps.dy_min, rx = taps.dx[tap] - taps.dx_min;
        int slot0[MT], slot1[MT];
#pragma unroll
        for (int r = 0; r < MT; ++r) {
            const int row = (warp * MT + r) * Cfg::STRIDE + ry;
            slot0[r] = (row * taps.IW + g * Cfg::STRIDE + rx) * CP;
            slot1[r] = (row * taps.IW + (g + 8) * Cfg::STRIDE + rx) * CP;
        }
#pragma unroll
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
            const int k0 = ks * 8;
            uint32_t a[MT][4], al[MT][4];
#pragma unroll
            for (int r = 0; r < MT; ++r) {
                a[r][0] = __float_as_uint(sA[slot0[r] + k0 + t]);
                a[r][1] = __float_as_uint(sA[slot1[r] + k0 + t]);
                a[r][2] = __float_as_uint(sA[slot0[r] + k0 + t + 4]);
                a[r][3] = __float_as_uint(sA[slot1[r] + k0 + t + 4]);
                if constexpr (Cfg::PASSES == 3) {
                    al[r][0] = __float_as_uint(sAlo[slot0[r] + k0 + t]);
                    al[r][1] = __float_as_uint(sAlo[slot1[r] + k0 + t]);
                    al[r][2] = __float_as_uint(sAlo[slot0[r] + k0 + t + 4]);
                    al[r][3] = __float_as_uint(sAlo[slot1[r] + k0 + t + 4]);
                }
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const uint32_t b0 = __float_as_uint(wb[(k0 + t) * NP + 8 * j + g]);
                const uint32_t b1 = __float_as_uint(wb[(k0 + t + 4) * NP + 8 * j + g]);
#pragma unroll
                for (int r = 0; r < MT; ++r) mma_tf32(acc[r][j], a[r][0], a[r][1], a[r][2], a[r][3], b0, b1);
                if constexpr (Cfg::PASSES == 3) {
                    const uint32_t bl0 = __float_as_uint(wbl[(k0 + t) * NP + 8 * j + g]);
                    const uint32_t bl1 = __float_as_uint(wbl[(k0 + t + 4) * NP + 8 * j + g]);
#pragma unroll
                    for (int r = 0; r < MT; ++r) {
                        mma_tf32(acc[r][j], a[r][0], a[r][1], a[r][2], a[r][3], bl0, bl1);
                        mma_tf32(acc[r][j], al[r][0], al[r][1], al[r][2], al[r][3], b0, b1);
                    }
                }
            }
        }
        if constexpr (!Cfg::WALL) {
            cp_async_wait_all();
            __syncthreads();
        }
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
            epi.template row<NT>(n, oy, ox0 + g + 8 * h, cb * NB, t, v);
        }
    }
}

template <class Cfg, class In, class Epi>
int launch_mma_conv(const char* name, const In& in, const Epi& epi, const MmaWeightSel& wsel,
                    const TapTable& taps, int N, int cout_total, int Hout, int Wout, int ncb, cudaStream_t st) {
    for (int i = 0; i < 3; ++i)
        IMVS_REQUIRE(wsel.hi[i] && (Cfg::PASSES == 1 || wsel.lo[i]), "%s: null weights", name);
    IMVS_REQUIRE(cout_total % 4 == 0 && ncb * Cfg::NB <= cout_total, "%s: cout_total=%d must be a multiple of 4 and >= %d", name,
                 cout_total, ncb * Cfg::NB);
    const size_t smem = Cfg::smem_bytes(taps);
    IMVS_REQUIRE(smem <= 200 * 1024, "%s: %zu bytes of shared memory needed", name, smem);
    auto kern = mma_conv_kernel<Cfg, In, Epi>;
    static int smem_ok = 0;
    IMVS_TRY(ensure_dynamic_smem(kern, smem, &smem_ok));
    dim3 grid(cdiv(Wout, 16), cdiv(Hout, Cfg::TH), N * ncb);
    IMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "%s: grid too large", name);
    kern<<<grid, Cfg::THREADS, smem, st>>>(in, epi, wsel, taps, cout_total, Hout, Wout, ncb);
    count_launch();
    IMVS_LAUNCH_CHECK(name);
    return 0;
}

int conv_passes();       // process-wide precision switch (imvs_set_conv_passes), defined in warp.cu

// precision dispatch: one call site, both instantiations
template <int CINP, int NB, int MT, int WARPS, int STRIDE, bool WALL, class In, class Epi>
int mma_conv(const char* name, const In& in, const Epi& epi, const MmaWeightSel& wsel, const TapTable& taps, int N,
             int cout_total, int Hout, int Wout, int ncb, cudaStream_t st) {
    if (conv_passes() == 3)
        return launch_mma_conv<MmaCfg<CINP, NB, MT, WARPS, STRIDE, 3, WALL>>(name, in, epi, wsel, taps, N, cout_total, Hout, Wout, ncb, st);
    return launch_mma_conv<MmaCfg<CINP, NB, MT, WARPS, STRIDE, 1, WALL>>(name, in, epi, wsel, taps, N, cout_total, Hout, Wout, ncb, st);
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
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT]) const {
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

// transposed-conv parity pass (see make_taps_tconv): grid position (oy, ox) -> output (2*oy + a, 2*ox + b),
// plus the U-Net skip connection (itermvs.py:374-377)
struct EpiTconvNHWC {
    float* out;
    const float* skip;
    int Hin, Win, C, a, b;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT]) const {
        if (oy >= Hin || ox >= Win) return;
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
