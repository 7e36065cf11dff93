// Register-blocked direct convolution (fp32 FFMA) for the small convolutions of the estimator.
//
// The spatial extents here are tiny (<= 160x128 per slice at the benchmark size) and channel counts
// are 8..64, so the design goal is enough independent threads to fill 148 SMs with a good
// FMA : load ratio, not tile reuse through shared memory:
//   * lanes run along x  -> every activation load / store is coalesced;
//   * a thread owns PX vertically adjacent output pixels x CO output channels (PX*CO accumulators);
//     per (cin, kx) it loads the NR input rows it needs once and reuses them for the KS vertical
//     taps and all CO channels;
//   * the block's weight slice [Cin][k*k][CB] is staged once in shared memory and read as
//     warp-uniform float4 broadcasts;
//   * activations are read straight through L1 (each input element is re-read by the 3 kx taps
//     and by the COG channel-group warps -- all L1 hits).
// Input and epilogue are functors, so concatenated inputs (ConvGRU), channels-last correlation
// volumes, residual adds, gates etc. fuse into the same kernel.
#pragma once
#include "common.cuh"

namespace imvs {

// Weight-set selector: slice n of a batched launch picks one of up to three weight sets
// (the three CorrNets of the iteration run as one launch).
struct WeightSel {
    const float* w[3];
    int period, split1, split2;
    __device__ __forceinline__ const float* pick(int n) const {
        int r = n % period;
        return r < split1 ? w[0] : (r < split2 ? w[1] : w[2]);
    }
    static WeightSel single(const float* p) {
        WeightSel s;
        s.w[0] = s.w[1] = s.w[2] = p;
        s.period = 1; s.split1 = 1; s.split2 = 1;
        return s;
    }
};

template <int COUT_, int CB_, int CO_, int PX_, int ROWG_, int KS_, int STRIDE_, int DIL_, int CHUNK_>
struct ConvCfg {
    static constexpr int COUT = COUT_;     // total output channels
    static constexpr int CB = CB_;         // output channels per block
    static constexpr int CO = CO_;         // output channels per thread
    static constexpr int PX = PX_;         // output rows per thread
    static constexpr int ROWG = ROWG_;     // row groups per block
    static constexpr int KS = KS_, STRIDE = STRIDE_, DIL = DIL_, CHUNK = CHUNK_;
    static constexpr int COG = CB / CO;
    static constexpr int NCB = COUT / CB;
    static constexpr int THREADS = 32 * ROWG * COG;
    static constexpr int PAD = DIL * (KS - 1) / 2;
    static constexpr int NR = (PX - 1) * STRIDE + (KS - 1) * DIL + 1;
    static constexpr int TILE_H = ROWG * PX;
    static_assert(CB % CO == 0 && COUT % CB == 0, "channel blocking");
    static_assert(CO % 4 == 0 || CO == 1 || CO == 2, "CO must allow float4 weight loads");
};

// ---- input functors: load CHUNK consecutive channels [ci0, ci0+CHUNK) at (n, iy, ix), zero outside
struct InPlanar {            // [N][Cin][H][W]
    const float* p;
    int Cin, H, W;
    template <int CHUNK>
    __device__ __forceinline__ void load(int n, int ci0, int iy, int ix, float (&v)[CHUNK]) const {
        static_assert(CHUNK == 1, "planar input uses CHUNK=1");
        v[0] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? ldg(p + (((size_t)n * Cin + ci0) * H + iy) * W + ix) : 0.f;
    }
};

struct InConcat2 {           // channels [0,CA) from a [N][CA][H][W], the rest from b [N][CBc][H][W]
    const float* a;
    const float* b;
    int CA, CBc, H, W;
    template <int CHUNK>
    __device__ __forceinline__ void load(int n, int ci0, int iy, int ix, float (&v)[CHUNK]) const {
        static_assert(CHUNK == 1, "concat input uses CHUNK=1");
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            const float* q = ci0 < CA ? a + (((size_t)n * CA + ci0) * H + iy) * W + ix
                                      : b + (((size_t)n * CBc + (ci0 - CA)) * H + iy) * W + ix;
            v[0] = ldg(q);
        } else {
            v[0] = 0.f;
        }
    }
};

struct InChannelsLast8 {     // [N][H][W][8]
    const float* p;
    int H, W;
    template <int CHUNK>
    __device__ __forceinline__ void load(int n, int ci0, int iy, int ix, float (&v)[CHUNK]) const {
        static_assert(CHUNK == 8, "channels-last-8 input uses CHUNK=8");
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            const float* q = p + (((size_t)n * H + iy) * W + ix) * 8;
            float4 a = ldg4(q), b = ldg4(q + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
    }
};

// ---- the kernel ------------------------------------------------------------------------------
// grid: (ceil(Wout/32), ceil(Hout/TILE_H), N*NCB); block: THREADS; dyn smem: Cin*KS*KS*CB floats
template <class Cfg, class In, class Epi>
__global__ void __launch_bounds__(Cfg::THREADS)
conv_kernel(const In in, const Epi epi, const WeightSel wsel, int Cin, int Hout, int Wout, int ncb) {
    constexpr int KK = Cfg::KS * Cfg::KS;
    extern __shared__ __align__(16) float sw[];
    const int n = blockIdx.z / ncb, cb = blockIdx.z % ncb;   // ncb <= NCB: only the first ncb channel blocks run
    {   // stage this block's weight slice: global [Cin][KK][COUT] -> smem [Cin][KK][CB]
        const float* wg = wsel.pick(n);
        const int rows = Cin * KK;
        if constexpr (Cfg::CB % 4 == 0) {
            constexpr int Q = Cfg::CB / 4;
            for (int i = threadIdx.x; i < rows * Q; i += Cfg::THREADS) {
                int row = i / Q, q = i % Q;
                reinterpret_cast<float4*>(sw)[i] = ldg4(wg + (size_t)row * Cfg::COUT + cb * Cfg::CB + q * 4);
            }
        } else {
            for (int i = threadIdx.x; i < rows * Cfg::CB; i += Cfg::THREADS) {
                int row = i / Cfg::CB, q = i % Cfg::CB;
                sw[i] = ldg(wg + (size_t)row * Cfg::COUT + cb * Cfg::CB + q);
            }
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int rg = (threadIdx.x >> 5) % Cfg::ROWG;
    const int cg = (threadIdx.x >> 5) / Cfg::ROWG;
    const int x = blockIdx.x * 32 + lane;
    const int y0 = (blockIdx.y * Cfg::ROWG + rg) * Cfg::PX;
    if (y0 >= Hout) return;

    float acc[Cfg::PX][Cfg::CO];
#pragma unroll
    for (int j = 0; j < Cfg::PX; ++j)
#pragma unroll
        for (int c = 0; c < Cfg::CO; ++c) acc[j][c] = 0.f;

    const int iy0 = y0 * Cfg::STRIDE - Cfg::PAD;
    const int ixb = x * Cfg::STRIDE - Cfg::PAD;
    for (int ci0 = 0; ci0 < Cin; ci0 += Cfg::CHUNK) {
#pragma unroll
        for (int kx = 0; kx < Cfg::KS; ++kx) {
            float v[Cfg::NR][Cfg::CHUNK];
            const int ix = ixb + kx * Cfg::DIL;
#pragma unroll
            for (int r = 0; r < Cfg::NR; ++r) in.template load<Cfg::CHUNK>(n, ci0, iy0 + r, ix, v[r]);
#pragma unroll
            for (int cc = 0; cc < Cfg::CHUNK; ++cc) {
#pragma unroll
                for (int ky = 0; ky < Cfg::KS; ++ky) {
                    const float* wp = sw + ((ci0 + cc) * KK + ky * Cfg::KS + kx) * Cfg::CB + cg * Cfg::CO;
                    float w[Cfg::CO];
                    if constexpr (Cfg::CO % 4 == 0) {
#pragma unroll
                        for (int q = 0; q < Cfg::CO / 4; ++q) {
                            float4 t = reinterpret_cast<const float4*>(wp)[q];
                            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < Cfg::CO; ++q) w[q] = wp[q];
                    }
#pragma unroll
                    for (int j = 0; j < Cfg::PX; ++j)
#pragma unroll
                        for (int c = 0; c < Cfg::CO; ++c)
                            acc[j][c] = fmaf(v[j * Cfg::STRIDE + ky * Cfg::DIL][cc], w[c], acc[j][c]);
                }
            }
        }
    }
    if (x < Wout) {
#pragma unroll
        for (int j = 0; j < Cfg::PX; ++j)
            if (y0 + j < Hout) epi.template store<Cfg::CO>(n, y0 + j, x, cb * Cfg::CB + cg * Cfg::CO, acc[j]);
    }
}

template <class Cfg, class In, class Epi>
int launch_conv(const char* name, const In& in, const Epi& epi, const WeightSel& wsel, int N, int Cin, int Hout,
                int Wout, cudaStream_t st, int ncb = Cfg::NCB) {
    constexpr int KK = Cfg::KS * Cfg::KS;
    IMVS_REQUIRE(Cin % Cfg::CHUNK == 0, "%s: Cin=%d not a multiple of %d", name, Cin, Cfg::CHUNK);
    size_t smem = (size_t)Cin * KK * Cfg::CB * sizeof(float);
    IMVS_REQUIRE(smem <= 200 * 1024, "%s: weight slice of %zu bytes does not fit shared memory", name, smem);
    auto kern = conv_kernel<Cfg, In, Epi>;
    static int smem_ok = 0;
    IMVS_TRY(ensure_dynamic_smem(kern, smem, &smem_ok));
    dim3 grid(cdiv(Wout, 32), cdiv(Hout, Cfg::TILE_H), N * ncb);
    IMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "%s: grid too large", name);
    kern<<<grid, Cfg::THREADS, smem, st>>>(in, epi, wsel, Cin, Hout, Wout, ncb);
    count_launch();
    IMVS_LAUNCH_CHECK(name);
    return 0;
}

// ---- transposed 3x3 convolution, stride 2, padding 1, output_padding 1 (CorrNet conv3/conv4) ----
// Each thread owns one INPUT pixel (iy, ix) -> the 2x2 output quad (2iy+a, 2ix+b) x CO channels:
//   out(2iy  ,2ix  ) = in(iy,ix) W[1][1]
//   out(2iy  ,2ix+1) = in(iy,ix+1) W[1][0] + in(iy,ix) W[1][2]
//   out(2iy+1,2ix  ) = in(iy+1,ix) W[0][1] + in(iy,ix) W[2][1]
//   out(2iy+1,2ix+1) = in(iy+1,ix+1) W[0][0] + in(iy+1,ix) W[0][2] + in(iy,ix+1) W[2][0] + in(iy,ix) W[2][2]
// (out[y] += in[iy] W[ky] with y = 2 iy - 1 + ky, PyTorch ConvTranspose2d semantics), plus the
// residual `skip` of the U-Net (itermvs.py:374-377).  Weights packed [Cin][9][COUT].
// grid: (ceil(Win/32), ceil(Hin/ROWS), N*NCB), block 32*ROWS*COG
template <int COUT, int CB, int CO, int ROWS>
__global__ void __launch_bounds__(32 * ROWS * (CB / CO))
tconv_kernel(const float* __restrict__ in, const float* __restrict__ skip, float* __restrict__ out,
             const WeightSel wsel, int Cin, int Hin, int Win) {
    constexpr int NCB = COUT / CB, THREADS = 32 * ROWS * (CB / CO);
    extern __shared__ __align__(16) float sw[];
    const int n = blockIdx.z / NCB, cb = blockIdx.z % NCB;
    {
        const float* wg = wsel.pick(n);
        constexpr int Q = CB / 4;
        for (int i = threadIdx.x; i < Cin * 9 * Q; i += THREADS) {
            int row = i / Q, q = i % Q;
            reinterpret_cast<float4*>(sw)[i] = ldg4(wg + (size_t)row * COUT + cb * CB + q * 4);
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int rg = (threadIdx.x >> 5) % ROWS;
    const int cg = (threadIdx.x >> 5) / ROWS;
    const int ix = blockIdx.x * 32 + lane, iy = blockIdx.y * ROWS + rg;
    if (iy >= Hin || ix >= Win) return;
    float acc[4][CO];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[q][c] = 0.f;
    const bool xr = ix + 1 < Win, yd = iy + 1 < Hin;
    for (int ci = 0; ci < Cin; ++ci) {
        const float* p = in + (((size_t)n * Cin + ci) * Hin + iy) * Win + ix;
        const float a = ldg(p);
        const float b = xr ? ldg(p + 1) : 0.f;
        const float c = yd ? ldg(p + Win) : 0.f;
        const float d = (xr && yd) ? ldg(p + Win + 1) : 0.f;
        const float* wp = sw + (size_t)ci * 9 * CB + cg * CO;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float w[CO];
#pragma unroll
            for (int q = 0; q < CO / 4; ++q) {
                float4 t = reinterpret_cast<const float4*>(wp + k * CB)[q];
                w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
            // k = ky*3 + kx : which (input, output-quad) pair this tap connects
            // quad index: 0 (0,0)  1 (0,1)  2 (1,0)  3 (1,1)
#pragma unroll
            for (int cc = 0; cc < CO; ++cc) {
                if (k == 4) acc[0][cc] = fmaf(a, w[cc], acc[0][cc]);            // W[1][1]
                else if (k == 3) acc[1][cc] = fmaf(b, w[cc], acc[1][cc]);       // W[1][0]
                else if (k == 5) acc[1][cc] = fmaf(a, w[cc], acc[1][cc]);       // W[1][2]
                else if (k == 1) acc[2][cc] = fmaf(c, w[cc], acc[2][cc]);       // W[0][1]
                else if (k == 7) acc[2][cc] = fmaf(a, w[cc], acc[2][cc]);       // W[2][1]
                else if (k == 0) acc[3][cc] = fmaf(d, w[cc], acc[3][cc]);       // W[0][0]
                else if (k == 2) acc[3][cc] = fmaf(c, w[cc], acc[3][cc]);       // W[0][2]
                else if (k == 6) acc[3][cc] = fmaf(b, w[cc], acc[3][cc]);       // W[2][0]
                else acc[3][cc] = fmaf(a, w[cc], acc[3][cc]);                   // k == 8, W[2][2]
            }
        }
    }
    const int Ho = 2 * Hin, Wo = 2 * Win;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int oy = 2 * iy + (q >> 1), ox = 2 * ix + (q & 1);
#pragma unroll
        for (int cc = 0; cc < CO; ++cc) {
            size_t o = (((size_t)n * COUT + cb * CB + cg * CO + cc) * Ho + oy) * Wo + ox;
            out[o] = acc[q][cc] + ldg(skip + o);
        }
    }
}

template <int COUT, int CB, int CO, int ROWS>
int launch_tconv(const char* name, const float* in, const float* skip, float* out, const WeightSel& wsel, int N,
                 int Cin, int Hin, int Win, cudaStream_t st) {
    size_t smem = (size_t)Cin * 9 * CB * sizeof(float);
    auto kern = tconv_kernel<COUT, CB, CO, ROWS>;
    static int smem_ok = 0;
    IMVS_TRY(ensure_dynamic_smem(kern, smem, &smem_ok));
    dim3 grid(cdiv(Win, 32), cdiv(Hin, ROWS), N * (COUT / CB));
    IMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "%s: grid too large", name);
    kern<<<grid, 32 * ROWS * (CB / CO), smem, st>>>(in, skip, out, wsel, Cin, Hin, Win);
    count_launch();
    IMVS_LAUNCH_CHECK(name);
    return 0;
}

// ---- common epilogues --------------------------------------------------------------------------
struct EpiPlanar {           // out [N][Cout][H][W] (+bias) (+relu)
    float* out;
    const float* bias;       // may be null
    int Cout, H, W;
    bool relu;
    template <int CO>
    __device__ __forceinline__ void store(int n, int y, int x, int co0, const float (&a)[CO]) const {
#pragma unroll
        for (int c = 0; c < CO; ++c) {
            float v = a[c] + (bias ? ldg(bias + co0 + c) : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            out[(((size_t)n * Cout + co0 + c) * H + y) * W + x] = v;
        }
    }
};

}  // namespace imvs
