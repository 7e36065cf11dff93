// CorrNet (reference models/itermvs.py:352-381) as ONE kernel that keeps a spatial tile resident in shared memory through all
// six layers: conv0 (8->8) -> conv1 (8->16, stride 2) -> conv2 (16->32, stride 2) -> conv3^T (32->16, + conv1) -> conv4^T (16->8,
// + conv0) -> conv5 (8->1, + bias).  The per-layer kernels are bound by launch + load -> compute -> store latency (7-14 us for
// <= 0.1 GFLOP each, 66 us per pass; the per-layer and the cooperative tcgen05 variants measured no better, profiles/README.md);
// here a CTA computes a 32 x 32 output tile of one slice from a 45 x 45 input window, recomputing the halo of every
// intermediate (1.9x / 1.7x / 1.6x / 1.3x / 1.1x of the tile's own pixels) instead of exchanging it through global memory:
// no inter-layer traffic, no inter-CTA dependency, one launch.  Plain fp32 FFMA (exact fp32, no operand split): the
// whole pass is 0.66 GFMA including the recomputation.  Intermediates are channel-planar in shared memory (thread = pixel ->
// conflict-free), the slice's six weight tensors (50.7 KB) sit next to them; values at positions outside the image are stored
// as zeros (= the zero padding the next layer sees).
//
// Regions (global coordinates of the layer's own resolution; tile origin (y0, x0) full res, h0 = y0 / 2, q0 = y0 / 4):
//   in   rows y0-8 .. y0+36 (45)      c0  y0-7 .. y0+35 (43)      c1  h0-3 .. h0+17 (21)      c2  q0-1 .. q0+8 (10)
//   x3   h0-1 .. h0+16 (18)           x4  y0-1 .. y0+32 (34)      out y0 .. y0+31 (32)
// Local index arithmetic: conv0 in[r+ky]; conv1 c0[2r+ky]; conv2 c1[2r+ky]; conv3^T / conv4^T in[(r-ky)/2 + 1] for the ky of
// r's parity (r even: ky = 0, 2; r odd: ky = 1), skips c1[r+2] / c0[r+6]; conv5 x4[r+ky].
#pragma once
#include "common.cuh"

namespace imvs {
namespace ctile {

constexpr int T = 32;                                  // output tile (full resolution)
constexpr int IN_R = 45, C0_R = 43, C1_R = 21, C2_R = 10, X3_R = 18, X4_R = 34;
constexpr int THREADS = 512;
constexpr int W0 = 9 * 8 * 8, W1 = 9 * 8 * 16, W2 = 9 * 16 * 32, W3 = 9 * 32 * 16, W4 = 9 * 16 * 8, W5 = 9 * 8 * 8;
constexpr int W_FLOATS = W0 + W1 + W2 + W3 + W4 + W5;                      // 12 672
constexpr int BUF_A = 8 * IN_R * IN_R;                                     // input window; later x3 | x4
constexpr int BUF_C0 = 8 * C0_R * C0_R, BUF_C1 = 16 * C1_R * C1_R, BUF_C2 = 32 * C2_R * C2_R;
constexpr int BUF_X3 = 16 * X3_R * X3_R, BUF_X4 = 8 * X4_R * X4_R;
static_assert(BUF_X3 + BUF_X4 <= BUF_A, "x3 and x4 reuse the input window's memory");
constexpr size_t SMEM_BYTES = sizeof(float) * (size_t)(BUF_A + BUF_C0 + BUF_C1 + BUF_C2 + W_FLOATS);      // 215 680

struct Params {
    const float* vol;            // [N][H][W][8]
    const float* w[3][6];        // [set][layer] fp32 [tap][CinP][CoutP] (imvs_wpair::fp32)
    const float* b5[3];          // conv5 bias per set
    int period, split1, split2;  // slice n -> set: r = n % period; r < split1 ? 0 : (r < split2 ? 1 : 2)
    float* out;                  // out[(n / period) * bstride + pixel * pstride + n % period]
    size_t bstride, pstride;
    int N, H, W, tiles_x, tiles_y;
};

// PX = output pixels of a COLUMN per work item (a weight vector is read once for PX pixels; lanes = consecutive columns); chosen per layer so that a layer has
// about one work item per thread: 473 / 462 / 400 / 360 / 340 / 512 items for the six layers

struct W8 { float4 a, b; };
__device__ __forceinline__ W8 ldw8(const float* __restrict__ w) {          // eight output channels' weights: two broadcast 16-byte reads
    return W8{*reinterpret_cast<const float4*>(w), *reinterpret_cast<const float4*>(w + 4)};
}
__device__ __forceinline__ void fma8(float (&acc)[8], float v, const W8& w) {
    acc[0] = fmaf(v, w.a.x, acc[0]); acc[1] = fmaf(v, w.a.y, acc[1]); acc[2] = fmaf(v, w.a.z, acc[2]); acc[3] = fmaf(v, w.a.w, acc[3]);
    acc[4] = fmaf(v, w.b.x, acc[4]); acc[5] = fmaf(v, w.b.y, acc[5]); acc[6] = fmaf(v, w.b.z, acc[6]); acc[7] = fmaf(v, w.b.w, acc[7]);
}

// stride-1 / stride-2 3x3 convolution + ReLU on shared-memory planes: out[co][r][c] = relu(sum in[ci][S r + ky][S c + kx] w[tap][ci][co]),
// zero where the output position (gy0 + r, gx0 + c) lies outside the Hl x Wl image of its resolution.
// item = (8 couts, strip of PX ROWS, column): consecutive lanes = consecutive columns, so the plane reads are conflict-free
// (stride S words); per (ci, kx) a column of S (PX - 1) + 3 inputs and 3 x 2 broadcast weight reads feed 3 * PX * 8 FMAs.
template <int CIN, int COUT, int S, int PX>
__device__ __forceinline__ void conv_relu(const float* __restrict__ in, int IR, float* __restrict__ out, int OR_, const float* __restrict__ w,
                                          int gy0, int gx0, int Hl, int Wl) {
    constexpr int G = COUT / 8, NV = S * (PX - 1) + 3;
    const int strips = (OR_ + PX - 1) / PX, n_items = G * strips * OR_;
    for (int item = threadIdx.x; item < n_items; item += THREADS) {
        const int g = item / (strips * OR_), q = item - g * strips * OR_, r0 = (q / OR_) * PX, c = q - (q / OR_) * OR_;
        float acc[PX][8];
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[p][k] = 0.f;
        if ((unsigned)(gx0 + c) < (unsigned)Wl) {
#pragma unroll 1
            for (int ci = 0; ci < CIN; ++ci) {
                const float* wp = w + ci * COUT + 8 * g;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float* ip = in + (ci * IR + S * r0) * IR + S * c + kx;
                    float v[NV];
#pragma unroll
                    for (int i = 0; i < NV; ++i) v[i] = (S * r0 + i < IR) ? ip[i * IR] : 0.f;      // rows below the window feed unsaved outputs only
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const W8 wv = ldw8(wp + (ky * 3 + kx) * CIN * COUT);
#pragma unroll
                        for (int p = 0; p < PX; ++p) fma8(acc[p], v[S * p + ky], wv);
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int r = r0 + p;
            if (r >= OR_) break;
            const bool inside = (unsigned)(gy0 + r) < (unsigned)Hl && (unsigned)(gx0 + c) < (unsigned)Wl;
#pragma unroll
            for (int k = 0; k < 8; ++k) out[((8 * g + k) * OR_ + r) * OR_ + c] = inside ? fmaxf(acc[p][k], 0.f) : 0.f;
        }
    }
}

// ConvTranspose2d(k 3, stride 2, padding 1, output_padding 1) + skip: out[co][r][c] = skip[co][r + so][c + so] +
// sum over the (ky, kx) of (r, c)'s parity of in[ci][(r - ky) / 2 + 1][(c - kx) / 2 + 1] w[tap][ci][co].  Items are ordered by
// parity class (a warp's lanes share their tap set); item = (8 couts, strip of PX rows of the class: r = 2 i + a, column c = 2 j + b),
// consecutive lanes = consecutive j: conflict-free input reads.
template <int CIN, int COUT, int PX>
__device__ __forceinline__ void tconv_skip(const float* __restrict__ in, int IR, float* __restrict__ out, int OR_, const float* __restrict__ skip,
                                           int SR, int so, const float* __restrict__ w, int gy0, int gx0, int Hl, int Wl) {
    constexpr int G = COUT / 8;
    const int half = OR_ / 2, strips = (half + PX - 1) / PX, per_class = G * strips * half;
    for (int item = threadIdx.x; item < 4 * per_class; item += THREADS) {
        const int cls = item / per_class, q0 = item - cls * per_class, g = q0 / (strips * half), q = q0 - g * strips * half;
        const int a = cls >> 1, b = cls & 1, i0 = (q / half) * PX, j = q % half, c = 2 * j + b;
        float acc[PX][8];
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[p][k] = 0.f;
        if ((unsigned)(gx0 + c) < (unsigned)Wl) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                if (jj == 1 && b == 1) break;                     // c odd: kx = 1 only; c even: kx = 0 and 2
                const int kx = b ? 1 : 2 * jj, ix = (c - kx) / 2 + 1;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i == 1 && a == 1) break;
                    const int ky = a ? 1 : 2 * i, iy0 = (2 * i0 + a - ky) / 2 + 1;          // input row of the strip's first pixel
                    const float* ip = in + iy0 * IR + ix;
                    const float* wp = w + (ky * 3 + kx) * CIN * COUT + 8 * g;
#pragma unroll 2
                    for (int ci = 0; ci < CIN; ++ci) {
                        const W8 wv = ldw8(wp + ci * COUT);
#pragma unroll
                        for (int p = 0; p < PX; ++p) fma8(acc[p], (iy0 + p < IR) ? ip[ci * IR * IR + p * IR] : 0.f, wv);
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int i = i0 + p, r = 2 * i + a;
            if (i >= half) break;
            const bool inside = (unsigned)(gy0 + r) < (unsigned)Hl && (unsigned)(gx0 + c) < (unsigned)Wl;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                out[((8 * g + k) * OR_ + r) * OR_ + c] = inside ? acc[p][k] + skip[((8 * g + k) * SR + r + so) * SR + c + so] : 0.f;
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1) corrnet_tile_kernel(const Params P) {
    extern __shared__ __align__(16) float sm[];
    float* bufA = sm;                        // in [8][45][45]  ->  x3 [16][18][18] | x4 [8][34][34]
    float* c0 = bufA + BUF_A;                // [8][43][43]
    float* c1 = c0 + BUF_C0;                 // [16][21][21]
    float* c2 = c1 + BUF_C1;                 // [32][10][10]
    float* sw = c2 + BUF_C2;                 // the slice's six weight tensors
    float* x3 = bufA;
    float* x4 = bufA + BUF_X3;
    const int tid = threadIdx.x;
    const int n = blockIdx.z, y0 = blockIdx.y * T, x0 = blockIdx.x * T, h0 = y0 / 2, hx0 = x0 / 2, q0 = y0 / 4, qx0 = x0 / 4;
    const int H = P.H, W = P.W;
    const int rr = n % P.period, set = rr < P.split1 ? 0 : (rr < P.split2 ? 1 : 2);
    pdl_trigger();
    {   // weights are constants of the forward pass: staged before the grid-dependency wait
        constexpr int sizes[6] = {W0, W1, W2, W3, W4, W5};
        int off = 0;
#pragma unroll
        for (int l = 0; l < 6; ++l) {
            const float4* src = reinterpret_cast<const float4*>(P.w[set][l]);
            for (int i = tid; i < sizes[l] / 4; i += THREADS) reinterpret_cast<float4*>(sw + off)[i] = __ldg(src + i);
            off += sizes[l];
        }
    }
    pdl_wait();
    // input window (channels-last in global memory -> channel planes), zero outside the image
    for (int i = tid; i < IN_R * IN_R; i += THREADS) {
        const int r = i / IN_R, c = i - r * IN_R, gy = y0 - 8 + r, gx = x0 - 8 + c;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) {
            const float* src = P.vol + (((size_t)n * H + gy) * W + gx) * 8;
            a = ldg4(src); b = ldg4(src + 4);
        }
        bufA[0 * IN_R * IN_R + i] = a.x; bufA[1 * IN_R * IN_R + i] = a.y; bufA[2 * IN_R * IN_R + i] = a.z; bufA[3 * IN_R * IN_R + i] = a.w;
        bufA[4 * IN_R * IN_R + i] = b.x; bufA[5 * IN_R * IN_R + i] = b.y; bufA[6 * IN_R * IN_R + i] = b.z; bufA[7 * IN_R * IN_R + i] = b.w;
    }
    __syncthreads();
    const float* w0 = sw;
    const float* w1 = w0 + W0;
    const float* w2 = w1 + W1;
    const float* w3 = w2 + W2;
    const float* w4 = w3 + W3;
    const float* w5 = w4 + W4;
    conv_relu<8, 8, 1, 4>(bufA, IN_R, c0, C0_R, w0, y0 - 7, x0 - 7, H, W);                              // itermvs.py:369
    __syncthreads();
    conv_relu<8, 16, 2, 2>(c0, C0_R, c1, C1_R, w1, h0 - 3, hx0 - 3, H / 2, W / 2);                     // :370
    __syncthreads();
    conv_relu<16, 32, 2, 1>(c1, C1_R, c2, C2_R, w2, q0 - 1, qx0 - 1, H / 4, W / 4);                    // :371
    __syncthreads();
    tconv_skip<32, 16, 2>(c2, C2_R, x3, X3_R, c1, C1_R, 2, w3, h0 - 1, hx0 - 1, H / 2, W / 2);         // :373  (bufA's input window is dead)
    __syncthreads();
    tconv_skip<16, 8, 4>(x3, X3_R, x4, X4_R, c0, C0_R, 6, w4, y0 - 1, x0 - 1, H, W);                   // :375
    __syncthreads();
    // conv5: 8 -> 1, + bias, scattered to the caller's layout (:378)
    const float bias = ldg(P.b5[set]);
    constexpr int PX = 2;
    for (int i = tid; i < (T / PX) * T; i += THREADS) {
        const int r0 = (i / T) * PX, c = i - (i / T) * T, gx = x0 + c;
        if (gx >= W) continue;
        float acc[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[p] = 0.f;
#pragma unroll 2
        for (int ci = 0; ci < 8; ++ci) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* ip = x4 + (ci * X4_R + r0) * X4_R + c + kx;
                float v[PX + 2];
#pragma unroll
                for (int k = 0; k < PX + 2; ++k) v[k] = ip[k * X4_R];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const float wv = w5[((ky * 3 + kx) * 8 + ci) * 8];
#pragma unroll
                    for (int p = 0; p < PX; ++p) acc[p] = fmaf(v[p + ky], wv, acc[p]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            const int gy = y0 + r0 + p;
            if (gy < H) P.out[(size_t)(n / P.period) * P.bstride + ((size_t)gy * W + gx) * P.pstride + rr] = acc[p] + bias;
        }
    }
}

}  // namespace ctile
}  // namespace imvs
