// FeatureNet (reference models/net.py:7-66) in eval mode on the tensor-core implicit-GEMM
// convolution: BatchNorm folded into weights/bias by the packer, ReLU / residual add / FPN
// "upsample + lateral" add fused into the epilogues, channels-last activations, and the three
// output convolutions write the pyramids directly in the layout the fused plane-sweep kernels read
// (no transposes).  All views of all reference views run as one batch (net.py:56-65 loops).
#include "common.cuh"
#include "mmaconv.cuh"
#include "tc5conv.cuh"
#ifndef CUSIM
#include "tc5pconv.cuh"
#endif

namespace imvs {

// lateral 1x1 conv + bias + bilinear x2 of the coarser map (net.py:46, 49, 61, 63)
struct EpiAddUp2 {
    float* out;              // [N][H][W][C]
    const float* bias;       // [C]
    const float* coarse;     // [N][H/2][W/2][C]
    int H, W, C;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT], int) const {
        if (oy >= H || ox >= W) return;
        const int Hc = H / 2, Wc = W / 2;
        int h0, h1, w0, w1;
        float lh, lw;
        up_index(oy, 0.5f, Hc, h0, h1, lh);
        up_index(ox, 0.5f, Wc, w0, w1, lw);
        const float* cb = coarse + (size_t)n * Hc * Wc * C;
        const size_t base = (((size_t)n * H + oy) * W + ox) * C;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = co0 + 8 * j + 2 * t;
            if (co >= C) continue;
            const float2 a = ldg2(cb + ((size_t)h0 * Wc + w0) * C + co), b = ldg2(cb + ((size_t)h0 * Wc + w1) * C + co);
            const float2 c = ldg2(cb + ((size_t)h1 * Wc + w0) * C + co), d = ldg2(cb + ((size_t)h1 * Wc + w1) * C + co);
            const float ux = (1.f - lh) * ((1.f - lw) * a.x + lw * b.x) + lh * ((1.f - lw) * c.x + lw * d.x);
            const float uy = (1.f - lh) * ((1.f - lw) * a.y + lw * b.y) + lh * ((1.f - lw) * c.y + lw * d.y);
            *reinterpret_cast<float2*>(out + base + co) =
                make_float2(ux + (v[2 * j] + ldg(bias + co)), uy + (v[2 * j + 1] + ldg(bias + co + 1)));
        }
    }
};

#ifndef CUSIM     // (the CPU emulation of the test-suite has no TMA / tcgen05 model)
// EpiAddUp2 for the tcgen05 / TMA path: the sum is written as split planes [N][C/8][H][W][8 halves] (operand of the
// output convolution, tc5pconv.cuh) and, when another lateral stage upsamples it, also as fp32 NHWC
struct EpiAddUp2H {
    float* out32;            // [N][H][W][C] or nullptr
    tc5p::Split out;         // split planes
    const float* bias;       // [C]
    const float* coarse;     // [N][H/2][W/2][C] fp32
    int H, W, C;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT], int) const {
        if (oy >= H || ox >= W) return;
        const int Hc = H / 2, Wc = W / 2;
        int h0, h1, w0, w1;
        float lh, lw;
        up_index(oy, 0.5f, Hc, h0, h1, lh);
        up_index(ox, 0.5f, Wc, w0, w1, lw);
        const float* cb = coarse + (size_t)n * Hc * Wc * C;
        const size_t plane = (size_t)H * W, pix = (size_t)oy * W + ox;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = co0 + 8 * j + 2 * t;
            if (co >= C) continue;
            const float2 a = ldg2(cb + ((size_t)h0 * Wc + w0) * C + co), b = ldg2(cb + ((size_t)h0 * Wc + w1) * C + co);
            const float2 c = ldg2(cb + ((size_t)h1 * Wc + w0) * C + co), d = ldg2(cb + ((size_t)h1 * Wc + w1) * C + co);
            const float ux = (1.f - lh) * ((1.f - lw) * a.x + lw * b.x) + lh * ((1.f - lw) * c.x + lw * d.x);
            const float uy = (1.f - lh) * ((1.f - lw) * a.y + lw * b.y) + lh * ((1.f - lw) * c.y + lw * d.y);
            const float ox_ = ux + (v[2 * j] + ldg(bias + co)), oy_ = uy + (v[2 * j + 1] + ldg(bias + co + 1));
            if (out32) *reinterpret_cast<float2*>(out32 + ((size_t)n * plane + pix) * C + co) = make_float2(ox_, oy_);
            store_split_pair(out.hi, out.lo, (((size_t)n * (C / 8) + (co >> 3)) * plane + pix) * 8 + (co & 7), ox_, oy_);
        }
    }
};

// The lateral stage on the TMA + tcgen05 kernel (tc5pconv.cuh, 1x1): out = conv1x1(x) + bias + up2(coarse), written as split
// planes (operand of the output convolution and, for intra2, the coarse map of the next lateral stage).  One thread = one
// pixel x NCH channels.  The coarse map is read from its SPLIT PLANES (value = hi + lo, 2^-23 relative from the fp32 value):
// in that chunk-planar layout the 32 pixels of a warp read 16 neighbouring texels x 16 bytes per instruction -- 2 L1
// wavefronts instead of the ~24 of the channels-last fp32 map (thread stride 192 bytes), which bounded this kernel (ablation in
// profiles/README.md: 37 of 50 us in the epilogue).
struct EpiLateral {
    static constexpr int kAhead = 1;
    __device__ __forceinline__ bool wants_prefetch() const { return false; }
    tc5p::Split out;         // [N][C/8][H][W][8]
    float* out32;            // [N][H][W][C] or nullptr
    const float* bias;       // [C]
    tc5p::Split coarse;      // [N][C/8][H/2][W/2][8] split planes
    int H, W;
    // the coarse map as fp32 in the same chunk-planar order, when its producer wrote one (IMVS_TUNE_LAT32P=1): no hi + lo
    // reconstruction (96 of the ~200 instructions per pixel and 8-channel chunk), same coalesced reads -- measured: inner1
    // 48.5 -> 49.2 us, inner2 16.7 -> 18.3 us (gpurun call r2c46), so the split planes stay the source
    const float* coarse32p = nullptr;
    float* out32p = nullptr; // this stage's output in that order (coarse map of the next lateral stage) or null
    template <int NCH> struct Pre {};
    template <int NB, int NCH>
    __device__ __forceinline__ void prefetch(int, int, int, int, Pre<NCH>&) const {}
    __device__ __forceinline__ static void unpack8(const uint4& h, const uint4& l, float (&o)[8]) {
        const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hh[q]));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&ll[q]));
            o[2 * q] = a.x + b.x;
            o[2 * q + 1] = a.y + b.y;
        }
    }
    template <int NB, int NCH>
    __device__ __forceinline__ void store(int n, int oy, int ox, int c0, float (&v)[NCH], const Pre<NCH>&, int* status) const {
        const int Hc = H / 2, Wc = W / 2;
        int h0, h1, w0, w1;
        float lh, lw;
        up_index(oy, 0.5f, Hc, h0, h1, lh);
        up_index(ox, 0.5f, Wc, w0, w1, lw);
        const size_t plane = (size_t)H * W, pix = (size_t)oy * W + ox, cplane = (size_t)Hc * Wc;
        const uint4* chi = reinterpret_cast<const uint4*>(coarse.hi);
        const uint4* clo = reinterpret_cast<const uint4*>(coarse.lo);
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < NCH / 8; ++j) {
            float* x = v + 8 * j;
            const size_t cb = ((size_t)n * (NB / 8) + (c0 / 8 + j)) * cplane;
            const size_t i00 = cb + (size_t)h0 * Wc + w0, i01 = cb + (size_t)h0 * Wc + w1, i10 = cb + (size_t)h1 * Wc + w0, i11 = cb + (size_t)h1 * Wc + w1;
            const float4 bi0 = ldg4(bias + c0 + 8 * j), bi1 = ldg4(bias + c0 + 8 * j + 4);
            const float bi[8] = {bi0.x, bi0.y, bi0.z, bi0.w, bi1.x, bi1.y, bi1.z, bi1.w};
            float a[8], b[8], c[8], d[8];
            if (coarse32p) {
                const float4* cp = reinterpret_cast<const float4*>(coarse32p);
                auto ld8 = [&](size_t i, float (&o)[8]) {
                    const float4 u = __ldg(cp + 2 * i), w = __ldg(cp + 2 * i + 1);
                    o[0] = u.x; o[1] = u.y; o[2] = u.z; o[3] = u.w; o[4] = w.x; o[5] = w.y; o[6] = w.z; o[7] = w.w;
                };
                ld8(i00, a); ld8(i01, b); ld8(i10, c); ld8(i11, d);
            } else {
                const uint4 ah = __ldg(chi + i00), al = __ldg(clo + i00), bh = __ldg(chi + i01), bl = __ldg(clo + i01);
                const uint4 ch = __ldg(chi + i10), cl = __ldg(clo + i10), dh = __ldg(chi + i11), dl = __ldg(clo + i11);
                unpack8(ah, al, a); unpack8(bh, bl, b); unpack8(ch, cl, c); unpack8(dh, dl, d);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                // same association as EpiAddUp2 / the oracle: up + (conv + bias)
                x[q] = ((1.f - lh) * ((1.f - lw) * a[q] + lw * b[q]) + lh * ((1.f - lw) * c[q] + lw * d[q])) + (x[q] + bi[q]);
                amax = fmaxf(amax, fabsf(x[q]));
            }
            uint4 h, l;
            split_f16(make_float2(x[0], x[1]), h.x, l.x);
            split_f16(make_float2(x[2], x[3]), h.y, l.y);
            split_f16(make_float2(x[4], x[5]), h.z, l.z);
            split_f16(make_float2(x[6], x[7]), h.w, l.w);
            const size_t idx = ((size_t)n * (NB / 8) + (c0 / 8 + j)) * plane + pix;
            reinterpret_cast<uint4*>(out.hi)[idx] = h;
            reinterpret_cast<uint4*>(out.lo)[idx] = l;
            if (out32p) {
                float4* o = reinterpret_cast<float4*>(out32p) + idx * 2;
                o[0] = make_float4(x[0], x[1], x[2], x[3]);
                o[1] = make_float4(x[4], x[5], x[6], x[7]);
            }
        }
        if (out32) {
            float* o = out32 + ((size_t)n * plane + pix) * NB + c0;
#pragma unroll
            for (int c = 0; c < NCH; c += 4) *reinterpret_cast<float4*>(o + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
        if (!(amax <= 65504.f) && status) atomicOr(status, 2);
    }
};

#endif  // !CUSIM
// block-0 of a residual stage: conv1 (ReLU) and downsample (no ReLU) read the same input with the same
// stride-2 stencil -> one implicit GEMM over the stacked output channels [conv1 | downsample]
struct EpiSplit2 {
    float* out_relu;         // couts [0, CO)      -> relu(v + bias)
    float* out_lin;          // couts [CO, 2*CO)   -> v + bias
    const float* bias;       // [2*CO]
    int H, W, CO;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT], int) const {
        if (oy >= H || ox >= W) return;
        const size_t base = (((size_t)n * H + oy) * W + ox) * CO;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = co0 + 8 * j + 2 * t;
            const float a = v[2 * j] + ldg(bias + co), b = v[2 * j + 1] + ldg(bias + co + 1);
            if (co < CO) *reinterpret_cast<float2*>(out_relu + base + co) = make_float2(fmaxf(a, 0.f), fmaxf(b, 0.f));
            else *reinterpret_cast<float2*>(out_lin + base + co - CO) = make_float2(a, b);
        }
    }
};

// conv1 of FeatureNet (net.py:13): 3 -> 8 channels, 3x3, pad 1, folded BN + ReLU; planar image in, NHWC-8 out.
// K = 27: the tensor-core engine pads this layer to K = 144 (one 16-channel k-step x 9 taps) and spends its
// time staging zeros (112 us at 5 x 640 x 512).  Here it is exact fp32 FFMA: a 32 x 8-pixel tile per
// 128-thread block, the three image planes with halo in shared memory, each thread two vertically adjacent
// pixels x 8 output channels (432 FFMA per 12 image + 54 broadcast weight LDS).
// PixT = float: the image as the reference's loaders deliver it; PixT = unsigned char: the raw 8-bit image, normalised here
// exactly as the loaders do (2 * np.array(img, float32) / 255. - 1, datasets/dtu_yao_eval.py:63-64: float32 multiply, divide,
// subtract in that order) -- a quarter of the H2D bytes.
constexpr int C0_TW = 32, C0_TH = 8, C0_PITCH = C0_TW + 2;
__device__ __forceinline__ float c0_pixel(const float* p) { return ldg(p); }
__device__ __forceinline__ float c0_pixel(const unsigned char* p) { return __fsub_rn(__fdiv_rn(2.0f * (float)__ldg(p), 255.0f), 1.0f); }
// PSPLIT: the output goes out as fp16 hi / lo PARITY PLANES [4N][1][H/2][W/2][8] (tc5pconv.cuh: operand of layer1's stride-2
// GEMM on the TMA + tcgen05 kernel) instead of fp32 NHWC-8; `out` then points at the hi plane, the lo plane follows it.
template <class PixT, bool PSPLIT = false>
__global__ void __launch_bounds__(128)
fnet_conv0_kernel(const PixT* __restrict__ img, const float* __restrict__ wgt, const float* __restrict__ bias,
                  float* __restrict__ out, int H, int W) {
    __shared__ float sI[3][C0_TH + 2][C0_PITCH];
    __shared__ __align__(16) float sW[27][8];       // [(ky*3 + kx)*3 + cin][cout]
    __shared__ __align__(16) float sB[8];
    const int tid = threadIdx.x, n = blockIdx.z;
    const int x0 = blockIdx.x * C0_TW, y0 = blockIdx.y * C0_TH;
    pdl_trigger();
    for (int i = tid; i < 27 * 8; i += 128) {       // packed weight: [tap][8 cin (3 used)][8 cout]
        const int row = i >> 3, co = i & 7, tap = row / 3, c = row % 3;
        sW[row][co] = ldg(wgt + (tap * 8 + c) * 8 + co);
    }
    if (tid < 8) sB[tid] = ldg(bias + tid);
    pdl_wait();
    const size_t plane = (size_t)H * W;
    const PixT* src = img + (size_t)n * 3 * plane;
    for (int i = tid; i < 3 * (C0_TH + 2) * C0_PITCH; i += 128) {
        const int c = i / ((C0_TH + 2) * C0_PITCH), rem = i % ((C0_TH + 2) * C0_PITCH);
        const int r = rem / C0_PITCH, col = rem % C0_PITCH;
        const int y = y0 - 1 + r, x = x0 - 1 + col;
        sI[c][r][col] = (y >= 0 && y < H && x >= 0 && x < W) ? c0_pixel(src + c * plane + (size_t)y * W + x) : 0.f;
    }
    __syncthreads();
    const int tx = tid & 31, ty = tid >> 5;
    float a0[8], a1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a0[k] = a1[k] = sB[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v[4][3];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) v[r][k] = sI[c][2 * ty + r][tx + k];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4 wa = *reinterpret_cast<const float4*>(&sW[(ky * 3 + kx) * 3 + c][0]);
                const float4 wb = *reinterpret_cast<const float4*>(&sW[(ky * 3 + kx) * 3 + c][4]);
                const float p = v[ky][kx], q = v[ky + 1][kx];
                a0[0] = fmaf(p, wa.x, a0[0]); a0[1] = fmaf(p, wa.y, a0[1]); a0[2] = fmaf(p, wa.z, a0[2]); a0[3] = fmaf(p, wa.w, a0[3]);
                a0[4] = fmaf(p, wb.x, a0[4]); a0[5] = fmaf(p, wb.y, a0[5]); a0[6] = fmaf(p, wb.z, a0[6]); a0[7] = fmaf(p, wb.w, a0[7]);
                a1[0] = fmaf(q, wa.x, a1[0]); a1[1] = fmaf(q, wa.y, a1[1]); a1[2] = fmaf(q, wa.z, a1[2]); a1[3] = fmaf(q, wa.w, a1[3]);
                a1[4] = fmaf(q, wb.x, a1[4]); a1[5] = fmaf(q, wb.y, a1[5]); a1[6] = fmaf(q, wb.z, a1[6]); a1[7] = fmaf(q, wb.w, a1[7]);
            }
    }
    const int x = x0 + tx, y = y0 + 2 * ty;
    if (x >= W) return;
#ifndef CUSIM     // (the CPU emulation of the test-suite has no TMA / tcgen05 model)
    if constexpr (PSPLIT) {
        uint4* ohi = reinterpret_cast<uint4*>(out);
        uint4* olo = ohi + (size_t)gridDim.z * H * W;            // 16-byte units: N * H * W pixels x one 8-channel chunk
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (y + r >= H) break;
            const float* a = r ? a1 : a0;
            uint4 h, l;
            split_f16(make_float2(fmaxf(a[0], 0.f), fmaxf(a[1], 0.f)), h.x, l.x);
            split_f16(make_float2(fmaxf(a[2], 0.f), fmaxf(a[3], 0.f)), h.y, l.y);
            split_f16(make_float2(fmaxf(a[4], 0.f), fmaxf(a[5], 0.f)), h.z, l.z);
            split_f16(make_float2(fmaxf(a[6], 0.f), fmaxf(a[7], 0.f)), h.w, l.w);
            const size_t idx = tc5p::parity_index(n, y + r, x, 0, 1, H, W);
            ohi[idx] = h;
            olo[idx] = l;
        }
        return;
    }
#endif  // !CUSIM
    if (y < H) {
        float4* o = reinterpret_cast<float4*>(out + (((size_t)n * H + y) * W + x) * 8);
        o[0] = make_float4(fmaxf(a0[0], 0.f), fmaxf(a0[1], 0.f), fmaxf(a0[2], 0.f), fmaxf(a0[3], 0.f));
        o[1] = make_float4(fmaxf(a0[4], 0.f), fmaxf(a0[5], 0.f), fmaxf(a0[6], 0.f), fmaxf(a0[7], 0.f));
    }
    if (y + 1 < H) {
        float4* o = reinterpret_cast<float4*>(out + (((size_t)n * H + y + 1) * W + x) * 8);
        o[0] = make_float4(fmaxf(a1[0], 0.f), fmaxf(a1[1], 0.f), fmaxf(a1[2], 0.f), fmaxf(a1[3], 0.f));
        o[1] = make_float4(fmaxf(a1[4], 0.f), fmaxf(a1[5], 0.f), fmaxf(a1[6], 0.f), fmaxf(a1[7], 0.f));
    }
}

// The same layer with FOUR vertically adjacent pixels per thread (32 x 16 tile): a tap's eight weights are read once for four
// pixels (864 FFMA per 54 + 54 shared-memory reads instead of 432 per 54 + 36), 1.9x fewer instructions per pixel -- and
// measured slower on B200 (51.5 us against 41.1 us: half as many CTAs / resident warps for a kernel that is bound by its
// load -> barrier -> compute -> store latency, not by issue slots).  Kept behind IMVS_TUNE_CONV0R4=1.
constexpr int C4_TH = 16;
template <class PixT, bool PSPLIT>
__global__ void __launch_bounds__(128)
fnet_conv0_r4_kernel(const PixT* __restrict__ img, const float* __restrict__ wgt, const float* __restrict__ bias,
                     float* __restrict__ out, int H, int W) {
    __shared__ float sI[3][C4_TH + 2][C0_PITCH];
    __shared__ __align__(16) float sW[27][8];       // [(ky*3 + kx)*3 + cin][cout]
    __shared__ __align__(16) float sB[8];
    const int tid = threadIdx.x, n = blockIdx.z;
    const int x0 = blockIdx.x * C0_TW, y0 = blockIdx.y * C4_TH;
    pdl_trigger();
    for (int i = tid; i < 27 * 8; i += 128) {       // packed weight: [tap][8 cin (3 used)][8 cout]
        const int row = i >> 3, co = i & 7, tap = row / 3, c = row % 3;
        sW[row][co] = ldg(wgt + (tap * 8 + c) * 8 + co);
    }
    if (tid < 8) sB[tid] = ldg(bias + tid);
    pdl_wait();
    const size_t plane = (size_t)H * W;
    const PixT* src = img + (size_t)n * 3 * plane;
    for (int i = tid; i < 3 * (C4_TH + 2) * C0_PITCH; i += 128) {
        const int c = i / ((C4_TH + 2) * C0_PITCH), rem = i % ((C4_TH + 2) * C0_PITCH);
        const int r = rem / C0_PITCH, col = rem % C0_PITCH;
        const int y = y0 - 1 + r, x = x0 - 1 + col;
        sI[c][r][col] = (y >= 0 && y < H && x >= 0 && x < W) ? c0_pixel(src + c * plane + (size_t)y * W + x) : 0.f;
    }
    __syncthreads();
    const int tx = tid & 31, ty = tid >> 5;          // rows 4 ty .. 4 ty + 3 of the tile
    float acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[r][k] = sB[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v[6][3];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) v[r][k] = sI[c][4 * ty + r][tx + k];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4 wa = *reinterpret_cast<const float4*>(&sW[(ky * 3 + kx) * 3 + c][0]);
                const float4 wb = *reinterpret_cast<const float4*>(&sW[(ky * 3 + kx) * 3 + c][4]);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float p = v[r + ky][kx];
                    acc[r][0] = fmaf(p, wa.x, acc[r][0]); acc[r][1] = fmaf(p, wa.y, acc[r][1]); acc[r][2] = fmaf(p, wa.z, acc[r][2]); acc[r][3] = fmaf(p, wa.w, acc[r][3]);
                    acc[r][4] = fmaf(p, wb.x, acc[r][4]); acc[r][5] = fmaf(p, wb.y, acc[r][5]); acc[r][6] = fmaf(p, wb.z, acc[r][6]); acc[r][7] = fmaf(p, wb.w, acc[r][7]);
                }
            }
    }
    const int x = x0 + tx;
    if (x >= W) return;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int y = y0 + 4 * ty + r;
        if (y >= H) break;
        const float* a = acc[r];
        if constexpr (PSPLIT) {
#ifndef CUSIM
            uint4* ohi = reinterpret_cast<uint4*>(out);
            uint4* olo = ohi + (size_t)gridDim.z * H * W;            // 16-byte units: N * H * W pixels x one 8-channel chunk
            uint4 h, l;
            split_f16(make_float2(fmaxf(a[0], 0.f), fmaxf(a[1], 0.f)), h.x, l.x);
            split_f16(make_float2(fmaxf(a[2], 0.f), fmaxf(a[3], 0.f)), h.y, l.y);
            split_f16(make_float2(fmaxf(a[4], 0.f), fmaxf(a[5], 0.f)), h.z, l.z);
            split_f16(make_float2(fmaxf(a[6], 0.f), fmaxf(a[7], 0.f)), h.w, l.w);
            const size_t idx = tc5p::parity_index(n, y, x, 0, 1, H, W);
            ohi[idx] = h;
            olo[idx] = l;
#endif
        } else {
            float4* o = reinterpret_cast<float4*>(out + (((size_t)n * H + y) * W + x) * 8);
            o[0] = make_float4(fmaxf(a[0], 0.f), fmaxf(a[1], 0.f), fmaxf(a[2], 0.f), fmaxf(a[3], 0.f));
            o[1] = make_float4(fmaxf(a[4], 0.f), fmaxf(a[5], 0.f), fmaxf(a[6], 0.f), fmaxf(a[7], 0.f));
        }
    }
}

struct FnetBuffers {
    float *a0, *l1[4], *l2[4], *l3[4], *intra2, *intra1;
    float *l1s, *l2s, *l3s, *intra2s;     // split-plane copies (tc5pconv.cuh) of the trunk outputs and of intra2, which are also read as fp32
    size_t total;
};

static FnetBuffers fnet_carve(float* base, size_t N, size_t H, size_t W) {
    FnetBuffers b;
    size_t c = 0;
    auto at = [&](size_t n) { float* p = base + c; c += (n + 63) / 64 * 64; return p; };
    const size_t hw = H * W;
    b.a0 = at(N * hw * 8);
    for (int i = 0; i < 4; ++i) b.l1[i] = at(N * (hw / 4) * 16);
    for (int i = 0; i < 4; ++i) b.l2[i] = at(N * (hw / 16) * 32);
    for (int i = 0; i < 4; ++i) b.l3[i] = at(N * (hw / 64) * 48);
    b.intra2 = at(N * (hw / 16) * 48);
    b.intra1 = at(N * (hw / 4) * 48);
    b.l1s = at(N * (hw / 4) * 16);
    b.l2s = at(N * (hw / 16) * 32);
    b.l3s = at(N * (hw / 64) * 48);
    b.intra2s = at(N * (hw / 16) * 48);
    b.total = c;
    return b;
}

// Stride-1 3x3 convolution CI -> CO (bias, optional residual, optional ReLU), NHWC.  Default (fp32-grade mode 4): the
// tcgen05 kernel with the fp16 hi/lo 3-product split (tc5conv.cuh:tc5h_conv_kernel); otherwise / under the CPU emulation
// of the test-suite the mma.sync engine with the tile shape passed in.
template <int CI, int CO, class Fallback>
static int conv3x3_s1(const char* name, const imvs_wpair& wp, const float* bias, const float* x, float* out, const float* residual,
                      int relu, int N, int H, int W, cudaStream_t st, Fallback&& fallback) {
#ifndef CUSIM
    if (conv_passes() == 4 && wp.f16umma && tune("TC5H", 1))
        return tc5::launch_h<CI, CO>(name, in_nhwc(x, H, W, CI), tc5::PixNHWC{out, bias, residual, H, W, CO, relu}, wp.f16umma, 3, 1, N,
                                     H, W, tc5_error_flag(), st);
#endif
    return fallback();
}

// one residual stage (two ResidualBlocks, module.py:32-50):  x [Hin][Win][CI] -> buf[3] [Hin/2][Win/2][CO]
// NBA / NBB: cout block of the stride-2 [conv1 | downsample] GEMM / of the CO -> CO convolutions; MT: row-tiles per warp
template <int CI, int CO, bool WALL_A, bool WALL_B, int NBA = 2 * CO, int NBB = CO, int MT = 2, int WARPS = 4>
static int res_stage(const imvs_featurenet_weights* w, int L, const float* x, float* const buf[4], int N, int Hin, int Win,
                     cudaStream_t st) {
    const int H = Hin / 2, W = Win / 2;
    const TapTables s2 = conv_tables(3, 2, 1, WARPS * MT), s1 = conv_tables(3, 1, 1, WARPS * MT);
    constexpr int NCA = 2 * CO / NBA, NCB = CO / NBB;
    // block 0: [conv1 (stride 2, relu) | downsample (stride 2)] as one GEMM, then conv2 + downsample -> relu
    const int LS = 21 + (L - 1) / 5;           // stacked [conv1 | downsample] weights of this stage
    IMVS_TRY((mma_conv<CI, NBA, MT, WARPS, 2, WALL_A>("fnet.block0.conv1|downsample", in_nhwc(x, Hin, Win, CI),
                                                  EpiSplit2{buf[0], buf[1], w->b[LS], H, W, CO}, WSets::single(w->w[LS]), s2, N, 2 * CO,
                                                  H, W, NCA, st)));
    IMVS_TRY((conv3x3_s1<CO, CO>("fnet.block0.conv2", w->w[L + 1], w->b[L + 1], buf[0], buf[2], buf[1], 1, N, H, W, st, [&] {
        return mma_conv<CO, NBB, MT, WARPS, 1, WALL_B>("fnet.block0.conv2", in_nhwc(buf[0], H, W, CO), EpiNHWC{buf[2], w->b[L + 1], buf[1], H, W, CO, CO, 1},
                                                       WSets::single(w->w[L + 1]), s1, N, CO, H, W, NCB, st); })));
    // block 1: conv1 relu, conv2 + x -> relu
    IMVS_TRY((conv3x3_s1<CO, CO>("fnet.block1.conv1", w->w[L + 3], w->b[L + 3], buf[2], buf[0], nullptr, 1, N, H, W, st, [&] {
        return mma_conv<CO, NBB, MT, WARPS, 1, WALL_B>("fnet.block1.conv1", in_nhwc(buf[2], H, W, CO), EpiNHWC{buf[0], w->b[L + 3], nullptr, H, W, CO, CO, 1},
                                                       WSets::single(w->w[L + 3]), s1, N, CO, H, W, NCB, st); })));
    IMVS_TRY((conv3x3_s1<CO, CO>("fnet.block1.conv2", w->w[L + 4], w->b[L + 4], buf[0], buf[3], buf[2], 1, N, H, W, st, [&] {
        return mma_conv<CO, NBB, MT, WARPS, 1, WALL_B>("fnet.block1.conv2", in_nhwc(buf[0], H, W, CO), EpiNHWC{buf[3], w->b[L + 4], buf[2], H, W, CO, CO, 1},
                                                       WSets::single(w->w[L + 4]), s1, N, CO, H, W, NCB, st); })));
    return 0;
}

#ifndef CUSIM     // (the CPU emulation of the test-suite has no TMA / tcgen05 model)
// The same residual stage on the persistent TMA + tcgen05 kernel (tc5pconv.cuh): every activation between the stride-2
// GEMM and the stage's last convolution lives as fp16 hi / lo split planes (written by the producers' epilogues, loaded
// by TMA, never converted by a thread); the trunk output is fp32 NHWC for its fp32 consumers (next stride-2 GEMM, lateral
// 1x1) plus, for stage 3, split planes for output3.
template <int CI, int CO, bool WALL_A, int NBA = 2 * CO, int MT = 2, int WARPS = 4>
static int res_stage_p(const imvs_featurenet_weights* w, int L, const float* x, float* const buf[4], float* trunk_split, int N, int Hin,
                       int Win, cudaStream_t st) {
    const int H = Hin / 2, W = Win / 2;
    const size_t elems = (size_t)N * H * W * CO;
    const tc5p::Split y1 = tc5p::split_at(buf[0], elems), ds = tc5p::split_at(buf[1], elems), b0 = tc5p::split_at(buf[2], elems);
    const tc5p::Split none{nullptr, nullptr}, ts = trunk_split ? tc5p::split_at(trunk_split, elems) : none;
    const TapTables s2 = conv_tables(3, 2, 1, WARPS * MT);
    constexpr int NCA = 2 * CO / NBA;
    const int LS = 21 + (L - 1) / 5;
    int* flag = tc5_error_flag();
    IMVS_TRY((mma_conv<CI, NBA, MT, WARPS, 2, WALL_A>("fnet.block0.conv1|downsample", in_nhwc(x, Hin, Win, CI),
                                                  EpiSplit2H{y1, ds, w->b[LS], H, W, CO}, WSets::single(w->w[LS]), s2, N, 2 * CO, H, W, NCA, st)));
    IMVS_TRY((tc5p::launch<CO, CO>("fnet.block0.conv2", y1, tc5p::Epi{b0, nullptr, ds, w->b[L + 1], H, W, 1}, w->w[L + 1].f16ummai, N, H, W, flag, st)));
    IMVS_TRY((tc5p::launch<CO, CO>("fnet.block1.conv1", b0, tc5p::Epi{y1, nullptr, none, w->b[L + 3], H, W, 1}, w->w[L + 3].f16ummai, N, H, W, flag, st)));
    IMVS_TRY((tc5p::launch<CO, CO>("fnet.block1.conv2", y1, tc5p::Epi{ts, buf[3], b0, w->b[L + 4], H, W, 1}, w->w[L + 4].f16ummai, N, H, W, flag, st)));
    return 0;
}

// ... and with the stride-2 GEMM on that kernel too: the stage's input arrives as PARITY PLANES (tc5pconv.cuh) from its
// producer (conv1 / the previous stage's last epilogue); the trunk goes out as fp32 NHWC (trunk32), split planes (trunk_split)
// and / or parity planes for the next stage (trunk_parity), whichever are non-null.  H, W: the stage's OUTPUT size.
template <int CI, int CO>
static int res_stage_p2(const imvs_featurenet_weights* w, int L, const tc5p::Split& xp, float* const buf[4], float* trunk32,
                        float* trunk_split, float* trunk_parity, int N, int H, int W, cudaStream_t st, float* trunk32p = nullptr) {
    const size_t elems = (size_t)N * H * W * CO;
    const tc5p::Split y1 = tc5p::split_at(buf[0], elems), ds = tc5p::split_at(buf[1], elems), b0 = tc5p::split_at(buf[2], elems);
    const tc5p::Split none{nullptr, nullptr}, ts = trunk_split ? tc5p::split_at(trunk_split, elems) : none,
                      tp = trunk_parity ? tc5p::split_at(trunk_parity, elems) : none;
    const int LS = 21 + (L - 1) / 5;
    int* flag = tc5_error_flag();
    if constexpr (2 * CO <= 64) {
        IMVS_TRY((tc5p::launch<CI, 2 * CO, 1, true, 3, 2>("fnet.block0.conv1|downsample", xp, tc5p::EpiStack2{y1, ds, w->b[LS], H, W, CO},
                                                          w->w[LS].f16ummai, N, H, W, flag, st)));
    } else {
        IMVS_TRY((tc5p::launch<CI, CO, 1, true, 3, 2>("fnet.block0.conv1", xp, tc5p::Epi{y1, nullptr, none, w->b[L], H, W, 1}, w->w[L].f16ummai,
                                                      N, H, W, flag, st)));
        IMVS_TRY((tc5p::launch<CI, CO, 1, true, 3, 2>("fnet.block0.downsample", xp, tc5p::Epi{ds, nullptr, none, w->b[L + 2], H, W, 0},
                                                      w->w[L + 2].f16ummai, N, H, W, flag, st)));
    }
    IMVS_TRY((tc5p::launch<CO, CO>("fnet.block0.conv2", y1, tc5p::Epi{b0, nullptr, ds, w->b[L + 1], H, W, 1}, w->w[L + 1].f16ummai, N, H, W, flag, st)));
    IMVS_TRY((tc5p::launch<CO, CO>("fnet.block1.conv1", b0, tc5p::Epi{y1, nullptr, none, w->b[L + 3], H, W, 1}, w->w[L + 3].f16ummai, N, H, W, flag, st)));
    IMVS_TRY((tc5p::launch<CO, CO>("fnet.block1.conv2", y1, tc5p::Epi{ts, trunk32, b0, w->b[L + 4], H, W, 1, tp, 0, trunk32p}, w->w[L + 4].f16ummai, N, H, W, flag, st)));
    return 0;
}

static bool fnet_tc5p_ready(const imvs_featurenet_weights* w) {
#ifdef CUSIM
    return false;
#else
    if (conv_passes() != 4 || !tune("TC5P", 1) || !tc5p::encode_tiled_fn()) return false;
    for (int L : {2, 4, 5, 7, 9, 10, 12, 14, 15, 16, 18, 20})
        if (!w->w[L].f16ummai) return false;
    return true;
#endif
}

namespace tc5p {

EncodeTiledFn encode_tiled_fn() {
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q{};
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return (EncodeTiledFn) nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int& c = cached[dev & 63];
    if (!c && cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) c = 148;
    return c > 0 ? c : 148;
}

// layout conversion for the operator-level entry point (inside FeatureNet the producers write split planes directly)
// Persistent launches take one CTA per SM and hold it (TMEM, up to 220 KB of shared memory) until their last tile.  A serving
// loop that keeps several reference views in flight on different streams can let each launch use 1 / share of the SMs, so
// that launches of different streams run side by side and each one's launch gap, prologue and tail overlap the other's work
// (imvs_set_sm_share; graph.StreamingPipeline sets it while it captures its slots).  1 = the whole GPU (latency mode).
static int g_sm_share = 1;
int grid_limit() {
    const int t = tune("TC5P_SHARE", 0);
    const int share = t > 0 ? t : g_sm_share;
    return std::max(1, sm_count() / std::max(1, share));
}
void set_sm_share(int share) { g_sm_share = share < 1 ? 1 : share; }
int get_sm_share() { return g_sm_share; }

__global__ void nhwc_to_split_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, size_t npix_total,
                                     int HW, int C) {
    pdl_trigger();
    pdl_wait();
    const int KC = C / 8;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // (n, kc, pixel)
    if (i >= npix_total * KC) return;
    const size_t pixg = i % ((size_t)HW), t = i / HW;
    const int kc = (int)(t % KC);
    const size_t n = t / KC;
    const float* src = x + ((n * HW + pixg) * C + kc * 8);
    const float4 a = ldg4(src), b = ldg4(src + 4);
    uint4 h, l;
    split_f16(make_float2(a.x, a.y), h.x, l.x);
    split_f16(make_float2(a.z, a.w), h.y, l.y);
    split_f16(make_float2(b.x, b.y), h.z, l.z);
    split_f16(make_float2(b.z, b.w), h.w, l.w);
    reinterpret_cast<uint4*>(hi)[i] = h;
    reinterpret_cast<uint4*>(lo)[i] = l;
}

int launch_nhwc_to_split(const float* x, const Split& dst, size_t npix_total, int HW, int C, cudaStream_t st) {
    IMVS_REQUIRE(x && dst.hi && dst.lo && C % 8 == 0, "nhwc_to_split: bad argument");
    const size_t n = npix_total * (C / 8);
    IMVS_REQUIRE(n < 4294967296ull * 256, "nhwc_to_split: too many pixels");
    IMVS_CUDA(launch_k(nhwc_to_split_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, x, dst.hi, dst.lo, npix_total, HW, C));
    return 0;
}

__global__ void split_to_nhwc_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ x, size_t npix_total,
                                     int HW, int C) {
    pdl_trigger();
    pdl_wait();
    const int KC = C / 8;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // (n, kc, pixel)
    if (i >= npix_total * KC) return;
    const size_t pixg = i % ((size_t)HW), t = i / HW;
    const int kc = (int)(t % KC);
    const size_t n = t / KC;
    const uint4 h = reinterpret_cast<const uint4*>(hi)[i], l = reinterpret_cast<const uint4*>(lo)[i];
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
    float o[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hh[q])), b = __half22float2(*reinterpret_cast<const __half2*>(&ll[q]));
        o[2 * q] = a.x + b.x; o[2 * q + 1] = a.y + b.y;
    }
    float* dst = x + ((n * HW + pixg) * C + kc * 8);
    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
}

}  // namespace tc5p

#endif  // !CUSIM
}  // namespace imvs

using namespace imvs;

extern "C" size_t imvs_featurenet_workspace_bytes(int N, int H, int W) {
    if (N < 1 || H < 8 || W < 8) return 0;
    return fnet_carve(nullptr, N, H, W).total * sizeof(float);
}

// 18 convolution launches; 19 in the default mode, where layer3's [conv1 | downsample] (96 stacked channels > the tcgen05 kernel's
// 64) are two launches
extern "C" int imvs_featurenet_launch_count(void) {
#ifdef CUSIM
    return 18;
#else
    return (conv_passes() == 4 && tune("TC5P", 1) && tune("TC5P_LAT", 1) && tune("TC5P_S2", 1) && tc5p::encode_tiled_fn()) ? 19 : 18;
#endif
}

static int featurenet_forward_impl(const imvs_featurenet_weights* w, const float* imgs, const unsigned char* imgs_u8, float* fea1,
                                   float* fea2, float* fea3, void* workspace, size_t workspace_bytes, int N, int H, int W, void* stream) {
    IMVS_REQUIRE(w && (imgs || imgs_u8) && fea1 && fea2 && fea3 && workspace, "featurenet_forward: null pointer");
    IMVS_REQUIRE(N >= 1 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0, "featurenet_forward: H, W must be multiples of 8 (H=%d W=%d)", H, W);
    IMVS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "featurenet_forward: workspace must be 256-byte aligned");
    FnetBuffers b = fnet_carve(static_cast<float*>(workspace), N, H, W);
    IMVS_REQUIRE(workspace_bytes >= b.total * sizeof(float), "featurenet_forward: workspace too small (%zu < %zu bytes)", workspace_bytes,
                 b.total * sizeof(float));
    cudaStream_t st = (cudaStream_t)stream;
    ApiScope api_;
    StageTimer tm_(ST_FEATURENET, stream);
    const int H1 = H / 2, W1 = W / 2, H2 = H / 4, W2 = W / 4, H3 = H / 8, W3 = W / 8;
    const TapTables s1 = conv_tables(3, 1, 1, 8), k1 = conv_tables(1, 1, 1, 8);
    // conv1: 3 -> 8, BN, ReLU on the planar image (net.py:13)
#ifndef CUSIM
    const bool p_ready = fnet_tc5p_ready(w);
    const bool lat = p_ready && tune("TC5P_LAT", 1) && w->w[17].f16ummai && w->w[19].f16ummai;
    // stride-2 GEMMs on the TMA + tcgen05 kernel too (then no layer of the default FeatureNet runs on mma.sync)
    const bool s2p = lat && tune("TC5P_S2", 1) && w->w[21].f16ummai && w->w[22].f16ummai && w->w[11].f16ummai && w->w[13].f16ummai;
#else
    const bool s2p = false;
#endif
    if (imgs_u8 || tune("CONV0", 1)) {
        IMVS_REQUIRE(w->w[0].fp32 && w->b[0], "featurenet_forward: conv1 weights missing");
        dim3 grid(cdiv(W, C0_TW), cdiv(H, C0_TH), N);
        IMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "fnet.conv1: grid too large");
        const bool r4 = tune("CONV0R4", 0) != 0;         // four rows per thread (32 x 16 tiles): measured SLOWER (51 vs 41 us, gpurun call r2c44), off
        dim3 grid4(cdiv(W, C0_TW), cdiv(H, C4_TH), N);
        if (s2p) {      // output as fp16 hi / lo parity planes (same bytes, in a0)
#ifndef CUSIM
            if (r4) {
                if (imgs_u8) IMVS_CUDA(launch_k(fnet_conv0_r4_kernel<unsigned char, true>, grid4, dim3(128), 0, st, imgs_u8, w->w[0].fp32, w->b[0], b.a0, H, W));
                else IMVS_CUDA(launch_k(fnet_conv0_r4_kernel<float, true>, grid4, dim3(128), 0, st, imgs, w->w[0].fp32, w->b[0], b.a0, H, W));
            } else {
                if (imgs_u8) IMVS_CUDA(launch_k(fnet_conv0_kernel<unsigned char, true>, grid, dim3(128), 0, st, imgs_u8, w->w[0].fp32, w->b[0], b.a0, H, W));
                else IMVS_CUDA(launch_k(fnet_conv0_kernel<float, true>, grid, dim3(128), 0, st, imgs, w->w[0].fp32, w->b[0], b.a0, H, W));
            }
#endif
        } else if (r4) {
            if (imgs_u8) IMVS_CUDA(launch_k(fnet_conv0_r4_kernel<unsigned char, false>, grid4, dim3(128), 0, st, imgs_u8, w->w[0].fp32, w->b[0], b.a0, H, W));
            else IMVS_CUDA(launch_k(fnet_conv0_r4_kernel<float, false>, grid4, dim3(128), 0, st, imgs, w->w[0].fp32, w->b[0], b.a0, H, W));
        } else {
            if (imgs_u8) IMVS_CUDA(launch_k(fnet_conv0_kernel<unsigned char>, grid, dim3(128), 0, st, imgs_u8, w->w[0].fp32, w->b[0], b.a0, H, W));
            else IMVS_CUDA(launch_k(fnet_conv0_kernel<float>, grid, dim3(128), 0, st, imgs, w->w[0].fp32, w->b[0], b.a0, H, W));
        }
    } else {
        IMVS_TRY((mma_conv<8, 8, 2, 4, 1, true>("fnet.conv1", InNCHW3{imgs, H, W}, EpiNHWC{b.a0, w->b[0], nullptr, H, W, 8, 8, 1},
                                                WSets::single(w->w[0]), s1, N, 8, H, W, 1, st)));
    }
#ifndef CUSIM
    if (p_ready) {
        // default: residual stages and output convolutions on the persistent TMA + tcgen05 kernel, split-plane activations
        int* flag = tc5_error_flag();
        const tc5p::Split none{nullptr, nullptr};
        if (s2p && (imgs_u8 || tune("CONV0", 1))) {
            // trunk of stage 1 / 2: split planes for the lateral 1x1 and parity planes (in the fp32 trunk's buffer) for the next stage
            IMVS_TRY((res_stage_p2<8, 16>(w, 1, tc5p::split_at(b.a0, (size_t)N * H * W * 8), b.l1, nullptr, b.l1s, b.l1[3], N, H1, W1, st)));
            IMVS_TRY((res_stage_p2<16, 32>(w, 6, tc5p::split_at(b.l1[3], (size_t)N * H1 * W1 * 16), b.l2, nullptr, b.l2s, b.l2[3], N, H2, W2, st)));
            IMVS_TRY((res_stage_p2<32, 48>(w, 11, tc5p::split_at(b.l2[3], (size_t)N * H2 * W2 * 32), b.l3, nullptr, b.l3s, nullptr, N, H3, W3, st, b.l3[3])));   // + fp32 chunk-planar copy in l3[3]: coarse map of inner2
        } else {
            IMVS_TRY((res_stage_p<8, 16, true>(w, 1, b.a0, b.l1, b.l1s, N, H, W, st)));                 // layer1 -> l1[3]  [H/2][W/2][16] (+ split)
            IMVS_TRY((res_stage_p<16, 32, true>(w, 6, b.l1[3], b.l2, b.l2s, N, H1, W1, st)));           // layer2 -> l2[3]  [H/4][W/4][32] (+ split)
            IMVS_TRY((res_stage_p<32, 48, false, 96, 1>(w, 11, b.l2[3], b.l3, b.l3s, N, H2, W2, st)));  // layer3 -> l3[3]  [H/8][W/8][48] (+ split)
        }
        const tc5p::Split l1s = tc5p::split_at(b.l1s, (size_t)N * H1 * W1 * 16), l2s = tc5p::split_at(b.l2s, (size_t)N * H2 * W2 * 32),
                          l3s = tc5p::split_at(b.l3s, (size_t)N * H3 * W3 * 48), i2s = tc5p::split_at(b.intra2s, (size_t)N * H2 * W2 * 48),
                          i1s = tc5p::split_at(b.intra1, (size_t)N * H1 * W1 * 48);
        IMVS_TRY((tc5p::launch<48, 48>("fnet.output3", l3s, tc5p::Epi{none, fea3, none, w->b[16], H3, W3, 0}, w->w[16].f16ummai, N, H3, W3, flag, st)));
        // intra2 = up2(f3) + inner2(f2) (net.py:60): 1x1 on the tensor core, bilinear taps in the epilogue
        const bool planar = s2p && (imgs_u8 || tune("CONV0", 1)) && tune("LAT32P", 0);      // fp32 chunk-planar coarse maps (in l3[3] / intra2): measured no gain (call r2c46), off
        if (lat) IMVS_TRY((tc5p::launch<32, 48, 1, true, 1>("fnet.inner2", l2s, EpiLateral{i2s, nullptr, w->b[17], l3s, H2, W2, planar ? b.l3[3] : nullptr, planar ? b.intra2 : nullptr}, w->w[17].f16ummai, N, H2, W2, flag, st)));
        else IMVS_TRY((mma_conv<32, 48, 2, 4, 1, true>("fnet.inner2", in_nhwc(b.l2[3], H2, W2, 32), EpiAddUp2H{b.intra2, i2s, w->b[17], b.l3[3], H2, W2, 48},
                                                        WSets::single(w->w[17]), k1, N, 48, H2, W2, 1, st)));
        IMVS_TRY((tc5p::launch<48, 32>("fnet.output2", i2s, tc5p::Epi{none, fea2, none, w->b[18], H2, W2, 0}, w->w[18].f16ummai, N, H2, W2, flag, st)));
        if (lat) IMVS_TRY((tc5p::launch<16, 48, 1, true, 1>("fnet.inner1", l1s, EpiLateral{i1s, nullptr, w->b[19], i2s, H1, W1, planar ? b.intra2 : nullptr, nullptr}, w->w[19].f16ummai, N, H1, W1, flag, st)));
        else IMVS_TRY((mma_conv<16, 48, 2, 4, 1, true>("fnet.inner1", in_nhwc(b.l1[3], H1, W1, 16), EpiAddUp2H{nullptr, i1s, w->b[19], b.intra2, H1, W1, 48},
                                                        WSets::single(w->w[19]), k1, N, 48, H1, W1, 1, st)));
        IMVS_TRY((tc5p::launch<48, 16>("fnet.output1", i1s, tc5p::Epi{none, fea1, none, w->b[20], H1, W1, 0}, w->w[20].f16ummai, N, H1, W1, flag, st)));
        return 0;
    }
#endif  // !CUSIM
    const int w8 = tune("FNETW", 0);     // 1: 8 warps x 1 row-tile per CTA instead of 4 x 2 (same tile, twice the resident warps)
    if (w8) IMVS_TRY((res_stage<8, 16, true, true, 32, 16, 1, 8>(w, 1, b.a0, b.l1, N, H, W, st)));
    else if (tune("FNET1", 0)) IMVS_TRY((res_stage<8, 16, true, true, 32, 16, 1, 4>(w, 1, b.a0, b.l1, N, H, W, st)));
    else IMVS_TRY((res_stage<8, 16, true, true>(w, 1, b.a0, b.l1, N, H, W, st)));          // layer1 -> l1[3]  [H/2][W/2][16]
    switch (tune("FNET2", 0)) {                                                          // layer2 -> l2[3]  [H/4][W/4][32]
        case 1: IMVS_TRY((res_stage<16, 32, true, false, 32, 16, 2>(w, 6, b.l1[3], b.l2, N, H1, W1, st))); break;
        case 2: IMVS_TRY((res_stage<16, 32, true, false, 64, 32, 1>(w, 6, b.l1[3], b.l2, N, H1, W1, st))); break;
        default: IMVS_TRY((res_stage<16, 32, true, false>(w, 6, b.l1[3], b.l2, N, H1, W1, st)));
    }
    const int t3 = tune("FNET3", 2);
    const EpiNHWC eo3{fea3, w->b[16], nullptr, H3, W3, 48, 48, 0};
    switch (t3) {                                                                         // layer3 -> l3[3]  [H/8][W/8][48]; output3 (net.py:59)
        case 1:
            IMVS_TRY((res_stage<32, 48, false, false, 32, 16, 2>(w, 11, b.l2[3], b.l3, N, H2, W2, st)));
            IMVS_TRY((mma_conv<48, 16, 2, 4, 1, false>("fnet.output3", in_nhwc(b.l3[3], H3, W3, 48), eo3, WSets::single(w->w[16]), s1, N, 48, H3, W3, 3, st)));
            break;
        case 2:
            IMVS_TRY((res_stage<32, 48, false, false, 96, 48, 1>(w, 11, b.l2[3], b.l3, N, H2, W2, st)));
            IMVS_TRY((conv3x3_s1<48, 48>("fnet.output3", w->w[16], w->b[16], b.l3[3], fea3, nullptr, 0, N, H3, W3, st, [&] {
                return mma_conv<48, 48, 1, 4, 1, false>("fnet.output3", in_nhwc(b.l3[3], H3, W3, 48), eo3, WSets::single(w->w[16]), conv_tables(3, 1, 1, 4), N, 48, H3, W3, 1, st); })));
            break;
        case 3:
            IMVS_TRY((res_stage<32, 48, false, false, 32, 16, 1>(w, 11, b.l2[3], b.l3, N, H2, W2, st)));
            IMVS_TRY((mma_conv<48, 16, 1, 4, 1, false>("fnet.output3", in_nhwc(b.l3[3], H3, W3, 48), eo3, WSets::single(w->w[16]), conv_tables(3, 1, 1, 4), N, 48, H3, W3, 3, st)));
            break;
        default:
            IMVS_TRY((res_stage<32, 48, false, false>(w, 11, b.l2[3], b.l3, N, H2, W2, st)));
            IMVS_TRY((mma_conv<48, 48, 2, 4, 1, false>("fnet.output3", in_nhwc(b.l3[3], H3, W3, 48), eo3, WSets::single(w->w[16]), s1, N, 48, H3, W3, 1, st)));
    }
    // intra2 = up2(f3) + inner2(f2); output2 (net.py:60-62); intra1 = up2(intra2) + inner1(f1); output1 (net.py:63-64)
    const EpiAddUp2 ei2{b.intra2, w->b[17], b.l3[3], H2, W2, 48}, ei1{b.intra1, w->b[19], b.intra2, H1, W1, 48};
    const EpiNHWC eo2{fea2, w->b[18], nullptr, H2, W2, 32, 32, 0}, eo1{fea1, w->b[20], nullptr, H1, W1, 16, 16, 0};
    if (w8) {
        IMVS_TRY((mma_conv<32, 48, 1, 8, 1, true>("fnet.inner2", in_nhwc(b.l2[3], H2, W2, 32), ei2, WSets::single(w->w[17]), k1, N, 48, H2, W2, 1, st)));
        IMVS_TRY((mma_conv<48, 32, 1, 8, 1, false>("fnet.output2", in_nhwc(b.intra2, H2, W2, 48), eo2, WSets::single(w->w[18]), s1, N, 32, H2, W2, 1, st)));
        IMVS_TRY((mma_conv<16, 48, 1, 8, 1, true>("fnet.inner1", in_nhwc(b.l1[3], H1, W1, 16), ei1, WSets::single(w->w[19]), k1, N, 48, H1, W1, 1, st)));
        IMVS_TRY((mma_conv<48, 16, 1, 8, 1, false>("fnet.output1", in_nhwc(b.intra1, H1, W1, 48), eo1, WSets::single(w->w[20]), s1, N, 16, H1, W1, 1, st)));
    } else if (tune("FNETO", 0)) {      // 4-row tiles: smaller haloed tile in shared memory, more CTAs per SM
        const TapTables s1m = conv_tables(3, 1, 1, 4), k1m = conv_tables(1, 1, 1, 4);
        IMVS_TRY((mma_conv<32, 48, 1, 4, 1, true>("fnet.inner2", in_nhwc(b.l2[3], H2, W2, 32), ei2, WSets::single(w->w[17]), k1m, N, 48, H2, W2, 1, st)));
        IMVS_TRY((mma_conv<48, 32, 1, 4, 1, false>("fnet.output2", in_nhwc(b.intra2, H2, W2, 48), eo2, WSets::single(w->w[18]), s1m, N, 32, H2, W2, 1, st)));
        IMVS_TRY((mma_conv<16, 48, 1, 4, 1, true>("fnet.inner1", in_nhwc(b.l1[3], H1, W1, 16), ei1, WSets::single(w->w[19]), k1m, N, 48, H1, W1, 1, st)));
        IMVS_TRY((mma_conv<48, 16, 1, 4, 1, false>("fnet.output1", in_nhwc(b.intra1, H1, W1, 48), eo1, WSets::single(w->w[20]), s1m, N, 16, H1, W1, 1, st)));
    } else {
        IMVS_TRY((mma_conv<32, 48, 2, 4, 1, true>("fnet.inner2", in_nhwc(b.l2[3], H2, W2, 32), ei2, WSets::single(w->w[17]), k1, N, 48, H2, W2, 1, st)));
        IMVS_TRY((conv3x3_s1<48, 32>("fnet.output2", w->w[18], w->b[18], b.intra2, fea2, nullptr, 0, N, H2, W2, st, [&] {
            return mma_conv<48, 32, 2, 4, 1, false>("fnet.output2", in_nhwc(b.intra2, H2, W2, 48), eo2, WSets::single(w->w[18]), s1, N, 32, H2, W2, 1, st); })));
        IMVS_TRY((mma_conv<16, 48, 2, 4, 1, true>("fnet.inner1", in_nhwc(b.l1[3], H1, W1, 16), ei1, WSets::single(w->w[19]), k1, N, 48, H1, W1, 1, st)));
        IMVS_TRY((conv3x3_s1<48, 16>("fnet.output1", w->w[20], w->b[20], b.intra1, fea1, nullptr, 0, N, H1, W1, st, [&] {
            return mma_conv<48, 16, 2, 4, 1, false>("fnet.output1", in_nhwc(b.intra1, H1, W1, 48), eo1, WSets::single(w->w[20]), s1, N, 16, H1, W1, 1, st); })));
    }
    return 0;
}

extern "C" int imvs_featurenet_forward(const imvs_featurenet_weights* w, const float* imgs, float* fea1, float* fea2, float* fea3,
                                       void* workspace, size_t workspace_bytes, int N, int H, int W, void* stream) {
    return featurenet_forward_impl(w, imgs, nullptr, fea1, fea2, fea3, workspace, workspace_bytes, N, H, W, stream);
}

extern "C" int imvs_featurenet_forward_u8(const imvs_featurenet_weights* w, const unsigned char* imgs, float* fea1, float* fea2,
                                          float* fea3, void* workspace, size_t workspace_bytes, int N, int H, int W, void* stream) {
    return featurenet_forward_impl(w, nullptr, imgs, fea1, fea2, fea3, workspace, workspace_bytes, N, H, W, stream);
}

// Operator-level entry point of the persistent TMA + tcgen05 convolution (csrc/tc5pconv.cuh): stride-1 3x3 (dilation dil)
// Cin -> Cout on fp32 NHWC tensors, fp32-grade (fp16 hi / lo split, three products, fp32 accumulation).  The fp32 operands
// are converted to split planes in the workspace first; inside FeatureNet the producers write that layout directly.
extern "C" int imvs_set_sm_share(int share) {
#ifndef CUSIM
    tc5p::set_sm_share(share);
#endif
    return 0;
}

extern "C" size_t imvs_conv3x3_tcgen05_workspace_bytes(int N, int H, int W, int Cin, int Cout) {
    if (N < 1 || H < 1 || W < 1 || Cin < 8 || Cout < 8) return 0;
    const size_t px = (size_t)N * H * W;
    return (px * Cin + 2 * px * Cout) * sizeof(float) + 3 * 256;
}

extern "C" int imvs_conv3x3_tcgen05(const float* x, const void* w_f16ummai, const float* bias, const float* residual, float* out,
                                    void* workspace, size_t workspace_bytes, int N, int H, int W, int Cin, int Cout, int dil, int relu,
                                    int via_split_output, void* stream) {
#ifdef CUSIM
    return fail("conv3x3_tcgen05: not available in the CPU emulation build");
#else
    IMVS_REQUIRE(x && w_f16ummai && out && workspace, "conv3x3_tcgen05: null pointer");
    IMVS_REQUIRE(N >= 1 && H >= 1 && W >= 1 && dil >= 1 && dil <= 3, "conv3x3_tcgen05: bad shape");
    IMVS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "conv3x3_tcgen05: workspace must be 256-byte aligned");
    IMVS_REQUIRE(workspace_bytes >= imvs_conv3x3_tcgen05_workspace_bytes(N, H, W, Cin, Cout), "conv3x3_tcgen05: workspace too small");
    ApiScope api_;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t px = (size_t)N * H * W;
    auto carve = [&](size_t& off, size_t bytes) { void* p = static_cast<char*>(workspace) + off; off += (bytes + 255) / 256 * 256; return p; };
    size_t off = 0;
    const tc5p::Split xin = tc5p::split_at(carve(off, px * Cin * 4), px * Cin);
    const tc5p::Split res = residual ? tc5p::split_at(carve(off, px * Cout * 4), px * Cout) : tc5p::Split{nullptr, nullptr};
    const tc5p::Split outs = via_split_output ? tc5p::split_at(carve(off, px * Cout * 4), px * Cout) : tc5p::Split{nullptr, nullptr};
    auto to_split = [&](const float* src, const tc5p::Split& dst, int C) -> int {
        const size_t n = px * (C / 8);
        IMVS_CUDA(launch_k(tc5p::nhwc_to_split_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, src, dst.hi, dst.lo, px, H * W, C));
        return 0;
    };
    IMVS_TRY(to_split(x, xin, Cin));
    if (residual) IMVS_TRY(to_split(residual, res, Cout));
    const tc5p::Epi epi{outs, via_split_output ? nullptr : out, res, bias, H, W, relu};
    int* flag = tc5_error_flag();
    const int key = (Cin * 100 + Cout) * 10 + dil;
    switch (key) {
        case 16161: IMVS_TRY((tc5p::launch<16, 16>("conv3x3_tcgen05", xin, epi, w_f16ummai, N, H, W, flag, st))); break;
        case 32321: IMVS_TRY((tc5p::launch<32, 32>("conv3x3_tcgen05", xin, epi, w_f16ummai, N, H, W, flag, st))); break;
        case 32322: IMVS_TRY((tc5p::launch<32, 32, 2>("conv3x3_tcgen05", xin, epi, w_f16ummai, N, H, W, flag, st))); break;
        case 48481: IMVS_TRY((tc5p::launch<48, 48>("conv3x3_tcgen05", xin, epi, w_f16ummai, N, H, W, flag, st))); break;
        case 48321: IMVS_TRY((tc5p::launch<48, 32>("conv3x3_tcgen05", xin, epi, w_f16ummai, N, H, W, flag, st))); break;
        case 48161: IMVS_TRY((tc5p::launch<48, 16>("conv3x3_tcgen05", xin, epi, w_f16ummai, N, H, W, flag, st))); break;
        default: return fail("conv3x3_tcgen05: (Cin, Cout, dilation) = (%d, %d, %d) is not instantiated", Cin, Cout, dil);
    }
    if (via_split_output) {
        const size_t n = px * (Cout / 8);
        IMVS_CUDA(launch_k(tc5p::split_to_nhwc_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, (const __half*)outs.hi,
                           (const __half*)outs.lo, out, px, H * W, Cout));
    }
    return 0;
#endif
}
