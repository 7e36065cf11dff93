// Update stage: ConvGRU (module.py:52-66) and depth / confidence heads with softmax, arg-max and
// clamped-window regression (itermvs.py:139-151, 171-190, 192-220).  Activations channels-last.
#include <algorithm>

#include "common.cuh"
#include "mmaconv.cuh"

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

// ------------------------------------------------------------------------------------ heads ----
// Per pixel:  t[32] (relu'd 3x3 output) -> fc1 32->64 relu -> fc2 64->256 + b -> softmax -> arg-max
// -> clamped +-4 window regression; optionally confidence = sigmoid(tc[32] . wc + bc).
// One warp owns 8 consecutive pixels per step; lane j owns logits {4j..4j+3, 128+4j..128+4j+3} of
// each of its 8 pixels (64 accumulators), so fc2 weights stream from shared memory as conflict-free
// float4 rows and activations as broadcasts; softmax / arg-max / window sums reduce with shuffles.
// The 256-bin logits never leave the SM (the reference writes and re-reads that 21 MB tensor ~6x).
constexpr int HEAD_THREADS = 256;
constexpr int HEAD_PXW = 8;            // pixels per warp step

struct HeadParams {
    const float* t;          // [B][P][64] channels-last: 0..31 depth head, 32..63 confidence head
    const float* fc1;        // [32][64]
    const float* fc2;        // [64][256]
    const float* fc2_b;      // [256]
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

__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(const HeadParams prm) {
    extern __shared__ __align__(16) float smem[];
    float* sW2 = smem;                      // [64][256]
    float* sW1 = sW2 + 64 * 256;            // [32][64]
    float* sB2 = sW1 + 32 * 64;             // [256]
    float* sWc = sB2 + 256;                 // [32] + 1
    float* sWarp = sWc + 36;                // per warp: st [64][8] + sh1 [64][8]
    for (int i = threadIdx.x; i < 64 * 256 / 4; i += HEAD_THREADS) reinterpret_cast<float4*>(sW2)[i] = ldg4(prm.fc2 + 4 * i);
    for (int i = threadIdx.x; i < 32 * 64 / 4; i += HEAD_THREADS) reinterpret_cast<float4*>(sW1)[i] = ldg4(prm.fc1 + 4 * i);
    for (int i = threadIdx.x; i < 256; i += HEAD_THREADS) sB2[i] = ldg(prm.fc2_b + i);
    const bool want_conf = (prm.conf != nullptr) || (prm.conf_logit != nullptr);
    if (want_conf && threadIdx.x < 33) sWc[threadIdx.x] = threadIdx.x < 32 ? ldg(prm.conf_w + threadIdx.x) : ldg(prm.conf_b);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* st = sWarp + warp * (2 * 64 * HEAD_PXW);
    float* sh1 = st + 64 * HEAD_PXW;
    const int P = prm.P;
    const int items = (prm.B * P) / HEAD_PXW;
    const int nwarps = gridDim.x * (HEAD_THREADS / 32);
    const int rows = want_conf ? 64 : 32;

    for (int item = blockIdx.x * (HEAD_THREADS / 32) + warp; item < items; item += nwarps) {
        const int gp = item * HEAD_PXW;          // flat pixel over [B][P]
        const int b = gp / P, p0 = gp % P;
        // ---- stage A: activations of 8 pixels (8 x 64 contiguous floats) -> smem, transposed to [k][px]
        {
            const float* tb = prm.t + ((size_t)b * P + p0) * 64;
            for (int i = lane; i < HEAD_PXW * rows / 4; i += 32) {
                const int px = i / (rows / 4), k4 = i % (rows / 4);
                const float4 v = ldg4(tb + (size_t)px * 64 + 4 * k4);
                st[(4 * k4 + 0) * HEAD_PXW + px] = v.x;
                st[(4 * k4 + 1) * HEAD_PXW + px] = v.y;
                st[(4 * k4 + 2) * HEAD_PXW + px] = v.z;
                st[(4 * k4 + 3) * HEAD_PXW + px] = v.w;
            }
        }
        __syncwarp();
        // ---- stage B: fc1 + relu: lane owns hidden channels 2*lane, 2*lane+1
        {
            float h0[HEAD_PXW], h1[HEAD_PXW];
#pragma unroll
            for (int i = 0; i < HEAD_PXW; ++i) { h0[i] = 0.f; h1[i] = 0.f; }
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const float4 ta = reinterpret_cast<const float4*>(st + k * HEAD_PXW)[0];
                const float4 tb4 = reinterpret_cast<const float4*>(st + k * HEAD_PXW)[1];
                const float2 w = reinterpret_cast<const float2*>(sW1 + k * 64)[lane];
                const float tv[8] = {ta.x, ta.y, ta.z, ta.w, tb4.x, tb4.y, tb4.z, tb4.w};
#pragma unroll
                for (int i = 0; i < HEAD_PXW; ++i) { h0[i] = fmaf(tv[i], w.x, h0[i]); h1[i] = fmaf(tv[i], w.y, h1[i]); }
            }
            float4* d0 = reinterpret_cast<float4*>(sh1 + (2 * lane) * HEAD_PXW);
            float4* d1 = reinterpret_cast<float4*>(sh1 + (2 * lane + 1) * HEAD_PXW);
            d0[0] = make_float4(fmaxf(h0[0], 0.f), fmaxf(h0[1], 0.f), fmaxf(h0[2], 0.f), fmaxf(h0[3], 0.f));
            d0[1] = make_float4(fmaxf(h0[4], 0.f), fmaxf(h0[5], 0.f), fmaxf(h0[6], 0.f), fmaxf(h0[7], 0.f));
            d1[0] = make_float4(fmaxf(h1[0], 0.f), fmaxf(h1[1], 0.f), fmaxf(h1[2], 0.f), fmaxf(h1[3], 0.f));
            d1[1] = make_float4(fmaxf(h1[4], 0.f), fmaxf(h1[5], 0.f), fmaxf(h1[6], 0.f), fmaxf(h1[7], 0.f));
        }
        __syncwarp();
        // ---- stage C: fc2 logits; lane owns channels ch(a) = a<4 ? 4*lane+a : 128+4*lane+(a-4)
        float acc[HEAD_PXW][8];
        {
            const float4 ba = reinterpret_cast<const float4*>(sB2)[lane];
            const float4 bb = reinterpret_cast<const float4*>(sB2 + 128)[lane];
#pragma unroll
            for (int i = 0; i < HEAD_PXW; ++i) {
                acc[i][0] = ba.x; acc[i][1] = ba.y; acc[i][2] = ba.z; acc[i][3] = ba.w;
                acc[i][4] = bb.x; acc[i][5] = bb.y; acc[i][6] = bb.z; acc[i][7] = bb.w;
            }
#pragma unroll 2
            for (int k = 0; k < 64; ++k) {
                const float4 ha = reinterpret_cast<const float4*>(sh1 + k * HEAD_PXW)[0];
                const float4 hb = reinterpret_cast<const float4*>(sh1 + k * HEAD_PXW)[1];
                const float4 wa = reinterpret_cast<const float4*>(sW2 + k * 256)[lane];
                const float4 wb = reinterpret_cast<const float4*>(sW2 + k * 256 + 128)[lane];
                const float hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int i = 0; i < HEAD_PXW; ++i)
#pragma unroll
                    for (int a = 0; a < 8; ++a) acc[i][a] = fmaf(hv[i], wv[a], acc[i][a]);
            }
        }
        // ---- stage D: softmax, arg-max (first maximum), clamped window regression
        float nd_mine = 0.f;
#pragma unroll
        for (int i = 0; i < HEAD_PXW; ++i) {
            float m = acc[i][0];
#pragma unroll
            for (int a = 1; a < 8; ++a) m = fmaxf(m, acc[i][a]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float e[8], s = 0.f;
#pragma unroll
            for (int a = 0; a < 8; ++a) { e[a] = expf(acc[i][a] - m); s += e[a]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            float pr[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) pr[a] = e[a] / s;
            // arg-max over probabilities, first index on ties (torch.argmax)
            float bv = -1.f;
            int bi = 0;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int ch = a < 4 ? 4 * lane + a : 128 + 4 * lane + (a - 4);
                if (pr[a] > bv) { bv = pr[a]; bi = ch; }      // channels visited in increasing order
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
                num = fmaf((float)(mult * ch), pr[a], num);
                den = fmaf((float)mult, pr[a], den);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                num += __shfl_xor_sync(0xffffffffu, num, o);
                den += __shfl_xor_sync(0xffffffffu, den, o);
            }
            const float ndv = (num / (1e-6f + den)) / (float)(IMVS_OUT_BINS - 1);
            if (lane == i) nd_mine = ndv;
            if (prm.prob) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int ch = a < 4 ? 4 * lane + a : 128 + 4 * lane + (a - 4);
                    prm.prob[((size_t)b * IMVS_OUT_BINS + ch) * P + p0 + i] = pr[a];
                }
            }
        }
        if (lane < HEAD_PXW) {
            prm.nd_out[(size_t)b * prm.nd_bstride + (size_t)(p0 + lane) * prm.nd_pstride] = nd_mine;
            if (prm.depth_out) {
                const float inv_min = 1.0f / prm.depth_min[b], inv_max = 1.0f / prm.depth_max[b];
                prm.depth_out[(size_t)b * P + p0 + lane] = unnormalize_depth(nd_mine, inv_min, inv_max);
            }
        }
        // ---- confidence head: sigmoid(tc . wc + bc), tc = rows 32..63 of t
        if (want_conf) {
            const int px = lane & 7, kq = lane >> 3;
            float s = 0.f;
#pragma unroll
            for (int k = kq; k < 32; k += 4) s = fmaf(st[(32 + k) * HEAD_PXW + px], sWc[k], s);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            s += sWc[32];
            if (lane < HEAD_PXW) {
                if (prm.conf_logit) prm.conf_logit[(size_t)b * P + p0 + lane] = s;
                if (prm.conf) prm.conf[(size_t)b * P + p0 + lane] = sigmoidf_(s);
            }
        }
        __syncwarp();
    }
}

}  // namespace imvs

using namespace imvs;

extern "C" int imvs_conv_gru(const imvs_weights* w, float* h, const float* x, float* scratch, int B, int H, int W, void* stream) {
    IMVS_REQUIRE(w && h && x && scratch, "conv_gru: null pointer");
    IMVS_REQUIRE(B >= 1 && H >= 1 && W >= 1, "conv_gru: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * 32 * H * W;
    float* z = scratch;
    float* rh = scratch + n;
    const TapTables taps = conv_tables(3, 1, 2, 8);
    IMVS_TRY((mma_conv<48, 32, 2, 4, 1, false>("gru.zr", InNHWC2{h, x, H, W, 32, IMVS_XCH}, EpiGruZR{w->gru_zr_b, h, z, rh, H, W},
                                               WSets::single(w->gru_zr), taps, B, 64, H, W, 2, st)));
    IMVS_TRY((mma_conv<48, 32, 2, 4, 1, false>("gru.q", InNHWC2{rh, x, H, W, 32, IMVS_XCH}, EpiGruQ{w->gru_q_b, z, h, H, W},
                                               WSets::single(w->gru_q), taps, B, 32, H, W, 1, st)));
    return 0;
}

extern "C" int imvs_depth_head(const imvs_weights* w, const float* hidden, float* nd_out, size_t nd_batch_stride,
                               size_t nd_pixel_stride, float* probability, float* conf, float* conf_logit, float* depth_out,
                               const float* depth_min, const float* depth_max, float* scratch,
                               int B, int H, int W, void* stream) {
    IMVS_REQUIRE(w && hidden && nd_out && scratch, "depth_head: null pointer");
    IMVS_REQUIRE(B >= 1 && H >= 1 && W >= 1 && (H * W) % HEAD_PXW == 0, "depth_head: H*W must be a multiple of %d", HEAD_PXW);
    IMVS_REQUIRE(!depth_out || (depth_min && depth_max), "depth_head: depth_out needs depth_min/depth_max");
    IMVS_REQUIRE(nd_pixel_stride >= 1, "depth_head: nd_pixel_stride must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    const bool want_conf = conf || conf_logit;
    const int P = H * W;
    float* t = scratch;                     // [B][P][64]
    // stacked weight [9][32][64]: channel block 0 = depth_head.0, block 1 = confidence_head.0; the
    // confidence block only runs when a confidence output is requested (itermvs.py:196-199)
    IMVS_TRY((mma_conv<32, 32, 2, 4, 1, false>("head.conv0", in_nhwc(hidden, H, W, 32), EpiNHWC{t, nullptr, nullptr, H, W, 64, 64, 1},
                                               WSets::single(w->head_conv0), conv_tables(3, 1, 2, 8), B, 64, H, W,
                                               want_conf ? 2 : 1, st)));
    HeadParams prm;
    prm.t = t;
    prm.fc1 = w->head_fc1; prm.fc2 = w->head_fc2; prm.fc2_b = w->head_fc2_b;
    prm.conf_w = w->conf_fc; prm.conf_b = w->conf_fc_b;
    prm.nd_out = nd_out; prm.nd_bstride = nd_batch_stride; prm.nd_pstride = nd_pixel_stride;
    prm.prob = probability; prm.conf = conf; prm.conf_logit = conf_logit; prm.depth_out = depth_out;
    prm.depth_min = depth_min; prm.depth_max = depth_max;
    prm.B = B; prm.P = P;
    const size_t smem = (size_t)(64 * 256 + 32 * 64 + 256 + 36 + (HEAD_THREADS / 32) * 2 * 64 * HEAD_PXW) * sizeof(float);
    static int smem_ok = 0;
    IMVS_TRY(ensure_dynamic_smem(head_kernel, smem, &smem_ok));
    int dev = 0, sms = 148;
    IMVS_CUDA(cudaGetDevice(&dev));
    IMVS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int items = (B * P) / HEAD_PXW;
    int blocks = std::min(cdiv(items, HEAD_THREADS / 32), 2 * sms);
    head_kernel<<<blocks, HEAD_THREADS, smem, st>>>(prm);
    count_launch();
    IMVS_LAUNCH_CHECK("head_kernel");
    return 0;
}
