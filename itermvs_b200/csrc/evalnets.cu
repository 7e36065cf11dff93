// The small convolutional nets of the evaluation stage and hidden-state initialisation:
//   CorrNet (itermvs.py:352-381), PixelViewWeight (itermvs.py:333-350, 53-57), hidden_init
//   (itermvs.py:153-164).  All built on the register-blocked direct convolution of conv.cuh.
#include "common.cuh"
#include "conv.cuh"

namespace imvs {

// ------------------------------------------------------------------------------- CorrNet ----
struct EpiCorrOut {          // conv5: + bias, scatter slice n -> out[(n/period)*batch_stride + (n%period)*HW + p]
    float* out;
    const float* bias[3];
    int period, split1, split2;
    size_t batch_stride;
    int H, W;
    template <int CO>
    __device__ __forceinline__ void store(int n, int y, int x, int, const float (&a)[CO]) const {
        int r = n % period;
        const float* b = r < split1 ? bias[0] : (r < split2 ? bias[1] : bias[2]);
        out[(size_t)(n / period) * batch_stride + (size_t)r * H * W + (size_t)y * W + x] = a[0] + ldg(b);
    }
};

using CfgCorr0 = ConvCfg<8, 8, 8, 4, 4, 3, 1, 1, 8>;      // cl8 -> 8, relu
using CfgCorr1 = ConvCfg<16, 16, 8, 4, 2, 3, 2, 1, 1>;    // 8 -> 16, stride 2, relu
using CfgCorr2 = ConvCfg<32, 32, 8, 4, 2, 3, 2, 1, 1>;    // 16 -> 32, stride 2, relu
using CfgCorr5 = ConvCfg<1, 1, 1, 4, 4, 3, 1, 1, 1>;      // 8 -> 1

static WeightSel sel_of(const imvs_corrnet_weights* sets, int period, int split1, int split2, int which) {
    WeightSel s;
    for (int i = 0; i < 3; ++i) {
        const imvs_corrnet_weights& c = sets[i];
        const float* p = which == 0 ? c.conv0 : which == 1 ? c.conv1 : which == 2 ? c.conv2
                       : which == 3 ? c.conv3 : which == 4 ? c.conv4 : c.conv5;
        s.w[i] = p;
    }
    s.period = period; s.split1 = split1; s.split2 = split2;
    return s;
}

// ------------------------------------------------------------------------ PixelViewWeight ----
struct EpiPvw {              // relu(16) . w1 + b1 -> logits[n][y][x]
    float* logits;
    const float* w1;
    const float* b1;
    int H, W;
    template <int CO>
    __device__ __forceinline__ void store(int n, int y, int x, int, const float (&a)[CO]) const {
        static_assert(CO == 16, "PixelViewWeight epilogue needs all 16 channels in one thread");
        float s = ldg(b1);
#pragma unroll
        for (int c = 0; c < CO; ++c) s = fmaf(fmaxf(a[c], 0.f), ldg(w1 + c), s);
        logits[((size_t)n * H + y) * W + x] = s;
    }
};
using CfgPvw = ConvCfg<16, 16, 16, 4, 4, 3, 1, 1, 8>;

// softmax over D then max over D == 1 / sum_d exp(l_d - max_d l)   (itermvs.py:347-348)
__global__ void pvw_reduce_kernel(const float* __restrict__ logits, float* __restrict__ vw3, int BS, int D, int P3) {
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

// F.interpolate(scale_factor=2, mode='bilinear')  (itermvs.py:56-57), maps [N][H][W] -> [N][2H][2W]
__global__ void upsample2x_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, bool apply_tanh) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int Ho = 2 * H, Wo = 2 * W;
    if (t >= (size_t)N * Ho * Wo) return;
    int ox = (int)(t % Wo), oy = (int)((t / Wo) % Ho);
    size_t n = t / ((size_t)Wo * Ho);
    int h0, h1, w0, w1;
    float lh, lw;
    up_index(oy, 0.5f, H, h0, h1, lh);
    up_index(ox, 0.5f, W, w0, w1, lw);
    const float* q = in + n * H * W;
    float v = (1.f - lh) * ((1.f - lw) * ldg(q + h0 * W + w0) + lw * ldg(q + h0 * W + w1)) +
              lh * ((1.f - lw) * ldg(q + h1 * W + w0) + lw * ldg(q + h1 * W + w1));
    out[t] = apply_tanh ? tanhf(v) : v;
}

// ----------------------------------------------------------------------------- hidden_init ----
using CfgHinit0 = ConvCfg<64, 32, 8, 4, 2, 3, 1, 1, 1>;   // D -> 64, relu
using CfgHinit1 = ConvCfg<32, 32, 8, 4, 2, 1, 1, 1, 1>;   // 1x1 64 -> 32, + bias

}  // namespace imvs

using namespace imvs;

extern "C" size_t imvs_corrnet_scratch_floats(int N, int H, int W) { return (size_t)N * 26 * H * W; }

extern "C" int imvs_corrnet(const imvs_corrnet_weights* sets, int period, int split1, int split2, const float* vol,
                            float* out, size_t out_batch_stride, float* scratch, int N, int H, int W, void* stream) {
    IMVS_REQUIRE(sets && vol && out && scratch, "corrnet: null pointer");
    IMVS_REQUIRE(N >= 1 && H >= 4 && W >= 4 && H % 4 == 0 && W % 4 == 0, "corrnet: H, W must be multiples of 4 (H=%d W=%d)", H, W);
    IMVS_REQUIRE(period >= 1 && N % period == 0, "corrnet: N=%d not a multiple of period=%d", N, period);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t HW = (size_t)H * W;
    float* c0 = scratch;                    // [N][8][H][W]
    float* c1 = c0 + (size_t)N * 8 * HW;    // [N][16][H/2][W/2]
    float* c2 = c1 + (size_t)N * 4 * HW;    // [N][32][H/4][W/4]
    float* x3 = c2 + (size_t)N * 2 * HW;    // [N][16][H/2][W/2]
    float* x4 = x3 + (size_t)N * 4 * HW;    // [N][8][H][W]
    const int H1 = H / 2, W1 = W / 2, H2 = H / 4, W2 = W / 4;
    IMVS_TRY((launch_conv<CfgCorr0>("corrnet.conv0", InChannelsLast8{vol, H, W}, EpiPlanar{c0, nullptr, 8, H, W, true},
                                    sel_of(sets, period, split1, split2, 0), N, 8, H, W, st)));
    IMVS_TRY((launch_conv<CfgCorr1>("corrnet.conv1", InPlanar{c0, 8, H, W}, EpiPlanar{c1, nullptr, 16, H1, W1, true},
                                    sel_of(sets, period, split1, split2, 1), N, 8, H1, W1, st)));
    IMVS_TRY((launch_conv<CfgCorr2>("corrnet.conv2", InPlanar{c1, 16, H1, W1}, EpiPlanar{c2, nullptr, 32, H2, W2, true},
                                    sel_of(sets, period, split1, split2, 2), N, 16, H2, W2, st)));
    IMVS_TRY((launch_tconv<16, 16, 8, 4>("corrnet.conv3", c2, c1, x3, sel_of(sets, period, split1, split2, 3), N, 32, H2, W2, st)));
    IMVS_TRY((launch_tconv<8, 8, 8, 4>("corrnet.conv4", x3, c0, x4, sel_of(sets, period, split1, split2, 4), N, 16, H1, W1, st)));
    EpiCorrOut e5;
    e5.out = out;
    for (int i = 0; i < 3; ++i) e5.bias[i] = sets[i].conv5_b;
    e5.period = period; e5.split1 = split1; e5.split2 = split2;
    e5.batch_stride = out_batch_stride; e5.H = H; e5.W = W;
    IMVS_TRY((launch_conv<CfgCorr5>("corrnet.conv5", InPlanar{x4, 8, H, W}, e5, sel_of(sets, period, split1, split2, 5), N, 8, H, W, st)));
    return 0;
}

extern "C" int imvs_pixel_view_weight(const imvs_weights* w, const float* corr, float* logits, float* vw3, float* vw2,
                                      int B, int S, int D, int H3, int W3, void* stream) {
    IMVS_REQUIRE(w && corr && logits && vw3 && vw2, "pixel_view_weight: null pointer");
    IMVS_REQUIRE(B >= 1 && S >= 1 && D >= 1 && H3 >= 1 && W3 >= 1, "pixel_view_weight: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = B * S * D, P3 = H3 * W3;
    IMVS_TRY((launch_conv<CfgPvw>("pvw.conv", InChannelsLast8{corr, H3, W3}, EpiPvw{logits, w->pvw_conv1, w->pvw_conv1_b, H3, W3},
                                  WeightSel::single(w->pvw_conv0), N, 8, H3, W3, st)));
    pvw_reduce_kernel<<<cdiv(B * S * P3, 128), 128, 0, st>>>(logits, vw3, B * S, D, P3);
    count_launch();
    IMVS_LAUNCH_CHECK("pvw_reduce_kernel");
    size_t total = (size_t)B * S * P3 * 4;
    upsample2x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(vw3, vw2, B * S, H3, W3, false);
    count_launch();
    IMVS_LAUNCH_CHECK("upsample2x_kernel");
    return 0;
}

extern "C" int imvs_hidden_init(const imvs_weights* w, const float* corr, float* hidden, float* scratch,
                                int B, int D, int H3, int W3, void* stream) {
    IMVS_REQUIRE(w && corr && hidden && scratch, "hidden_init: null pointer");
    IMVS_REQUIRE(B >= 1 && D >= 1 && H3 >= 1 && W3 >= 1, "hidden_init: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t P3 = (size_t)H3 * W3;
    float* t = scratch;                       // [B][64][P3]
    float* u = scratch + (size_t)B * 64 * P3; // [B][32][P3]
    IMVS_TRY((launch_conv<CfgHinit0>("hidden_init.conv0", InPlanar{corr, D, H3, W3}, EpiPlanar{t, nullptr, 64, H3, W3, true},
                                     WeightSel::single(w->hinit_conv0), B, D, H3, W3, st)));
    IMVS_TRY((launch_conv<CfgHinit1>("hidden_init.fc", InPlanar{t, 64, H3, W3}, EpiPlanar{u, w->hinit_fc_b, 32, H3, W3, false},
                                     WeightSel::single(w->hinit_fc), B, 64, H3, W3, st)));
    size_t total = (size_t)B * 32 * P3 * 4;
    upsample2x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(u, hidden, B * 32, H3, W3, true);
    count_launch();
    IMVS_LAUNCH_CHECK("upsample2x_kernel(tanh)");
    return 0;
}
