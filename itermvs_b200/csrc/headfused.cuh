// Fused depth head on the 5th-generation tensor core (tcgen05 + TMEM), fp32-grade:
//
//     t (conv0 output, 64 ch) -> fc1 1x1 32->64 + ReLU -> fc2 1x1 64->256 + bias -> softmax over the 256 bins ->
//     arg-max -> clamped +-4 window regression        (+ the confidence head's 1x1 + sigmoid)
//     reference: models/itermvs.py:139-151, 171-190, 196-219
//
// One CTA = 128 consecutive pixels = the 128 TMEM lanes of one UMMA M-block.  Both 1x1 convolutions are GEMMs whose
// fp32-grade product x*w = hi*hi + hi*lo + lo*hi (x = hi + lo in fp16, 22 significant bits, fp32 accumulation in TMEM;
// the same 3-product split the mma.sync engine uses in mode 4) is three UMMA chains over the same accumulator:
//     fc1: D1[128 x 64]  (TMEM columns 0.. 63)  = A1[128 x 32] * W1^T        6 x tcgen05.mma.kind::f16 (K = 16)
//     fc2: D2[128 x 256] (TMEM columns 0..255)  = A2[128 x 64] * W2^T       12 x tcgen05.mma.kind::f16
// A2 = split(ReLU(D1)) is produced by the CTA itself (tcgen05.ld -> registers -> shared memory, over A1's buffer);
// D2 then overwrites D1's columns.  The 256 logits of a pixel sit in ONE TMEM lane: thread = (pixel, column half) reads
// them twice (max / arg-max, then exp-sum and the window sums; the bias is added from shared memory on the way) -- the
// 21 MB logits tensor of the unfused path (written by fc2, re-read by the regression kernel) and the 5 MB fc1 activation
// never exist.  109 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM, so the 160 CTAs of a 640x512
// reference view are one wave on 148 SMs.  Operands live in shared memory in the UMMA K-major / no-swizzle canonical layout
// [K/8][rows][8 halves] (core matrix = 8 rows x 16 bytes, SBO = 128 B, LBO = rows * 16 B); the weights arrive pre-split
// and pre-ordered from the host (itermvs_b200/_pack.py:pack_head_fused) with four 1-D bulk copies (TMA engine).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc5conv.cuh"

namespace imvs {
namespace hf {

constexpr int HF_THREADS = 256;
constexpr int HF_M = 128;
constexpr int KC1 = 4, N1 = 64;                  // fc1: K = 32 = 4 chunks of 8 halves
constexpr int KC2 = 8, N2 = 256;                 // fc2: K = 64 = 8 chunks
constexpr uint32_t W1_BYTES = KC1 * N1 * 16, W2_BYTES = KC2 * N2 * 16;
constexpr uint32_t BIAS_BYTES = N2 * 4;
constexpr uint32_t BLOB_BYTES = 2 * W1_BYTES + 2 * W2_BYTES + BIAS_BYTES;   // [W1 hi | W1 lo | W2 hi | W2 lo | bias fp32] = 74 752
constexpr uint32_t A1_BYTES = KC1 * HF_M * 16, A2_BYTES = KC2 * HF_M * 16;
constexpr uint32_t X_BYTES = 2 * HF_M * 4 * sizeof(float);
constexpr uint32_t SMEM_BYTES = BLOB_BYTES + 2 * A2_BYTES + X_BYTES + 64;      // A1 (hi | lo) aliases the head of A2: 111 680
constexpr uint32_t TMEM_COLS = 256;

struct Params {
    const float* t;          // [n_px][64]: ReLU(conv0); channels 0..31 depth head, 32..63 confidence head
    const void* blob;        // packed fp16 weights, BLOB_BYTES
    const float* conf_w;     // [32]
    const float* conf_b;     // [1]
    float* nd_out;           // normalized depth, element (b, p) at b * nd_bstride + p * nd_pstride
    size_t nd_bstride, nd_pstride;
    float* conf;             // [n_px] or null
    float* conf_logit;       // [n_px] or null
    float* depth_out;        // [n_px] or null
    const float* depth_min;
    const float* depth_max;
    int n_px, P;             // B * P pixels, P per batch item
    int* err_flag;
};

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_f16(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// 8 consecutive channels -> 16 bytes of fp16 roundings and 16 bytes of fp16-rounded remainders (mmaconv.cuh:split_f16)
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
    split_f16(make_float2(x[0], x[1]), hi.x, lo.x);
    split_f16(make_float2(x[2], x[3]), hi.y, lo.y);
    split_f16(make_float2(x[4], x[5]), hi.z, lo.z);
    split_f16(make_float2(x[6], x[7]), hi.w, lo.w);
}

__global__ void __launch_bounds__(HF_THREADS, 2) head_fused_kernel(const Params prm) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sW = smem;                                  // W1 hi | W1 lo | W2 hi | W2 lo | bias
    const float* sBias = reinterpret_cast<const float*>(sW + 2 * W1_BYTES + 2 * W2_BYTES);
    unsigned char* sA2 = sW + BLOB_BYTES;                      // hi | lo
    unsigned char* sA1 = sA2;                                  // hi | lo in A2's first 16 KB: dead once fc1 has completed
    float* sX = reinterpret_cast<float*>(sA2 + 2 * A2_BYTES);  // [2 halves][128 pixels][4]
    uint64_t* sBar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sX) + X_BYTES);   // weights, fc1, fc2
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 3);   // TMEM base: fc1 accumulator at column 0, fc2 at column 64
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_w = tc5::smem_u32(sBar), bar_1 = tc5::smem_u32(sBar + 1), bar_2 = tc5::smem_u32(sBar + 2);
    const int row0 = blockIdx.x * HF_M;

    if (warp == 0) tc5::tmem_alloc(tc5::smem_u32(sTmem), TMEM_COLS);
    if (tid == 32) {
        tc5::mbar_init(bar_w, 1);
        tc5::mbar_init(bar_1, 1);
        tc5::mbar_init(bar_2, 1);
        tc5::fence_mbar_init();
        tc5::mbar_expect_tx(bar_w, BLOB_BYTES);
        const unsigned char* g = static_cast<const unsigned char*>(prm.blob);
        tc5::bulk_g2s(tc5::smem_u32(sW), g, 2 * W1_BYTES, bar_w);
        tc5::bulk_g2s(tc5::smem_u32(sW + 2 * W1_BYTES), g + 2 * W1_BYTES, W2_BYTES, bar_w);
        tc5::bulk_g2s(tc5::smem_u32(sW + 2 * W1_BYTES + W2_BYTES), g + 2 * W1_BYTES + W2_BYTES, W2_BYTES + BIAS_BYTES, bar_w);
    }
    pdl_trigger();
    pdl_wait();                  // TMEM allocation and the weight copies overlap the predecessor's tail
    // ---- A1 = split(t[:, 0:32]): thread -> (pixel row, two 8-channel chunks); consecutive threads = consecutive rows;
    //      all four 16-byte loads of a thread are in flight before the first conversion
    {
        static_assert(HF_M * KC1 == 2 * HF_THREADS, "two chunks per thread");
        const int row = tid & (HF_M - 1), kc0 = tid >> 7;                  // chunks kc0 and kc0 + 2
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + row < prm.n_px) {
            const float* src = prm.t + (size_t)(row0 + row) * 64 + kc0 * 8;
            q[0] = ldg4(src); q[1] = ldg4(src + 4); q[2] = ldg4(src + 16); q[3] = ldg4(src + 20);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float x[8] = {q[2 * u].x, q[2 * u].y, q[2 * u].z, q[2 * u].w, q[2 * u + 1].x, q[2 * u + 1].y, q[2 * u + 1].z, q[2 * u + 1].w};
            uint4 hi, lo;
            split8(x, hi, lo);
            const int kc = kc0 + 2 * u;
            *reinterpret_cast<uint4*>(sA1 + (kc * HF_M + row) * 16) = hi;
            *reinterpret_cast<uint4*>(sA1 + A1_BYTES + (kc * HF_M + row) * 16) = lo;
        }
    }
    tc5::fence_async_shared();           // generic-proxy writes -> visible to the tensor core's async proxy
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tm1 = sTmem[0], tm2 = tm1;        // D2 overwrites D1's columns once A2 has been built from them

    // ---- fc1: one thread issues 3 products x 2 K-steps and commits
    bool ok = true;
    if (tid == 0) {
        ok = tc5::mbar_wait_bounded(bar_w, 0);
        if (ok) {
            constexpr uint32_t idesc = make_idesc_f16(N1);
            constexpr uint32_t lbo_a = HF_M * 16, lbo_b = N1 * 16;
            const uint64_t a_hi = tc5::make_desc(tc5::smem_u32(sA1), lbo_a, 128u), a_lo = tc5::make_desc(tc5::smem_u32(sA1 + A1_BYTES), lbo_a, 128u);
            const uint64_t b_hi = tc5::make_desc(tc5::smem_u32(sW), lbo_b, 128u), b_lo = tc5::make_desc(tc5::smem_u32(sW + W1_BYTES), lbo_b, 128u);
            uint32_t acc = 0;
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
                const uint64_t da = prod == 2 ? a_lo : a_hi, db = prod == 1 ? b_lo : b_hi;
#pragma unroll
                for (int ks = 0; ks < KC1 / 2; ++ks) {
                    umma_f16(tm1, da + (uint64_t)((2u * lbo_a * ks) >> 4), db + (uint64_t)((2u * lbo_b * ks) >> 4), idesc, acc);
                    acc = 1;
                }
            }
        }
        tc5::umma_commit(bar_1);
    }
    // ---- A2 = split(ReLU(D1)): warp w owns TMEM lanes 32 * (w & 3) .. (pixels) and the column half (w >> 2)
    const int lg = warp & 3, half = warp >> 2, m = lg * 32 + lane;
    const uint32_t lane_base = (uint32_t)(lg * 32) << 16;
    bool done = tc5::mbar_wait_bounded(bar_1, 0);
    tc5::fence_after_sync();
    if (done) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float v[16];
            tc5::tmem_ld16(tm1 + lane_base + (uint32_t)(half * 32 + c * 16), v);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaxf(v[8 * j + i], 0.f);
                float amax = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) amax = fmaxf(amax, x[i]);
                if (!(amax <= 65504.f) && prm.err_flag) atomicOr(prm.err_flag, 2);      // fp16 range guard (imvs_device_status)
                uint4 hi, lo;
                split8(x, hi, lo);
                const int kc = half * 4 + c * 2 + j;
                *reinterpret_cast<uint4*>(sA2 + (kc * HF_M + m) * 16) = hi;
                *reinterpret_cast<uint4*>(sA2 + A2_BYTES + (kc * HF_M + m) * 16) = lo;
            }
        }
    }
    tc5::fence_async_shared();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();

    // ---- fc2: 3 products x 4 K-steps into the same TMEM columns (every thread has read D1: the barrier above)
    if (tid == 0) {
        if (ok && done) {
            constexpr uint32_t idesc = make_idesc_f16(N2);
            constexpr uint32_t lbo_a = HF_M * 16, lbo_b = N2 * 16;
            const unsigned char* w2 = sW + 2 * W1_BYTES;
            const uint64_t a_hi = tc5::make_desc(tc5::smem_u32(sA2), lbo_a, 128u), a_lo = tc5::make_desc(tc5::smem_u32(sA2 + A2_BYTES), lbo_a, 128u);
            const uint64_t b_hi = tc5::make_desc(tc5::smem_u32(w2), lbo_b, 128u), b_lo = tc5::make_desc(tc5::smem_u32(w2 + W2_BYTES), lbo_b, 128u);
            uint32_t acc = 0;
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
                const uint64_t da = prod == 2 ? a_lo : a_hi, db = prod == 1 ? b_lo : b_hi;
#pragma unroll
                for (int ks = 0; ks < KC2 / 2; ++ks) {
                    umma_f16(tm2, da + (uint64_t)((2u * lbo_a * ks) >> 4), db + (uint64_t)((2u * lbo_b * ks) >> 4), idesc, acc);
                    acc = 1;
                }
            }
        }
        tc5::umma_commit(bar_2);
    }
    done = tc5::mbar_wait_bounded(bar_2, 0) && done;
    tc5::fence_after_sync();

    // ---- regression: thread = (pixel m, column half): 128 of the pixel's 256 logits
    const int col0 = half * 128;
    float mx = -3.0e38f;
    int bi = col0;
    if (done) {
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            float v[16];
            tc5::tmem_ld16(tm2 + lane_base + (uint32_t)(col0 + c * 16), v);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 bq = *reinterpret_cast<const float4*>(sBias + col0 + c * 16 + i);       // broadcast read
                v[i] += bq.x; v[i + 1] += bq.y; v[i + 2] += bq.z; v[i + 3] += bq.w;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (v[i] > mx) { mx = v[i]; bi = col0 + c * 16 + i; }          // first maximum, like torch.argmax
        }
    }
    sX[(half * HF_M + m) * 4 + 0] = mx;
    sX[(half * HF_M + m) * 4 + 1] = __int_as_float(bi);
    __syncthreads();
    {
        const float omx = sX[((half ^ 1) * HF_M + m) * 4 + 0];
        const int obi = __float_as_int(sX[((half ^ 1) * HF_M + m) * 4 + 1]);
        // ties go to the lower index = the lower half
        const bool take = half == 0 ? (omx > mx) : (omx >= mx);
        if (take) { mx = omx; bi = obi; }
    }
    // exp(v - mx) = 2^(v * log2e - mx * log2e): one FFMA + one MUFU.EX2 per bin.  The rounding of the common offset
    // mx * log2e scales every bin of the pixel by the same factor, which cancels in e / sum(e); what remains is the
    // 2-ulp error of ex2.approx, the same grade as expf.
    float s = 0.f, num = 0.f, den = 0.f;
    if (done) {
        const float L2E = 1.4426950408889634f, mL = -mx * L2E;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            float v[16];
            tc5::tmem_ld16(tm2 + lane_base + (uint32_t)(col0 + c * 16), v);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 bq = *reinterpret_cast<const float4*>(sBias + col0 + c * 16 + i);
                v[i] += bq.x; v[i + 1] += bq.y; v[i + 2] += bq.z; v[i + 3] += bq.w;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v[i], L2E, mL)));
                v[i] = e;
                s += e;
            }
            // window: indices clamp(bi-4 .. bi+4, 0, 255); clamped duplicates are counted repeatedly (itermvs.py:205-218).
            // Only the (at most two) 16-column chunks that intersect the window take this branch.
            const int ch0 = col0 + c * 16;
            if (ch0 + 15 >= bi - IMVS_RADIUS && ch0 <= bi + IMVS_RADIUS) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int ch = ch0 + i;
                    int mult = (ch >= bi - IMVS_RADIUS && ch <= bi + IMVS_RADIUS) ? 1 : 0;
                    if (ch == 0) mult = max(0, IMVS_RADIUS + 1 - bi);
                    if (ch == IMVS_OUT_BINS - 1) mult = max(0, bi - (IMVS_OUT_BINS - 2 - IMVS_RADIUS));
                    num = fmaf((float)(mult * ch), v[i], num);
                    den = fmaf((float)mult, v[i], den);
                }
            }
        }
    }
    __syncthreads();                     // the arg-max exchange has been read by everyone
    sX[(half * HF_M + m) * 4 + 0] = s;
    sX[(half * HF_M + m) * 4 + 1] = num;
    sX[(half * HF_M + m) * 4 + 2] = den;
    __syncthreads();
    const int gw = row0 + m;
    if (gw < prm.n_px) {
        if (half == 0) {
            const float* o = sX + (HF_M + m) * 4;
            const float st = s + o[0];
            // sum_k idx_k p_k / (1e-6 + sum_k p_k) with p = e / sum(e)
            const float nume = (num + o[1]) / st, dene = (den + o[2]) / st;
            const float ndv = (nume / (1e-6f + dene)) / (float)(IMVS_OUT_BINS - 1);
            const int b = gw / prm.P, p = gw - b * prm.P;
            prm.nd_out[(size_t)b * prm.nd_bstride + (size_t)p * prm.nd_pstride] = ndv;
            if (prm.depth_out) {
                const float inv_min = 1.0f / prm.depth_min[b], inv_max = 1.0f / prm.depth_max[b];
                prm.depth_out[gw] = unnormalize_depth(ndv, inv_min, inv_max);
            }
        } else if (prm.conf || prm.conf_logit) {
            // confidence head: 1x1 over conv0's channels 32..63 (+ bias), sigmoid (itermvs.py:147-151, 197-199)
            const float* tc = prm.t + (size_t)gw * 64 + 32;
            float cs = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 a = ldg4(tc + c), w = ldg4(prm.conf_w + c);
                cs = fmaf(a.x, w.x, cs); cs = fmaf(a.y, w.y, cs); cs = fmaf(a.z, w.z, cs); cs = fmaf(a.w, w.w, cs);
            }
            cs += ldg(prm.conf_b);
            if (prm.conf_logit) prm.conf_logit[gw] = cs;
            if (prm.conf) prm.conf[gw] = sigmoidf_(cs);
        }
    }
    if ((!ok || !done) && lane == 0 && prm.err_flag) atomicExch(prm.err_flag, 1);
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tm1, TMEM_COLS);
}

inline int launch(const Params& prm, cudaStream_t st) {
    static int smem_ok = 0;
    IMVS_TRY(ensure_dynamic_smem(head_fused_kernel, SMEM_BYTES, &smem_ok));
    const int blocks = cdiv(prm.n_px, HF_M);
    if (launch_k(head_fused_kernel, dim3(blocks), dim3(HF_THREADS), SMEM_BYTES, st, prm) != cudaSuccess)
        return fail("launch of head_fused_kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

}  // namespace hf
}  // namespace imvs
