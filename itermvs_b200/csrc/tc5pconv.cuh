// Persistent, warp-specialised tcgen05 implicit-GEMM convolution fed by TMA from PRE-SPLIT activations.
//
// The fp32-grade mode computes every convolution as three fp16 products (hi*hi + hi*lo + lo*hi, fp32 accumulation in
// TMEM) of operands split x = hi + lo.  tc5conv.cuh:tc5h_conv_kernel reads fp32 activations and splits them while
// staging: 17 M warp instructions per 16 -> 16 layer, one CTA per 128-pixel block, load -> convert -> MMA -> epilogue
// strictly one after the other (ncu: long_scoreboard 9-13 per issue, no unit above 45 %).  Here the PRODUCER of an
// activation writes it already split, as two fp16 tensors ("split planes")
//         [N][C/8][H][W][8 halves]          (hi plane, lo plane; together the bytes of the fp32 tensor)
// which is exactly the tcgen05 K-major / no-swizzle canonical operand order of a haloed tile: one
// cp.async.bulk.tensor (TMA, 4-D tile {32 px * 8 halves, rows, C/8, 1}) per plane drops the tile
//         element (slot, k)  at  (k / 8) * LBO + slot * 16 B + (k % 8) * 2 B,   slot = tile_row * 32 + tile_col
// into shared memory, zero-filled outside the image (= the convolution's padding), with no thread touching it.
// A stencil tap (dy, dx) is the same tile addressed (dy * 32 + dx) * 16 bytes further (tc5conv.cuh), so the A operand
// of every tap is a shifted shared-memory descriptor.
//
//   grid = min(#tiles, #SMs) persistent CTAs of 320 .. 576 threads:
//     warp 0      TMA producer: tile t+1.. into a ring of NSTAGES shared-memory stages (full / empty mbarriers)
//     warp 1      MMA issuer: MB (1 or 2) M-blocks of 128 slots x 9 taps x 3 products x CINP/16 k-steps of
//                 tcgen05.mma.kind::f16 into one of TWO TMEM accumulator sets; tcgen05.commit frees the stage and
//                 publishes the accumulator
//     warps 2..   epilogue (8 .. 16 warps): tcgen05.ld, bias / residual (read from split planes) / ReLU, fp16 range guard, and the
//                 stores: split planes for the next tcgen05 layer and / or fp32 NHWC for the other consumers.  The
//                 residual of tile t+2 is requested before tile t is processed.
//   weights ([tap][CINP/8][hi | lo][NB][8 halves], _pack.py:pack_umma_f16i) stay resident in shared memory.
// Tile = 32 slots wide (30 valid output columns for a 3x3), 4 * MB output rows.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "mmaconv.cuh"
#include "tc5conv.cuh"

namespace imvs {
namespace tc5p {

using tc5::smem_u32;
using tc5::mbar_init;
using tc5::fence_mbar_init;
using tc5::mbar_expect_tx;
using tc5::mbar_wait_bounded;
using tc5::bulk_g2s;
using tc5::tmem_alloc;
using tc5::tmem_dealloc;
using tc5::fence_before_sync;
using tc5::fence_after_sync;
using tc5::umma_commit;
using tc5::umma_f16k;
using tc5::make_desc;
using tc5::make_idesc_f16k;

constexpr int WT = 32;                 // slots per tile row
// CTA: warp 0 producer, warp 1 MMA issuer, then 4 * MB * CS epilogue warps: a warp reads the TMEM lanes 32 * (warp % 4) ..,
// its group index selects the M-block and one of CS channel slices (the epilogue, not the tensor core, was the pipeline's
// slowest stage with 8 warps: ncu r2c17 -- the MMA warp spinning on the accumulator-empty barrier)
// TC: transposed convolution -- the four slices are the output parities.  K1 (1x1 layers): 8-channel slices, i.e. 24 epilogue
// warps, were measured for the lateral layers (gpurun call r2c36 / r2c37): same 49 us with twice the executed instructions
// (32.7 M vs 17 M: the per-thread fixed part), issue slots 67 % busy -- that epilogue is instruction-bound, so K1 changes nothing
template <int NB, int MB, bool TC = false, bool K1 = false> struct Shape {
    static constexpr int CS = TC ? 4 : (MB == 2 ? 2 : (NB % 32 == 0 ? 4 : (NB == 48 ? 3 : 2)));      // slices per M-block
    static constexpr int NCH = TC ? NB : NB / CS;                                        // channels per epilogue thread (8 or 16)
    static constexpr int EPI_WARPS = 4 * MB * CS;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static_assert(NCH % 8 == 0, "epilogue channel slice");
};
constexpr int MAX_STAGES = 6;

// one activation tensor stored as split planes [N][C/8][H][W][8 halves]
struct Split {
    __half* hi;
    __half* lo;
};
inline Split split_at(void* base, size_t elems) {     // the two planes of a tensor of `elems` values packed back to back
    return Split{static_cast<__half*>(base), static_cast<__half*>(base) + elems};
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// weights of a launch: one packed set (_pack.py:pack_umma_f16i) or three, chosen per image n as the mma.sync engine's
// MmaWeightSel does (CorrNet batches the slices of three pyramid levels, each with its own network)
struct WSel {
    const void* w[3];
    int nsets, period, split1, split2;
    __host__ __device__ int pick(int n) const {
        if (nsets == 1) return 0;
        const int r = n % period;
        return r < split1 ? 0 : (r < split2 ? 1 : 2);
    }
};
inline WSel wsel_single(const void* w) { return WSel{{w, w, w}, 1, 1, 1, 1}; }

struct Geo {
    int tiles_x, tiles_y, n_tiles;
    int THo;             // output rows per tile = 4 * MB
    int rows;            // tile rows staged = THo + 2 * pad
    int pad, dil, ks;
    int valid;           // valid output columns per tile = WT - 2 * pad
    int nstages;
    int dbg;             // timing experiments only (IMVS_TUNE_TC5P_DBG, results wrong): 1 = epilogue does not store, 2 = one tap's MMAs only
    uint32_t a_bytes;    // bytes of one plane of one stage = KC * rows * WT * 16
};

// ---- epilogue: one thread = one pixel, channels [c0, c0 + NCH) ------------------------------------------------------
// Interface of an epilogue class: members H, W; Pre<NCH>; prefetch<NB, NCH>(n, oy, ox, c0, pre) (global reads that do not
// depend on the accumulator, issued kAhead tiles ahead when wants_prefetch()); store<NB, NCH>(n, oy, ox, c0, v, pre, status).
// out = relu?(acc + bias + residual); written as split planes (the next tcgen05 layer's operand) and / or fp32 NHWC.
// index (in 16-byte units) of chunk kc of pixel (oy, ox) of image n in the PARITY-PLANE layout [4N][KCo][H/2][W/2][8] that a
// stride-2 layer reads (image 4n + 2 (oy & 1) + (ox & 1) holds the pixels of that parity); H, W even
__device__ __forceinline__ size_t parity_index(int n, int oy, int ox, int kc, int KCo, int H, int W) {
    return (((size_t)(4 * n + 2 * (oy & 1) + (ox & 1)) * KCo + kc) * (H >> 1) + (oy >> 1)) * (W >> 1) + (ox >> 1);
}

struct Epi {
    static constexpr int kAhead = 2;
    Split out;               // [N][NB/8][H][W][8] or {nullptr, nullptr}
    float* out32;            // [N][H][W][NB] or nullptr
    Split res;               // residual split planes (same shape as out) or {nullptr, nullptr}
    const float* bias;       // [NB] or nullptr
    int H, W, relu;
    Split outp = {nullptr, nullptr};     // the same values as parity planes (operand of a following stride-2 layer) or null
    int kco = 0;             // 8-channel chunks of the output / residual tensors when fewer than NB / 8 (cout padded to 16); 0: NB / 8
    float* out32p = nullptr; // the same values as fp32 in the chunk-planar order [N][NB/8][H][W][8] (coarse map of a lateral stage) or null

    template <int NCH> struct Pre { uint4 h[NCH / 8], l[NCH / 8]; };
    __device__ __forceinline__ bool wants_prefetch() const { return res.hi != nullptr; }

    template <int NB, int NCH>
    __device__ __forceinline__ void prefetch(int n, int oy, int ox, int c0, Pre<NCH>& p) const {
        if (!res.hi) return;
        const size_t plane = (size_t)H * W, pix = (size_t)oy * W + ox;
        const int KCo = kco ? kco : NB / 8;
#pragma unroll
        for (int j = 0; j < NCH / 8; ++j) {
            if (c0 / 8 + j >= KCo) break;
            const size_t idx = ((size_t)n * KCo + (c0 / 8 + j)) * plane + pix;
            p.h[j] = __ldg(reinterpret_cast<const uint4*>(res.hi) + idx);
            p.l[j] = __ldg(reinterpret_cast<const uint4*>(res.lo) + idx);
        }
    }

    template <int NB, int NCH>
    __device__ __forceinline__ void store(int n, int oy, int ox, int c0, float (&v)[NCH], const Pre<NCH>& p, int* status) const {
        const size_t plane = (size_t)H * W, pix = (size_t)oy * W + ox;
        const int KCo = kco ? kco : NB / 8;
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < NCH / 8; ++j) {
            if (c0 / 8 + j >= KCo) break;                // padded output channels
            float* x = v + 8 * j;
            if (bias) {
                const float4 b0 = ldg4(bias + c0 + 8 * j), b1 = ldg4(bias + c0 + 8 * j + 4);
                x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w; x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
            }
            if (res.hi) {
                const uint32_t hh[4] = {p.h[j].x, p.h[j].y, p.h[j].z, p.h[j].w}, ll[4] = {p.l[j].x, p.l[j].y, p.l[j].z, p.l[j].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hh[q]));
                    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&ll[q]));
                    x[2 * q] += a.x + b.x;
                    x[2 * q + 1] += a.y + b.y;
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (relu) x[q] = fmaxf(x[q], 0.f);
                amax = fmaxf(amax, fabsf(x[q]));
            }
            if (out.hi || outp.hi) {
                uint4 h, l;
                split_f16(make_float2(x[0], x[1]), h.x, l.x);
                split_f16(make_float2(x[2], x[3]), h.y, l.y);
                split_f16(make_float2(x[4], x[5]), h.z, l.z);
                split_f16(make_float2(x[6], x[7]), h.w, l.w);
                if (out.hi) {
                    const size_t idx = ((size_t)n * KCo + (c0 / 8 + j)) * plane + pix;
                    reinterpret_cast<uint4*>(out.hi)[idx] = h;
                    reinterpret_cast<uint4*>(out.lo)[idx] = l;
                }
                if (outp.hi) {
                    const size_t idx = parity_index(n, oy, ox, c0 / 8 + j, KCo, H, W);
                    reinterpret_cast<uint4*>(outp.hi)[idx] = h;
                    reinterpret_cast<uint4*>(outp.lo)[idx] = l;
                }
            }
            if (out32p) {
                float4* o = reinterpret_cast<float4*>(out32p) + (((size_t)n * KCo + (c0 / 8 + j)) * plane + pix) * 2;
                o[0] = make_float4(x[0], x[1], x[2], x[3]);
                o[1] = make_float4(x[4], x[5], x[6], x[7]);
            }
        }
        if (out32) {
            float* o = out32 + ((size_t)n * plane + pix) * (8 * KCo) + c0;
#pragma unroll
            for (int c = 0; c < NCH; c += 4)
                if (c0 + c < 8 * KCo) *reinterpret_cast<float4*>(o + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
        // a value beyond +-65504 saturates in the split above / in the next layer's split: raise the flag (imvs_device_status bit 1)
        if (!(amax <= 65504.f) && status) atomicOr(status, 2);
    }
};

// [conv1 | downsample] of a ResidualBlock's first stage as one GEMM over the stacked output channels (module.py:36-49):
// channels [0, CO) -> relu(v + bias) -> y, channels [CO, 2 CO) -> v + bias -> ds; both written as split planes
struct EpiStack2 {
    static constexpr int kAhead = 1;
    __device__ __forceinline__ bool wants_prefetch() const { return false; }
    Split y, ds;             // [N][CO/8][H][W][8] each
    const float* bias;       // [2 CO]
    int H, W, CO;
    template <int NCH> struct Pre {};
    template <int NB, int NCH>
    __device__ __forceinline__ void prefetch(int, int, int, int, Pre<NCH>&) const {}
    template <int NB, int NCH>
    __device__ __forceinline__ void store(int n, int oy, int ox, int c0, float (&v)[NCH], const Pre<NCH>&, int* status) const {
        const size_t plane = (size_t)H * W, pix = (size_t)oy * W + ox;
        const bool first = c0 < CO;                      // a thread's channel slice lies in one half (NCH divides CO)
        const Split& o = first ? y : ds;
        const int cb = first ? c0 : c0 - CO;
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < NCH / 8; ++j) {
            float* x = v + 8 * j;
            const float4 b0 = ldg4(bias + c0 + 8 * j), b1 = ldg4(bias + c0 + 8 * j + 4);
            x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w; x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (first) x[q] = fmaxf(x[q], 0.f);
                amax = fmaxf(amax, fabsf(x[q]));
            }
            uint4 h, l;
            split_f16(make_float2(x[0], x[1]), h.x, l.x);
            split_f16(make_float2(x[2], x[3]), h.y, l.y);
            split_f16(make_float2(x[4], x[5]), h.z, l.z);
            split_f16(make_float2(x[6], x[7]), h.w, l.w);
            const size_t idx = ((size_t)n * (CO / 8) + (cb / 8 + j)) * plane + pix;
            reinterpret_cast<uint4*>(o.hi)[idx] = h;
            reinterpret_cast<uint4*>(o.lo)[idx] = l;
        }
        if (!(amax <= 65504.f) && status) atomicOr(status, 2);
    }
};

__device__ __forceinline__ uint32_t elect_one() {       // one lane of the (converged) warp; the compiler keeps the region uniform
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xffffffffu));
    return pred;
}

// 3x3, dilation DIL.  Per (tap, k-step) TWO MMAs instead of three: the weights of a tap sit in shared memory as
// [CINP/8][hi rows 0..NB) | lo rows NB..2NB)][8 halves], so  A_hi x [B_hi | B_lo]  is ONE N = 2*NB instruction filling the
// accumulator columns [0, NB) (hi*hi) and [NB, 2NB) (hi*lo), and  A_lo x B_hi  (N = NB) adds the third product to columns
// [0, NB); the epilogue sums the two column halves.  (The tensor core's cost per M=128, K=16 instruction is set by the A
// operand it reads, not by N at these sizes -- tools/ubench/umma_chain.cu.)
// STRIDE = 2 (3x3, padding 1): the input is stored as PARITY PLANES -- image (n, row parity rp, column parity cp) holds the
// pixels (2i + rp, 2j + cp) as an ordinary split-plane image [4N][C/8][H/2][W/2][8] -- so that every tap of the strided stencil
// is again a shifted view of a dense tile: tap (ky, kx) reads parity (ky != 1, kx != 1) at offset (ky == 2, kx == 2).  A stage
// holds the four parity sub-tiles (4 MB + 1 rows x 32 slots each, the odd ones starting one row / column earlier): eight TMA
// loads per tile, 31 valid output columns per 32.  CINP = 8 (one K chunk): the MMA's second K chunk aliases the first
// (descriptor LBO = 0) against zero weights.
// STRIDE = 0: ConvTranspose2d(k = 3, stride 2, padding 1, output_padding 1).  The tile lives on the INPUT grid (one extra row /
// column below / right); output parity (a, b) of input pixel (iy, ix) -> output (2 iy + a, 2 ix + b) is a stride-1 stencil over
// the input with 1, 2, 2 or 4 taps (mmaconv.cuh:tconv_tables): four accumulators per M-block, one per parity, nine tap MMAs in
// total as for a 3x3; the epilogue's four warp groups are the four parities (Epi::store gets c0 = parity * NB).
// The body of one layer, shared by the stand-alone kernel below and by multi-layer persistent kernels (evalnets.cu: the fused
// CorrNet): smem_raw = 128-byte aligned shared memory ([0, 256): mbarriers; then the stages; then the weights), tmem_d = 256
// allocated TMEM columns' base.  (Re-)initialises its mbarriers, stages its weights, runs the three roles over the layer's
// tiles and returns after a block barrier.  STANDALONE: the grid-dependency wait sits between the weight requests and the
// first access to the input.  Threads beyond Shape::THREADS (a fused kernel's block is sized for its widest layer) idle.
template <int CINP, int NB, int MB, int DIL, int KS, int STRIDE, bool STANDALONE, class Epi>
__device__ __forceinline__ void layer_body(const CUtensorMap& map_hi, const CUtensorMap& map_lo, const Epi& epi, const WSel& wsel,
                                           const Geo& geo, int* err_flag, unsigned char* smem_raw, const uint32_t tmem_d) {
    static_assert((CINP % 16 == 0 || CINP == 8) && NB % 16 == 0 && NB <= 64, "UMMA kind::f16 shape");
    static_assert(MB == 1 || (MB == 2 && NB <= 32), "M-blocks per tile (TMEM: 2 sets x MB x 2*NB columns <= 256)");
    static_assert(KS == 1 || KS == 3, "1x1 or 3x3");
    static_assert(STRIDE == 1 || ((STRIDE == 2 || STRIDE == 0) && KS == 3 && DIL == 1), "stride 2 / transposed: 3x3, no dilation");
    constexpr bool TCONV = STRIDE == 0;
    static_assert(!TCONV || (MB == 1 && NB == 16), "transposed: four parity accumulators of 2 * NB columns, two sets, in 256 TMEM columns");
    constexpr int KC = CINP / 8, KCW = (CINP + 15) / 16 * 2, KSTEPS = (CINP + 15) / 16, TAPS = KS * KS;
    constexpr int PAD = STRIDE != 1 ? 0 : DIL * (KS - 1) / 2;                         // halo columns lost per tile side
    constexpr int VALID = STRIDE != 1 ? WT - 1 : WT - 2 * PAD;                        // valid output columns per tile
    constexpr int ROWS = STRIDE != 1 ? 4 * MB + 1 : 4 * MB + 2 * PAD;                 // rows of one (sub-)tile
    constexpr int NSUB = STRIDE == 2 ? 4 : 1, SUBSLOT = ROWS * WT, NSLOT = NSUB * SUBSLOT;
    constexpr int NBS = TCONV ? 8 * NB : 2 * NB;                 // TMEM columns per M-block: [hi*hi + lo*hi | hi*lo] (x 4 parities)
    constexpr int ACC_COLS = MB * NBS;                           // one accumulator set
    constexpr int TMEM_COLS = 2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : 256);
    static_assert(2 * ACC_COLS <= 256, "TMEM budget");
    constexpr int SMEM_HEAD = 256;                               // = SMEM_HEAD_BYTES (mbarriers + TMEM slot)
    constexpr uint32_t SUB_BYTES = KC * SUBSLOT * 16;            // one plane of one (sub-)tile
    constexpr uint32_t A_BYTES = NSUB * SUB_BYTES;               // one plane of one stage
    constexpr uint32_t B_TAP_BYTES = KCW * 2 * NB * 16;          // hi and lo of one tap
    constexpr uint32_t LBO_A = CINP == 8 ? 0u : SUBSLOT * 16u, LBO_B = 2 * NB * 16;
    using Sh = Shape<NB, MB, TCONV, KS == 1>;
    constexpr int CS = Sh::CS, NCH = Sh::NCH, THREADS = Sh::THREADS;
    const int nst = geo.nstages;
    // barriers: [0, S) full, [S, 2S) empty, 2S + {0,1} accumulator full, 2S + {2,3} accumulator empty, 2S + 4 weights
    uint64_t* sBar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned char* sA = smem_raw + SMEM_HEAD;                                   // [nst][hi | lo][KC][NSLOT][16 B]
    unsigned char* sB = sA + (size_t)nst * 2 * A_BYTES;                         // [set][tap][KC][hi NB | lo NB][16 B]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(sBar);
    auto bar_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_empty = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
    auto bar_tfull = [&](int a) { return bar0 + 8u * (2 * MAX_STAGES + a); };
    auto bar_tempty = [&](int a) { return bar0 + 8u * (2 * MAX_STAGES + 2 + a); };
    const uint32_t bar_w = bar0 + 8u * (2 * MAX_STAGES + 4);

    static_assert(TMEM_COLS <= 256, "layers share a 256-column TMEM allocation");
    if (tid == 32) {
        if constexpr (!STANDALONE) {        // a previous layer used these barriers: all of its phases are complete
            for (int b = 0; b < 2 * MAX_STAGES + 5; ++b) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar0 + 8u * b) : "memory");
        }
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), THREADS - 64); }
        mbar_init(bar_w, 1);
        fence_mbar_init();
        // weights are constants of the forward pass: requested before the grid-dependency wait; global order = shared
        // order [tap][KC][hi | lo][NB][8] (_pack.py:pack_umma_f16i), one bulk copy per tap
        mbar_expect_tx(bar_w, (uint32_t)wsel.nsets * TAPS * B_TAP_BYTES);
        for (int set = 0; set < wsel.nsets; ++set)
            for (int tap = 0; tap < TAPS; ++tap)
                bulk_g2s(smem_u32(sB) + (set * TAPS + tap) * B_TAP_BYTES,
                         static_cast<const unsigned char*>(wsel.w[set]) + (size_t)tap * B_TAP_BYTES, B_TAP_BYTES, bar_w);
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if constexpr (STANDALONE) pdl_wait();

    const int tiles_x = geo.tiles_x, tiles_y = geo.tiles_y, n_tiles = geo.n_tiles;
    auto decode = [&](int tile, int& n, int& oy0, int& ox0) {
        const int tx = tile % tiles_x, t2 = tile / tiles_x;
        n = t2 / tiles_y;
        oy0 = (t2 - n * tiles_y) * (4 * MB);
        ox0 = tx * VALID;
    };

    if (warp == 0) {
        // ---- TMA producer (the whole warp walks the tile list; one elected lane issues) ----------------------------
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int s = it % nst, ph = (it / nst) & 1;
            if (!mbar_wait_bounded(bar_empty(s), ph ^ 1)) { if (lane == 0 && err_flag) atomicOr(err_flag, 1); break; }
            int n, oy0, ox0;
            decode(tile, n, oy0, ox0);
            if (elect_one()) {
                mbar_expect_tx(bar_full(s), 2 * A_BYTES);
                const uint32_t dst = smem_u32(sA) + (uint32_t)s * 2 * A_BYTES;
                if constexpr (STRIDE != 2) {
                    tma_load_4d(dst, &map_hi, bar_full(s), (ox0 - PAD) * 8, oy0 - PAD, 0, n);
                    tma_load_4d(dst + A_BYTES, &map_lo, bar_full(s), (ox0 - PAD) * 8, oy0 - PAD, 0, n);
                } else {
#pragma unroll
                    for (int par = 0; par < 4; ++par) {          // par = 2 * rp + cp; odd parities start one row / column earlier
                        tma_load_4d(dst + par * SUB_BYTES, &map_hi, bar_full(s), (ox0 - (par & 1)) * 8, oy0 - (par >> 1), 0, 4 * n + par);
                        tma_load_4d(dst + A_BYTES + par * SUB_BYTES, &map_lo, bar_full(s), (ox0 - (par & 1)) * 8, oy0 - (par >> 1), 0, 4 * n + par);
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---- MMA issuer (whole warp, one elected lane issues a tile's MMAs and both commits) -----------------------
        bool ok = mbar_wait_bounded(bar_w, 0);
        constexpr uint32_t idesc2 = make_idesc_f16k(2 * NB), idesc1 = make_idesc_f16k(NB);
        constexpr uint64_t dconst_a = ((uint64_t)(LBO_A >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
        constexpr uint64_t dconst_b = ((uint64_t)(LBO_B >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
        const uint64_t db0 = dconst_b | (uint64_t)((smem_u32(sB) & 0x3FFFFu) >> 4);
        int it = 0;
        for (int tile = blockIdx.x; ok && tile < n_tiles; tile += gridDim.x, ++it) {
            const int s = it % nst, ph = (it / nst) & 1, acc = it & 1, aph = (it >> 1) & 1;
            ok = mbar_wait_bounded(bar_full(s), ph) && mbar_wait_bounded(bar_tempty(acc), aph ^ 1);
            if (!ok) break;
            fence_after_sync();
            int nimg, oy0_, ox0_;
            decode(tile, nimg, oy0_, ox0_);
            const uint32_t set16 = (uint32_t)wsel.pick(nimg) * (TAPS * (B_TAP_BYTES >> 4));      // this image's weight set
            if (elect_one()) {
                const uint64_t da_hi = dconst_a | (uint64_t)(((smem_u32(sA) + (uint32_t)s * 2 * A_BYTES) & 0x3FFFFu) >> 4);
                const uint64_t da_lo = da_hi + (A_BYTES >> 4);
                const uint32_t dcol = tmem_d + (uint32_t)(acc * ACC_COLS);
                constexpr uint32_t KA = (2u * LBO_A) >> 4, KB = (2u * LBO_B) >> 4;     // k-step advance (two K chunks), 16-byte units
                if constexpr (!TCONV) {
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap) {
                        if ((geo.dbg & 2) && tap > 0) break;
#pragma unroll
                        for (int k16 = 0; k16 < KSTEPS; ++k16) {
                            // slots == 16-byte units: tap shift + k-step advance
                            constexpr uint32_t SUB16 = SUB_BYTES >> 4;
                            const int ky = tap / KS, kx = tap % KS;
                            const uint32_t tapoff = STRIDE == 1 ? (uint32_t)(ky * DIL * WT + kx * DIL)
                                                                : (uint32_t)(2 * (ky != 1) + (kx != 1)) * SUB16 + (uint32_t)((ky == 2) * WT + (kx == 2));
                            const uint32_t shift = tapoff + (uint32_t)k16 * KA;
                            const uint64_t db = db0 + (uint64_t)(set16 + tap * (B_TAP_BYTES >> 4) + k16 * KB);
                            const uint32_t accum = (tap | k16) != 0;
#pragma unroll
                            for (int mb = 0; mb < MB; ++mb)          // A_hi x [B_hi | B_lo]
                                umma_f16k(dcol + mb * NBS, da_hi + (shift + mb * 128), db, idesc2, accum);
#pragma unroll
                            for (int mb = 0; mb < MB; ++mb)          // A_lo x B_hi
                                umma_f16k(dcol + mb * NBS, da_lo + (shift + mb * 128), db, idesc1, 1u);
                        }
                    }
                } else {
#pragma unroll
                    for (int par = 0; par < 4; ++par) {              // output parity (a, b) = (par >> 1, par & 1)
                        const int a = par >> 1, b = par & 1;
#pragma unroll
                        for (int i = 0; i <= a; ++i) {               // a = 0: ky = 1 (dy 0);  a = 1: ky = 0 (dy +1), ky = 2 (dy 0)
#pragma unroll
                            for (int j = 0; j <= b; ++j) {
                                const int ky = a == 0 ? 1 : (i == 0 ? 0 : 2), dy = (a == 1 && i == 0) ? 1 : 0;
                                const int kx = b == 0 ? 1 : (j == 0 ? 0 : 2), dx = (b == 1 && j == 0) ? 1 : 0;
                                const int tap = ky * 3 + kx;
#pragma unroll
                                for (int k16 = 0; k16 < KSTEPS; ++k16) {
                                    const uint32_t shift = (uint32_t)(dy * WT + dx) + (uint32_t)k16 * KA;
                                    const uint64_t db = db0 + (uint64_t)(set16 + tap * (B_TAP_BYTES >> 4) + k16 * KB);
                                    const uint32_t accum = (i | j | k16) != 0;
                                    umma_f16k(dcol + par * 2 * NB, da_hi + shift, db, idesc2, accum);
                                    umma_f16k(dcol + par * 2 * NB, da_lo + shift, db, idesc1, 1u);
                                }
                            }
                        }
                    }
                }
                umma_commit(bar_empty(s));          // the stage may be refilled once these MMAs have read it
                umma_commit(bar_tfull(acc));        // ... and the accumulator set is complete
            }
            __syncwarp();
        }
        if (!ok && lane == 0 && err_flag) atomicOr(err_flag, 1);
    } else if (warp < THREADS / 32) {
        // ---- epilogue -------------------------------------------------------------------------------------------
        const int ew = warp - 2, lg = warp & 3, grp = ew >> 2;       // TMEM lanes 32 * (warp % 4) ..; grp: (M-block, channel slice)
        const int mb = grp / CS, c0 = (grp % CS) * NCH;               // transposed: c0 = parity * NB
        const int col0 = TCONV ? (grp % CS) * 2 * NB : c0;            // the slice's first accumulator column inside its M-block
        const int slot = mb * 128 + lg * 32 + lane, r = slot >> 5, c = slot & 31;
        auto pixel = [&](int tile, int& n, int& oy, int& ox) {
            int oy0, ox0;
            decode(tile, n, oy0, ox0);
            oy = oy0 + r; ox = ox0 + c;
            return c < VALID && oy < epi.H && ox < epi.W;
        };
        // accumulator-independent global reads (residual, gate operands) are requested Epi::kAhead (1 or 2) tiles ahead
        constexpr int AHEAD = Epi::kAhead;
        const bool wants_pre = epi.wants_prefetch();
        typename Epi::template Pre<NCH> pre{}, pre1{}, pre2{};
        if (wants_pre) {
            int n, oy, ox;
            if (blockIdx.x < n_tiles && pixel(blockIdx.x, n, oy, ox)) epi.template prefetch<NB, NCH>(n, oy, ox, c0, pre);
            if (AHEAD == 2 && blockIdx.x + gridDim.x < n_tiles && pixel(blockIdx.x + gridDim.x, n, oy, ox)) epi.template prefetch<NB, NCH>(n, oy, ox, c0, pre1);
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1, aph = (it >> 1) & 1;
            int n, oy, ox;
            const bool okp = pixel(tile, n, oy, ox);
            if (wants_pre) {
                int n2, oy2, ox2;
                const int nxt = tile + AHEAD * gridDim.x;
                if (nxt < n_tiles && pixel(nxt, n2, oy2, ox2)) epi.template prefetch<NB, NCH>(n2, oy2, ox2, c0, AHEAD == 2 ? pre2 : pre1);
            }
            if (!mbar_wait_bounded(bar_tfull(acc), aph)) { if (lane == 0 && err_flag) atomicOr(err_flag, 1); break; }
            fence_after_sync();
            float v[NCH], u[NCH];
            const uint32_t taddr = tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * ACC_COLS + mb * NBS + col0);
#pragma unroll
            for (int j = 0; j < NCH / 8; ++j) {
                tmem_ld8_nowait(taddr + 8 * j, v + 8 * j);
                tmem_ld8_nowait(taddr + NB + 8 * j, u + 8 * j);
            }
            tmem_ld_wait();
            fence_before_sync();
            mbar_arrive(bar_tempty(acc));           // the accumulator set is in registers: the MMA warp may overwrite it
#pragma unroll
            for (int q = 0; q < NCH; ++q) v[q] += u[q];
            if (okp && !(geo.dbg & 1)) epi.template store<NB, NCH>(n, oy, ox, c0, v, pre, err_flag);
            pre = pre1;
            if (AHEAD == 2) pre1 = pre2;
        }
    }
    fence_before_sync();
    __syncthreads();
}

constexpr int SMEM_HEAD_BYTES = 256;       // mbarriers (2 * MAX_STAGES + 5) + the TMEM base slot, in front of the stages
template <int NB, int MB, int STRIDE> constexpr int tmem_cols() {      // two accumulator sets of MB x (2 NB, or 8 NB transposed) columns
    constexpr int c = 2 * MB * (STRIDE == 0 ? 8 * NB : 2 * NB);
    return c <= 64 ? 64 : (c <= 128 ? 128 : 256);
}

template <int CINP, int NB, int MB, int DIL, int KS, int STRIDE, class Epi>
__global__ void __launch_bounds__((Shape<NB, MB, STRIDE == 0, KS == 1>::THREADS), 1)
tc5p_conv_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const Epi epi,
                 const WSel wsel, const Geo geo, int* err_flag) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem_raw = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem_raw + SMEM_HEAD_BYTES - 16);
    constexpr int COLS = tmem_cols<NB, MB, STRIDE>();
    if (threadIdx.x < 32) tmem_alloc(smem_u32(sTmem), COLS);
    pdl_trigger();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = *sTmem;
    layer_body<CINP, NB, MB, DIL, KS, STRIDE, true, Epi>(map_hi, map_lo, epi, wsel, geo, err_flag, smem_raw, tmem_d);
    if (threadIdx.x < 32) tmem_dealloc(tmem_d, COLS);
}

// ---- host side ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();       // cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (featurenet.cu); nullptr if absent
int sm_count();                        // multiprocessors of the current device (cached per device)
int grid_limit();                      // CTAs a persistent launch may use: sm_count() / imvs_set_sm_share() (featurenet.cu)
void set_sm_share(int share);
int get_sm_share();

// tensor map over one split plane [N][KC][H][W][8 halves] as the 4-D tensor {W*8, H, KC, N}; box {256, rows, KC, 1}
// (parity planes of a stride-2 layer's input: N = 4 * images, H and W = half the image's)
inline int make_plane_map(CUtensorMap* map, const __half* plane, int N, int KC, int H, int W, int rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    IMVS_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)KC, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)KC * H * W * 16};
    const cuuint32_t box[4] = {256, (cuuint32_t)rows, (cuuint32_t)KC, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(plane), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IMVS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for a [%d][%d][%d][%d][8] plane", (int)r, N, KC, H, W);
    return 0;
}

// tile geometry, stage count and shared-memory need of one layer on `ctas` persistent CTAs (budget: bytes of dynamic shared memory)
template <int CINP, int NB, int MB, int DIL, int KS, int STRIDE>
int plan_layer(const char* name, const WSel& wsel, int N, int H, int W, int ctas, size_t budget, Geo& g, size_t& smem) {
    constexpr int KC = CINP / 8, KCW = (CINP + 15) / 16 * 2, PAD = STRIDE != 1 ? 0 : DIL * (KS - 1) / 2;
    g = Geo{};
    g.ks = KS; g.dil = DIL; g.pad = PAD;
    g.valid = STRIDE != 1 ? WT - 1 : WT - 2 * PAD;
    g.THo = 4 * MB;
    g.rows = STRIDE != 1 ? g.THo + 1 : g.THo + 2 * PAD;
    g.tiles_x = cdiv(W, g.valid); g.tiles_y = cdiv(H, g.THo);
    g.n_tiles = g.tiles_x * g.tiles_y * N;
    g.a_bytes = (uint32_t)(STRIDE == 2 ? 4 : 1) * KC * g.rows * WT * 16;
    g.dbg = tune("TC5P_DBG", 0);
    const size_t fixed = (size_t)wsel.nsets * KS * KS * 2 * KCW * NB * 16 + SMEM_HEAD_BYTES + 128 + 256;   // weights, barriers + TMEM slot, alignment, overshoot
    IMVS_REQUIRE(fixed + 2 * (size_t)2 * g.a_bytes <= budget, "%s: tile does not fit shared memory", name);
    const int per_cta = cdiv(g.n_tiles, std::min(g.n_tiles, ctas));
    g.nstages = (int)std::min<size_t>(std::min(std::min(MAX_STAGES, std::max(2, tune("TC5P_ST", MAX_STAGES))), std::max(2, per_cta)),
                                      (budget - fixed) / (2 * (size_t)g.a_bytes));
    smem = fixed + (size_t)g.nstages * 2 * g.a_bytes;
    return 0;
}

// H, W: OUTPUT size (= input size for STRIDE 1; the input of a STRIDE 2 layer is 2H x 2W, stored as parity planes; transposed:
// the INPUT grid, the output is 2H x 2W)
template <int CINP, int NB, int MB, int DIL, int KS, int STRIDE, class Epi>
int launch_mb(const char* name, const Split& in, const Epi& epi, const WSel& wsel, int N, int H, int W, int* err_flag, cudaStream_t st) {
    constexpr int KC = CINP / 8;
    Geo g;
    size_t smem = 0;
    IMVS_TRY((plan_layer<CINP, NB, MB, DIL, KS, STRIDE>(name, wsel, N, H, W, grid_limit(), 220 * 1024, g, smem)));   // ensure_dynamic_smem() opts in to 220 KB
    CUtensorMap mh, ml;
    IMVS_TRY(make_plane_map(&mh, in.hi, STRIDE == 2 ? 4 * N : N, KC, H, W, g.rows));     // stride 2: parity planes are H x W (= the output size)
    IMVS_TRY(make_plane_map(&ml, in.lo, STRIDE == 2 ? 4 * N : N, KC, H, W, g.rows));
    auto kern = tc5p_conv_kernel<CINP, NB, MB, DIL, KS, STRIDE, Epi>;
    static int smem_ok = 0;
    IMVS_TRY(ensure_dynamic_smem(kern, smem, &smem_ok));
    const int grid = std::min(g.n_tiles, grid_limit());
    if (launch_k(kern, dim3(grid), dim3(Shape<NB, MB, STRIDE == 0, KS == 1>::THREADS), smem, st, mh, ml, epi, wsel, g, err_flag) != cudaSuccess)
        return fail("launch of %s failed: %s", name, cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// stride-1 KS x KS (3x3 with dilation DIL, or 1x1) convolution CINP -> NB of a split-plane tensor; the epilogue class decides
// what happens to the accumulators (Epi: bias / residual / ReLU; featurenet.cu:EpiLateral; update.cu: the GRU gates)
template <int CINP, int NB, int DIL = 1, bool ALLOW_MB2 = true, int KS = 3, int STRIDE = 1, class Epi>
int launch(const char* name, const Split& in, const Epi& epi, const WSel& wsel, int N, int H, int W, int* err_flag, cudaStream_t st) {
    IMVS_REQUIRE(in.hi && in.lo, "%s: null tcgen05 operand", name);
    for (int i = 0; i < wsel.nsets; ++i) IMVS_REQUIRE(wsel.w[i], "%s: null tcgen05 weights", name);
    IMVS_REQUIRE((double)N * (CINP / 8) * H * W * 16 * (STRIDE == 2 ? 4 : 1) < 1.8e19 && W >= 1 && H >= 1, "%s: bad shape", name);
    if constexpr (NB <= 32 && ALLOW_MB2 && STRIDE != 0) {
        const int tiles2 = cdiv(W, STRIDE == 2 ? WT - 1 : WT - DIL * (KS - 1)) * cdiv(H, 8) * N;
        const int force = tune("TC5P_MB", 0);
        // 8-row tiles (two M-blocks share one haloed tile: 1.25x instead of 1.5x halo rows) when they still fill the machine
        if (force == 2 || (force == 0 && tiles2 >= 2 * sm_count())) return launch_mb<CINP, NB, 2, DIL, KS, STRIDE, Epi>(name, in, epi, wsel, N, H, W, err_flag, st);
    }
    return launch_mb<CINP, NB, 1, DIL, KS, STRIDE, Epi>(name, in, epi, wsel, N, H, W, err_flag, st);
}
template <int CINP, int NB, int DIL = 1, bool ALLOW_MB2 = true, int KS = 3, int STRIDE = 1, class Epi>
int launch(const char* name, const Split& in, const Epi& epi, const void* w_f16, int N, int H, int W, int* err_flag, cudaStream_t st) {
    return launch<CINP, NB, DIL, ALLOW_MB2, KS, STRIDE, Epi>(name, in, epi, wsel_single(w_f16), N, H, W, err_flag, st);
}

// fp32 NHWC [N][H][W][C] -> split planes (featurenet.cu); C a multiple of 8
int launch_nhwc_to_split(const float* x, const Split& dst, size_t npix_total, int HW, int C, cudaStream_t st);

}  // namespace tc5p

// ---- mma.sync epilogues that WRITE split planes (producers of a tcgen05 layer's operand) ------------------------------
__device__ __forceinline__ void store_split_pair(__half* hi, __half* lo, size_t half_index, float a, float b) {
    uint32_t h, l;
    split_f16(make_float2(a, b), h, l);
    *reinterpret_cast<uint32_t*>(hi + half_index) = h;
    *reinterpret_cast<uint32_t*>(lo + half_index) = l;
}

// EpiSplit2 (featurenet.cu) with both halves written as split planes [N][CO/8][H][W][8]
struct EpiSplit2H {
    tc5p::Split out_relu;    // couts [0, CO)      -> relu(v + bias)
    tc5p::Split out_lin;     // couts [CO, 2*CO)   -> v + bias
    const float* bias;       // [2*CO]
    int H, W, CO;
    template <int NT>
    __device__ __forceinline__ void row(int n, int oy, int ox, int co0, int t, const float (&v)[2 * NT], int) const {
        if (oy >= H || ox >= W) return;
        const size_t plane = (size_t)H * W, pix = (size_t)oy * W + ox;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = co0 + 8 * j + 2 * t;
            float a = v[2 * j] + ldg(bias + co), b = v[2 * j + 1] + ldg(bias + co + 1);
            const bool first = co < CO;
            if (first) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            const int c = first ? co : co - CO;
            const size_t idx = (((size_t)n * (CO / 8) + (c >> 3)) * plane + pix) * 8 + (c & 7);
            const tc5p::Split& o = first ? out_relu : out_lin;
            store_split_pair(o.hi, o.lo, idx, a, b);
        }
    }
};

}  // namespace imvs
