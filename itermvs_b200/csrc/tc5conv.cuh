// tcgen05 (5th-generation tensor core, TMEM accumulator) implicit-GEMM convolution, TF32.
//
// Used for the stride-1 32..64-channel convolutions when conv_passes == 1 (single-pass TF32).
//
//   * a CTA computes ONE UMMA M-block: 128 consecutive "slots" of a haloed tile that is WT slots wide
//     (slot = tile_row * WT + tile_col), for NB output channels: D[128 x NB] fp32 lives in TMEM.
//   * the input tile (TH_o + 2*pad rows, rounded to TF32) sits in shared memory in the UMMA
//     K-major / no-swizzle canonical layout with the 8-row core matrices laid out CONTIGUOUSLY:
//         element (slot, k)  at  (k / 4) * LBO + slot * 16 B + (k % 4) * 4 B ,  SBO = 128 B.
//     Because consecutive slots are 16 bytes apart for every K chunk, a stencil tap (dy, dx) is the
//     SAME shared-memory tile addressed (dy*WT + dx) * 16 bytes further: the A operand of each tap
//     is just a shared-memory descriptor with a shifted start address.  No im2col, the tile is
//     fetched from L2 once.  Output columns >= WT - 2*pad of every tile row are wrap-around garbage
//     and are discarded by the epilogue.
//   * weights are pre-packed on the host in the same canonical layout ([tap][k/4][n][4], TF32) so the
//     CTA stages them with one 1-D bulk copy per tap through the TMA engine (cp.async.bulk + mbarrier).
//   * one elected thread issues KS*KS*CINP/8 tcgen05.mma (kind::tf32, M=128, N=NB, K=8) back to back
//     and commits to an mbarrier; four warps then pull their 32 TMEM lanes with tcgen05.ld and run the
//     fused epilogue (one thread = one pixel = NB contiguous output channels).
#pragma once
#include "common.cuh"
#include "mmaconv.cuh"

namespace imvs {
long long* tc5_clock_buffer();
namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the async proxy (TMA engine, no tensor map), completes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor):
//   [0,14) start >> 4, [16,30) leading byte offset >> 4 (between the two 16-byte K chunks of one MMA),
//   [32,46) stride byte offset >> 4 (between 8-row core matrices), [46,48) version = 1, [61,64) layout = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Geometry {      // host-computed tile geometry
    int WT;            // slots per tile row (power of two: 32 or 64); valid output columns = WT - 2*pad
    int THo;           // output rows per CTA = 128 / WT
    int pad;           // dil * (ks - 1) / 2
    int dil, ks;
    int nslot;         // slots staged >= (THo + 2*pad) * WT + 2*pad, = 2 (mod 8)
};

inline Geometry make_geometry(int ks, int dil, int Wout) {
    Geometry g;
    g.ks = ks; g.dil = dil; g.pad = dil * (ks - 1) / 2;
    // narrower tiles waste fewer halo columns on narrow images, wider ones fewer halo rows
    g.WT = (Wout + 2 * g.pad <= 32 || (Wout % (64 - 2 * g.pad) != 0 && Wout % (32 - 2 * g.pad) == 0)) ? 32 : 64;
    g.THo = 128 / g.WT;
    g.nslot = ((g.THo + 2 * g.pad) * g.WT + 2 * g.pad + 7) / 8 * 8 + 2;      // = 2 (mod 8): conflict-free staging stores
    return g;
}

// error flag (device int): 1 = an mbarrier wait timed out (should never happen; keeps a broken build from hanging the GPU)
// Epi::part<NC>(n, oy, ox, c0, v): output channels [c0, c0+NC) of pixel (oy, ox), in range, one thread.
constexpr int TC5_THREADS = 256;      // 8 warps: all stage the tile; warp w reads TMEM lanes 32*(w&3).. and columns half (w>>2)

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it) done = mbar_try_wait(bar, parity);
    return done != 0;
}

template <int CINP, int NB, class In, class Epi>
__global__ void __launch_bounds__(TC5_THREADS)
tc5_conv_kernel(const In in, const Epi epi, const float* __restrict__ w_umma, const Geometry geo, int Hout, int Wout, int* err_flag,
                long long* clk) {
    static_assert(CINP % 16 == 0 && NB % 32 == 0 && NB <= 256, "UMMA shape / staging pattern");
    constexpr int KC = CINP / 4;                                 // 16-byte K chunks
    constexpr int KG = KC / 4;                                   // chunk groups of 4 (one per lane & 3)
    constexpr int TMEM_COLS = NB <= 32 ? 32 : (NB <= 64 ? 64 : (NB <= 128 ? 128 : 256));
    constexpr int NC = NB / 2;                                   // epilogue columns per thread
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const bool stamp = clk != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
    if (stamp) clk[0] = clock64();
    const int nslot = geo.nslot, WT = geo.WT, pad = geo.pad, dil = geo.dil, ks = geo.ks;
    const int ntaps = ks * ks;
    float* sA = reinterpret_cast<float*>(smem_raw);              // [KC][nslot][4]
    float* sB = sA + (size_t)KC * nslot * 4;                     // [ntaps][KC][NB][4]
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sB + (size_t)ntaps * KC * NB * 4);   // [0] weights landed, [1] MMAs done
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z;
    const int wvalid = WT - 2 * pad;
    const int oy0 = blockIdx.y * geo.THo, ox0 = blockIdx.x * wvalid;
    const uint32_t bar_w = smem_u32(sBar), bar_d = smem_u32(sBar + 1);

    if (warp == 0) tmem_alloc(smem_u32(sTmem), TMEM_COLS);
    if (tid == 32) {
        mbar_init(bar_w, 1);
        mbar_init(bar_d, 1);
        fence_mbar_init();
        // weights (already TF32, canonical order): one bulk copy per tap through the TMA engine
        constexpr uint32_t tap_bytes = KC * NB * 16;
        mbar_expect_tx(bar_w, tap_bytes * (uint32_t)ntaps);
        for (int tap = 0; tap < ntaps; ++tap)
            bulk_g2s(smem_u32(sB) + tap * tap_bytes, w_umma + (size_t)tap * (tap_bytes / 4), tap_bytes, bar_w);
    }
    pdl_trigger();
    pdl_wait();         // TMEM allocation and the weight copies above overlap the predecessor's tail
    // ---- stage the haloed input tile, rounding to TF32 (round-to-nearest) on the way:
    //      slot s -> pixel (oy0 - pad + s / WT, ox0 - pad + s % WT).  A quarter-warp covers 2 slots x 4 chunks;
    //      nslot = 2 (mod 8) makes the 16-byte stores of a quarter-warp hit 8 distinct bank groups.
    {
        // all global loads of a thread are issued before the first conversion / store (<= 7 x KG float4 in flight)
        constexpr int MAXIT = 7;                                  // ceil(nslot_max / 64), nslot_max = 394 + padding
        const int kcl = lane & 3, sl = lane >> 2;
        float4 v[MAXIT][KG];
        const auto img = in.image(n);
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            const int s = (it * (TC5_THREADS / 32) + warp) * 8 + sl;
            const int iy = oy0 - pad + s / WT, ix = ox0 - pad + (s & (WT - 1));
#pragma unroll
            for (int kg = 0; kg < KG; ++kg) {
                bool valid;
                const float* src = img.ptr4(iy, ix, kg * 4 + kcl, valid);
                v[it][kg] = (valid && s < nslot) ? ldg4(src) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            const int s = (it * (TC5_THREADS / 32) + warp) * 8 + sl;
            if (s < nslot) {
#pragma unroll
                for (int kg = 0; kg < KG; ++kg) {
                    uint4 r = make_uint4(f2tf32(v[it][kg].x), f2tf32(v[it][kg].y), f2tf32(v[it][kg].z), f2tf32(v[it][kg].w));
                    *reinterpret_cast<uint4*>(sA + ((size_t)(kg * 4 + kcl) * nslot + s) * 4) = r;
                }
            }
        }
    }
    if (stamp) clk[1] = clock64();
    fence_async_shared();                 // generic-proxy writes -> visible to the tensor core's async proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = *sTmem;
    if (stamp) clk[3] = clock64();

    // ---- MMAs: one thread issues everything, then commits to the mbarrier
    bool ok_w = true;
    if (tid == 0) {
        ok_w = mbar_wait_bounded(bar_w, 0);                       // weights landed (async proxy)
        if (stamp) clk[4] = clock64();
        constexpr uint32_t idesc = make_idesc_tf32(NB);
        const uint32_t lbo_a = (uint32_t)nslot * 16u, lbo_b = (uint32_t)NB * 16u;
        const uint64_t da0 = make_desc(smem_u32(sA), lbo_a, 128u), db0 = make_desc(smem_u32(sB), lbo_b, 128u);
        const uint32_t ka = (2u * lbo_a) >> 4, kb = (2u * lbo_b) >> 4;            // per K-step (8 elements) advance, in 16 B units
        uint32_t acc = 0;
        if (ok_w) {
            for (int tap = 0; tap < ntaps; ++tap) {
                const uint32_t shift = (uint32_t)((tap / ks) * dil * WT + (tap % ks) * dil);       // slots == 16 B units
                uint64_t da = da0 + shift, db = db0 + (uint32_t)(tap * KC * NB);
#pragma unroll
                for (int k8 = 0; k8 < CINP / 8; ++k8) {
                    umma_tf32(tmem_d, da, db, idesc, acc);
                    acc = 1;
                    da += ka; db += kb;
                }
            }
        }
        umma_commit(bar_d);
        if (stamp) clk[5] = clock64();
    }
    // ---- epilogue: warp w owns TMEM lanes 32*(w&3).. (= slots) and the column half (w>>2); thread = one pixel
    {
        const bool done = mbar_wait_bounded(bar_d, 0);
        fence_after_sync();
        if (stamp) clk[6] = clock64();
        if (!done) {
            if (lane == 0 && err_flag) atomicExch(err_flag, 1);
        } else {
            const int lg = warp & 3, half = warp >> 2;
            const int m = lg * 32 + lane;
            const int oy = oy0 + m / WT, oxl = m & (WT - 1), ox = ox0 + oxl;
            const bool ok = (oxl < wvalid) && (oy < Hout) && (ox < Wout);
            float v[NC];
#pragma unroll
            for (int c = 0; c < NC; c += 16) {
                float t16[16];
                tmem_ld16(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(half * NC + c), t16);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[c + i] = t16[i];
            }
            if (ok) epi.template part<NC>(n, oy, ox, half * NC, v);
        }
    }
    if (stamp) clk[7] = clock64();
    if (tid == 0 && !ok_w && err_flag) atomicExch(err_flag, 1);
    fence_before_sync();
    __syncthreads();
    if (stamp) clk[8] = clock64();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

template <int CINP, int NB>
inline size_t smem_bytes(const Geometry& g) {
    return sizeof(float) * ((size_t)(CINP / 4) * g.nslot * 4 + (size_t)g.ks * g.ks * (CINP / 4) * NB * 4) + 64;
}

template <int CINP, int NB, class In, class Epi>
int launch(const char* name, const In& in, const Epi& epi, const float* w_umma, int ks, int dil, int N, int Hout, int Wout,
           int* err_flag, cudaStream_t st) {
    IMVS_REQUIRE(w_umma, "%s: null tcgen05 weights", name);
    const Geometry g = make_geometry(ks, dil, Wout);
    const size_t smem = smem_bytes<CINP, NB>(g);
    IMVS_REQUIRE(smem <= 220 * 1024, "%s: %zu bytes of shared memory needed", name, smem);
    IMVS_REQUIRE(g.nslot <= 7 * 64, "%s: tile of %d slots exceeds the staging pattern", name, g.nslot);
    auto kern = tc5_conv_kernel<CINP, NB, In, Epi>;
    static int smem_ok = 0;
    IMVS_TRY(ensure_dynamic_smem(kern, smem, &smem_ok));
    dim3 grid(cdiv(Wout, g.WT - 2 * g.pad), cdiv(Hout, g.THo), N);
    IMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "%s: grid too large", name);
    if (launch_k(kern, grid, dim3(TC5_THREADS), smem, st, in, epi, w_umma, g, Hout, Wout, err_flag, tc5_clock_buffer()) != cudaSuccess)
        return fail("launch of %s failed: %s", name, cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// ==================================================================================================
// fp32-grade variant (the DEFAULT precision mode 4): the same slot-shifted implicit GEMM with the operands split
// x = hi + lo in fp16 (hi = fp16(x), lo = fp16(x - hi)) and three tcgen05.mma.kind::f16 chains per tap into the same TMEM
// accumulator -- hi*hi + hi*lo + lo*hi, fp32 accumulation: 22 significant bits, the products of the mma.sync engine's
// mode 4.  hi + lo of a value take the 4 bytes its TF32 copy took, so the shared-memory budget is unchanged; K per MMA
// is 16 instead of 8, so the three products cost 1.5x the TF32 MMA count -- on a tensor pipe that idles either way
// (these layers are staging / epilogue bound).
//   tile  : [hi | lo] x [CINP/8 chunks][nslot][8 halves]   element (slot, k) at (k/8)*LBO + slot*16 B + (k%8)*2 B
//   weights (host-packed, _pack.py:pack_umma_f16): [cout block][tap][hi | lo][CINP/8][NB][8 halves]
// ==================================================================================================
__device__ __forceinline__ void umma_f16k(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_f16k(int n) {       // D = F32, A = B = F16, K-major, M = 128
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

inline Geometry make_geometry_h(int ks, int dil, int Wout, int kc) {
    Geometry g = make_geometry(ks, dil, Wout);
    // tile width by staged slots per output row: ceil(W / valid) * WT * (THo + 2 pad) / THo
    {
        const int p2 = 2 * g.pad;
        auto cost = [&](int wt) { const int tho = 128 / wt; return (double)((Wout + wt - p2 - 1) / (wt - p2)) * wt * (tho + p2) / tho; };
        g.WT = (p2 < 32 && cost(32) < cost(64)) ? 32 : 64;
        g.THo = 128 / g.WT;
    }
    // staging stores of a quarter-warp (8 lanes = (8 / min(kc,4)) slots x min(kc,4) chunks) hit 8 distinct 16-byte bank
    // groups when nslot = 2 (mod 8) for >= 4 chunks, 4 (mod 8) for 2 chunks
    const int want = kc >= 4 ? 2 : 4;
    int n = (g.THo + 2 * g.pad) * g.WT + 2 * g.pad;
    while (n % 8 != want) ++n;
    g.nslot = n;
    return g;
}

template <int CINP, int NB, class In, class Epi>
__global__ void __launch_bounds__(TC5_THREADS, 5)
tc5h_conv_kernel(const In in, const Epi epi, const void* __restrict__ w_f16, const Geometry geo, int Hout, int Wout, int* err_flag) {
    static_assert(CINP % 16 == 0 && NB % 16 == 0 && NB <= 256, "UMMA kind::f16 shape");
    constexpr int KC = CINP / 8;                                 // 16-byte K chunks (8 halves)
    constexpr int CPL = KC >= 4 ? 4 : 2;                         // chunks handled by consecutive lanes
    constexpr int SPW = 32 / CPL;                                // slots per warp and staging iteration
    constexpr int TMEM_COLS = NB <= 32 ? 32 : (NB <= 64 ? 64 : (NB <= 128 ? 128 : 256));
    constexpr int NC = NB / 2;                                   // epilogue columns per thread
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nslot = geo.nslot, WT = geo.WT, pad = geo.pad, dil = geo.dil, ks = geo.ks;
    const int wsh = WT == 64 ? 6 : 5;                            // WT is 32 or 64
    const int ntaps = ks * ks;
    const uint32_t a_bytes = (uint32_t)KC * nslot * 16;          // one of hi / lo
    constexpr uint32_t b_tap_bytes = KC * NB * 16;               // one of hi / lo, one tap
    unsigned char* sA = smem_raw;                                // hi | lo
    unsigned char* sB = sA + 2 * a_bytes;                        // [tap][hi | lo][KC][NB][8]
    uint64_t* sBar = reinterpret_cast<uint64_t*>(sB + (size_t)ntaps * 2 * b_tap_bytes);
    uint32_t* sTmem = reinterpret_cast<uint32_t*>(sBar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z;
    const int wvalid = WT - 2 * pad;
    const int oy0 = blockIdx.y * geo.THo, ox0 = blockIdx.x * wvalid;
    const uint32_t bar_w = smem_u32(sBar), bar_d = smem_u32(sBar + 1);

    if (warp == 0) tmem_alloc(smem_u32(sTmem), TMEM_COLS);
    if (tid == 32) {
        mbar_init(bar_w, 1);
        mbar_init(bar_d, 1);
        fence_mbar_init();
        const uint32_t total = (uint32_t)ntaps * 2 * b_tap_bytes;
        mbar_expect_tx(bar_w, total);
        // one bulk copy per tap (hi and lo of a tap are adjacent): <= 2 * KC * NB * 16 bytes each
        for (int tap = 0; tap < ntaps; ++tap)
            bulk_g2s(smem_u32(sB) + tap * 2 * b_tap_bytes, static_cast<const unsigned char*>(w_f16) + (size_t)tap * 2 * b_tap_bytes,
                     2 * b_tap_bytes, bar_w);
    }
    pdl_trigger();
    pdl_wait();
    // ---- stage the haloed input tile, splitting into fp16 hi / lo on the way: slot s -> pixel (oy0 - pad + s / WT,
    //      ox0 - pad + s % WT).  lane -> (slot sl = lane / CPL, chunk kcl = lane % CPL [+ CPL * g]): a quarter-warp reads
    //      128 contiguous bytes per pixel; UI iterations (32 registers of loads) are in flight before the first conversion.
    {
        constexpr int G = (KC + CPL - 1) / CPL;                  // chunk groups per slot
        constexpr int UI = CPL == 2 ? 2 : (4 / G > 0 ? 4 / G : 1);   // iterations in flight: 256 (G = 1) or 128 slots per round
        constexpr int PER_IT = (TC5_THREADS / 32) * SPW;         // slots per iteration of the whole CTA
        const int kcl = lane % CPL, sl = lane / CPL;
        const auto img = in.image(n);
#pragma unroll 1
        for (int s0 = 0; s0 < nslot; s0 += UI * PER_IT) {
            float4 v[UI][G][2];
#pragma unroll
            for (int u = 0; u < UI; ++u) {
                const int s = s0 + u * PER_IT + warp * SPW + sl;
                const int iy = oy0 - pad + (s >> wsh), ix = ox0 - pad + (s & (WT - 1));
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int kc = g * CPL + kcl;
                    bool valid;
                    const float* src = img.ptr4(iy, ix, 2 * (kc < KC ? kc : 0), valid);
                    const bool ok = valid && s < nslot && kc < KC;
                    v[u][g][0] = ok ? ldg4(src) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[u][g][1] = ok ? ldg4(src + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < UI; ++u) {
                const int s = s0 + u * PER_IT + warp * SPW + sl;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int kc = g * CPL + kcl;
                    if (s < nslot && kc < KC) {
                        uint4 hi, lo;
                        split_f16(make_float2(v[u][g][0].x, v[u][g][0].y), hi.x, lo.x);
                        split_f16(make_float2(v[u][g][0].z, v[u][g][0].w), hi.y, lo.y);
                        split_f16(make_float2(v[u][g][1].x, v[u][g][1].y), hi.z, lo.z);
                        split_f16(make_float2(v[u][g][1].z, v[u][g][1].w), hi.w, lo.w);
                        *reinterpret_cast<uint4*>(sA + ((size_t)kc * nslot + s) * 16) = hi;
                        *reinterpret_cast<uint4*>(sA + a_bytes + ((size_t)kc * nslot + s) * 16) = lo;
                    }
                }
            }
        }
    }
    fence_async_shared();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = *sTmem;

    bool ok_w = true;
    if (tid == 0) {
        ok_w = mbar_wait_bounded(bar_w, 0);
        constexpr uint32_t idesc = make_idesc_f16k(NB);
        const uint32_t lbo_a = (uint32_t)nslot * 16u, lbo_b = (uint32_t)NB * 16u;
        const uint64_t da_hi = make_desc(smem_u32(sA), lbo_a, 128u), da_lo = make_desc(smem_u32(sA + a_bytes), lbo_a, 128u);
        const uint32_t ka = (2u * lbo_a) >> 4, kb = (2u * lbo_b) >> 4;            // per K-step (16 elements) advance, 16 B units
        uint32_t acc = 0;
        if (ok_w) {
            for (int tap = 0; tap < ntaps; ++tap) {
                const uint32_t shift = (uint32_t)((tap / ks) * dil * WT + (tap % ks) * dil);       // slots == 16 B units
                const uint64_t db_hi = make_desc(smem_u32(sB) + tap * 2 * b_tap_bytes, lbo_b, 128u);
                const uint64_t db_lo = make_desc(smem_u32(sB) + tap * 2 * b_tap_bytes + b_tap_bytes, lbo_b, 128u);
#pragma unroll
                for (int prod = 0; prod < 3; ++prod) {
                    uint64_t da = (prod == 2 ? da_lo : da_hi) + shift, db = prod == 1 ? db_lo : db_hi;
#pragma unroll
                    for (int k16 = 0; k16 < CINP / 16; ++k16) {
                        umma_f16k(tmem_d, da, db, idesc, acc);
                        acc = 1;
                        da += ka; db += kb;
                    }
                }
            }
        }
        umma_commit(bar_d);
    }
    // ---- epilogue: warp w owns TMEM lanes 32*(w&3).. (= slots) and the column half (w>>2); thread = one pixel.  What the
    //      epilogue reads from global memory besides the accumulator (residual) is requested BEFORE the wait on the MMAs.
    {
        const int lg = warp & 3, half = warp >> 2;
        const int m = lg * 32 + lane;
        const int oy = oy0 + (m >> wsh), oxl = m & (WT - 1), ox = ox0 + oxl;
        const bool ok = (oxl < wvalid) && (oy < Hout) && (ox < Wout);
        typename Epi::template Pre<NC> pre;
        if (ok) epi.template prefetch<NC>(n, oy, ox, half * NC, pre);
        const bool done = mbar_wait_bounded(bar_d, 0);
        fence_after_sync();
        if (!done) {
            if (lane == 0 && err_flag) atomicExch(err_flag, 1);
        } else {
            float v[NC];
            if constexpr (NC % 16 == 0) {
#pragma unroll
                for (int c = 0; c < NC; c += 16) {
                    float t16[16];
                    tmem_ld16(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(half * NC + c), t16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[c + i] = t16[i];
                }
            } else {
                static_assert(NC % 8 == 0, "epilogue column split");
#pragma unroll
                for (int c = 0; c < NC; c += 8) {
                    uint32_t r[8];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                                 : "r"(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(half * NC + c)));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[c + i] = __uint_as_float(r[i]);
                }
            }
            if (ok) {
                float amax = 0.f;
#pragma unroll
                for (int c = 0; c < NC; ++c) amax = fmaxf(amax, fabsf(v[c]));
                if (!(amax <= 65504.f) && err_flag) atomicOr(err_flag, 2);       // fp16 range guard, see imvs_device_status
                epi.template part_pre<NC>(n, oy, ox, half * NC, v, pre);
            }
        }
    }
    if (tid == 0 && !ok_w && err_flag) atomicExch(err_flag, 1);
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

template <int CINP, int NB>
inline size_t smem_bytes_h(const Geometry& g) {
    return (size_t)2 * (CINP / 8) * g.nslot * 16 + (size_t)g.ks * g.ks * 2 * (CINP / 8) * NB * 16 + 64;
}

// w_f16: packed weights of ALL cout blocks ([cout block][tap][hi|lo][KC][NB][8 halves]); blockIdx.z = n (one cout block per launch
// when cout_total == NB; otherwise launch once per block with the pointer advanced -- the FeatureNet / estimator layers served
// here have cout_total == NB)
template <int CINP, int NB, class In, class Epi>
int launch_h(const char* name, const In& in, const Epi& epi, const void* w_f16, int ks, int dil, int N, int Hout, int Wout,
             int* err_flag, cudaStream_t st) {
    IMVS_REQUIRE(w_f16, "%s: null tcgen05 fp16 weights", name);
    const Geometry g = make_geometry_h(ks, dil, Wout, CINP / 8);
    const size_t smem = smem_bytes_h<CINP, NB>(g);
    IMVS_REQUIRE(smem <= 220 * 1024, "%s: %zu bytes of shared memory needed", name, smem);
    auto kern = tc5h_conv_kernel<CINP, NB, In, Epi>;
    static int smem_ok = 0;
    IMVS_TRY(ensure_dynamic_smem(kern, smem, &smem_ok));
    dim3 grid(cdiv(Wout, g.WT - 2 * g.pad), cdiv(Hout, g.THo), N);
    IMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "%s: grid too large", name);
    if (launch_k(kern, grid, dim3(TC5_THREADS), smem, st, in, epi, w_f16, g, Hout, Wout, err_flag) != cudaSuccess)
        return fail("launch of %s failed: %s", name, cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// ---- epilogues (one thread = one pixel, NC contiguous channels starting at c0) ---------------------
// TF32 mode: gates use the fast exponential (ex2.approx; relative error ~1e-6, far below TF32's 5e-4)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

struct PixNHWC {          // out[n][oy][ox][c0..c0+NC) = (v + bias) (+ residual) (relu)
    float* out;
    const float* bias;
    const float* residual;
    int H, W, C, relu;
    template <int NC> struct Pre { float4 r[NC / 4]; };
    template <int NC>
    __device__ __forceinline__ void prefetch(int n, int oy, int ox, int c0, Pre<NC>& p) const {
        if (!residual) return;
        const size_t base = (((size_t)n * H + oy) * W + ox) * C + c0;
#pragma unroll
        for (int c = 0; c < NC; c += 4) p.r[c / 4] = ldg4(residual + base + c);
    }
    template <int NC>
    __device__ __forceinline__ void part_pre(int n, int oy, int ox, int c0, const float (&v)[NC], const Pre<NC>& p) const {
        const size_t base = (((size_t)n * H + oy) * W + ox) * C + c0;
#pragma unroll
        for (int c = 0; c < NC; c += 4) {
            float4 o = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            if (bias) { const float4 b = ldg4(bias + c0 + c); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
            if (residual) { const float4 r = p.r[c / 4]; o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(out + base + c) = o;
        }
    }
    template <int NC>
    __device__ __forceinline__ void part(int n, int oy, int ox, int c0, const float (&v)[NC]) const {
        const size_t base = (((size_t)n * H + oy) * W + ox) * C + c0;
#pragma unroll
        for (int c = 0; c < NC; c += 4) {
            float4 o = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            if (bias) { const float4 b = ldg4(bias + c0 + c); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
            if (residual) { const float4 r = ldg4(residual + base + c); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(out + base + c) = o;
        }
    }
};

struct PixGruZR {         // 64 stacked channels: c0 = 0 -> z = sigmoid (32), c0 = 32 -> r = sigmoid, store r*h   (module.py:61-62)
    const float* bias;
    const float* h;
    float* z;
    float* rh;
    int H, W;
    int precise = 0;      // 1: expf-based sigmoid (the fp32-grade mode); 0: ex2.approx (TF32 mode)
    __device__ __forceinline__ float sg(float x) const { return precise ? sigmoidf_(x) : fast_sigmoid(x); }
    template <int NC>
    __device__ __forceinline__ void part(int n, int oy, int ox, int c0, const float (&v)[NC]) const {
        static_assert(NC == 32, "z | r halves");
        const size_t base = (((size_t)n * H + oy) * W + ox) * 32;
        if (c0 == 0) {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 b = ldg4(bias + c);
                *reinterpret_cast<float4*>(z + base + c) = make_float4(sg(v[c] + b.x), sg(v[c + 1] + b.y), sg(v[c + 2] + b.z), sg(v[c + 3] + b.w));
            }
        } else {
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 b = ldg4(bias + 32 + c), hh = ldg4(h + base + c);
                *reinterpret_cast<float4*>(rh + base + c) = make_float4(sg(v[c] + b.x) * hh.x, sg(v[c + 1] + b.y) * hh.y,
                                                                         sg(v[c + 2] + b.z) * hh.z, sg(v[c + 3] + b.w) * hh.w);
            }
        }
    }
    template <int NC> struct Pre {};
    template <int NC> __device__ __forceinline__ void prefetch(int, int, int, int, Pre<NC>&) const {}
    template <int NC>
    __device__ __forceinline__ void part_pre(int n, int oy, int ox, int c0, const float (&v)[NC], const Pre<NC>&) const { part<NC>(n, oy, ox, c0, v); }
};

struct PixGruQ {          // 32 channels in two halves: q = tanh, h <- (1-z) h + z q in place   (module.py:63-64)
    const float* bias;
    const float* z;
    float* h;
    int H, W;
    int precise = 0;      // 1: tanhf (the fp32-grade mode); 0: ex2.approx based (TF32 mode)
    __device__ __forceinline__ float th(float x) const { return precise ? tanhf(x) : fast_tanh(x); }
    template <int NC>
    __device__ __forceinline__ void part(int n, int oy, int ox, int c0, const float (&v)[NC]) const {
        const size_t base = (((size_t)n * H + oy) * W + ox) * 32 + c0;
#pragma unroll
        for (int c = 0; c < NC; c += 4) {
            const float4 b = ldg4(bias + c0 + c), zz = ldg4(z + base + c);
            float4 hh = *reinterpret_cast<const float4*>(h + base + c);
            hh.x = (1.f - zz.x) * hh.x + zz.x * th(v[c] + b.x);
            hh.y = (1.f - zz.y) * hh.y + zz.y * th(v[c + 1] + b.y);
            hh.z = (1.f - zz.z) * hh.z + zz.z * th(v[c + 2] + b.z);
            hh.w = (1.f - zz.w) * hh.w + zz.w * th(v[c + 3] + b.w);
            *reinterpret_cast<float4*>(h + base + c) = hh;
        }
    }
    template <int NC> struct Pre {};
    template <int NC> __device__ __forceinline__ void prefetch(int, int, int, int, Pre<NC>&) const {}
    template <int NC>
    __device__ __forceinline__ void part_pre(int n, int oy, int ox, int c0, const float (&v)[NC], const Pre<NC>&) const { part<NC>(n, oy, ox, c0, v); }
};

}  // namespace tc5

int tc5_enabled();        // imvs_set_tcgen05(): use the tcgen05 kernels when conv_passes == 1 (default on)
long long* tc5_clock_buffer();   // non-null only while imvs_tc5_debug_clocks(1): CTA 0 stamps its phases
int* tc5_error_flag();    // process-wide device int (lazily allocated outside graph capture), or nullptr

}  // namespace imvs
