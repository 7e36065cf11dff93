// cusim runtime: one ucontext fiber per CUDA thread, blocks executed one after the other.
// TEST INFRASTRUCTURE ONLY -- see the header comment of tests/cusim/cuda_runtime.h.
#include "cuda_runtime.h"

#include <ucontext.h>

#include <vector>

uint3 threadIdx;
dim3 blockIdx, blockDim, gridDim;

namespace cusim {

Thread* cur = nullptr;

namespace {

constexpr size_t STACK_BYTES = 96 * 1024;
constexpr size_t DYN_SMEM_BYTES = 228 * 1024;

struct Fiber {
    Thread t;
    ucontext_t ctx;
    bool done = false;
    const unsigned long long* wait_gen = nullptr;      // parked on a barrier: runnable again once *wait_gen != wait_val
    unsigned long long wait_val = 0;
};
struct WarpState {
    unsigned long long gen = 0;
    int arrived = 0, live = 0;
    uint32_t slot[32];
    uint32_t xchg[32][12];
};
struct BlockState {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    unsigned long long gen = 0;
    int arrived = 0, live = 0;
};

BlockState* g_blk = nullptr;
Fiber* g_fiber = nullptr;
ucontext_t g_main;
const std::function<void()>* g_body = nullptr;
alignas(128) char g_dyn_smem[DYN_SMEM_BYTES];
std::vector<char> g_stacks;
unsigned long long g_events = 0;       // barrier releases + thread exits: the scheduler's notion of progress

void yield_() { swapcontext(&g_fiber->ctx, &g_main); }
void park_until_changed(const unsigned long long& gen, unsigned long long seen) {
    Fiber* f = g_fiber;
    f->wait_gen = &gen;
    f->wait_val = seen;
    while (gen == seen) yield_();                  // the scheduler does not switch to a parked fiber before the release
    f->wait_gen = nullptr;
}

void release_warp_if_complete(WarpState& w) {
    if (w.live > 0 && w.arrived == w.live) { w.arrived = 0; ++w.gen; ++g_events; }
}
void release_block_if_complete(BlockState& b) {
    if (b.live > 0 && b.arrived == b.live) { b.arrived = 0; ++b.gen; ++g_events; }
}

void trampoline() {
    (*g_body)();
    Fiber* f = g_fiber;
    f->done = true;
    // an exited thread no longer takes part in barriers (what the hardware does for exited warps / lanes)
    WarpState& w = g_blk->warps[f->t.warp];
    --w.live;
    release_warp_if_complete(w);
    --g_blk->live;
    release_block_if_complete(*g_blk);
    swapcontext(&f->ctx, &g_main);
}

}  // namespace

void die(const char* what) {
    fprintf(stderr, "cusim: %s (block %u,%u,%u thread %u,%u,%u)\n", what, blockIdx.x, blockIdx.y, blockIdx.z,
            cur ? cur->tid.x : 0, cur ? cur->tid.y : 0, cur ? cur->tid.z : 0);
    abort();
}

void* dyn_smem() { return g_dyn_smem; }
uint32_t (*warp_xchg())[12] { return g_blk->warps[cur->warp].xchg; }

void warp_barrier() {
    WarpState& w = g_blk->warps[cur->warp];
    const unsigned long long g = w.gen;
    ++w.arrived;
    release_warp_if_complete(w);
    if (w.gen == g) park_until_changed(w.gen, g);
}

void block_barrier() {
    BlockState& b = *g_blk;
    const unsigned long long g = b.gen;
    ++b.arrived;
    release_block_if_complete(b);
    if (b.gen == g) park_until_changed(b.gen, g);
}

uint32_t shfl(uint32_t v, int src_lane) {
    WarpState& w = g_blk->warps[cur->warp];
    w.slot[cur->lane] = v;
    warp_barrier();
    const uint32_t r = w.slot[src_lane & 31];
    warp_barrier();
    return r;
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > 1024) die("bad block size");
    if (smem > DYN_SMEM_BYTES) die("dynamic shared memory request exceeds 228 KB");
    if (g_blk) die("nested launch");
    const int nwarps = (nthreads + 31) / 32;
    g_stacks.resize((size_t)nthreads * STACK_BYTES);
    BlockState blk;
    g_blk = &blk;
    g_body = &body;
    gridDim = grid;
    blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = dim3(bx, by, bz);
                // poison the dynamic shared memory: a kernel must not rely on what a previous block left there
                std::memset(g_dyn_smem, 0xA5, smem);
                blk.fibers.assign(nthreads, Fiber());
                blk.warps.assign(nwarps, WarpState());
                blk.gen = 0; blk.arrived = 0; blk.live = nthreads;
                for (int i = 0; i < nthreads; ++i) {
                    Fiber& f = blk.fibers[i];
                    f.t.tid.x = i % block.x;
                    f.t.tid.y = (i / block.x) % block.y;
                    f.t.tid.z = i / (block.x * block.y);
                    f.t.warp = i / 32;
                    f.t.lane = i % 32;
                    blk.warps[f.t.warp].live++;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = g_stacks.data() + (size_t)i * STACK_BYTES;
                    f.ctx.uc_stack.ss_size = STACK_BYTES;
                    f.ctx.uc_link = &g_main;
                    makecontext(&f.ctx, trampoline, 0);
                }
                int remaining = nthreads;
                while (remaining > 0) {
                    const unsigned long long before = g_events;
                    for (int i = 0; i < nthreads; ++i) {
                        Fiber& f = blk.fibers[i];
                        if (f.done) continue;
                        if (f.wait_gen && *f.wait_gen == f.wait_val) continue;      // still parked
                        g_fiber = &f;
                        cur = &f.t;
                        threadIdx = f.t.tid;
                        swapcontext(&g_main, &f.ctx);
                        if (f.done) { --remaining; ++g_events; }
                    }
                    // every live fiber is parked on a barrier that cannot complete (divergent barrier, lost lane)
                    if (remaining > 0 && g_events == before) die("deadlock: no fiber makes progress");
                }
            }
    cur = nullptr;
    g_fiber = nullptr;
    g_blk = nullptr;
    g_body = nullptr;
}

}  // namespace cusim
