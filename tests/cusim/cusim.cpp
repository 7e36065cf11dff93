// cusim runtime: one ucontext fiber per CUDA thread, blocks executed one after the other.
// TEST INFRASTRUCTURE ONLY -- see the header comment of tests/cusim/cuda_runtime.h.
#include "cuda_runtime.h"

#include <ucontext.h>

#include <algorithm>
#include <map>
#include <string>
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

// ---- access-pattern tracer ---------------------------------------------------------------------------------------
bool g_trace = false;
namespace {
struct LoadRec { uint64_t site; uint32_t k; uint16_t warp; uint16_t bytes; uint64_t addr; };
struct SiteStats { uint64_t requests = 0, lanes = 0, bytes = 0, sectors = 0, lines = 0, wavefronts = 0; };
std::vector<LoadRec> g_recs;                                   // of the block being executed
std::vector<std::map<uint64_t, uint32_t>> g_site_count;        // per thread: executions of each load site so far
std::map<uint64_t, SiteStats> g_sites;                         // per launch
std::string g_trace_path;
int g_launch_index = 0;

void trace_flush_block() {
    std::sort(g_recs.begin(), g_recs.end(), [](const LoadRec& a, const LoadRec& b) {
        return a.warp != b.warp ? a.warp < b.warp : (a.site != b.site ? a.site < b.site : a.k < b.k);
    });
    size_t i = 0;
    std::vector<uint64_t> sectors, lines;
    while (i < g_recs.size()) {
        size_t j = i;
        sectors.clear(); lines.clear();
        uint64_t bytes = 0;
        while (j < g_recs.size() && g_recs[j].warp == g_recs[i].warp && g_recs[j].site == g_recs[i].site && g_recs[j].k == g_recs[i].k) {
            for (uint64_t a = g_recs[j].addr >> 5; a <= (g_recs[j].addr + g_recs[j].bytes - 1) >> 5; ++a) sectors.push_back(a);
            bytes += g_recs[j].bytes;
            ++j;
        }
        std::sort(sectors.begin(), sectors.end());
        sectors.erase(std::unique(sectors.begin(), sectors.end()), sectors.end());
        for (uint64_t s : sectors) lines.push_back(s >> 2);
        lines.erase(std::unique(lines.begin(), lines.end()), lines.end());
        SiteStats& st = g_sites[g_recs[i].site];
        st.requests += 1; st.lanes += j - i; st.bytes += bytes; st.sectors += sectors.size(); st.lines += lines.size();
        // data-stage model: a wavefront returns at most 128 bytes of register data and touches one 128-byte line
        st.wavefronts += std::max<uint64_t>(lines.size(), (bytes + 127) / 128);
        i = j;
    }
    g_recs.clear();
}

void trace_end_launch(dim3 grid, dim3 block) {
    FILE* f = fopen(g_trace_path.c_str(), "a");
    if (!f) return;
    fprintf(f, "{\"launch\": %d, \"grid\": [%u, %u, %u], \"block\": [%u, %u, %u], \"sites\": [", g_launch_index++, grid.x, grid.y, grid.z,
            block.x, block.y, block.z);
    bool first = true;
    for (auto& kv : g_sites) {
        const SiteStats& s = kv.second;
        fprintf(f, "%s{\"site\": \"%llx\", \"requests\": %llu, \"lanes\": %llu, \"bytes\": %llu, \"sectors\": %llu, \"lines\": %llu, \"wavefronts\": %llu}",
                first ? "" : ", ", (unsigned long long)kv.first, (unsigned long long)s.requests, (unsigned long long)s.lanes,
                (unsigned long long)s.bytes, (unsigned long long)s.sectors, (unsigned long long)s.lines, (unsigned long long)s.wavefronts);
        first = false;
    }
    fprintf(f, "]}\n");
    fclose(f);
    g_sites.clear();
}
}  // namespace

void trace_load(const void* p, int bytes, const void* site) {
    const int tid = (int)(g_fiber - g_blk->fibers.data());
    const uint32_t k = g_site_count[tid][(uint64_t)site]++;
    g_recs.push_back(LoadRec{(uint64_t)site, k, (uint16_t)cur->warp, (uint16_t)bytes, (uint64_t)p});
}

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
    const char* tr = getenv("CUSIM_TRACE");
    g_trace = tr && *tr;
    if (g_trace) g_trace_path = tr;
    // CUSIM_SHUFFLE=<seed>: run the fibers of a block in a random order that changes every scheduling round -- a missing
    // barrier between a producer and a consumer then shows up as a result that depends on the seed
    const char* sh = getenv("CUSIM_SHUFFLE");
    const bool shuffle = sh && *sh;
    unsigned long long rng = shuffle ? strtoull(sh, nullptr, 10) * 0x9E3779B97F4A7C15ULL + 1 : 0;
    std::vector<int> order(nthreads);
    for (int i = 0; i < nthreads; ++i) order[i] = i;
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
                if (g_trace) g_site_count.assign(nthreads, {});
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
                    if (shuffle) {                      // a different legal interleaving every round
                        for (int i = nthreads - 1; i > 0; --i) {
                            rng = rng * 6364136223846793005ULL + 1442695040888963407ULL;
                            std::swap(order[i], order[(rng >> 33) % (unsigned)(i + 1)]);
                        }
                    }
                    for (int oi = 0; oi < nthreads; ++oi) {
                        const int i = order[oi];
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
                if (g_trace) trace_flush_block();
            }
    if (g_trace) trace_end_launch(grid, block);
    cur = nullptr;
    g_fiber = nullptr;
    g_blk = nullptr;
    g_body = nullptr;
}

}  // namespace cusim
