// simt_engine.cpp (SIMT emulator) -- TEST INFRASTRUCTURE, not product code.
//
// Executes a CUDA grid on the host: every CUDA thread is a fiber with its own stack, the fibers of one
// block run on one host thread, blocks are handed out to a small pool of host threads.  A fiber runs until it
// reaches a warp collective that is not complete yet (it then yields to the other lanes of its warp), a
// __syncthreads() (it yields until every live thread of the block has arrived) or the end of the kernel.
// If no fiber of a block can make progress the block has deadlocked -- a collective inside divergent code,
// a barrier not reached by every thread -- and the process aborts with a message saying where every warp
// stands.  Execution order inside a warp is lane order, warps run one after the other: the schedule is
// deterministic, so a failure reproduces.
#include "cuda_runtime.h"
#include <condition_variable>
#include <mutex>
#include <sys/mman.h>
#include <thread>
#include <vector>

extern "C" void emu_ctx_switch(void **from_sp, void *to_sp);
asm(R"(
.text
.globl emu_ctx_switch
.type emu_ctx_switch,@function
emu_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_ctx_switch, .-emu_ctx_switch
)");

namespace emu {

thread_local ThreadCtx tctx;

namespace {
constexpr size_t kStackBytes = 64 * 1024;
constexpr int kMaxThreads = 1024;
enum State : int { kRunnable = 0, kWaitWarp = 1, kWaitBlock = 2, kDone = 3 };

struct Lane {
    void *sp = nullptr;
    char *stack = nullptr;
    State state = kDone;
    unsigned wait_gen = 0;
};
struct Warp {
    uint64_t buf[2][32];
    unsigned present[2];
    unsigned gen = 0;
    int arrived = 0, live = 0;
};
struct BlockCtx {
    Lane lanes[kMaxThreads];
    Warp warps[kMaxThreads / 32];
    void *sched_sp = nullptr;
    int nthreads = 0, ndone = 0, at_barrier = 0;
    int current = -1;
    dim3 bdim, gdim;
    uint3 bid;
    void (*invoke)(void *) = nullptr;
    void *closure = nullptr;
    std::vector<char> dyn;
    char *stacks = nullptr;
};
thread_local BlockCtx *g_block = nullptr;

void release_warp(Warp &w) {
    w.arrived = 0;
    w.present[(w.gen + 1) & 1] = 0;
    ++w.gen;
}

void yield_to_scheduler(BlockCtx *b, Lane &L) { emu_ctx_switch(&L.sp, b->sched_sp); }

void lane_entry() {
    BlockCtx *b = g_block;
    const int t = b->current;
    b->invoke(b->closure);
    Lane &L = b->lanes[t];
    Warp &w = b->warps[t >> 5];
    L.state = kDone;
    ++b->ndone;
    --w.live;
    if (w.live > 0 && w.arrived == w.live) release_warp(w); // the others were only waiting for lanes that still ran
    yield_to_scheduler(b, L);
    fail("resumed a finished CUDA thread");
}

void set_thread(BlockCtx *b, int t) {
    b->current = t;
    tctx.lane_linear = t;
    const unsigned bx = b->bdim.x, by = b->bdim.y;
    tctx.tid.x = (unsigned)t % bx;
    tctx.tid.y = ((unsigned)t / bx) % by;
    tctx.tid.z = (unsigned)t / (bx * by);
}

void describe_and_abort(BlockCtx *b) {
    fprintf(stderr, "[simt-emu] DEADLOCK in block (%u,%u,%u): no thread can make progress\n", b->bid.x, b->bid.y, b->bid.z);
    for (int w = 0; w * 32 < b->nthreads; ++w) {
        int c[4] = {0, 0, 0, 0};
        for (int l = w * 32; l < std::min(b->nthreads, w * 32 + 32); ++l) ++c[b->lanes[l].state];
        fprintf(stderr, "  warp %2d: %d at a warp collective, %d at __syncthreads, %d finished, %d runnable\n", w, c[kWaitWarp],
                c[kWaitBlock], c[kDone], c[kRunnable]);
    }
    abort();
}

void run_block(BlockCtx *b) {
    const int nt = b->nthreads;
    b->ndone = 0;
    b->at_barrier = 0;
    for (int w = 0; w * 32 < nt; ++w) {
        Warp &W = b->warps[w];
        W.gen = 0;
        W.arrived = 0;
        W.live = std::min(32, nt - w * 32);
        W.present[0] = W.present[1] = 0;
    }
    for (int t = 0; t < nt; ++t) {
        Lane &L = b->lanes[t];
        L.stack = b->stacks + (size_t)t * kStackBytes;
        uintptr_t top = ((uintptr_t)L.stack + kStackBytes) & ~(uintptr_t)15;
        void **sp = (void **)(top - 64);
        for (int k = 0; k < 6; ++k) sp[k] = nullptr; // r15 r14 r13 r12 rbx rbp
        sp[6] = (void *)&lane_entry;                  // return address; rsp = top - 8 on entry
        sp[7] = nullptr;
        L.sp = sp;
        L.state = kRunnable;
        *(uint64_t *)L.stack = 0x5AFE5AFE5AFE5AFEull; // overflow canary at the low end
    }
    tctx.bid = b->bid;
    tctx.bdim = b->bdim;
    tctx.gdim = b->gdim;
    while (b->ndone < nt) {
        bool progressed = false;
        for (int w = 0; w * 32 < nt; ++w) {
            Warp &W = b->warps[w];
            const int l0 = w * 32, l1 = std::min(nt, l0 + 32);
            bool any = true;
            while (any) {
                any = false;
                for (int t = l0; t < l1; ++t) {
                    Lane &L = b->lanes[t];
                    if (L.state == kDone || L.state == kWaitBlock) continue;
                    if (L.state == kWaitWarp) {
                        if (W.gen == L.wait_gen) continue;
                        L.state = kRunnable;
                    }
                    set_thread(b, t);
                    emu_ctx_switch(&b->sched_sp, L.sp);
                    if (*(uint64_t *)L.stack != 0x5AFE5AFE5AFE5AFEull) fail("CUDA thread stack overflow in the emulator");
                    any = progressed = true;
                }
            }
        }
        if (b->at_barrier > 0 && b->at_barrier == nt - b->ndone) {
            for (int t = 0; t < nt; ++t)
                if (b->lanes[t].state == kWaitBlock) b->lanes[t].state = kRunnable;
            b->at_barrier = 0;
            progressed = true;
        }
        if (!progressed) describe_and_abort(b);
    }
}

// ---- host thread pool ------------------------------------------------------------------------------
struct Grid {
    dim3 grid, block;
    size_t smem;
    void (*invoke)(void *);
    void *closure;
    std::atomic<uint64_t> next{0};
    uint64_t nblocks;
};

BlockCtx *my_block_ctx() {
    static thread_local BlockCtx *ctx = nullptr;
    if (!ctx) {
        ctx = new BlockCtx();
        ctx->stacks = (char *)mmap(nullptr, kStackBytes * kMaxThreads, PROT_READ | PROT_WRITE,
                                   MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (ctx->stacks == (char *)MAP_FAILED) fail("mmap of the fiber stacks failed");
    }
    return ctx;
}

void work_on(Grid *g) {
    BlockCtx *b = my_block_ctx();
    g_block = b;
    while (true) {
        const uint64_t k = g->next.fetch_add(1);
        if (k >= g->nblocks) break;
        b->nthreads = (int)(g->block.x * g->block.y * g->block.z);
        b->bdim = g->block;
        b->gdim = g->grid;
        b->bid.x = (unsigned)(k % g->grid.x);
        b->bid.y = (unsigned)((k / g->grid.x) % g->grid.y);
        b->bid.z = (unsigned)(k / ((uint64_t)g->grid.x * g->grid.y));
        b->invoke = g->invoke;
        b->closure = g->closure;
        if (b->dyn.size() < g->smem + 64) b->dyn.resize(g->smem + 64);
        run_block(b);
    }
}

struct Pool {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> threads;
    Grid *job = nullptr;
    uint64_t job_id = 0;
    int busy = 0;
    bool stop = false;
    int nworkers;
    Pool() {
        const char *e = getenv("SIMT_EMU_THREADS");
        int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        nworkers = std::max(1, std::min(n, 64));
        for (int k = 1; k < nworkers; ++k) threads.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &t : threads) t.join();
    }
    void loop() {
        uint64_t seen = 0;
        while (true) {
            Grid *g;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || (job && job_id != seen); });
                if (stop) return;
                seen = job_id;
                g = job;
                ++busy;
            }
            work_on(g);
            {
                std::lock_guard<std::mutex> lk(mu);
                --busy;
            }
            cv_done.notify_all();
        }
    }
    void run(Grid *g) {
        if (g->nblocks > 1 && nworkers > 1) {
            {
                std::lock_guard<std::mutex> lk(mu);
                job = g;
                ++job_id;
            }
            cv_work.notify_all();
        }
        work_on(g);
        std::unique_lock<std::mutex> lk(mu);
        job = nullptr; // late wakers must not pick up a finished grid
        cv_done.wait(lk, [&] { return busy == 0; });
    }
};
Pool &pool() {
    static Pool *p = new Pool(); // leaked on purpose: no destructor order issues at process exit
    return *p;
}
std::mutex g_launch_mu; // one grid at a time (the emulated device executes every stream synchronously)
} // namespace

int num_workers() { return pool().nworkers; }

void fail(const char *msg) {
    fprintf(stderr, "[simt-emu] %s\n", msg);
    abort();
}

void *dyn_smem_raw() {
    BlockCtx *b = g_block;
    return (void *)(((uintptr_t)b->dyn.data() + 63) & ~(uintptr_t)63);
}

const uint64_t *warp_exchange(uint64_t v, unsigned *present) {
    BlockCtx *b = g_block;
    const int t = b->current;
    Warp &w = b->warps[t >> 5];
    const int lane = t & 31;
    const unsigned g = w.gen;
    w.buf[g & 1][lane] = v;
    w.present[g & 1] |= 1u << lane;
    if (++w.arrived == w.live) {
        release_warp(w);
    } else {
        Lane &L = b->lanes[t];
        L.state = kWaitWarp;
        L.wait_gen = g;
        yield_to_scheduler(b, L);
    }
    *present = w.present[g & 1];
    return w.buf[g & 1];
}

void block_barrier() {
    BlockCtx *b = g_block;
    Lane &L = b->lanes[b->current];
    L.state = kWaitBlock;
    ++b->at_barrier;
    yield_to_scheduler(b, L);
}

void run_grid(dim3 grid, dim3 block, size_t smem, void (*invoke)(void *), void *closure) {
    const uint64_t nthreads = (uint64_t)block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > kMaxThreads) fail("block size out of range");
    if (g_block && g_block->current >= 0 && g_block->ndone < g_block->nthreads) fail("kernel launch from inside a kernel");
    std::lock_guard<std::mutex> lk(g_launch_mu);
    Grid g;
    g.grid = grid;
    g.block = block;
    g.smem = smem;
    g.invoke = invoke;
    g.closure = closure;
    g.nblocks = (uint64_t)grid.x * grid.y * grid.z;
    if (g.nblocks == 0) return;
    pool().run(&g);
}

} // namespace emu
