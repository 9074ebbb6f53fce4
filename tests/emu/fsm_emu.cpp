// Fiber-based CUDA block emulator (TEST INFRASTRUCTURE ONLY).
// One CUDA thread == one ucontext fiber. Blocks are distributed over a few OS threads.
// Barriers (__syncthreads, __syncwarp, bar.sync id,count) are cooperative yields; the
// scheduler releases a barrier when every participating fiber has arrived.
#define FSM_EMU 1
#include "fsm_compat.h"
#include <ucontext.h>
#include <vector>
#include <thread>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <memory>

namespace fsm_emu {

thread_local Ctx g_ctx;

enum State { RUNNABLE, WAIT_BLOCK, WAIT_WARP, WAIT_NAMED, DONE };

struct Fiber {
    ucontext_t ctx;
    std::unique_ptr<char[]> stack;
    State state = RUNNABLE;
    int named_id = -1, named_count = 0;
    fsm_dim3 tid;
};

struct BlockRunner {
    ucontext_t main_ctx;
    std::vector<Fiber> fibers;
    int current = -1;
    const std::function<void()>* body = nullptr;
};

static thread_local BlockRunner* t_runner = nullptr;
static const size_t kStack = 256 * 1024;

static void fiber_entry() {
    BlockRunner* r = t_runner;
    (*r->body)();
    r->fibers[r->current].state = DONE;
    swapcontext(&r->fibers[r->current].ctx, &r->main_ctx);
}

static void yield_with(State s, int id = -1, int count = 0) {
    BlockRunner* r = t_runner;
    Fiber& f = r->fibers[r->current];
    f.state = s;
    f.named_id = id;
    f.named_count = count;
    swapcontext(&f.ctx, &r->main_ctx);
}

void barrier_block() { yield_with(WAIT_BLOCK); }
void barrier_warp() { yield_with(WAIT_WARP); }
void barrier_named(int id, int count) { yield_with(WAIT_NAMED, id, count); }

static void run_block(BlockRunner& r, fsm_dim3 grid, fsm_dim3 block, fsm_dim3 bid, char* smem,
                      const std::function<void()>& body) {
    const int nthreads = block.x * block.y * block.z;
    r.body = &body;
    t_runner = &r;
    if ((int)r.fibers.size() != nthreads) {
        r.fibers.clear();
        r.fibers.resize(nthreads);
        for (auto& f : r.fibers) f.stack.reset(new char[kStack]);
    }
    for (int i = 0; i < nthreads; ++i) {
        Fiber& f = r.fibers[i];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.get();
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &r.main_ctx;
        makecontext(&f.ctx, fiber_entry, 0);
        f.state = RUNNABLE;
        f.tid = fsm_dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
    }
    int done = 0;
    while (done < nthreads) {
        bool progressed = false;
        for (int i = 0; i < nthreads; ++i) {
            Fiber& f = r.fibers[i];
            if (f.state != RUNNABLE) continue;
            progressed = true;
            r.current = i;
            g_ctx.tid = f.tid;
            g_ctx.bid = bid;
            g_ctx.bdim = block;
            g_ctx.gdim = grid;
            g_ctx.smem = smem;
            swapcontext(&r.main_ctx, &f.ctx);
            if (f.state == DONE) ++done;
        }
        // release barriers
        int n_wait_block = 0;
        for (auto& f : r.fibers) n_wait_block += (f.state == WAIT_BLOCK);
        if (n_wait_block > 0 && n_wait_block == nthreads - done) {
            for (auto& f : r.fibers) if (f.state == WAIT_BLOCK) f.state = RUNNABLE;
            progressed = true;
        }
        for (int w = 0; w * 32 < nthreads; ++w) {
            int lo = w * 32, hi = std::min(nthreads, lo + 32), nw = 0, alive = 0;
            for (int i = lo; i < hi; ++i) {
                nw += (r.fibers[i].state == WAIT_WARP);
                alive += (r.fibers[i].state != DONE);
            }
            if (nw > 0 && nw == alive) {
                for (int i = lo; i < hi; ++i) if (r.fibers[i].state == WAIT_WARP) r.fibers[i].state = RUNNABLE;
                progressed = true;
            }
        }
        for (int id = 0; id < 16; ++id) {
            int n = 0, want = 0;
            for (auto& f : r.fibers) if (f.state == WAIT_NAMED && f.named_id == id) { ++n; want = f.named_count; }
            if (n > 0 && n >= want) {
                for (auto& f : r.fibers) if (f.state == WAIT_NAMED && f.named_id == id) f.state = RUNNABLE;
                progressed = true;
            }
        }
        if (!progressed) {
            std::fprintf(stderr, "fsm_emu: deadlock in block (%u,%u,%u)\n", bid.x, bid.y, bid.z);
            std::abort();
        }
    }
}

void launch(fsm_dim3 grid, fsm_dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    const long nblocks = (long)grid.x * grid.y * grid.z;
    int nworkers = (int)std::min<long>(nblocks, std::max(1u, std::thread::hardware_concurrency()));
    const char* env = std::getenv("FSM_EMU_THREADS");
    if (env) nworkers = std::max(1, std::min(nworkers, std::atoi(env)));
    std::atomic<long> next(0);
    auto worker = [&]() {
        BlockRunner runner;
        std::vector<char> smem(smem_bytes + 64);
        for (;;) {
            long b = next.fetch_add(1);
            if (b >= nblocks) break;
            fsm_dim3 bid(b % grid.x, (b / grid.x) % grid.y, b / ((long)grid.x * grid.y));
            run_block(runner, grid, block, bid, smem.data(), body);
        }
    };
    if (nworkers <= 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nworkers; ++i) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
}

}  // namespace fsm_emu
