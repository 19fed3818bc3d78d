// TEST INFRASTRUCTURE (see tests/emu/include/cuda_runtime.h): runs a kernel launch on the host.  Blocks run one after the
// other; the threads of a block are ucontext fibers scheduled round-robin, parked at __syncthreads() until every live
// thread of the block has arrived and at a warp shuffle until every live lane of the warp has.
#include <cuda_runtime.h>

EmuThread* emu_cur = nullptr;
uint3 emu_blockIdx, emu_blockDim, emu_gridDim;

namespace {
constexpr size_t STACK = 256 * 1024;
std::vector<EmuThread> g_threads;
std::vector<char> g_stacks;
std::vector<char> g_smem;
ucontext_t g_sched;
const std::function<void()>* g_body = nullptr;

void trampoline()
{
    (*g_body)();
    emu_cur->state = 3;
    swapcontext(&emu_cur->ctx, &g_sched);
}
} // namespace

void* emu_dyn_smem() { return g_smem.data(); }

void emu_yield(int state)
{
    emu_cur->state = state;
    swapcontext(&emu_cur->ctx, &g_sched);
}

void emu_launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body)
{
    g_body = &body;
    if (g_threads.size() < block) {
        g_threads.resize(block);
        g_stacks.resize((size_t)block * STACK);
    }
    g_smem.assign(smem + 16, 0);
    emu_gridDim = uint3{ grid, 1, 1 };
    emu_blockDim = uint3{ block, 1, 1 };
    for (unsigned b = 0; b < grid; ++b) {
        emu_blockIdx = uint3{ b, 0, 0 };
        for (unsigned t = 0; t < block; ++t) {
            EmuThread& th = g_threads[t];
            getcontext(&th.ctx);
            th.ctx.uc_stack.ss_sp = g_stacks.data() + (size_t)t * STACK;
            th.ctx.uc_stack.ss_size = STACK;
            th.ctx.uc_link = &g_sched;
            makecontext(&th.ctx, trampoline, 0);
            th.tid = uint3{ t, 0, 0 };
            th.state = 0;
        }
        for (;;) {
            unsigned done = 0, at_sync = 0, progressed = 0;
            for (unsigned t = 0; t < block; ++t) {
                EmuThread& th = g_threads[t];
                if (th.state == 0) {
                    emu_cur = &th;
                    swapcontext(&g_sched, &th.ctx);
                    ++progressed;
                }
                done += th.state == 3;
                at_sync += th.state == 1;
            }
            if (done == block) {
                break;
            }
            // warp shuffles: a warp whose live lanes all wait exchanges and resumes
            for (unsigned w = 0; w < block; w += 32) {
                const unsigned end = std::min(block, w + 32);
                bool all = true, any = false;
                for (unsigned t = w; t < end; ++t) {
                    all = all && (g_threads[t].state == 2 || g_threads[t].state == 3);
                    any = any || g_threads[t].state == 2;
                }
                if (all && any) {
                    unsigned long long vote = 0; // __any_sync: shfl_off == ~0u marks a vote; every waiting lane gets the OR
                    for (unsigned t = w; t < end; ++t) {
                        if ((g_threads[t].state == 2) && (g_threads[t].shfl_off == ~0u)) {
                            vote |= g_threads[t].shfl_in;
                        }
                    }
                    for (unsigned t = w; t < end; ++t) {
                        if (g_threads[t].shfl_off == ~0u) {
                            g_threads[t].shfl_out = vote;
                            continue;
                        }
                        const unsigned src = t + g_threads[t].shfl_off;
                        g_threads[t].shfl_out = (src < end) ? g_threads[src].shfl_in : g_threads[t].shfl_in;
                    }
                    for (unsigned t = w; t < end; ++t) {
                        if (g_threads[t].state == 2) {
                            g_threads[t].state = 0;
                            ++progressed;
                        }
                    }
                }
            }
            if (at_sync && (at_sync + done == block)) {
                for (unsigned t = 0; t < block; ++t) {
                    if (g_threads[t].state == 1) {
                        g_threads[t].state = 0;
                    }
                }
                ++progressed;
            }
            if (!progressed) {
                fprintf(stderr, "[emu] deadlock: %u at __syncthreads, %u done of %u threads\n", at_sync, done, block);
                abort();
            }
        }
    }
    emu_cur = nullptr;
    g_body = nullptr;
}
