// tests/cpusim/cpusim.cpp - TEST INFRASTRUCTURE ONLY.  See cpusim.h.
#include "cpusim.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" void cpusim_switch(void** save_sp, void* new_sp);
asm(".text\n"
    ".globl cpusim_switch\n"
    ".type cpusim_switch,@function\n"
    "cpusim_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n");

namespace cpusim {

Fiber* g_cur = nullptr;
Dim3 g_block_dim, g_grid_dim;
unsigned g_concurrent_ctas = 3;
unsigned long long g_switches = 0;

static void* g_sched_sp = nullptr;
static const std::function<void()>* g_body = nullptr;
static const size_t kStack = 96 * 1024;

void yield()
{
    ++g_switches;
    cpusim_switch(&g_cur->sp, g_sched_sp);
}

static void fiber_entry()
{
    (*g_body)();
    g_cur->done = true;
    // exited threads still count as arrived for later barriers of their warp/CTA
    for (;;) {
        yield();
    }
}

static void init_fiber(Fiber& f)
{
    if (!f.stack) f.stack = (char*)aligned_alloc(64, kStack);
    uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                 // fake return address of fiber_entry
    *--sp = (void*)&fiber_entry;     // popped by cpusim_switch's ret
    for (int i = 0; i < 6; ++i) *--sp = nullptr;
    f.sp = sp;
    f.done = false;
}

void launch(Dim3 grid, Dim3 block, size_t smem_bytes, const std::function<void()>& body)
{
    if (g_cur) { fprintf(stderr, "cpusim: nested launch\n"); abort(); }
    const unsigned nthreads = block.x * block.y * block.z;
    const unsigned long long nctas = (unsigned long long)grid.x * grid.y * grid.z;
    const unsigned nwarps = (nthreads + 31) / 32;
    const unsigned conc = g_concurrent_ctas < 1 ? 1 : g_concurrent_ctas;
    g_block_dim = block;
    g_grid_dim = grid;
    g_body = &body;

    struct Slot { Cta cta; std::vector<Warp> warps; std::vector<Fiber> fibers; bool live = false; unsigned ndone = 0; };
    static std::vector<Slot> slots;  // stacks are reused across launches
    if (slots.size() < conc) slots.resize(conc);
    std::vector<char> smem(conc * (smem_bytes + 64));

    unsigned long long next_cta = 0, finished = 0;
    auto start_cta = [&](Slot& s, unsigned slot_idx) {
        unsigned long long b = next_cta++;
        s.live = true;
        s.ndone = 0;
        s.cta = Cta();
        s.cta.nthreads = nthreads;
        s.cta.smem = smem.data() + slot_idx * (smem_bytes + 64);
        memset(s.cta.smem, 0xCD, smem_bytes);
        s.cta.bid = Dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y)));
        s.warps.assign(nwarps, Warp());
        if (s.fibers.size() < nthreads) s.fibers.resize(nthreads);
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber& f = s.fibers[t];
            init_fiber(f);
            f.tid = Dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            f.bid = s.cta.bid;
            f.cta = &s.cta;
            f.warp = &s.warps[t / 32];
            f.lane = t % 32;
            s.warps[t / 32].nlanes++;
        }
    };
    for (unsigned i = 0; i < conc && next_cta < nctas; ++i) start_cta(slots[i], i);
    while (finished < nctas) {
        for (unsigned i = 0; i < conc; ++i) {
            Slot& s = slots[i];
            if (!s.live) continue;
            for (unsigned t = 0; t < nthreads; ++t) {
                Fiber& f = s.fibers[t];
                if (f.done) continue;
                g_cur = &f;
                cpusim_switch(&g_sched_sp, f.sp);
                if (f.done) {
                    s.ndone++;
                    // a finished thread no longer takes part in barriers
                    f.cta->nthreads--;
                    f.warp->nlanes--;
                    if (f.cta->arrived && f.cta->arrived == f.cta->nthreads) { f.cta->arrived = 0; f.cta->gen++; }
                    if (f.warp->arrived && f.warp->arrived == f.warp->nlanes) { f.warp->arrived = 0; f.warp->gen++; }
                }
            }
            if (s.ndone == nthreads) {
                s.live = false;
                finished++;
                if (next_cta < nctas) start_cta(s, i);
            }
        }
    }
    g_cur = nullptr;
    g_body = nullptr;
}

}  // namespace cpusim
