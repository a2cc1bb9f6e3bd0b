// warp_emu.cpp -- test infrastructure: a 32-lane warp on the CPU.
//
// Each lane is a fiber (own stack, cooperative switch); a warp collective (shuffle / ballot / syncwarp with a
// member mask) is a rendezvous of the lanes named in the mask, so warp-synchronous CUDA code -- including
// code that diverges between sub-warp groups and uses group masks -- runs unchanged on the host.  Between
// collectives a fiber runs alone, so memory is always consistent (a missing __syncwarp is NOT detected here;
// that is what compute-sanitizer racecheck on the GPU is for).  A collective that can never complete (lanes of
// the mask waiting at different collectives, or exited) aborts with a dump instead of hanging.
//
// Used by tests/emul/grp_emul.cu to run c3poa_b200/csrc/poa_grp.cuh against the oracle without a GPU.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <csignal>
#include <execinfo.h>
#include <unistd.h>

extern "C" void c3emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl c3emu_switch
.type c3emu_switch, @function
c3emu_switch:
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
.size c3emu_switch, .-c3emu_switch
)");

namespace {
constexpr int kLanes = 32;
constexpr size_t kStack = 1 << 20;

struct Fiber { void *sp; char *stack; bool done; unsigned wait_mask; };
struct Coll { unsigned mask; int arrived; unsigned gen; long long val[2][kLanes]; int tag[2]; };

Fiber g_f[kLanes];
void *g_sched_sp;
int g_cur = -1;
void (*g_body)(void *, int);
void *g_arg;
Coll g_coll[64];
int g_ncoll;
unsigned long long g_progress;

void fiber_entry()
{
    g_body(g_arg, g_cur);
    g_f[g_cur].done = true;
    ++g_progress;
    c3emu_switch(&g_f[g_cur].sp, g_sched_sp);
    abort();
}

void yield() { c3emu_switch(&g_f[g_cur].sp, g_sched_sp); }

Coll *find(unsigned mask)
{
    for (int i = 0; i < g_ncoll; ++i) if (g_coll[i].mask == mask) return &g_coll[i];
    if (g_ncoll >= 64) { fprintf(stderr, "warp_emu: too many distinct masks\n"); abort(); }
    Coll *c = &g_coll[g_ncoll++];
    memset(c, 0, sizeof(*c));
    c->mask = mask;
    return c;
}

long long collective(unsigned mask, long long v, int src, bool want_ballot, int tag)
{
    const int lane = g_cur;
    if (!((mask >> lane) & 1u)) { fprintf(stderr, "warp_emu: lane %d not in its own mask %08x\n", lane, mask); abort(); }
    Coll *c = find(mask);
    const unsigned my = c->gen;
    // every lane of a rendezvous must come from the same call site: lanes meeting at DIFFERENT collectives is a
    // deadlock (or undefined behaviour) on the GPU
    if (c->arrived == 0) c->tag[my & 1] = tag;
    else if (c->tag[my & 1] != tag) {
        fprintf(stderr, "warp_emu: lanes of mask %08x meet at different collectives (source lines %d and %d, lane %d)\n", mask, c->tag[my & 1], tag, lane);
        abort();
    }
    c->val[my & 1][lane] = v;
    if (++c->arrived == __builtin_popcount(mask)) { c->arrived = 0; c->gen++; ++g_progress; }
    else {
        g_f[lane].wait_mask = mask;
        while (c->gen == my) yield();
        g_f[lane].wait_mask = 0;
    }
    if (want_ballot) {
        unsigned b = 0;
        for (int l = 0; l < kLanes; ++l) if (((mask >> l) & 1u) && c->val[my & 1][l]) b |= 1u << l;
        return (long long)b;
    }
    if (src < 0) return 0;
    src &= 31;
    if (!((mask >> src) & 1u)) { fprintf(stderr, "warp_emu: lane %d shuffles from lane %d outside mask %08x\n", lane, src, mask); abort(); }
    return c->val[my & 1][src];
}
}  // namespace

static void segv_handler(int sig)
{
    void *bt[32];
    const int n = backtrace(bt, 32);
    fprintf(stderr, "warp_emu: signal %d in lane %d\n", sig, g_cur);
    backtrace_symbols_fd(bt, n, 2);
    _exit(139);
}
extern "C" {
int c3emu_lane(void) { return g_cur; }
int c3emu_shfl(unsigned mask, int v, int src, int tag) { return (int)collective(mask, v, src, false, tag); }
unsigned c3emu_ballot(unsigned mask, int pred, int tag) { return (unsigned)collective(mask, pred ? 1 : 0, -1, true, tag); }
void c3emu_sync(unsigned mask, int tag) { (void)collective(mask, 0, -1, false, tag); }

// runs body(arg, lane) for the 32 lanes of one warp to completion; returns 0, or -1 on a deadlock
int c3emu_run_warp(void (*body)(void *, int), void *arg)
{
    g_body = body; g_arg = arg; g_ncoll = 0; g_progress = 0;
    if (getenv("C3EMU_BACKTRACE")) { signal(SIGSEGV, segv_handler); signal(SIGBUS, segv_handler); }
    for (int l = 0; l < kLanes; ++l) {
        Fiber &f = g_f[l];
        if (!f.stack) f.stack = (char *)aligned_alloc(64, kStack);
        f.done = false; f.wait_mask = 0;
        uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
        void **sp = (void **)(top - 64);
        for (int i = 0; i < 6; ++i) sp[i] = nullptr;
        sp[6] = (void *)&fiber_entry;
        sp[7] = nullptr;
        f.sp = sp;
    }
    int live = kLanes;
    unsigned long long last = ~0ull;
    int stale = 0;
    while (live > 0) {
        const unsigned long long before = g_progress;
        live = 0;
        for (int l = 0; l < kLanes; ++l) {
            if (g_f[l].done) continue;
            g_cur = l;
            c3emu_switch(&g_sched_sp, g_f[l].sp);
            if (!g_f[l].done) ++live;
        }
        if (live > 0 && g_progress == before && before == last) {
            if (++stale > 4) {
                fprintf(stderr, "warp_emu: deadlock; lanes waiting on masks:");
                for (int l = 0; l < kLanes; ++l) fprintf(stderr, " %d:%s%08x", l, g_f[l].done ? "done/" : "", g_f[l].wait_mask);
                fprintf(stderr, "\n");
                return -1;
            }
        } else stale = 0;
        last = g_progress;
    }
    g_cur = -1;
    return 0;
}
}
