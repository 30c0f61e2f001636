// tests/simt/shim/simt_core.h — TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory).
// Fiber-based SIMT interpreter: CTAs run on worker OS threads, a CTA's threads are fibers, warp / CTA collectives block a
// fiber until its partners arrive.  x86-64 System V only (the container and the GPU box's host).
#pragma once
#include <sys/mman.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <ucontext.h>
#include <dlfcn.h>

namespace simt {

inline int env_int(const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; }

// ---- context switch ---------------------------------------------------------------------------------------------------------
extern "C" void simt_switch(void** saveSp, void* loadSp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
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
.size simt_switch,.-simt_switch
)");

enum Op { OP_BALLOT, OP_ADD, OP_SHFL, OP_SYNC };

struct Warp
{
    unsigned int arrived = 0, exited = 0, present = 0;        // lane masks
    unsigned int gen = 0;
    Op op = OP_SYNC; unsigned int mask = 0;
    unsigned long long vals[32], snap[32];
    unsigned int snapArrived = 0; unsigned long long snapSum = 0; unsigned int snapBallot = 0;
};

struct Fiber { void* sp = nullptr; bool done = false; };

constexpr size_t FIBER_STACK = 96 * 1024;

struct Cta
{
    dim3 grid, block; uint3 bidx; unsigned int nThreads = 0;
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    unsigned int syncGen = 0, syncArrived = 0, liveThreads = 0;
    void* schedSp = nullptr;
    unsigned int cur = 0;
    unsigned long long progress = 0;                          // bumped whenever a collective completes or a thread exits
    const std::function<void()>* body = nullptr;
    char* stacks = nullptr; size_t stacksBytes = 0;           // reused across the CTAs this worker runs
};

inline thread_local Cta* tl_cta = nullptr;
inline thread_local uint3 tl_threadIdx, tl_blockIdx;
inline thread_local dim3 tl_blockDim, tl_gridDim;

inline void set_builtins(Cta* c, unsigned int tid)
{
    tl_threadIdx.x = tid % c->block.x; tl_threadIdx.y = (tid / c->block.x) % c->block.y; tl_threadIdx.z = tid / (c->block.x * c->block.y);
}

inline void yield()
{
    Cta* c = tl_cta;
    simt_switch(&c->fibers[c->cur].sp, c->schedSp);
}

// schedule profiler (RT_SIMT_PROFILE builds): run-time scheduling knobs and counters, exported for tools/simt_schedule_profile.py
extern "C" { inline int simtKnob[8] = {1, 1, 1, 2, 1, 0, 0, 0}; inline unsigned long long simtProf[64] = {}; }
inline int* const knob = simtKnob;
inline void prof_add(int i, unsigned long long v) { __atomic_fetch_add(&simtProf[i], v, __ATOMIC_RELAXED); }

inline unsigned int lane_id() { return tl_cta->cur & 31u; }

inline unsigned int activemask()
{
    // any subset of the converged lanes that contains the caller is a legal answer; the smallest one needs no rendezvous
    return 1u << (tl_cta->cur & 31u);
}

inline unsigned long long collective(Op op, unsigned int mask, unsigned int val, int src)
{
    Cta* c = tl_cta;
    const unsigned int lane = c->cur & 31u, bit = 1u << lane;
    Warp& w = c->warps[c->cur >> 5];
    if (!(mask & bit)) { fprintf(stderr, "simt: lane %u calls a collective whose mask %08x does not name it\n", lane, mask); abort(); }
    mask &= w.present;                                        // lanes beyond blockDim do not exist
    if (mask == bit)
    {
        switch (op) { case OP_BALLOT: return val ? bit : 0u; case OP_ADD: return val; case OP_SHFL: return val; default: return 0; }
    }
    if (w.arrived == 0) { w.op = op; w.mask = mask; }
    else if (w.op != op || w.mask != mask)
    {
        fprintf(stderr, "simt: lanes of one warp wait in different collectives (op %d mask %08x vs op %d mask %08x): divergent collective\n", (int)w.op, w.mask, (int)op, mask);
        abort();
    }
    const unsigned int myGen = w.gen;
    w.vals[lane] = val;
    w.arrived |= bit;
    for (;;)
    {
        if (w.gen != myGen) break;
        if (((w.arrived | w.exited) & mask) == mask)
        {
            // last one in: publish the snapshot every participant reads its result from
            unsigned long long sum = 0; unsigned int ballot = 0;
            for (int l = 0; l < 32; l++)
                if (w.arrived & (1u << l)) { w.snap[l] = w.vals[l]; sum += (unsigned int)w.vals[l]; if (w.vals[l]) ballot |= 1u << l; }
            w.snapArrived = w.arrived; w.snapSum = sum; w.snapBallot = ballot;
            w.arrived = 0; w.gen++; c->progress++;
            break;
        }
        yield();
    }
    switch (op)
    {
    case OP_BALLOT: return w.snapBallot;
    case OP_ADD: return (unsigned int)w.snapSum;
    case OP_SHFL:
    {
        const unsigned int s = (unsigned int)src & 31u;
        return (w.snapArrived & (1u << s)) ? w.snap[s] : val;     // reading an inactive lane is undefined on hardware; keep own value
    }
    default: return 0;
    }
}

inline void syncthreads()
{
    Cta* c = tl_cta;
    const unsigned int myGen = c->syncGen;
    c->syncArrived++;
    for (;;)
    {
        if (c->syncGen != myGen) break;
        if (c->syncArrived >= c->liveThreads) { c->syncArrived = 0; c->syncGen++; c->progress++; break; }
        yield();
    }
}

inline void fiber_entry()
{
    Cta* c = tl_cta;
    (*c->body)();
    // thread exit: it no longer takes part in collectives
    Fiber& f = c->fibers[c->cur];
    f.done = true;
    c->warps[c->cur >> 5].exited |= 1u << (c->cur & 31u);
    c->liveThreads--; c->progress++;
    void* dummy;
    simt_switch(&dummy, c->schedSp);
    abort();                                                   // never resumed
}

inline void run_cta(Cta& c)
{
    const unsigned int n = c.nThreads;
    const size_t need = (size_t)n * FIBER_STACK;
    if (c.stacksBytes < need)
    {
        if (c.stacks) munmap(c.stacks, c.stacksBytes);
        c.stacks = (char*)mmap(nullptr, need, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (c.stacks == MAP_FAILED) { fprintf(stderr, "simt: cannot map fiber stacks\n"); abort(); }
        c.stacksBytes = need;
    }
    c.fibers.assign(n, Fiber());
    c.warps.assign((n + 31) / 32, Warp());
    for (unsigned int t = 0; t < n; t++)
    {
        c.warps[t >> 5].present |= 1u << (t & 31u);
        uintptr_t top = ((uintptr_t)(c.stacks + (size_t)(t + 1) * FIBER_STACK)) & ~(uintptr_t)15;
        void** sp = (void**)top - 8;
        for (int i = 0; i < 6; i++) sp[i] = nullptr;           // r15 r14 r13 r12 rbx rbp
        sp[6] = (void*)&fiber_entry;                           // `ret` lands here with rsp = top - 8 (ABI: rsp + 8 is 16-aligned at entry)
        sp[7] = nullptr;
        c.fibers[t].sp = sp;
    }
    c.syncGen = 0; c.syncArrived = 0; c.liveThreads = n; c.progress = 0;
    tl_cta = &c;
    tl_blockIdx = c.bidx; tl_blockDim = c.block; tl_gridDim = c.grid;
    unsigned long long lastProgress = ~0ull; unsigned long long idleRounds = 0;
    // Order in which the scheduler visits the threads of the CTA in a round (RT_SIMT_ORDER): 0 = ascending (default), 1 = descending,
    // 2 = a fresh pseudo-random order every round.  A race between lanes that is masked by one order shows up under another.
    const int orderMode = env_int("RT_SIMT_ORDER", 0);
    std::vector<unsigned int> order(n);
    for (unsigned int t = 0; t < n; t++) order[t] = orderMode == 1 ? n - 1 - t : t;
    unsigned long long rng = 0x9E3779B97F4A7C15ull ^ ((unsigned long long)c.bidx.x << 32) ^ c.bidx.y;
    while (c.liveThreads > 0)
    {
        if (orderMode == 2)
            for (unsigned int i = n; i > 1; i--)
            {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                const unsigned int j = (unsigned int)((rng >> 33) % i);
                const unsigned int tmp = order[i - 1]; order[i - 1] = order[j]; order[j] = tmp;
            }
        for (unsigned int k = 0; k < n; k++)
        {
            const unsigned int t = order[k];
            if (c.fibers[t].done) continue;
            c.cur = t;
            set_builtins(&c, t);
            simt_switch(&c.schedSp, c.fibers[t].sp);
        }
        if (c.progress == lastProgress) { if (++idleRounds > 2000000ull) { fprintf(stderr, "simt: deadlock — no thread of the CTA makes progress (a collective some lane never reaches?)\n"); abort(); } }
        else { lastProgress = c.progress; idleRounds = 0; }
    }
    tl_cta = nullptr;
}

// A crash inside a fiber would otherwise die silently: say where (the CTA / thread and a raw backtrace for addr2line).
inline void crash_handler(int sig, siginfo_t* si, void* uc)
{
    char buf[512];
    Cta* c = tl_cta;
    const unsigned long long rip = (unsigned long long)((ucontext_t*)uc)->uc_mcontext.gregs[REG_RIP];
    Dl_info info; memset(&info, 0, sizeof(info)); dladdr((void*)rip, &info);
    int n = snprintf(buf, sizeof(buf), "simt: signal %d at pc %s+0x%llx (address %p) in block (%u,%u,%u) thread %u\n", sig, info.dli_fname ? info.dli_fname : "?",
                     rip - (unsigned long long)info.dli_fbase, si->si_addr, c ? c->bidx.x : 0u, c ? c->bidx.y : 0u, c ? c->bidx.z : 0u, c ? c->cur : 0u);
    if (write(2, buf, (size_t)n) < 0) { }
    void* bt[48]; const int k = backtrace(bt, 48);
    backtrace_symbols_fd(bt, k, 2);
    _exit(139);
}
inline void install_crash_handler()
{
    static std::atomic<bool> done(false);
    if (done.exchange(true) || !env_int("RT_SIMT_BACKTRACE", 0)) return;
    static char altstack[65536];
    stack_t ss; ss.ss_sp = altstack; ss.ss_size = sizeof(altstack); ss.ss_flags = 0; sigaltstack(&ss, nullptr);
    struct sigaction sa; memset(&sa, 0, sizeof(sa)); sa.sa_sigaction = crash_handler; sa.sa_flags = SA_ONSTACK | SA_SIGINFO;
    sigaction(SIGSEGV, &sa, nullptr); sigaction(SIGBUS, &sa, nullptr); sigaction(SIGFPE, &sa, nullptr);
}

// Synchronous launch: the CTAs of the grid are taken from a shared counter by a few worker threads.
inline void launch(dim3 grid, dim3 block, size_t smemBytes, const std::function<void()>& body)
{
    if (smemBytes > 232448) { fprintf(stderr, "simt: %zu bytes of dynamic shared memory exceed the 227 KB of an SM\n", smemBytes); abort(); }
    install_crash_handler();
    const unsigned long long total = (unsigned long long)grid.x * grid.y * grid.z;
    const unsigned int nThreads = block.x * block.y * block.z;
    if (total == 0 || nThreads == 0 || nThreads > 1024) { fprintf(stderr, "simt: invalid launch configuration\n"); abort(); }
    std::atomic<unsigned long long> next(0);
    auto worker = [&]()
    {
        Cta c;
        c.grid = grid; c.block = block; c.nThreads = nThreads; c.body = &body;
        for (;;)
        {
            const unsigned long long b = next.fetch_add(1);
            if (b >= total) break;
            c.bidx.x = (unsigned int)(b % grid.x); c.bidx.y = (unsigned int)((b / grid.x) % grid.y); c.bidx.z = (unsigned int)(b / ((unsigned long long)grid.x * grid.y));
            run_cta(c);
        }
        if (c.stacks) munmap(c.stacks, c.stacksBytes);
    };
    unsigned int nWorkers = (unsigned int)env_int("RT_SIMT_WORKERS", 0);
    if (nWorkers == 0) { nWorkers = std::thread::hardware_concurrency(); if (nWorkers == 0) nWorkers = 4; if (nWorkers > 16) nWorkers = 16; }
    if ((unsigned long long)nWorkers > total) nWorkers = (unsigned int)total;
    if (nWorkers <= 1) { worker(); return; }
    std::vector<std::thread> th;
    for (unsigned int i = 0; i < nWorkers; i++) th.emplace_back(worker);
    for (auto& t : th) t.join();
}

} // namespace simt
