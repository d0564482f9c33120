// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): the execution model behind the host stand-in.
//
// hostcc::launch runs the blocks of a grid one after the other; inside a block every CUDA thread is a coroutine (ucontext)
// on the calling OS thread, scheduled round-robin.  A thread runs until it reaches a rendezvous (block barrier, warp
// collective) that is not complete yet or calls yield() from a spin loop.  Threads that have returned count as arrived, the
// way exited threads do on the hardware.  One coroutine runs at a time, so plain read-modify-writes are atomic and a data
// race between two barriers shows up as an order dependence at worst -- this model checks logic, not the memory model.
// Persistent kernels whose blocks wait for each other must make progress with one resident block (the library's ticket /
// look-back / work-queue kernels do: a block only ever waits for work handed out earlier).
#include <cuda_runtime.h>

#include <sys/mman.h>
#include <ucontext.h>

#include <utility>
#include <vector>

// AddressSanitizer has to be told about every stack switch (hostcc_build.build(extra=('-fsanitize=address',)))
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#define HOSTCC_ASAN 1
#else
#define HOSTCC_ASAN 0
#endif

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace hostcc {

namespace {

constexpr size_t kStackBytes = 256 * 1024;

struct Warp {
    unsigned alive = 0, arrived = 0, want = 0;      // lanes not yet returned / deposited in the open rendezvous / its mask
    unsigned long long val[32] = {}, snap[32] = {};
    unsigned snapMask = 0;
    unsigned long long gen = 0;
};

struct Lane {
    ucontext_t uc;
    uint3 tid;
    unsigned warp, lane;
    bool done = false;
    // blocked until *waitGen != waitAt (nullptr: runnable)
    const unsigned long long* waitGen = nullptr;
    unsigned long long waitAt = 0;
};

ucontext_t g_sched;
std::vector<Lane> g_lanes;
std::vector<Warp> g_warps;
char* g_stacks = nullptr;                       // kMaxLanes stacks, reserved once, touched on demand
constexpr unsigned kMaxLanes = 1024;
Lane* g_cur = nullptr;
const std::function<void()>* g_body = nullptr;
unsigned g_blockAlive = 0, g_blockArrived = 0;
unsigned long long g_blockGen = 0;

#if HOSTCC_ASAN
void* g_schedFake = nullptr; const void* g_schedBottom = nullptr; size_t g_schedSize = 0;
#endif

void switch_out()
{
#if HOSTCC_ASAN
    Lane* me = g_cur;
    void* fake = nullptr;
    __sanitizer_start_switch_fiber(me->done ? nullptr : &fake, g_schedBottom, g_schedSize);
    swapcontext(&me->uc, &g_sched);
    __sanitizer_finish_switch_fiber(fake, &g_schedBottom, &g_schedSize);
#else
    swapcontext(&g_cur->uc, &g_sched);
#endif
}

void wait_for(const unsigned long long* gen, unsigned long long at)
{
    Lane* me = g_cur;
    while (*gen == at) { me->waitGen = gen; me->waitAt = at; switch_out(); }
    me->waitGen = nullptr;
}

void finish_rendezvous(Warp& w)
{
    memcpy(w.snap, w.val, sizeof(w.val));
    w.snapMask = w.arrived;
    w.arrived = 0; w.want = 0;
    w.gen++;
}

void lane_entry()
{
    Lane* me = g_cur;
#if HOSTCC_ASAN
    __sanitizer_finish_switch_fiber(nullptr, &g_schedBottom, &g_schedSize);
#endif
    (*g_body)();
    me->done = true;
    // a returned thread no longer takes part in barriers / collectives: it may complete the ones its siblings wait in
    Warp& w = g_warps[me->warp];
    w.alive &= ~(1u << me->lane);
    if (w.arrived && (w.arrived & (w.want & w.alive)) == (w.want & w.alive)) finish_rendezvous(w);
    g_blockAlive--;
    if (g_blockArrived && g_blockArrived == g_blockAlive) { g_blockArrived = 0; g_blockGen++; }
    switch_out();
    abort();        // never resumed
}

void run_block(dim3 block)
{
    const unsigned n = block.x * block.y * block.z;
    g_lanes.assign(n, Lane());
    g_warps.assign((n + 31) / 32, Warp());
    if (n > kMaxLanes) { fprintf(stderr, "hostcc: block of %u threads\n", n); abort(); }
    if (!g_stacks) {
        g_stacks = (char*)mmap(nullptr, (size_t)kMaxLanes * kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_stacks == MAP_FAILED) { perror("hostcc: mmap"); abort(); }
    }
    g_blockAlive = n; g_blockArrived = 0;
    for (unsigned i = 0; i < n; i++) {
        Lane& l = g_lanes[i];
        l.tid = uint3{i % block.x, (i / block.x) % block.y, i / (block.x * block.y)};
        l.warp = i / 32; l.lane = i % 32;
        g_warps[l.warp].alive |= 1u << l.lane;
        getcontext(&l.uc);
        l.uc.uc_stack.ss_sp = g_stacks + (size_t)i * kStackBytes;
        l.uc.uc_stack.ss_size = kStackBytes;
        l.uc.uc_link = nullptr;
        makecontext(&l.uc, lane_entry, 0);
    }
    // Scheduling order of the runnable threads: a correctly synchronised kernel does not care.  HOSTCC_ORDER=reverse | random
    // (seeded) runs them back to front / shuffled per round: a missing barrier shows up as a changed result.
    static const int order = [] { const char* e = getenv("HOSTCC_ORDER"); return !e ? 0 : !strcmp(e, "reverse") ? 1 : !strcmp(e, "random") ? 2 : 0; }();
    static unsigned long long rng = 0x9E3779B97F4A7C15ull;
    std::vector<unsigned> perm(n);
    for (unsigned i = 0; i < n; i++) perm[i] = order == 1 ? n - 1 - i : i;
    unsigned remaining = n;
    while (remaining) {
        bool ran = false;
        if (order == 2)
            for (unsigned i = n - 1; i > 0; i--) { rng = rng * 6364136223846793005ull + 1442695040888963407ull; std::swap(perm[i], perm[(unsigned)((rng >> 33) % (i + 1))]); }
        for (unsigned k = 0; k < n; k++) {
            const unsigned i = perm[k];
            Lane& l = g_lanes[i];
            if (l.done) continue;
            if (l.waitGen && *l.waitGen == l.waitAt) continue;           // still blocked
            g_cur = &l; threadIdx = l.tid;
#if HOSTCC_ASAN
            __sanitizer_start_switch_fiber(&g_schedFake, g_stacks + (size_t)i * kStackBytes, kStackBytes);
            swapcontext(&g_sched, &l.uc);
            __sanitizer_finish_switch_fiber(g_schedFake, nullptr, nullptr);
#else
            swapcontext(&g_sched, &l.uc);
#endif
            ran = true;
            if (l.done) remaining--;
        }
        if (!ran) { fprintf(stderr, "hostcc: deadlock - every live thread of block (%u,%u) waits in a barrier / collective\n", blockIdx.x, blockIdx.y); abort(); }
    }
    g_cur = nullptr;
}

}  // namespace

bool in_kernel() { return g_cur != nullptr; }
unsigned lane_id() { return g_cur ? g_cur->lane : 0u; }

void yield()
{
    if (!g_cur) return;
    g_cur->waitGen = nullptr;
    switch_out();
}

void sync_block()
{
    if (!g_cur) return;
    const unsigned long long gen = g_blockGen;
    if (++g_blockArrived == g_blockAlive) { g_blockArrived = 0; g_blockGen++; return; }
    wait_for(&g_blockGen, gen);
}

const unsigned long long* exchange(unsigned mask, unsigned long long v, unsigned* arrivedMask)
{
    static unsigned long long self[32];
    if (!g_cur) { self[0] = v; *arrivedMask = 1u; return self; }
    Warp& w = g_warps[g_cur->warp];
    const unsigned long long gen = w.gen;
    w.val[g_cur->lane] = v;
    w.arrived |= 1u << g_cur->lane;
    w.want |= mask;
    if ((w.arrived & (w.want & w.alive)) == (w.want & w.alive)) finish_rendezvous(w);
    else wait_for(&w.gen, gen);
    *arrivedMask = w.snapMask;
    return w.snap;
}

void launch(dim3 grid, dim3 block, const std::function<void()>& body)
{
    if (g_cur) { fprintf(stderr, "hostcc: nested launch\n"); abort(); }
    gridDim = grid; blockDim = block;
    g_body = &body;
    for (unsigned z = 0; z < grid.z; z++)
        for (unsigned y = 0; y < grid.y; y++)
            for (unsigned x = 0; x < grid.x; x++) {
                blockIdx = uint3{x, y, z};
                run_block(block);
            }
    g_body = nullptr;
}

}  // namespace hostcc
