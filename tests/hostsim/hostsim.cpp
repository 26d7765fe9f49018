// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE: the scheduler behind tests/hostsim/cuda_runtime.h.
//
// A launch runs its blocks one after the other; the threads of a block are cooperative fibers (ucontext) that
// run until they finish or reach a barrier (__syncthreads, or the exchange step of a warp shuffle / ballot).
// A barrier releases once every thread of the block (warp) that has not yet exited has arrived, as on the
// hardware.  If no fiber can run and some have not finished, the launch reports an error instead of hanging.
#include <sys/mman.h>
#include <ucontext.h>

#include <cstdio>
#include <unordered_map>
#include <vector>

#include "cuda_runtime.h"

namespace hostsim {

ThreadCtx* cur = nullptr;
int last_error = 0;

namespace {

constexpr size_t kStackBytes = 256 << 10;

struct Warp {
    int alive = 0;
    int count = 0;
    unsigned gen = 0;
    uint64_t slot[32];
};

struct Fiber {
    ucontext_t ctx;
    ThreadCtx tc;
    bool done = false;
    const unsigned* wait_ptr = nullptr;     // blocked while *wait_ptr == wait_val
    unsigned wait_val = 0;
    int warp = 0, lane = 0;
    int blk = 0;                            // index of this thread's CTA within the running cluster
    unsigned or_calls = 0;                  // __syncthreads_or calls so far (every thread of a CTA makes the same number)
};

struct Block {
    int alive = 0;
    int bar_count = 0;
    unsigned bar_gen = 0;
    int or_acc[3] = {0, 0, 0};              // __syncthreads_or: call n accumulates in slot n % 3 (see syncthreads_or)
    std::vector<Warp> warps;
    std::vector<char> store;                // backing store of this CTA's dynamic shared memory
    char* smem = nullptr;                   // 128-byte aligned in every CTA: the CTAs of a cluster lay it out identically,
    size_t smem_size = 0;                   // as on the hardware (cluster_map is "same offset in another CTA's window")
};

// The CTAs of a thread-block cluster run CONCURRENTLY (all their threads are fibers of one scheduling loop), each with
// its own barriers, warps and shared memory; a plain launch is a cluster of one.
struct Cluster {
    int alive = 0;
    int bar_count = 0;
    unsigned bar_gen = 0;
};

ucontext_t g_sched;
std::vector<Fiber> g_fibers;
std::vector<char*> g_stacks;
Fiber* g_self = nullptr;
std::vector<Block> g_blocks;
Cluster g_cluster;
const std::function<void()>* g_body = nullptr;
#define g_blk (g_blocks[g_self->blk])

char* stack_for(size_t i) {
    while (g_stacks.size() <= i) {
        void* p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("hostsim: mmap"); abort(); }
        g_stacks.push_back((char*)p);
    }
    return g_stacks[i];
}

void yield() { swapcontext(&g_self->ctx, &g_sched); }

void block_wait(unsigned* gen, int* count, int alive) {
    Fiber* f = g_self;
    const unsigned g = *gen;
    if (++*count >= alive) {
        *count = 0;
        ++*gen;
        return;
    }
    f->wait_ptr = gen;
    f->wait_val = g;
    yield();
}

void thread_exit(Fiber* f) {
    f->done = true;
    // a thread that leaves no longer counts towards the barriers of its block / warp
    --g_blk.alive;
    if (g_blk.bar_count > 0 && g_blk.bar_count >= g_blk.alive) {
        g_blk.bar_count = 0;
        ++g_blk.bar_gen;
    }
    --g_cluster.alive;
    if (g_cluster.bar_count > 0 && g_cluster.bar_count >= g_cluster.alive) {
        g_cluster.bar_count = 0;
        ++g_cluster.bar_gen;
    }
    Warp& w = g_blk.warps[f->warp];
    --w.alive;
    if (w.count > 0 && w.count >= w.alive) {
        w.count = 0;
        ++w.gen;
    }
}

void trampoline() {
    (*g_body)();
    thread_exit(g_self);
    swapcontext(&g_self->ctx, &g_sched);
}

}  // namespace

void* dyn_smem() { return g_blk.smem; }
unsigned lane_id() { return (unsigned)g_self->lane; }

// ---- thread-block clusters (die_b200/csrc/die_cluster.cuh) ------------------------------------------------------
unsigned cluster_rank() { return (unsigned)g_self->blk; }
void cluster_sync() { block_wait(&g_cluster.bar_gen, &g_cluster.bar_count, g_cluster.alive); }
void* cluster_map(void* p, unsigned rank) {
    char* lo = g_blk.smem;
    if ((char*)p < lo || (char*)p >= lo + g_blk.smem_size || rank >= g_blocks.size()) {
        fprintf(stderr, "hostsim: cluster_map of a pointer outside this CTA's dynamic shared memory, or of a rank outside the cluster\n");
        abort();
    }
    return g_blocks[rank].smem + ((char*)p - lo);
}

void syncthreads() { block_wait(&g_blk.bar_gen, &g_blk.bar_count, g_blk.alive); }

// __syncthreads_or: call n ORs into slot n % 3 and reads it after the barrier.  The slot of call n + 1 is cleared BEFORE
// the barrier of call n: its last user (call n - 2) was read by everybody before they arrived at barrier n - 1, and
// nobody writes it for call n + 1 before passing barrier n.
int syncthreads_or(int pred) {
    const unsigned n = g_self->or_calls++;
    g_blk.or_acc[(n + 1) % 3] = 0;
    if (pred) g_blk.or_acc[n % 3] = 1;
    syncthreads();
    return g_blk.or_acc[n % 3];
}

void warp_exchange(uint64_t mine, uint64_t* all32) {
    Warp& w = g_blk.warps[g_self->warp];
    w.slot[g_self->lane] = mine;
    block_wait(&w.gen, &w.count, w.alive);
    memcpy(all32, w.slot, sizeof(w.slot));
    block_wait(&w.gen, &w.count, w.alive);      // nobody deposits the next value before everybody has read
}

// ---- mbarrier + bulk copies (die_b200/csrc/die_async.cuh) ------------------------------------------------------
// A bulk copy is QUEUED on its barrier and its destination poisoned; the bytes move when somebody first waits on the
// barrier -- as late as the protocol allows -- so a consumer that reads without waiting sees NaNs, and a phase whose
// announced byte count does not match the copies never completes (reported as a deadlock).
namespace {
struct PendingCopy { void* dst; const void* src; uint32_t bytes; };
struct MbarState {
    int init = 0, pending = 0;
    long tx = 0;
    unsigned phase = 0;
    int waiters = 0;
    std::vector<PendingCopy> queue;
};
std::unordered_map<void*, MbarState> g_mbars;

void mbar_maybe_complete(MbarState& st) {
    if (st.pending == 0 && st.tx == 0) {
        ++st.phase;
        st.pending = st.init;
    }
}

MbarState& mbar_of(void* bar) {
    auto it = g_mbars.find(bar);
    if (it == g_mbars.end()) { fprintf(stderr, "hostsim: mbarrier used before mbarrier.init\n"); abort(); }
    return it->second;
}
}  // namespace

void mbar_init(void* bar, int count) {
    MbarState st;
    st.init = st.pending = count;
    g_mbars[bar] = st;
}

void mbar_arrive_expect_tx(void* bar, uint32_t bytes) {
    MbarState& st = mbar_of(bar);
    if (st.pending <= 0) { fprintf(stderr, "hostsim: more arrivals than the mbarrier was initialised for\n"); abort(); }
    st.tx += bytes;
    --st.pending;
    mbar_maybe_complete(st);
}

void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
    if (bytes == 0 || bytes % 16 != 0 || ((uintptr_t)dst & 15) != 0 || ((uintptr_t)src & 15) != 0) {
        fprintf(stderr, "hostsim: cp.async.bulk needs a size that is a multiple of 16 and 16-byte aligned addresses "
                        "(dst %p, src %p, %u bytes)\n", dst, src, bytes);
        abort();
    }
    char* lo = g_blk.smem;
    if ((char*)dst < lo || (char*)dst + bytes > lo + g_blk.smem_size) {
        fprintf(stderr, "hostsim: bulk copy destination outside the dynamic shared memory of the block\n");
        abort();
    }
    MbarState& st = mbar_of(bar);
    if (st.waiters > 0) {                               // somebody already waits for this phase: the bytes land now
        memcpy(dst, src, bytes);
        st.tx -= bytes;
        mbar_maybe_complete(st);
        return;
    }
    memset(dst, 0xFF, bytes);
    st.queue.push_back(PendingCopy{dst, src, bytes});
}

void mbar_wait(void* bar, uint32_t parity) {
    for (;;) {
        MbarState& st = mbar_of(bar);
        if ((st.phase & 1u) != parity) return;          // the phase with this parity has completed
        if (!st.queue.empty()) {
            for (const PendingCopy& c : st.queue) {
                memcpy(c.dst, c.src, c.bytes);
                st.tx -= c.bytes;
            }
            st.queue.clear();
            mbar_maybe_complete(st);
            continue;
        }
        ++st.waiters;
        g_self->wait_ptr = &st.phase;
        g_self->wait_val = st.phase;
        yield();
        --mbar_of(bar).waiters;
    }
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) { run_grid(grid, block, 1, smem, body); }

void run_grid(dim3 grid, dim3 block, unsigned cluster, size_t smem, const std::function<void()>& body) {
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    if (nthreads == 0 || nthreads > 1024 || grid.x == 0 || cluster == 0 || cluster > 16 || nblocks % cluster != 0 ||
        (cluster > 1 && (grid.y != 1 || grid.z != 1))) {
        last_error = cudaErrorInvalidValue;
        return;
    }
    g_body = &body;
    const size_t nfibers = nthreads * cluster;
    if (g_fibers.size() < nfibers) g_fibers.resize(nfibers);
    const int nwarps = (int)((nthreads + 31) / 32);
    // Clusters (plain launches: blocks) run one after the other.  HOSTSIM_BLOCK_ORDER = reverse | shuffle runs them in
    // another order: results that change with it mean a kernel depends on the order in which its CTAs execute, which
    // the hardware does not promise.
    const size_t nclusters = nblocks / cluster;
    std::vector<size_t> order(nclusters);
    for (size_t i = 0; i < nclusters; ++i) order[i] = i;
    static const char* mode = getenv("HOSTSIM_BLOCK_ORDER");
    if (mode != nullptr && strcmp(mode, "reverse") == 0) {
        for (size_t i = 0; i < nclusters; ++i) order[i] = nclusters - 1 - i;
    } else if (mode != nullptr && strcmp(mode, "shuffle") == 0) {
        static uint64_t state = 0x9E3779B97F4A7C15ull;
        for (size_t i = nclusters; i > 1; --i) {              // Fisher-Yates with a fixed-seed xorshift
            state ^= state << 13; state ^= state >> 7; state ^= state << 17;
            std::swap(order[i - 1], order[state % i]);
        }
    }
    g_blocks.resize(cluster);
    for (size_t oc = 0; oc < nclusters; ++oc) {
        g_mbars.clear();
        g_cluster = Cluster();
        g_cluster.alive = (int)nfibers;
        for (unsigned q = 0; q < cluster; ++q) {
            Block& blk = g_blocks[q];
            blk.alive = (int)nthreads;
            blk.bar_count = 0;
            blk.or_acc[0] = blk.or_acc[1] = blk.or_acc[2] = 0;
            blk.warps.assign(nwarps, Warp());
            blk.store.assign(smem + 16 + 128, (char)0xFF);   // NaN-poisoned: a read of unwritten shared memory shows up in the results
            blk.smem = (char*)(((uintptr_t)blk.store.data() + 127) & ~(uintptr_t)127);
            blk.smem_size = smem + 16;
            const size_t bid = order[oc] * cluster + q;
            const unsigned bx = (unsigned)(bid % grid.x), by = (unsigned)((bid / grid.x) % grid.y),
                           bz = (unsigned)(bid / ((size_t)grid.x * grid.y));
            for (size_t t = 0; t < nthreads; ++t) {
                Fiber& f = g_fibers[q * nthreads + t];
                f.done = false;
                f.wait_ptr = nullptr;
                f.or_calls = 0;
                f.blk = (int)q;
                f.warp = (int)(t / 32);
                f.lane = (int)(t % 32);
                ++blk.warps[f.warp].alive;
                f.tc.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / ((size_t)block.x * block.y))};
                f.tc.bid = uint3{bx, by, bz};
                f.tc.bdim = block;
                f.tc.gdim = grid;
                getcontext(&f.ctx);
                f.ctx.uc_stack.ss_sp = stack_for(q * nthreads + t);
                f.ctx.uc_stack.ss_size = kStackBytes;
                f.ctx.uc_link = &g_sched;
                makecontext(&f.ctx, trampoline, 0);
            }
        }
        // HOSTSIM_THREAD_ORDER = reverse: the fibers are scheduled from the last thread down.  A missing barrier
        // between a producer and a consumer phase can go unnoticed when producers happen to run first.
        static const char* torder = getenv("HOSTSIM_THREAD_ORDER");
        const bool treverse = torder != nullptr && strcmp(torder, "reverse") == 0;
        size_t remaining = nfibers;
        while (remaining > 0) {
            bool progressed = false;
            for (size_t tt = 0; tt < nfibers; ++tt) {
                const size_t t = treverse ? nfibers - 1 - tt : tt;
                Fiber& f = g_fibers[t];
                if (f.done) continue;
                if (f.wait_ptr != nullptr && *f.wait_ptr == f.wait_val) continue;
                f.wait_ptr = nullptr;
                g_self = &f;
                cur = &f.tc;
                swapcontext(&g_sched, &f.ctx);
                progressed = true;
                if (f.done) --remaining;
            }
            if (!progressed) {                  // every live thread waits at a barrier that cannot complete
                last_error = cudaErrorLaunchFailure;
                g_self = nullptr;
                cur = nullptr;
                return;
            }
        }
    }
    g_self = nullptr;
    cur = nullptr;
}

// "Device" allocations are electric-fence style: the buffer ENDS at a page boundary followed by an inaccessible page,
// and an inaccessible page precedes its first page, so a kernel that reads or writes past the end of an array (or
// before its first page) faults on the spot -- the emulator's stand-in for compute-sanitizer's memcheck.
namespace {
constexpr size_t kPage = 4096;
struct Mapping { void* base; size_t bytes; };
std::unordered_map<void*, Mapping>& mappings() {
    static std::unordered_map<void*, Mapping> m;
    return m;
}
}  // namespace

void* dev_alloc(size_t bytes) {
    const size_t need = ((bytes ? bytes : 1) + 15) & ~(size_t)15;          // keep 16-byte alignment of the start
    const size_t body = (need + kPage - 1) / kPage * kPage;
    const size_t total = body + 2 * kPage;
    char* base = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) return nullptr;
    mprotect(base, kPage, PROT_NONE);
    mprotect(base + kPage + body, kPage, PROT_NONE);
    memset(base + kPage, 0xCD, body);           // cudaMalloc does not zero either
    char* p = base + kPage + body - need;
    mappings()[p] = Mapping{base, total};
    return p;
}

void dev_free(void* p) {
    if (p == nullptr) return;
    auto it = mappings().find(p);
    if (it == mappings().end()) { fprintf(stderr, "hostsim: free of a pointer that was never allocated\n"); abort(); }
    munmap(it->second.base, it->second.bytes);
    mappings().erase(it);
}

}  // namespace hostsim

// numpy-side arrays (medium, agents, action ...) can live in fenced memory too: tests/hostsim/sim.py
extern "C" void* hostsim_alloc(size_t bytes) { return hostsim::dev_alloc(bytes); }
extern "C" void hostsim_free(void* p) { hostsim::dev_free(p); }
