// die_forward_memo.cuh -- PhysarumAgent.forward (core/agent/gradient.py:96-124, 168-208) for the steady-state configuration,
// with the float64 trigonometry of a slot looked up instead of evaluated.
//
// The LEAN forward kernel is bound by its instruction stream (~350 warp-instructions per slot), and ~125 of
// them evaluate functions of ONE number, the slot's heading theta: the float32 sin / cos for the sensed cell (30), and for
// the new heading theta' = renormalize(theta +- turn) the float64 sin, cos and angle (16 + 79) that become the action.
// Headings of a discrete-turn agent live on a lattice: multiples of the turn angle plus a few ulps of accumulated rounding
// -- a 256 x 256 environment holds ~200 distinct values after 100 steps, ~520 after 1000, ~630 after 2000.  So every CTA
// keeps a table in shared memory
//      theta (bits)  ->  { heading, sin, cos } of renormalize(theta - turn) and of renormalize(theta + turn)
// filled by the same device functions the LEAN kernel calls (die_renormalize_radians, die_sincos_angle), so a looked-up
// value IS the evaluated value, bit for bit.  sin / cos of theta itself (float32, for the sensed cell and the guard-banded
// turn decision) follow from the two entries: sin(theta - t) + sin(theta + t) = 2 sin(theta) cos(t).
//
// CTAs are persistent (four of 256 threads per SM, as the LEAN kernel, each with a 768-entry table of 56-byte entries =
// 42 KB) and walk over the 2048-slot chunks of gradient_forward_kernel with a grid stride: (chunk, thread, item) -> slot is
// that kernel's mapping, so the Philox coin of a slot is the same bit.  During its first kMemoWarmChunks chunks a CTA
// LEARNS: it evaluates everything in line, exactly as the LEAN kernel, and inserts the headings it meets (atomicCAS on the
// key, linear probing, entries are never replaced) -- nobody reads the table.  One block barrier later the table is
// frozen and read-only for the rest of the launch (~200 chunks): no reader ever runs beside a writer.  Whatever the
// frozen table does not hold (a rare heading, |theta| > 64, NaN) takes the in-line arithmetic.
#pragma once
#include "die_agent_kernels.cuh"

namespace die {

constexpr int kMemoEntries = 768;
constexpr int kMemoProbes = 4;
#if defined(DIE_HOSTSIM)
constexpr int kMemoWarmChunks = 1;                  // (the emulator runs 8 CTAs: freeze early so that tests exercise the lookups
                                                    //  AND the in-line path for headings the first chunk did not hold)
#else
constexpr int kMemoWarmChunks = 8;                  // 16 384 slots of (usually) eight environments
#endif
constexpr unsigned long long kMemoEmpty = 0xFFFFFFFFFFFFFFFFull;    // a NaN: never the key of an inserted entry

struct MemoEntry {
    unsigned long long key;                         // bits of theta
    double h[2], s[2], c[2];                        // [0]: theta - turn, [1]: theta + turn (renormalised): new heading, sin, cos
};
constexpr size_t kMemoSmemBytes = sizeof(MemoEntry) * kMemoEntries;

__device__ __forceinline__ unsigned memo_hash(unsigned long long bits) {
    const unsigned x = ((unsigned)bits ^ ((unsigned)(bits >> 32) * 0x85EBCA6Bu)) * 0x9E3779B1u;
    return __umulhi(x, (unsigned)kMemoEntries);     // [0, kMemoEntries)
}

__device__ __forceinline__ bool memo_cacheable(double th) { return fabs(th) <= DIE_SINCOSF_MAX; }   // (false for NaN)

// -> index of theta's entry, or -1 (frozen table only)
__device__ __forceinline__ int memo_find(const MemoEntry* tab, double th) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(th);
    unsigned e = memo_hash(bits);
#pragma unroll
    for (int probe = 0; probe < kMemoProbes; ++probe) {
        const unsigned long long k = tab[e].key;
        if (k == bits) return (int)e;
        if (k == kMemoEmpty) return -1;
        e = (e + 1 == kMemoEntries) ? 0u : e + 1;
    }
    return -1;
}

// Learning phase only (no reader is active): claim a free slot of theta's probe window and fill it.
__device__ __noinline__ void memo_learn(MemoEntry* tab, double th, double turn_radians) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(th);
    unsigned e = memo_hash(bits);
    for (int probe = 0; probe < kMemoProbes; ++probe) {
        const unsigned long long old = atomicCAS(&tab[e].key, kMemoEmpty, bits);
        if (old == bits) return;                    // somebody else owns this heading's entry (and fills it)
        if (old == kMemoEmpty) {
#pragma unroll
            for (int sgn = 0; sgn < 2; ++sgn) {
                const double turn = sgn ? turn_radians : -turn_radians;
                double s, c, h;
                die_sincos_angle(renormalize_radians(th + turn), &s, &c, &h);
                tab[e].h[sgn] = h;
                tab[e].s[sgn] = s;
                tab[e].c[sgn] = c;
            }
            return;
        }
        e = (e + 1 == kMemoEntries) ? 0u : e + 1;
    }
}

// FH: the food under the agent per slot from the last step's feed kernel (pair mode, die_api.cu) instead of a gather.
template <bool FH>
__global__ void __launch_bounds__(kAgentThreads, 4)
physarum_forward_memo_kernel(const GradientArgs a, const int total_chunks, const double inv_2cos_turn) {
    extern __shared__ unsigned long long memo_smem[];
    MemoEntry* tab = (MemoEntry*)memo_smem;
    for (int e = threadIdx.x; e < kMemoEntries; e += kAgentThreads) tab[e].key = kMemoEmpty;
    __syncthreads();

    const die_gradient_params_t& p = a.p;
    const Axis ax = a.ax, ay = a.ay;
    const int64_t M = a.M;
    const int64_t C = (int64_t)a.H * a.W;
    const int H = a.H, W = a.W;
    const uint64_t step = (a.step_dev != nullptr) ? *a.step_dev : a.step;
    const double atol = p.turn_radians * p.turn_tolerance;
    const bool quick_sense = a.sense_guard_x > 0.0;

    int learned = 0;
    for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x, ++learned) {
        const bool learning = learned < kMemoWarmChunks;
        const unsigned b = (unsigned)chunk / (unsigned)a.nchunk;
        const int64_t first = (int64_t)((unsigned)chunk - b * (unsigned)a.nchunk) * (kAgentThreads * kFwdItems) + threadIdx.x;
        const double* ag_x = a.agents + (int64_t)b * 4 * M + first;
        const double* ag_y = ag_x + M;
        const double* food = a.medium + ((int64_t)b * 3 + 1) * C;
        const double* chem = a.medium + ((int64_t)b * 3 + 2) * C;
        double* th_p = a.theta + (int64_t)b * M + first;
        double* ab = a.action + (int64_t)b * 3 * M + first;
        double* ab_y = ab + M;
        double* ab_dep = ab + 2 * M;
        const float2* grad32 = a.grad32 + (int64_t)b * C;
        const int32_t* cl_p = a.cells + (int64_t)b * M + first;
        const double* fh_p = FH ? a.food_here + (int64_t)b * M + first : nullptr;
        // coin of slot (chunk, thread, k) = bit k of this word, as in gradient_forward_kernel
        const uint32_t coin_bits = philox_draw(a.seed, step, ((uint64_t)chunk + (uint64_t)a.b0 * a.nchunk) * kAgentThreads + threadIdx.x, 2u).x;
        const int left = (int)((M - first < (int64_t)kFwdItems * kAgentThreads) ? (M - first) : (int64_t)kFwdItems * kAgentThreads);

        bool nvalid = left > 0;
        double nx = 0.0, ny = 0.0, nth = 0.0, nfh = 0.0;
        int ncell = 0;
        if (nvalid) {
            nx = ag_x[0];
            ny = ag_y[0];
            nth = th_p[0];
            if (FH) nfh = fh_p[0];
            else ncell = cl_p[0];
        }
        for (int k = 0; k < kFwdItems; ++k) {
            if (!nvalid) break;
            const int i = k * kAgentThreads;
            const double x = nx, y = ny, th = nth;
            const double food_here = FH ? nfh : food[ncell];
            nvalid = (k + 1 < kFwdItems) && (i + kAgentThreads < left);
            if (nvalid) {                                          // next item's coalesced loads
                nx = ag_x[i + kAgentThreads];
                ny = ag_y[i + kAgentThreads];
                nth = th_p[i + kAgentThreads];
                if (FH) nfh = fh_p[i + kAgentThreads];
                else ncell = cl_p[i + kAgentThreads];
            }
            const bool cacheable = memo_cacheable(th);
            int en = -1;
            if (cacheable) {
                if (learning) memo_learn(tab, th, p.turn_radians);
                else en = memo_find(tab, th);
            }

            // sin / cos of the heading in float32 for the sensed cell and the guard-banded decision (see gradient_forward_kernel)
            float sf, cf;
            int sx, sy;
            bool cell_ok = false;
            double s_lo = 0.0, s_hi = 0.0, c_lo = 0.0, c_hi = 0.0;
            if (en >= 0) {
                s_lo = tab[en].s[0]; s_hi = tab[en].s[1];
                c_lo = tab[en].c[0]; c_hi = tab[en].c[1];
            }
            if (quick_sense && cacheable) {
                if (en >= 0) {
                    sf = (float)((s_lo + s_hi) * inv_2cos_turn);
                    cf = (float)((c_lo + c_hi) * inv_2cos_turn);
                } else {
                    die_sincosf_approx(th, &sf, &cf);
                }
                const bool okx = nearest_cell_guarded(x + p.sense_offset * (double)cf, ax, a.sense_guard_x, &sx);
                const bool oky = nearest_cell_guarded(y + p.sense_offset * (double)sf, ay, a.sense_guard_y, &sy);
                cell_ok = okx & oky;
            }
            if (!cell_ok) {
                double sn, cs;
                die_sincos(th, &sn, &cs);
                sx = nearest_cell(x + p.sense_offset * cs, ax);
                sy = nearest_cell(y + p.sense_offset * sn, ay);
                sf = (float)sn;
                cf = (float)cs;
            }
            const int sc = sx * W + sy;
            const float2 g2 = grad32[sc];

            die_turn_t tr;
            if (!die_turn_quick_ff(&a.plan, g2.x, g2.y, sf, cf, th, atol, p.sense_radians, &tr)) {
                double gx, gy;
                sample_gradient(chem, sx, sy, H, W, gx, gy);                 // the exact path wants all 53 bits
                tr = turn_exact_call(gx, gy, th, atol, p.sense_radians, 1, p.use_grad_clip, p.grad_clip);
            }
            const int c = (int)((coin_bits >> k) & 1u);
            const int tsign = (tr.turn != 0) ? tr.turn : 2 * c - 1;
            double s2, c2, heading;
            if (en >= 0) {
                const int sgn = tsign > 0;
                heading = tab[en].h[sgn];
                s2 = sgn ? s_hi : s_lo;
                c2 = sgn ? c_hi : c_lo;
            } else {
                const double turn = (tsign < 0) ? -p.turn_radians : p.turn_radians;
                die_sincos_angle(renormalize_radians(th + turn), &s2, &c2, &heading);
            }
            th_p[i] = heading;
            double dep = p.deposit * food_here;
            dep = dep * ((tr.deposit_mask != 0) ? 1.0 : 0.1);
            ab[i] = c2 * p.scale;
            ab_y[i] = s2 * p.scale;
            ab_dep[i] = dep;
        }
        // the table freezes after the last learning chunk: every insert is complete before the first lookup
        if (learned == kMemoWarmChunks - 1) __syncthreads();
    }
}

}  // namespace die
