// die_agent_kernels.cuh -- per-slot kernels: policies (Brownian / Const / Gradient / Physarum)
// and the agent half of Env.step (move + claim, feed + reward partials, final reduction).
//
// One thread handles a few slots of one environment (a CTA owns a chunk of consecutive slots, see
// slot_chunk); every per-slot array is channel-major ([B][ch][M]) so consecutive threads read
// consecutive doubles (coalesced, 256 B per warp per channel).  Field accesses are data-dependent
// gathers (one 32 B sector each).  Citations are file:line under /root/reference.
#pragma once
#include "die_device.cuh"
#include "die_slab.cuh"
#include "../../include/die_b200.h"

namespace die {

constexpr int kAgentThreads = 256;

// Per-slot kernels cut every environment's M slots into chunks of kAgentThreads * ITEMS
// consecutive slots, one chunk per CTA (thread t handles slots base + k*256 + t): coalesced,
// no 64-bit division per slot, and a fixed slot -> (CTA, thread, k) map that the in-kernel RNG
// is keyed on (so the random stream does not depend on the launch geometry or the SM count).
struct SlotChunk {
    int64_t b;        // environment
    int64_t base;     // first slot of the chunk within the environment
};

template <int ITEMS>
__device__ __forceinline__ SlotChunk slot_chunk(int nchunk) {
    SlotChunk c;
    const unsigned blk = blockIdx.x;
    const unsigned b = blk / (unsigned)nchunk;
    c.b = b;
    c.base = (int64_t)(blk - b * (unsigned)nchunk) * (kAgentThreads * ITEMS);
    return c;
}

static inline int chunks_for(int64_t M, int items) {
    return (int)((M + (int64_t)kAgentThreads * items - 1) / ((int64_t)kAgentThreads * items));
}

// ---------------------------------------------------------------------------------------------
// BrownianAgent.forward  (core/agent/static.py:40-50; core/data_init.py:159-169,218-220,248-253)
//   chan = ((b - a) * round(u, 3) + a) * alive,  u drawn in the order dx, dy, deposit1.
// ---------------------------------------------------------------------------------------------
constexpr int kBrownItems = 4;

__global__ void __launch_bounds__(kAgentThreads)
brownian_forward_kernel(const double* __restrict__ agents, double* __restrict__ action,
                        int64_t M, int nchunk, double s, double dep_scale,
                        const double* __restrict__ u, uint64_t seed, uint64_t step,
                        const uint64_t* __restrict__ step_dev) {
    // step_dev: the call counter lives in device memory (a CUDA-graph replay cannot change a kernel argument)
    if (step_dev != nullptr) step = *step_dev;
    const SlotChunk ch = slot_chunk<kBrownItems>(nchunk);
    const double* alive_p = agents + (ch.b * 4 + 2) * M;
    const double* ub = (u != nullptr) ? u + ch.b * 3 * M : nullptr;
    double* ab = action + ch.b * 3 * M;
    const double span = s - (-s);                           // (b - a) with a = -s, b = s
#pragma unroll
    for (int k = 0; k < kBrownItems; ++k) {
        const int64_t i = ch.base + k * kAgentThreads + threadIdx.x;
        if (i >= M) break;
        const double alive = alive_p[i];
        double u0, u1, u2;
        if (ub != nullptr) {
            u0 = ub[i];
            u1 = ub[M + i];
            u2 = ub[2 * M + i];
        } else {                                            // one Philox block = 3 x 32-bit uniforms
            const uint4 r = philox_draw(seed, step, (uint64_t)(ch.b * M + i), 0u);
            u0 = u32(r.x);
            u1 = u32(r.y);
            u2 = u32(r.z);
        }
        // np.round(u, 3) == rint(u * 1000) / 1000   
        const double q0 = rint(u0 * 1000.0) / 1000.0;
        const double q1 = rint(u1 * 1000.0) / 1000.0;
        const double q2 = rint(u2 * 1000.0) / 1000.0;
        ab[i]         = (span * q0 + (-s)) * alive;
        ab[M + i]     = (span * q1 + (-s)) * alive;
        ab[2 * M + i] = ((dep_scale - 0.0) * q2 + 0.0) * alive;
    }
}

// ConstAgent.forward (core/agent/static.py:19-28)
__global__ void __launch_bounds__(kAgentThreads)
const_forward_kernel(double* __restrict__ action, int64_t M, int nchunk,
                     double dx, double dy, double dep) {
    const SlotChunk ch = slot_chunk<kBrownItems>(nchunk);
    double* ab = action + ch.b * 3 * M;
#pragma unroll
    for (int k = 0; k < kBrownItems; ++k) {
        const int64_t i = ch.base + k * kAgentThreads + threadIdx.x;
        if (i >= M) break;
        ab[i] = dx;
        ab[M + i] = dy;
        ab[2 * M + i] = dep;
    }
}

// ---------------------------------------------------------------------------------------------
// JonesAgent.forward -- the classic three-sensor Physarum particle (Jones 2010) on the reference's Env protocol.
// NOT a reference class ("optional ... not parity-checkable -- the reference has none"): the
// specification is oracle/die_ref.py:JonesAgent, whose every operation this kernel repeats in the same order --
//   sensors FL, F, FR at heading + SA, heading, heading - SA, distance `sense_offset`, each reading chem1 at the
//   nearest cell of its position, CLAMPED like every sense lookup of the reference (core/utils.py:39-54);
//   F > FL and F > FR: straight on;  F < FL and F < FR: +-RA by a coin;  FL < FR: -RA;  FR < FL: +RA;  else straight on;
//   heading' = renormalize_radians(heading + turn) (core/utils.py:177-179); action = (scale cos, scale sin) of heading',
//   deposit * food under the agent -- unmasked, all M slots, like PhysarumAgent (core/agent/gradient.py:113-124, Q8).
// Comparisons of float64 values gathered from identical cells: bit-exact given die_math.h's sin / cos on both sides.
// ---------------------------------------------------------------------------------------------
constexpr int kJonesItems = 4;

struct JonesArgs {
    Axis ax, ay;
    int W;
    int64_t M, C;
    int nchunk;
    const double* agents;       // [B][4][M]
    const double* medium;       // [B][3][H][W] (FT elements)
    double* theta;              // [B][M]
    double* action;             // [B][3][M]
    const uint8_t* coin;        // may be null -> Philox
    double scale, deposit, sense_offset, sense_radians, turn_radians;
    uint64_t seed, step;
    const uint64_t* step_dev;   // may be null; else the call counter is read from device memory (CUDA-graph replays)
};

template <typename FT>
__global__ void __launch_bounds__(kAgentThreads)
jones_forward_kernel(const JonesArgs a) {
    const SlotChunk ch = slot_chunk<kJonesItems>(a.nchunk);
    const int64_t M = a.M;
    const int W = a.W;
    const double* ag = a.agents + ch.b * 4 * M;
    const FT* food = (const FT*)a.medium + (ch.b * 3 + 1) * a.C;
    const FT* chem = (const FT*)a.medium + (ch.b * 3 + 2) * a.C;
    double* th_p = a.theta + ch.b * M;
    double* ab = a.action + ch.b * 3 * M;
    const uint8_t* coin_p = (a.coin != nullptr) ? a.coin + ch.b * M : nullptr;
    const uint64_t step = (a.step_dev != nullptr) ? *a.step_dev : a.step;
#pragma unroll 1
    for (int k = 0; k < kJonesItems; ++k) {
        const int64_t i = ch.base + k * kAgentThreads + threadIdx.x;
        if (i >= M) break;
        const double x = ag[i], y = ag[M + i], th = th_p[i];
        const double food_here = (double)food[nearest_cell(x, a.ax) * W + nearest_cell(y, a.ay)];
        double v[3];                                        // FL, F, FR
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const double ang = (s == 0) ? th + a.sense_radians : (s == 1 ? th : th - a.sense_radians);
            double sn, cs;
            die_sincos(ang, &sn, &cs);
            const int sx = nearest_cell(x + a.sense_offset * cs, a.ax);
            const int sy = nearest_cell(y + a.sense_offset * sn, a.ay);
            v[s] = (double)chem[sx * W + sy];
        }
        double turn = 0.0;
        if (v[1] > v[0] && v[1] > v[2]) {
            turn = 0.0;
        } else if (v[1] < v[0] && v[1] < v[2]) {
            const int c = (coin_p != nullptr) ? (coin_p[i] ? 1 : 0)
                                              : (int)(philox_draw(a.seed, step, (uint64_t)(ch.b * M + i), 5u).x & 1u);
            turn = c ? a.turn_radians : -a.turn_radians;
        } else if (v[0] < v[2]) {
            turn = -a.turn_radians;
        } else if (v[2] < v[0]) {
            turn = a.turn_radians;
        }
        const double heading = renormalize_radians(th + turn);
        double s2, c2;
        die_sincos(heading, &s2, &c2);
        th_p[i] = heading;
        ab[i] = c2 * a.scale;
        ab[M + i] = s2 * a.scale;
        ab[2 * M + i] = a.deposit * food_here;
    }
}

// ---------------------------------------------------------------------------------------------
// GradientAgent.forward / PhysarumAgent.forward  (core/agent/gradient.py:96-124)
//
// The reference materialises the normalised gradient of the whole chem1 field
// (_get_gradient, :55-71) and then samples it at ONE cell per slot.  Here the np.gradient
// stencil is evaluated only at the sampled cell: identical arithmetic per sample, 4 chem
// gathers instead of ~10 full-field passes.  sin / cos / atan2 are die_math.h's
// bit-reproducible routines (see that file for why).
// ---------------------------------------------------------------------------------------------
constexpr int kFwdItems = 8;     // one 32-bit Philox word serves a thread's 8 coin flips
// MINB (template): minimum resident CTAs per SM = register cap 65536 / (256 * MINB): 3 -> 80, 4 -> 64, 5 -> 48

struct GradientArgs {
    die_gradient_params_t p;
    die_turn_plan_t plan;       // guard-banded thresholds of the quick turn decision (die_turn.h)
    Axis ax, ay;                // axis 0 (H cells, coordinate x) and axis 1 (W cells, coordinate y)
    int H, W;
    int64_t M;
    int nchunk;
    const double* agents;
    const double* medium;
    double* theta;
    double* prev_grad;          // may be null (inertia == 0 && noise_scale == 0)
    double* action;
    const uint8_t* coin;        // may be null -> Philox
    const double* noise;        // may be null -> Philox Box-Muller (only if noise_scale != 0)
    int32_t* sense_cells;       // may be null
    const double2* grad;        // may be null: np.gradient(chem1) per cell, published by Env.step
    const float2* grad32;       // may be null: the same, rounded to float32 (only offered when the gradient is used for
                                // the guard-banded turn DECISION alone: discrete turn, plan enabled, normalised)
    const int32_t* cells;       // may be null: linear cell of every slot, cached by Env.step
    const double* food_here;    // FH instantiation: the env_food under every slot, gathered by the last step's feed kernel
    uint64_t seed, step;
    const uint64_t* step_dev;   // may be null; else the call counter is read from device memory (CUDA-graph replays)
    double sense_guard_x, sense_guard_y;   // > 0: the sense position may be formed with the float32 sin / cos of
                                // die_sincosf_approx; a sensed cell within this many cells of a cell boundary is re-evaluated
                                // with die_sincos (|sense_offset| DIE_SINCOSF_ERR (n - 1) + 1e-9); 0: always die_sincos
    int b0;                     // the launch covers environments [b0, b0 + B') of a larger batch (pointers already offset):
                                // only the in-kernel RNG needs to know, so that chunked launches draw the same numbers
    // MOVE instantiation: Env._agent_move + the claim, evaluated speculatively for the action being written
    int32_t* winner;            // [B][H*W] claim table
    int32_t* cells_out;         // [B][M] post-move cell of every slot (the env's OTHER cell buffer)
    double* burned_out;         // may be null: [B][M] out, linear_action_cost of the action being written (cost hint)
    double cost_w_dep, cost_w_dist;
    die_sqrt_near_t cost_sqrt;  // enabled where |(dx, dy)| = |scale| to a few ulps (a unit direction): die_math.h
    double* commit_xy;          // null: the move stays speculative (the feed kernel commits the positions once the step
                                // adopts it).  Else = `agents`, writable: the moved x, y are stored in place by THIS
                                // launch (DIE_FWD_COMMIT_MOVE: the caller promises that very action to the next step)
    const uint32_t* alive_bits; // [B][Mw] bit i&31 of word i>>5 = (alive[i] > 0)
    int64_t Mw;
    int boundary;
    SlabGeom sg;                // SLAB instantiation only: H, W above are the GLOBAL field, M the LOCAL slots
    SlabTables st;
};

// Env._agent_move_handle_boundary (core/env.py:152-161) for one coordinate
__device__ __forceinline__ double apply_boundary(double v, int boundary) {
    if (boundary == DIE_BOUNDARY_WRAP) return mod1(v);
    if (boundary == DIE_BOUNDARY_LIMIT) return fmin(fmax(v, 0.0), 1.0);          // np.clip(0., 1.)
    return v;
}

// Software pipeline: the coalesced loads (x, y, theta, cached cell) of item k+1 are issued before the
// arithmetic of item k, and the food gather of item k right at its start, so the only exposed memory
// latency per item is the gradient gather at the sensed cell.
// The reference arithmetic of the turn decision as an out-of-line call: it runs for about one warp in 300
// (die_turn_quick defers to it), and kept out of line its registers (atan2, sqrt, two divisions) do not count
// against the hot path of the LEAN kernel.
__device__ __noinline__ die_turn_t turn_exact_call(double gx, double gy, double th, double atol, double sense_radians,
                                                   int normalized, int use_clip, double clip) {
    die_normalize_gradient(&gx, &gy, normalized, use_clip, clip);
    return die_turn_exact(gx, gy, th, atol, sense_radians);
}

// np.gradient of chem at cell (sx, sy): central (f[i+1] - f[i-1]) / 2 inside, one-sided f[1] - f[0] /
// f[n-1] - f[n-2] at the edges, non-periodic (Q5): clamped neighbours give both forms
template <typename FT>
__device__ __forceinline__ void sample_gradient(const FT* __restrict__ chem, int sx, int sy, int H, int W,
                                                double& gx, double& gy) {
    const int sc = sx * W + sy;
    const int xm = (sx > 0) ? -W : 0, xp = (sx < H - 1) ? W : 0;
    const int ym = (sy > 0) ? -1 : 0, yp = (sy < W - 1) ? 1 : 0;
    gx = (double)chem[sc + xp] - (double)chem[sc + xm];
    gy = (double)chem[sc + yp] - (double)chem[sc + ym];
    if (xp - xm == 2 * W) gx *= 0.5;      // central difference: / 2.0 exactly
    if (yp - ym == 2) gy *= 0.5;
}

// LEAN: the steady-state configuration of the Physarum loop is known at compile time -- in-kernel Philox coins, no
// momentum state, no recorded sense cells, the env's published gradient and cell cache valid -- so the per-item null
// checks and parameter reloads of the general kernel fold away.  Same arithmetic, same results.
// G32 (LEAN only): the published gradient is the float32 one.  die_turn_quick rounds the gradient to float32 first
// thing, so every decision it settles is the same; a slot it defers re-samples the float64 gradient from chem1.
// FT: element type of the medium the agent observes (float64, or float32 in the env's float32 field mode): gathered
// values are widened, all arithmetic stays float64.
// FH (LEAN only): the food under the agent comes per SLOT from the array the last step's feed kernel filled (it gathered at
// that very cell anyway) -- a coalesced 8-byte load instead of a random gather.  For one LARGE field, whose ghost slots
// are spread over tables far larger than L2, every random 8-byte gather costs ~80-90 B of DRAM traffic.
template <bool DISCRETE_TURN, bool SLAB, bool MOVE, int MINB, bool LEAN = false, bool G32 = false, typename FT = double,
          bool FH = false>
__global__ void __launch_bounds__(kAgentThreads, MINB)
gradient_forward_kernel(const GradientArgs a) {
    static_assert(!SLAB || sizeof(FT) == 8, "the slab decomposition runs float64 fields");
    static_assert(!FH || LEAN, "food_here is a hint of the steady-state instantiation");
    const die_gradient_params_t& p = a.p;
    const Axis ax = a.ax, ay = a.ay;
    const int64_t M = a.M;
    const int64_t C = (int64_t)a.H * a.W;
    const int H = a.H, W = a.W;
    const SlotChunk ch = slot_chunk<kFwdItems>(a.nchunk);
    const int64_t first = ch.base + threadIdx.x;          // this thread's slots: first + k*256

    // per-thread base pointers, advanced by kAgentThreads per item
    const double* ag_x = a.agents + ch.b * 4 * M + first;
    const double* ag_y = ag_x + M;       // (own base pointers: `ag_x[M + i]` costs a 64-bit index sum per access)
    const FT* food = (const FT*)a.medium + (ch.b * 3 + 1) * C;
    const FT* chem = (const FT*)a.medium + (ch.b * 3 + 2) * C;
    double* th_p = a.theta + ch.b * M + first;
    double* ab = a.action + ch.b * 3 * M + first;
    double* ab_y = ab + M;
    double* ab_dep = ab + 2 * M;
    double* pg = (!LEAN && a.prev_grad != nullptr) ? a.prev_grad + ch.b * 2 * M + first : nullptr;
    const uint8_t* coin_p = (!LEAN && a.coin != nullptr) ? a.coin + ch.b * M + first : nullptr;
    const double* nz = (!LEAN && a.noise != nullptr) ? a.noise + ch.b * 2 * M + first : nullptr;
    int32_t* sc_p = (!LEAN && a.sense_cells != nullptr) ? a.sense_cells + ch.b * M + first : nullptr;
    const double2* grad = ((LEAN && !G32) || a.grad != nullptr) ? a.grad + ch.b * C : nullptr;
    const float2* grad32 = (G32 || (!LEAN && !SLAB && a.grad32 != nullptr)) ? a.grad32 + ch.b * C : nullptr;
    const int32_t* cl_p = (LEAN || a.cells != nullptr) ? a.cells + ch.b * M + first : nullptr;
    const double* fh_p = FH ? a.food_here + ch.b * M + first : nullptr;
    int32_t* win = MOVE ? a.winner + ch.b * C : nullptr;
    int32_t* co_p = MOVE ? a.cells_out + ch.b * M + first : nullptr;
    double* cx_p = (MOVE && a.commit_xy != nullptr) ? a.commit_xy + ch.b * 4 * M + first : nullptr;
    double* bo_p = (a.burned_out != nullptr) ? a.burned_out + ch.b * M + first : nullptr;
    // all 32 slots of a warp-item share one word of the alive bitmask (first - lane is a multiple of 32)
    const uint32_t* bits_p = MOVE ? a.alive_bits + ch.b * a.Mw + (first >> 5) : nullptr;

    const uint64_t step = (a.step_dev != nullptr) ? *a.step_dev : a.step;
    uint32_t coin_bits = 0;
    if (DISCRETE_TURN && (LEAN || coin_p == nullptr))      // coin of slot (CTA, t, k) = bit k of this word
        coin_bits = philox_draw(a.seed, step,
                                ((uint64_t)blockIdx.x + (uint64_t)a.b0 * a.nchunk) * kAgentThreads + threadIdx.x, 2u).x;
    const double atol = p.turn_radians * p.turn_tolerance;
    // with an identity momentum step (no inertia, no noise) and a unit-length direction the new
    // heading angle(cos d + i sin d) comes out of die_sincos_angle together with cos d, sin d
    const bool fused_heading = LEAN || (DISCRETE_TURN && pg == nullptr && p.normalized_grad);

    // slots left from this thread's first one (32 bits: one register carried through the loop instead of the 64-bit
    // `first`, which the compiler otherwise rebuilds from blockIdx for every item's bounds check)
    const int left = (int)((M - first < (int64_t)kFwdItems * kAgentThreads) ? (M - first) : (int64_t)kFwdItems * kAgentThreads);
    bool nvalid = left > 0;
    double nx = 0.0, ny = 0.0, nth = 0.0, nfh = 0.0;
    int ncell = 0;
    if (nvalid) {
        nx = ag_x[0];
        ny = ag_y[0];
        nth = th_p[0];
        if (FH) nfh = fh_p[0];
        else if (LEAN || cl_p != nullptr) ncell = cl_p[0];
    }

    for (int k = 0; k < kFwdItems; ++k) {
        if (!nvalid) break;
        const int i = k * kAgentThreads;                       // offset from this thread's first slot
        const double x = nx, y = ny, th = nth;
        // food under the agent (:113-115), issued first: independent of the turn arithmetic
        double food_here;
        if (FH) {
            food_here = nfh;
        } else {
            const int here = (LEAN || cl_p != nullptr) ? ncell : nearest_cell(x, ax) * W + nearest_cell(y, ay);
            food_here = SLAB ? slab_load_food(a.st, a.sg, here) : (double)food[here];
        }
        uint32_t alive_word = 0;
        if (MOVE) alive_word = bits_p[i >> 5];
        nvalid = (k + 1 < kFwdItems) && (i + kAgentThreads < left);
        if (nvalid) {                                          // next item's coalesced loads
            nx = ag_x[i + kAgentThreads];
            ny = ag_y[i + kAgentThreads];
            nth = th_p[i + kAgentThreads];
            if (FH) nfh = fh_p[i + kAgentThreads];
            else if (LEAN || cl_p != nullptr) ncell = cl_p[i + kAgentThreads];
        }

        // _sense_offset (:73-76): polar2xy(r, theta) = (r cos, r sin); field_by_agents(grad_field, offset) (:105): nearest,
        // CLAMPED not wrapped (Q4).  sin / cos of the heading are needed for the sensed CELL and for the guard-banded
        // turn decision only, so a float32 evaluation (+-4e-7) serves both wherever the cell is not within the guard of a
        // cell boundary; elsewhere (about one slot in 10^5) the float64 die_sincos decides, as it used to for every slot.
        // (Tried in round 2 and not kept: running this float32 sense evaluation ONE ITEM AHEAD, so that the next item's
        //  gradient gather is in flight during the current item's float64 sin / cos -- 64 registers with spills, and
        //  slower in every regime: 4.67 vs 4.36 ms batched, 0.411 vs 0.380 ms at 4096^2 early, 0.672 vs 0.641 ms at its
        //  steady state; profiles/r02s_forward_gather_one_item_ahead_not_kept.txt.)
        float sf, cf;
        int sx, sy;
        bool cell_ok = false;
        if (a.sense_guard_x > 0.0 && fabs(th) <= DIE_SINCOSF_MAX) {
            die_sincosf_approx(th, &sf, &cf);
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

        // np.gradient at (sx, sy): central (f[i+1] - f[i-1]) / 2 inside, one-sided f[1] - f[0] /
        // f[n-1] - f[n-2] at the edges, non-periodic (Q5): clamped neighbours give both forms
        const int sc = sx * W + sy;                                // H*W < 2^31 (die_env_create)
        if (!LEAN && sc_p != nullptr) sc_p[i] = sc;
        double gx, gy;
        float gxf = 0.f, gyf = 0.f;
        const bool from32 = G32 || (!LEAN && grad32 != nullptr);
        if (from32) {                                                    // published as float32: one 8-byte gather
            const float2 g2 = grad32[sc];
            gxf = g2.x;
            gyf = g2.y;
            gx = (double)g2.x;
            gy = (double)g2.y;
        } else if (LEAN || (SLAB ? a.st.grad != nullptr : grad != nullptr)) {   // published by the field pass: one 16-byte gather
            // read-only for the whole launch: ld.global.nc lets L1 cache lines that live on a peer GPU
            // (the ghost slots of every rank all look at the same few cells near the corners)
            const double2 g2 = SLAB ? slab_load_grad(a.st, a.sg, sc) : grad[sc];
            gx = g2.x;
            gy = g2.y;
        } else {
            const int xm = (sx > 0) ? -W : 0, xp = (sx < H - 1) ? W : 0;
            const int ym = (sy > 0) ? -1 : 0, yp = (sy < W - 1) ? 1 : 0;
            if (SLAB) {
                gx = __ldg(slab_chan(a.st.medium_in, a.sg, 2, sc + xp)) - __ldg(slab_chan(a.st.medium_in, a.sg, 2, sc + xm));
                gy = __ldg(slab_chan(a.st.medium_in, a.sg, 2, sc + yp)) - __ldg(slab_chan(a.st.medium_in, a.sg, 2, sc + ym));
            } else {
                gx = (double)chem[sc + xp] - (double)chem[sc + xm];
                gy = (double)chem[sc + yp] - (double)chem[sc + ym];
            }
            if (xp - xm == 2 * W) gx *= 0.5;      // central difference: / 2.0 exactly (as sample_gradient)
            if (yp - ym == 2) gy *= 0.5;
        }

        bool deposit_mask = true;
        double heading = 0.0;
        if (DISCRETE_TURN) {
            // PhysarumAgent._discrete_turn / _choose_turn (:168-208).  die_turn_quick settles the turn from
            // the raw gradient and (sin, cos) of the heading whenever no threshold is within its guard band;
            // the rest (about one warp in 300) runs the reference's own arithmetic, die_turn_exact.
            die_turn_t tr;
            double dr = 1.0;
            // (G32: the float32 pair goes to the decision as it is; die_turn_quick_f would round-trip it through double)
            if (!((LEAN || a.plan.enabled) &&
                  (G32 ? die_turn_quick_ff(&a.plan, gxf, gyf, sf, cf, th, atol, p.sense_radians, &tr)
                       : die_turn_quick_f(&a.plan, gx, gy, sf, cf, th, atol, p.sense_radians, &tr)))) {
                if (from32) sample_gradient(chem, sx, sy, H, W, gx, gy);     // the exact path wants all 53 bits
                if (LEAN) {           // (normalised gradient: dr = 1)
                    tr = turn_exact_call(gx, gy, th, atol, p.sense_radians, 1, p.use_grad_clip, p.grad_clip);
                } else {
                    // _get_gradient (:59-65): scipy.linalg.norm, nan_to_num(grad / norm), grad *= (norm >= clip)
                    die_normalize_gradient(&gx, &gy, p.normalized_grad, p.use_grad_clip, p.grad_clip);
                    if (!p.normalized_grad) dr = hypot(gx, gy);
                    tr = die_turn_exact(gx, gy, th, atol, p.sense_radians);
                }
            }
            const int c = (!LEAN && coin_p != nullptr) ? (coin_p[i] ? 1 : 0) : (int)((coin_bits >> k) & 1u);
            // turn = (tr.turn != 0 ? tr.turn : ((double)c - 0.5) * 2.0) * turn_radians: the factor is exactly +-1
            const int tsign = (tr.turn != 0) ? tr.turn : 2 * c - 1;
            const double turn = (tsign < 0) ? -p.turn_radians : p.turn_radians;
            deposit_mask = tr.deposit_mask != 0;
            const double dirn = renormalize_radians(th + turn);
            double s2, c2;
            if (fused_heading) die_sincos_angle(dirn, &s2, &c2, &heading);
            else die_sincos(dirn, &s2, &c2);
            if (LEAN || p.normalized_grad) {      // dr = 1: 1 * c2 - 0 * s2 = c2 exactly (neither is ever zero)
                gx = c2;
                gy = s2;
            } else {
                // polar2xy(dr, d) = dr * exp(1j d) with a REAL array dr (core/utils.py:154-164): numpy promotes it to
                // dr + 0j and multiplies complex numbers, (dr c - 0 s) + 1j (dr s + 0 c).  The same numbers as
                // (dr c, dr s) unless dr == 0 (a zero gradient), where only the zeros' signs differ -- and those decide
                // whether the next heading angle(gx + 1j gy) is 0 or pi.
                gx = __dsub_rn(__dmul_rn(dr, c2), __dmul_rn(0.0, s2));
                gy = __dadd_rn(__dmul_rn(dr, s2), __dmul_rn(0.0, c2));
            }
        } else {
            die_normalize_gradient(&gx, &gy, p.normalized_grad, p.use_grad_clip, p.grad_clip);
        }

        // _process_momentum (:82-91)
        if (!LEAN && pg != nullptr) {
            gx = (1.0 - p.inertia) * gx + p.inertia * pg[i];
            gy = (1.0 - p.inertia) * gy + p.inertia * pg[M + i];
            if (nz != nullptr) {
                gx += p.noise_scale * nz[i];
                gy += p.noise_scale * nz[M + i];
            } else if (p.noise_scale != 0.0) {
                const uint4 r = philox_draw(a.seed, step, (uint64_t)((ch.b + a.b0) * M + first + i), 3u);
                const double u1 = 1.0 - u53(r.x, r.y), u2 = u53(r.z, r.w);
                const double rad = 0.4 * sqrt(-2.0 * log(u1));
                double sn3, cs3;
                die_sincos(kTwoPi * u2, &sn3, &cs3);
                gx += p.noise_scale * (rad * cs3);
                gy += p.noise_scale * (rad * sn3);
            }
            pg[i] = gx;
            pg[M + i] = gy;
        }
        th_p[i] = fused_heading ? heading : angle_xy<false>(gx, gy);   // :110

        // deposit relative to the food under the agent (:113-117, :210-214)
        double dep = p.deposit * food_here;
        if (DISCRETE_TURN) dep = dep * (deposit_mask ? 1.0 : 0.1);

        const double adx = gx * p.scale, ady = gy * p.scale;  // unmasked (Q8)
        ab[i] = adx;
        ab_y[i] = ady;
        ab_dep[i] = dep;
        // cost hint: linear_action_cost (core/env.py:29-35) of this very action, the expression of agent_feed_kernel; the
        // step that receives the action unmodified reads these 8 bytes per slot instead of dx, dy, deposit again (24)
        // (die_sqrt_near: the correctly rounded root from one Newton step where that is provable, sqrt() elsewhere)
        if (bo_p != nullptr)
            bo_p[i] = a.cost_w_dep * fabs(dep) + a.cost_w_dist * die_sqrt_near(&a.cost_sqrt, adx * adx + ady * ady);

        if (MOVE) {
            // Env._agent_move (core/env.py:152-172) of THIS action + cell resolution + claim: what
            // move_claim_kernel would compute from (agents, action); the positions themselves are
            // committed by the feed kernel once Env.step adopts the action (die_env_step_flags, DIE_STEP_ADOPT_MOVE)
            const double mx = apply_boundary(x + adx, a.boundary);
            const double my = apply_boundary(y + ady, a.boundary);
            const int cell = nearest_cell(mx, ax) * W + nearest_cell(my, ay);
            co_p[i] = cell;
            if (cx_p != nullptr) {                 // committed move: this thread alone reads and writes its slots' x, y
                cx_p[i] = mx;
                cx_p[M + i] = my;
            }
            if ((alive_word >> (threadIdx.x & 31)) & 1u) atomicMax(win + cell, (int32_t)(first + i));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Env._agent_move (core/env.py:152-172) + cell resolution (core/utils.py:39-54) + the claim
// that resolves "last writer wins" (core/env.py:211): the winner of a cell is the
// highest slot index among the alive agents on it -- atomicMax is order-independent, so the
// result is deterministic and bit-exact.
// ---------------------------------------------------------------------------------------------
constexpr int kMoveItems = 4;

template <bool SLAB, bool BITS>
__global__ void __launch_bounds__(kAgentThreads)
move_claim_kernel(double* __restrict__ agents, const double* __restrict__ action,
                  int32_t* __restrict__ winner, int32_t* __restrict__ cells,
                  const Axis ax, const Axis ay, int64_t M, int nchunk, int boundary,
                  const uint32_t* __restrict__ alive_bits, int64_t Mw,
                  const SlabGeom sg, const SlabTables st) {
    const int W = ay.n;
    const int64_t C = (int64_t)ax.n * ay.n;
    const SlotChunk ch = slot_chunk<kMoveItems>(nchunk);
    double* ag = agents + ch.b * 4 * M;
    const double* ac = action + ch.b * 3 * M;
    int32_t* win = winner + ch.b * C;
    int32_t* cl = cells + ch.b * M;
    // BITS: alive-ness from the env's bitmask (1 bit instead of 8 bytes per slot)
    const uint32_t* bits_p = BITS ? alive_bits + ch.b * Mw : nullptr;
#pragma unroll
    for (int k = 0; k < kMoveItems; ++k) {
        const int64_t i = ch.base + k * kAgentThreads + threadIdx.x;
        if (i >= M) break;
        const double x = apply_boundary(ag[i] + ac[i], boundary);
        const double y = apply_boundary(ag[M + i] + ac[M + i], boundary);
        const bool alive = BITS ? ((bits_p[i >> 5] >> (threadIdx.x & 31)) & 1u) != 0 : ag[2 * M + i] > 0.0;
        ag[i] = x;
        ag[M + i] = y;
        const int cell = nearest_cell(x, ax) * W + nearest_cell(y, ay);
        cl[i] = cell;
        if (alive) {
            if (SLAB) atomicMax(slab_cell(st.claim, sg, cell), (int32_t)slab_slot_global(sg, sg.rank, i));
            else atomicMax(win + cell, (int32_t)i);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Env._agent_feed per-slot part (core/env.py:220-234, cost :29-35) + the reward / num_agents
// reduction (:118-121).  consumed[M] = consumed_field[cell(slot)] for ALL M slots (Q1, Q7), where
// consumed_field = rate_feed * food * occ was written per cell by the field pass.  Alive slots also
// clear their cell's claim for the next step.  Block partials are written in a fixed layout and
// summed in a fixed order by finalize_stats_kernel, so the reward is run-to-run deterministic.
// ---------------------------------------------------------------------------------------------
constexpr int kFeedItems = 4;      // slots per thread

// MOVE: the step adopted a move that the forward kernel evaluated speculatively (cells + claims are
// in place); this kernel then also commits the positions, pos = boundary(pos + action[dx, dy])
// (core/env.py:152-172, the same two operations as move_claim_kernel).
// BITS: alive-ness comes from the env's bitmask instead of the float64 channel (MOVE implies BITS).
// (register caps were tried in round 2 on a B200: 64 / 48 registers = 4 / 5 CTAs per SM run 10 % / 20 % SLOWER than
//  the compiler's own 74 registers, profiles/r02a_staged_ab_summary.txt -- removed)
struct FeedArgs {
    double* agents;
    const double* action;
    const double* consumed_field;    // [B][C] rate_feed * food * occ of this step (the field pass wrote it); element type FT
    const double2* cell_pairs;       // PAIR: [B][C] {consumed_field, new env_food} per cell, written by the field pass instead
    double* food_here;               // PAIR: [B][M] out: the new env_food under every slot (the next forward's food_here)
    int32_t* winner;
    int32_t* cells;                  // read; DIE also resets the cached cell of a slot it puts back at (0, 0)
    double* part_gain;
    int32_t* part_alive;
    int64_t C, M;
    int nblk;
    double w_dep, w_dist;
    const uint32_t* alive_bits;
    int64_t Mw;
    int boundary;
    const double* burned;            // COST: [B][M] linear_action_cost per slot, written by the forward kernel for this action
};

// DIE: Dynamics.agents_die -- Env._agent_lifecycle (core/env.py:245-250) folded in: a slot whose stock after feeding is
// not above 1e-4 has ALL its channels zeroed (agents.where(agent_food > 1e-4, 0): dead, back at (0, 0), no stock), which
// holds for every ghost slot every step; num_agents counts the survivors (core/env.py:118, after the lifecycle).
// PAIR: the per-cell scratch is {consumed_field, new env_food} (16 bytes, one sector): the slot's gather brings both, and
// the food is stored per slot for the next forward pass (see FH there).
// COST: the action cost of every slot comes from the array the forward kernel filled for exactly this action (the caller
// proved the identity, as for the speculative move): one 8-byte load per slot instead of dx, dy, deposit (24 bytes).
// MINB: minimum resident CTAs per SM (a register cap: 6 -> 42, 8 -> 32 registers), tried on the COST instantiation only.
template <bool SLAB, bool MOVE, bool BITS, bool DIE = false, typename FT = double, bool PAIR = false, bool COST = false,
          int MINB = 1>
__global__ void __launch_bounds__(kAgentThreads, MINB)
agent_feed_kernel(const FeedArgs a, const SlabGeom sg, const SlabTables st) {
    static_assert(!PAIR || (!SLAB && !DIE && sizeof(FT) == 8), "the pair table serves the plain float64 step");
    static_assert(!COST || (!SLAB && !MOVE && !DIE), "the cost hint serves the plain step (a speculative move needs dx, dy)");
    const int64_t M = a.M;
    const int nblk = a.nblk;
    const int64_t b = blockIdx.x / (unsigned)nblk;
    const int blk = blockIdx.x - (int)b * nblk;
    const int64_t first = (int64_t)blk * (kAgentThreads * kFeedItems) + threadIdx.x;
    double* __restrict__ ag_x = a.agents + b * 4 * M + first;  // x; y, alive, agent_food are + M, 2M, 3M
    const double* __restrict__ ac = a.action + b * 3 * M + first;
    const FT* __restrict__ cf = (SLAB || PAIR) ? nullptr : (const FT*)a.consumed_field + b * a.C;
    const double2* __restrict__ cp = PAIR ? a.cell_pairs + b * a.C : nullptr;
    double* __restrict__ fh = PAIR ? a.food_here + b * M + first : nullptr;
    int32_t* __restrict__ win = a.winner + b * a.C;
    int32_t* __restrict__ cl = a.cells + b * M + first;
    const uint32_t* __restrict__ bits_p = BITS ? a.alive_bits + b * a.Mw + (first >> 5) : nullptr;
    const double* __restrict__ bu_p = COST ? a.burned + b * M + first : nullptr;
    const double w_dep = a.w_dep, w_dist = a.w_dist;
    const int boundary = a.boundary;
    double* const part_gain = a.part_gain;
    int32_t* const part_alive = a.part_alive;

    int cell[kFeedItems];
    double dx[kFeedItems], dy[kFeedItems], dep[kFeedItems], stock[kFeedItems], eaten[kFeedItems];
    double px[kFeedItems], py[kFeedItems];
    double under[PAIR ? kFeedItems : 1];
    bool alive[kFeedItems], valid[kFeedItems];
#pragma unroll
    for (int k = 0; k < kFeedItems; ++k) {
        valid[k] = first + k * kAgentThreads < M;
        cell[k] = valid[k] ? cl[k * kAgentThreads] : 0;
    }
#pragma unroll
    for (int k = 0; k < kFeedItems; ++k) {
        const int i = k * kAgentThreads;
        if (PAIR) {
            const double2 pr = valid[k] ? cp[cell[k]] : make_double2(0.0, 0.0);
            eaten[k] = pr.x;
            under[k] = pr.y;                       // (stored after every gather of this thread is in flight)
        } else {
            eaten[k] = valid[k] ? (SLAB ? slab_load_consumed(st, sg, cell[k]) : (double)cf[cell[k]]) : 0.0;
        }
        if (BITS) alive[k] = valid[k] && ((bits_p[i >> 5] >> (threadIdx.x & 31)) & 1u);
        else alive[k] = valid[k] && ag_x[2 * M + i] > 0.0;
        if (MOVE) {
            px[k] = valid[k] ? ag_x[i] : 0.0;
            py[k] = valid[k] ? ag_x[M + i] : 0.0;
        }
        stock[k] = valid[k] ? ag_x[3 * M + i] : 0.0;
        if (COST) {
            dx[k] = valid[k] ? bu_p[i] : 0.0;              // (the cost itself)
            dy[k] = dep[k] = 0.0;
        } else {
            dx[k] = valid[k] ? ac[i] : 0.0;
            dy[k] = valid[k] ? ac[M + i] : 0.0;
            dep[k] = valid[k] ? ac[2 * M + i] : 0.0;
        }
    }
    double gain_sum = 0.0;
    int alive_cnt = 0;
#pragma unroll
    for (int k = 0; k < kFeedItems; ++k) {
        if (valid[k]) {
            const int i = k * kAgentThreads;
            if (PAIR) fh[i] = under[k];
            if (MOVE) {
                ag_x[i] = apply_boundary(px[k] + dx[k], boundary);
                ag_x[M + i] = apply_boundary(py[k] + dy[k], boundary);
            }
            const double burned = COST ? dx[k] : w_dep * fabs(dep[k]) + w_dist * sqrt(dx[k] * dx[k] + dy[k] * dy[k]);
            const double gained = eaten[k] - burned;
            const double stock_new = stock[k] + gained;
            gain_sum += gained;
            bool survives = true;
            if (DIE) {
                survives = stock_new > 1e-4;
                if (!survives) {                   // every channel of the slot <- 0 (core/env.py:249-250)
                    ag_x[i] = 0.0;
                    ag_x[M + i] = 0.0;
                    ag_x[2 * M + i] = 0.0;
                    cl[i] = 0;                     // the cell of (0, 0): the next forward's cached "cell under the agent"
                }
            }
            ag_x[3 * M + i] = survives ? stock_new : 0.0;
            if (alive[k]) {                        // claim table back to empty for the next step
                if (SLAB) *slab_cell(st.claim, sg, cell[k]) = -1;
                else win[cell[k]] = -1;
                if (survives) ++alive_cnt;
            }
        }
    }
    __shared__ double s_gain[kAgentThreads / 32];
    __shared__ int s_alive[kAgentThreads / 32];
    gain_sum = warp_sum(gain_sum);
    alive_cnt = warp_sum(alive_cnt);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        s_gain[wid] = gain_sum;
        s_alive[wid] = alive_cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double g = 0.0;
        int n = 0;
#pragma unroll
        for (int k = 0; k < kAgentThreads / 32; ++k) {
            g += s_gain[k];
            n += s_alive[k];
        }
        part_gain[blockIdx.x] = g;
        part_alive[blockIdx.x] = n;
    }
}

// alive bitmask of the env: bit (i & 31) of word [b][i >> 5] = (agents[b][2][i] > 0)
__global__ void __launch_bounds__(256)
alive_bits_kernel(const double* __restrict__ agents, uint32_t* __restrict__ bits, int64_t M, int64_t Mw, int B) {
    const int64_t total = (int64_t)B * Mw * 32;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const int64_t b = t / (Mw * 32), i = t - b * (Mw * 32);
        const bool alive = i < M && agents[(b * 4 + 2) * M + i] > 0.0;
        const uint32_t word = __ballot_sync(0xffffffffu, alive);
        if ((threadIdx.x & 31) == 0) bits[b * Mw + (i >> 5)] = word;
    }
}

// Refresh of the corner mirror (die_slab.cuh): the four corner_r x corner_r corner patches of the published
// gradient, the NEW medium's env_food and consumed_field, pulled from their owners (peer loads over NVLink)
// into local memory.
__global__ void __launch_bounds__(256)
slab_corner_copy_kernel(const SlabGeom sg, const SlabTables st, double2* __restrict__ corner_grad,
                        double* __restrict__ corner_food, double* __restrict__ corner_cons, int with_grad) {
    const int R = st.corner_r, P = 2 * R;
    const int total = P * P;
    for (int b = blockIdx.x * 256 + threadIdx.x; b < total; b += gridDim.x * 256) {
        const int pr = b / P, pc = b - pr * P;
        const int row = (pr < R) ? pr : pr + (sg.H - P);
        const int col = (pc < R) ? pc : pc + (sg.W - P);
        const int cell = row * sg.W + col;
        if (with_grad) corner_grad[b] = __ldg(slab_cell(st.grad, sg, cell));
        corner_food[b] = __ldg(slab_chan(st.medium_out, sg, 1, cell));
        corner_cons[b] = __ldg(slab_cell(st.consumed, sg, cell));
    }
}

constexpr int kFinalThreads = 1024;

// Sum of an environment's block partials in a fixed order (thread t takes partials t, t + 1024, ... in four
// interleaved accumulators, then warp shuffles, then the 32 warp sums in index order): run-to-run deterministic.
// NT physical threads execute the 1024 logical ones in passes, so the cluster-fused step (die_env_fused.cuh), whose
// CTAs are smaller, produces the same bits.  CG: read the partials with ld.global.cg (they were written by other CTAs
// of the running kernel).  s_g / s_n: 32 entries of shared memory each.
template <int NT, bool CG>
__device__ __forceinline__ void finalize_sum(const double* __restrict__ pgn, const int32_t* __restrict__ pal, int nblk,
                                             double* s_g, long long* s_n, double* reward, int64_t* alive) {
    static_assert(kFinalThreads % NT == 0 && NT % 32 == 0, "logical threads are emulated in whole passes");
    for (int pass = 0; pass < kFinalThreads / NT; ++pass) {
        const int t = pass * NT + threadIdx.x;                 // the logical thread
        double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
        long long n = 0;
        int k = t;
        for (; k + 3 * kFinalThreads < nblk; k += 4 * kFinalThreads) {
            const double a0 = CG ? __ldcg(pgn + k) : pgn[k];
            const double a1 = CG ? __ldcg(pgn + k + kFinalThreads) : pgn[k + kFinalThreads];
            const double a2 = CG ? __ldcg(pgn + k + 2 * kFinalThreads) : pgn[k + 2 * kFinalThreads];
            const double a3 = CG ? __ldcg(pgn + k + 3 * kFinalThreads) : pgn[k + 3 * kFinalThreads];
            if (CG) n += (long long)__ldcg(pal + k) + __ldcg(pal + k + kFinalThreads) + __ldcg(pal + k + 2 * kFinalThreads) +
                         __ldcg(pal + k + 3 * kFinalThreads);
            else n += (long long)pal[k] + pal[k + kFinalThreads] + pal[k + 2 * kFinalThreads] + pal[k + 3 * kFinalThreads];
            g0 += a0;
            g1 += a1;
            g2 += a2;
            g3 += a3;
        }
        for (; k < nblk; k += kFinalThreads) {
            g0 += CG ? __ldcg(pgn + k) : pgn[k];
            n += CG ? __ldcg(pal + k) : pal[k];
        }
        double g = (g0 + g1) + (g2 + g3);
        g = warp_sum(g);
        n = warp_sum(n);
        if ((threadIdx.x & 31) == 0) {
            s_g[t >> 5] = g;
            s_n[t >> 5] = n;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double gg = 0.0;
        long long nn = 0;
        for (int w = 0; w < kFinalThreads / 32; ++w) {
            gg += s_g[w];
            nn += s_n[w];
        }
        *reward = gg;
        *alive = nn;
    }
}

__global__ void __launch_bounds__(kFinalThreads)
finalize_stats_kernel(const double* __restrict__ part_gain, const int32_t* __restrict__ part_alive,
                      int nblk, double* __restrict__ reward, int64_t* __restrict__ alive) {
    __shared__ double s_g[kFinalThreads / 32];
    __shared__ long long s_n[kFinalThreads / 32];
    const int b = blockIdx.x;
    finalize_sum<kFinalThreads, false>(part_gain + (int64_t)b * nblk, part_alive + (int64_t)b * nblk, nblk, s_g, s_n,
                                       reward + b, alive + b);
}

}  // namespace die
