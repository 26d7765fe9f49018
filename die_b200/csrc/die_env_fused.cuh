// die_env_fused.cuh -- Env.step (core/env.py:101-131) of a SMALL periodic environment as ONE kernel: a thread-block
// cluster of S CTAs owns one environment of the batch for the whole step, so everything that the three-kernel path
// (move_claim -> field_step -> agent_feed -> finalize_stats) hands from kernel to kernel through HBM stays on chip:
//
//   claim table      int32 per cell ("last writer wins"): CTA q holds rows [q rows_per, (q+1) rows_per) in its
//                    shared memory; agents on any CTA claim with atomicMax through DISTRIBUTED SHARED MEMORY, the field
//                    phase reads its own rows + the halo rows of its neighbours, the feed phase reads the cell under
//                    every slot.  No global claim table, no occupancy bitmap, no clearing pass.
//   chem1 rows       the CTA's rows + halo arrive by TMA bulk copies (cp.async.bulk global -> shared on an mbarrier,
//                    issued before the move phase so they land while it runs).  An environment's rows are contiguous in
//                    memory, so a CTA's whole block is ONE copy (two where it wraps around the field).
//   action           dx, dy are read by the move phase and again (from L2) by the feed phase of the same CTA.
//   reward           block partials in the layout and order of agent_feed_kernel, summed by CTA 0 in the order of
//                    finalize_stats_kernel -- the reward is bit-identical to the three-kernel path.
//
// HBM traffic per cell-update (M = C, float64): R{x,y,dx,dy,dep,agent_food,chem,food} W{x,y,agent_food,cell,occ,food,
// chem,grad32} = 64 + 60 = 124 B against 56.9 + 61.8 + 54.2 = 172.9 B measured for the three kernels (ncu, r02a).
//
// Phases (cluster barriers between them; a CTA never leaves before the last one, its shared memory is read remotely):
//   0  claims <- -1; thread 0 issues the bulk copies                                        | cluster_sync
//   1  Env._agent_move + cell resolution + claim   (core/env.py:152-172, core/utils.py:39-54) | cluster_sync
//   2  deposit + occupancy + food + diffusion * decay (+ np.gradient of the new chem1)      (core/env.py:204-228,136-150)
//   3  Env._agent_feed per slot + block partials   (core/env.py:220-234, 29-35)            | cluster_sync
//   4  CTA 0: reward, num_agents                   (core/env.py:118-121)
// Arithmetic is operation for operation that of move_claim_kernel / field_step_kernel / agent_feed_kernel.
// Every phase issues all the loads of a round before its first dependent instruction (a CTA has only its own warps
// to hide latency with: the first version, one load -> use -> store chain per slot, ran at 22 % of the DRAM bandwidth).
// Applies when: periodic diffusion, blur radius 1..4, an even row length (16-byte rows for the bulk copies) of at least
// R cells, the alive bitmask, H divisible by a cluster size whose shared-memory footprint fits and whose CTAs get at
// most kFusedMaxRounds rounds of feed blocks; everything else takes the three kernels.
#pragma once
#include "die_async.cuh"
#include "die_cluster.cuh"
#include "die_agent_kernels.cuh"
#include "die_field_kernels.cuh"

namespace die {

struct FusedArgs {
    // ---- agent side
    double* agents;              // [B][4][M]
    const double* action;        // [B][3][M]
    int32_t* cells;              // [B][M] linear cell of every slot after the move (the next forward's hint)
    double* part_gain;           // [B][nblk] the feed blocks' partial sums (same layout as agent_feed_kernel)
    int32_t* part_alive;
    double* reward;              // [B]
    int64_t* alive_out;          // [B]
    const uint32_t* alive_bits;  // [B][Mw] (BITS)
    int64_t Mw;
    Axis ax, ay;
    int boundary;
    double w_dep, w_dist;
    int64_t M;
    int nblk;                    // ceil(M / 1024): agent_feed_kernel's blocks per environment
    int bpc;                     // blocks per CTA = ceil(nblk / S)
    // ---- field side
    const double* medium_in;     // [B][3][H][W]
    double* medium_out;
    double2* grad;               // [B][H*W] or null
    float2* grad32;              // [B][H*W] or null
    int H, W;
    int S;                       // CTAs per environment (the cluster size)
    int rows_per;                // H / S
    int slab_shift;              // log2(rows_per * W) when that is a power of two, else -1
    double rate_feed, keep;
    int food_infinite;
    const double* flow_rwave;    // op_food_flow, as FieldArgs
    const double* flow_col;
    const double* flow_row;
    double flow_t, flow_scale, flow_keep;
    const double* flow_frame;
    BlurWeights bw;
};

// shared memory of one CTA (bytes), the same carving as in the kernel
static inline size_t fused_smem_bytes(int rows_per, int W, int R, bool grad, int NT) {
    const int G = grad ? 1 : 0, HALO = R + G;
    const size_t LH = rows_per + 2 * HALO, OH = rows_per + 2 * G, LWP = W + 2 * R;
    size_t doubles = LH * W + OH * LWP + 4 * (NT / 32) + 32 + 32 + 1;    // s_chem, s_v, s_gain, s_fg, s_fn, mbarrier
    size_t ints = (size_t)rows_per * W + 4 * (NT / 32);                   // s_claim, s_alive
    return 128 + doubles * 8 + ints * 4;
}

constexpr int kFusedMaxRounds = 4;      // feed blocks a thread group works through (cells stay in registers between phases)

template <int R, int NT, bool GRAD, bool PLAIN>
__global__ void __launch_bounds__(NT, (NT <= 512) ? 2 : 1)
env_step_fused_kernel(const FusedArgs a) {
    static_assert(NT % kAgentThreads == 0, "the feed phase works in agent_feed_kernel's blocks of 256 threads");
    constexpr int G = GRAD ? 1 : 0, HALO = R + G;
    constexpr int NW = NT / 32;
    constexpr int VB = NT / kAgentThreads;                 // feed blocks in flight per CTA
    constexpr int kBlockSlots = kAgentThreads * kFeedItems;
    constexpr int MAXT = 128 / NW;                         // output tasks per warp whose food is prefetched
    static_assert(kFeedItems == kMoveItems, "one slot mapping for the move and the feed phase");
    const int H = a.H, W = a.W, S = a.S, rows_per = a.rows_per;
    const int LH = rows_per + 2 * HALO, OH = rows_per + 2 * G, LWP = W + 2 * R;
    const int64_t C = (int64_t)H * W, M = a.M;
    const int slab_cells = rows_per * W;
    const int rank = (int)cluster_rank();
    const int64_t b = blockIdx.x / (unsigned)S;
    const int r0 = rank * rows_per;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int segs = (W + 31) >> 5;                        // 32-column pieces of a row: one warp task each

    extern __shared__ double smem[];
    char* base = (char*)(((uintptr_t)smem + 127) & ~(uintptr_t)127);
    double* s_chem = (double*)base;                        // [LH][W] staged chem1 (+ deposits); then s_out [OH][W]
    double* s_v = s_chem + (size_t)LH * W;                 // [OH][LWP] after the axis-0 pass, periodic halo columns
    double* s_gain = s_v + (size_t)OH * LWP;               // [kFusedMaxRounds][NW]
    double* s_fg = s_gain + kFusedMaxRounds * NW;          // [32]
    long long* s_fn = (long long*)(s_fg + 32);             // [32]
    mbar_t* bar = (mbar_t*)(s_fn + 32);
    int32_t* s_claim = (int32_t*)(bar + 1);                // [rows_per][W] this CTA's rows of the claim table
    int* s_alive = (int*)(s_claim + slab_cells);           // [kFusedMaxRounds][NW]
    double* s_out = s_chem;

    const double* __restrict__ food_in = a.medium_in + (b * 3 + 1) * C;
    const double* chem_in = a.medium_in + (b * 3 + 2) * C;
    double* __restrict__ occ_out = a.medium_out + b * 3 * C;
    double* __restrict__ food_out = occ_out + C;
    double* __restrict__ chem_out = occ_out + 2 * C;
    double* ag = a.agents + b * 4 * M;
    const double* __restrict__ ac = a.action + b * 3 * M;
    int32_t* __restrict__ cl = a.cells + b * M;
    const uint32_t* __restrict__ abits = a.alive_bits + b * a.Mw;

    // ---- phase 0 ---------------------------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < slab_cells; i += NT) s_claim[i] = -1;
    __syncthreads();
    if (tid == 0) {
        // staged row lr stands for field row (r0 - HALO + lr) mod H: maximal runs of consecutive rows, one copy each
        mbar_arrive_expect_tx(bar, (uint32_t)((size_t)LH * W * sizeof(double)));
        int lr = 0;
        while (lr < LH) {
            const int gr = wrap_index(r0 - HALO + lr, H);
            const int n = min(LH - lr, H - gr);
            bulk_g2s(s_chem + (size_t)lr * W, chem_in + (int64_t)gr * W, (uint32_t)((size_t)n * W * sizeof(double)), bar);
            lr += n;
        }
    }
    cluster_sync();                                        // every slice of the claim table is empty before anybody claims

    // ---- phase 1: move + claim ---------------------------------------------------------------------------------
    // A thread handles the same slots here and in the feed phase (agent_feed_kernel's map: block blk, item k, thread
    // t256), so their cells stay in registers.  Every round issues all its loads before the first store.
    const int blk_lo = rank * a.bpc, blk_hi = min(blk_lo + a.bpc, a.nblk);
    const int sub = tid / kAgentThreads, t256 = tid - sub * kAgentThreads;
    const int rounds = (a.bpc + VB - 1) / VB;              // <= kFusedMaxRounds (checked by the launcher)
    int mycell[kFusedMaxRounds][kFeedItems];
    uint32_t myalive = 0;                                  // bit j * kFeedItems + k
#pragma unroll
    for (int j = 0; j < kFusedMaxRounds; ++j) {
        const int blk = blk_lo + j * VB + sub;
        const bool active = j < rounds && blk < blk_hi;
        const int64_t first = (int64_t)blk * kBlockSlots + t256;
        double x[kMoveItems], y[kMoveItems], dx[kMoveItems], dy[kMoveItems];
        bool valid[kMoveItems];
#pragma unroll
        for (int k = 0; k < kMoveItems; ++k) {
            const int64_t i = first + k * kAgentThreads;
            valid[k] = active && i < M;
            x[k] = valid[k] ? ag[i] : 0.0;
            y[k] = valid[k] ? ag[M + i] : 0.0;
            dx[k] = valid[k] ? ac[i] : 0.0;
            dy[k] = valid[k] ? ac[M + i] : 0.0;
            if (valid[k] && ((abits[i >> 5] >> (i & 31)) & 1u)) myalive |= 1u << (j * kFeedItems + k);
        }
#pragma unroll
        for (int k = 0; k < kMoveItems; ++k) {
            mycell[j][k] = 0;
            if (valid[k]) {
                const int64_t i = first + k * kAgentThreads;
                const double nx = apply_boundary(x[k] + dx[k], a.boundary);
                const double ny = apply_boundary(y[k] + dy[k], a.boundary);
                ag[i] = nx;
                ag[M + i] = ny;
                const int cx = nearest_cell(nx, a.ax), cy = nearest_cell(ny, a.ay);
                const int cell = cx * W + cy;
                mycell[j][k] = cell;
                cl[i] = cell;
                if ((myalive >> (j * kFeedItems + k)) & 1u) {
                    const int q = cx / rows_per;
                    atomicMax(cluster_map(s_claim, (unsigned)q) + (cx - q * rows_per) * W + cy, (int32_t)i);
                }
            }
        }
    }
    // the food under this CTA's output cells is needed at the very end of the field phase: request it now
    const int out_tasks = rows_per * segs;
    const bool food_pre = out_tasks <= MAXT * NW;
    double fpre[MAXT];
    if (food_pre) {
        int r = 0, seg = warp;
#pragma unroll
        for (int u = 0; u < MAXT; ++u) {
            while (seg >= segs) { seg -= segs; ++r; }
            const int c = seg * 32 + lane;
            fpre[u] = (r < rows_per && c < W) ? food_in[(int64_t)(r0 + r) * W + c] : 0.0;
            seg += NW;
        }
    }
    cluster_sync();                                        // all claims are in

    // ---- phase 2: the field rows of this CTA ---------------------------------------------------------------------
    mbar_wait(bar, 0);
    {   // deposit: chem[cell] = chem[cell] + deposit1[winner]   (core/env.py:211)
        // Warp tasks (staged row lr, 32-column piece seg) are walked incrementally: no integer division in any loop of
        // this phase (the first version spent a third of its instructions on them).  The claims of staged row lr live
        // in the CTA found by stepping rows_per rows at a time around the ring of CTAs (H = S rows_per).
        const double* __restrict__ dep = ac + 2 * M;
        constexpr int U = 6;
        int lr = 0, seg = warp;
        while (seg >= segs) { seg -= segs; ++lr; }
        while (lr < LH) {
            int w6[U], at[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                w6[u] = -1;
                at[u] = 0;
                const int c = seg * 32 + lane;
                if (lr < LH && c < W) {
                    int q = rank, lrow = lr - HALO;
                    while (lrow < 0) { lrow += rows_per; q = (q == 0) ? S - 1 : q - 1; }
                    while (lrow >= rows_per) { lrow -= rows_per; q = (q + 1 == S) ? 0 : q + 1; }
                    const int32_t* slice = (q == rank) ? s_claim : cluster_map(s_claim, (unsigned)q);
                    w6[u] = slice[lrow * W + c];
                    at[u] = lr * W + c;
                }
                seg += NW;
                while (seg >= segs) { seg -= segs; ++lr; }
            }
            double d6[U];
#pragma unroll
            for (int u = 0; u < U; ++u) d6[u] = (w6[u] >= 0) ? dep[w6[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (w6[u] >= 0) s_chem[at[u]] = s_chem[at[u]] + d6[u];
        }
    }
    __syncthreads();
    // axis-0 pass (scipy's operation order, as field_step_kernel); the periodic halo columns of s_v are filled on the
    // way (W >= R, checked by the launcher: every halo column is the copy of exactly one real column)
    for (int r = 0, seg = warp;; seg += NW) {
        while (seg >= segs) { seg -= segs; ++r; }
        if (r >= OH) break;
        const int c = seg * 32 + lane;
        if (c < W) {
            const double* p = s_chem + (size_t)(r + R) * W + c;
            double acc = p[0] * a.bw.w[R];
#pragma unroll
            for (int k = R; k >= 1; --k) acc += (p[-k * W] + p[k * W]) * a.bw.w[R - k];
            double* row = s_v + r * LWP;
            row[R + c] = acc;
            if (c < R) row[R + W + c] = acc;
            if (c >= W - R) row[c - (W - R)] = acc;
        }
    }
    __syncthreads();
    // axis-1 pass into s_out (aliases the staged chem, which nobody reads any more)
    for (int r = 0, seg = warp;; seg += NW) {
        while (seg >= segs) { seg -= segs; ++r; }
        if (r >= OH) break;
        const int c = seg * 32 + lane;
        if (c < W) {
            const double* p = s_v + r * LWP + R + c;
            double acc = p[0] * a.bw.w[R];
#pragma unroll
            for (int k = R; k >= 1; --k) acc += (p[-k] + p[k]) * a.bw.w[R - k];
            s_out[r * W + c] = acc * a.keep;
        }
    }
    __syncthreads();
    {   // outputs: new chem1 (+ its np.gradient), occupancy, food
        double2* __restrict__ grad = (GRAD && a.grad != nullptr) ? a.grad + b * C : nullptr;
        float2* __restrict__ grad32 = (GRAD && a.grad32 != nullptr) ? a.grad32 + b * C : nullptr;
        auto emit = [&](int r, int c, double f) {
            const int gi = r0 + r;
            const int g = gi * W + c;
            const double* q = s_out + (r + G) * W + c;
            chem_out[g] = q[0];
            if (GRAD) {       // np.gradient: (f[i+1] - f[i-1]) / 2 inside, one-sided at the edges (non-periodic, Q5)
                const int um = (gi > 0) ? -W : 0, up = (gi < H - 1) ? W : 0;
                const int lm = (c > 0) ? -1 : 0, lp = (c < W - 1) ? 1 : 0;
                double gx = q[up] - q[um];
                double gy = q[lp] - q[lm];
                if (up - um == 2 * W) gx *= 0.5;
                if (lp - lm == 2) gy *= 0.5;
                if (grad32 != nullptr) grad32[g] = make_float2((float)gx, (float)gy);
                else grad[g] = make_double2(gx, gy);
            }
            const double occ = (s_claim[r * W + c] >= 0) ? 1.0 : 0.0;
            const double cf = (a.rate_feed * f) * occ;              // consumed_field, core/env.py:224
            food_out[g] = next_food<PLAIN>(a, f, cf, gi, c, g);
            occ_out[g] = occ;
        };
        if (food_pre) {
            int r = 0, seg = warp;
#pragma unroll
            for (int u = 0; u < MAXT; ++u) {
                while (seg >= segs) { seg -= segs; ++r; }
                const int c = seg * 32 + lane;
                if (r < rows_per && c < W) emit(r, c, fpre[u]);
                seg += NW;
            }
        } else {
            for (int r = 0, seg = warp;; seg += NW) {
                while (seg >= segs) { seg -= segs; ++r; }
                if (r >= rows_per) break;
                const int c = seg * 32 + lane;
                if (c < W) emit(r, c, food_in[(int64_t)(r0 + r) * W + c]);
            }
        }
    }

    // ---- phase 3: feed (agent_feed_kernel's arithmetic and block partials) ----------------------------------------
#pragma unroll
    for (int j = 0; j < kFusedMaxRounds; ++j) {
        if (j < rounds) {                                  // (uniform over the CTA)
            const int blk = blk_lo + j * VB + sub;
            const bool active = blk < blk_hi;
            const int64_t first = (int64_t)blk * kBlockSlots + t256;
            double gain_sum = 0.0;
            int alive_cnt = 0;
            double dx[kFeedItems], dy[kFeedItems], dep[kFeedItems], stock[kFeedItems], fd[kFeedItems];
            int32_t claim[kFeedItems];
            bool valid[kFeedItems];
#pragma unroll
            for (int k = 0; k < kFeedItems; ++k) {
                const int64_t i = first + k * kAgentThreads;
                valid[k] = active && i < M;
                const int cell = mycell[j][k];
                const int q = (a.slab_shift >= 0) ? (cell >> a.slab_shift) : (cell / slab_cells);
                const int32_t* slice = (q == rank) ? s_claim : cluster_map(s_claim, (unsigned)q);
                claim[k] = valid[k] ? slice[cell - q * slab_cells] : -1;
                fd[k] = valid[k] ? food_in[cell] : 0.0;
                stock[k] = valid[k] ? ag[3 * M + i] : 0.0;
                dx[k] = valid[k] ? ac[i] : 0.0;
                dy[k] = valid[k] ? ac[M + i] : 0.0;
                dep[k] = valid[k] ? ac[2 * M + i] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < kFeedItems; ++k) {
                if (valid[k]) {
                    const int64_t i = first + k * kAgentThreads;
                    // (rate_feed * food) * occ: the field phase's own expression for consumed_field
                    const double eaten = (a.rate_feed * fd[k]) * ((claim[k] >= 0) ? 1.0 : 0.0);
                    const double burned = a.w_dep * fabs(dep[k]) + a.w_dist * sqrt(dx[k] * dx[k] + dy[k] * dy[k]);
                    const double gained = eaten - burned;
                    ag[3 * M + i] = stock[k] + gained;
                    gain_sum += gained;
                    if ((myalive >> (j * kFeedItems + k)) & 1u) ++alive_cnt;
                }
            }
            gain_sum = warp_sum(gain_sum);
            alive_cnt = warp_sum(alive_cnt);
            if (lane == 0) {
                s_gain[j * NW + warp] = gain_sum;
                s_alive[j * NW + warp] = alive_cnt;
            }
        }
    }
    __syncthreads();
    if (t256 == 0) {                                       // the leader of each feed block: its 8 warp sums in order
        for (int j = 0; j < rounds; ++j) {
            const int blk = blk_lo + j * VB + sub;
            if (blk < blk_hi) {
                double gsum = 0.0;
                int n = 0;
#pragma unroll
                for (int k = 0; k < kAgentThreads / 32; ++k) {
                    gsum += s_gain[j * NW + sub * (kAgentThreads / 32) + k];
                    n += s_alive[j * NW + sub * (kAgentThreads / 32) + k];
                }
                a.part_gain[b * a.nblk + blk] = gsum;
                a.part_alive[b * a.nblk + blk] = n;
            }
        }
    }
    cluster_sync();                                        // partials written; nobody reads my shared memory any more

    // ---- phase 4: reward, num_agents -----------------------------------------------------------------------------
    if (rank == 0)
        finalize_sum<NT, true>(a.part_gain + b * a.nblk, a.part_alive + b * a.nblk, a.nblk, s_fg, s_fn,
                               a.reward + b, a.alive_out + b);
}

}  // namespace die
