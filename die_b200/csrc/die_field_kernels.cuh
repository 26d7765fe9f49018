// die_field_kernels.cuh -- the dense half of Env.step: one pass over every cell that fuses
//   pheromone deposit     chem[cell] = chem[cell] + deposit[winner]   core/env.py:211  (Q2)
//   occupancy layout      medium['agents'] = 0; [cells] = 1            core/env.py:214-215
//   food consumption      food -= rate_feed * food * (occ > 0)         core/env.py:224-228
//   food flow             identity                                     core/env.py:147-150
//   diffusion * decay     gaussian(chem, sigma, 'wrap') * (1-d)        core/env.py:136-145
// reading medium_in (never written) and the claim table, writing medium_out and the per-cell
// consumed_field (rate_feed * food * occ, core/env.py:224) that the agent feed kernel gathers.
// Algorithmic traffic: chem R+W, food R+W, occupancy W (+ claim table 4 B R, consumed 8 B W).
// (Round 2 tried to drop the consumed_field scratch: the field pass wrote one occupancy bit per cell and the feed
//  kernel formed rate_feed * food * occ from medium_in's food and that bit.  8 B/cell less written, bit-identical -- and
//  slower on a B200: the feed kernel is bound by its gathers, and two of them per slot (food + bit word) cost more than
//  the coalesced 8-byte store saved: feed 2.61 -> 2.99 ms, field 3.26 -> 3.26 ms batched; 0.224 -> 0.261 / 0.225 -> 0.212 ms
//  at 4096^2, profiles/r02e_bench_after_consumed_removal.txt.  Reverted.)
//
// The deposit is applied while the halo tile is staged: every staged cell reads its claim
// (the winning slot, or -1) and, when claimed, gathers that slot's deposit1 -- so the previous
// medium buffer stays intact and no scatter into the field is needed.
//
// The blur is scipy.ndimage's separable correlate1d in its exact operation order (axis 0 then
// axis 1, each  x0*w0 + (x[-r]+x[+r])*w[-r] + ... + (x[-1]+x[+1])*w[-1], periodic), so the result
// is bit-identical to skimage.filters.gaussian on the same input.
#pragma once
#include "die_device.cuh"
#include "die_slab.cuh"
#include "../../include/die_b200.h"

namespace die {

struct BlurWeights {
    double w[2 * DIE_MAX_RADIUS + 1];
};

struct FieldArgs {
    const double* medium_in;     // element type = the kernel's FT (float64, or float32 in the env's float32 field mode:
    double* medium_out;          // the three field arrays medium_in / medium_out / consumed are then arrays of float)
    const int32_t* winner;       // [B][H*W] claim table (read only here; the feed kernel clears it)
    const double* action;        // [B][3][M]; channel 2 = deposit1
    double* consumed;            // [B][H*W] consumed_field out
    double2* cell_pairs;         // non-null: [B][H*W] {consumed_field, new env_food} out INSTEAD of consumed (large fields: the
                                 // feed kernel's one gather then also brings the food the next forward pass needs)
    double2* grad;               // [B][H*W] np.gradient(chem_out) (raw d/dx, d/dy) or null
    float2* grad32;              // the same pairs rounded to float32 (tuning "grad_f32"): what the guard-banded quick turn
                                 // decision reads anyway (die_turn.h); at most one of grad / grad32 is set
    int H, W;
    int64_t M;
    int tiles_i, tiles_j;
    double rate_feed;
    double keep;                 // 1. - rate_decay_chem
    int food_infinite;
    int diffuse_mode;            // DIE_DIFFUSE_* (the slab instantiation is wrap only)
    int prefetch_food;           // tile kernel: L2 prefetch of the output tile's food lines (tuning switch)
    // op_food_flow = WaveSequence(...).get_flow_operator(scale, decay), core/data_init.py:29-38, 71-89 (null = identity)
    const double* flow_rwave;    // [H*W]  r + cos(pi x) + sin(0.4 pi y): the time-independent part of the wave phase
    const double* flow_col;      // [W]    sin(pi x 3 + t) of the current time step, per column
    const double* flow_row;      // [H]    cos(pi y 3 + t) of the current time step, per row
    double flow_t, flow_scale, flow_keep;
    const double* flow_frame;    // [H*W]  F_t of the current time step, tabulated by the host (any FieldSequence: the
                                 //        reference's PerlinNoiseSequence, a user's own); null = wave closed form / identity
    BlurWeights bw;              // centre at [R]
    SlabGeom sg;                 // SLAB instantiation only (H, W above are then the GLOBAL field)
    SlabTables st;
};

// Env._agent_feed's subtraction (core/env.py:227-228) followed by Env._medium_resource_dynamics (:147-150):
//   food = op_food_flow(food - consumed_field),  identity or  scale * F_t + (1 - decay) * food  with
//   F_t = (1 - mix) cos(1 pi (rwave + t)) + mix (sin(pi x 3 + t) + cos(pi y 3 + t)),  mix = 0.25   (WaveSequence)
// The cosine is die_math.h's (<= 0.7 ulp, bit-identical to the oracle's portable backend); everything
// time-independent or separable was tabulated by numpy on the host.
template <bool PLAIN = false, typename ARGS = FieldArgs>
__device__ __forceinline__ double next_food(const ARGS& a, double f, double cf, int row, int col, int64_t g) {
    double food = a.food_infinite ? f : f - cf;
    if (!PLAIN && a.flow_frame != nullptr) {           // scale * next(it) + (1 - decay) * current, core/data_init.py:35
        food = a.flow_scale * a.flow_frame[g] + a.flow_keep * food;
    } else if (!PLAIN && a.flow_rwave != nullptr) {
        double sn, cs;
        die_sincos(kPi * (a.flow_rwave[g] + a.flow_t), &sn, &cs);
        const double islands = a.flow_col[col] + a.flow_row[row];
        const double z = 0.75 * cs + 0.25 * islands;
        food = a.flow_scale * z + a.flow_keep * food;
    }
    return food;
}

__device__ __forceinline__ int wrap_index(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}

// the same for -n <= i < 2n (a halo tile on a field at least as large as the tile): no integer division.  The
// modulo by a run-time n costs ~20 instructions, twice per staged cell -- a fifth of the field pass's instructions,
// and the pass turned out to be issue-bound, not DRAM-bound (float32 fields made it 10 % faster, not 40 %).
__device__ __forceinline__ int wrap_near(int i, int n) {
    if (i < 0) i += n;
    else if (i >= n) i -= n;
    return i;
}

// scipy.ndimage boundary extension (skimage.filters.gaussian(mode=...), core/env.py:140-143): the source index an
// out-of-range index i stands for, or -1 for 'constant' (cval = 0).
//   wrap      a b c d | a b c d | a b c d        nearest   a a a a | a b c d | d d d d
//   reflect   d c b a | a b c d | d c b a        mirror    d c b | a b c d | c b a
__device__ __forceinline__ int extend_index(int i, int n, int mode) {
    if (mode == DIE_DIFFUSE_WRAP) return wrap_index(i, n);
    if (i >= 0 && i < n) return i;
    if (mode == DIE_DIFFUSE_NEAREST) return i < 0 ? 0 : n - 1;
    if (mode == DIE_DIFFUSE_REFLECT) {
        const int p = 2 * n;
        i = wrap_index(i, p);
        return i < n ? i : p - 1 - i;
    }
    if (mode == DIE_DIFFUSE_MIRROR) {
        if (n == 1) return 0;
        const int p = 2 * n - 2;
        i = wrap_index(i, p);
        return i < n ? i : p - i;
    }
    return -1;                                     // DIE_DIFFUSE_CONSTANT
}

// Tile TH x TW outputs per CTA.  With GRAD the CTA blurs a one-cell ring more ((TH+2) x (TW+2))
// so that np.gradient of the NEW chem field can be formed per output cell (central differences
// inside, one-sided at the global edges, non-periodic -- core/agent/gradient.py:57) and published
// for GradientAgent / PhysarumAgent.forward, which then needs ONE 16-byte gather per slot instead
// of four 8-byte ones.  G = ring width (0 or 1).
//   s_in  [(TH+2G+2R)][(TW+2G+2R)]  staged chem (+deposit), reused for the blurred ring tile
//   s_v   [(TH+2G)][(TW+2G+2R)]     after the axis-0 pass
// Staging is done in two sweeps so that all of a thread's (independent) chem + claim loads are
// in flight before the first dependent deposit gather.
// (register caps were tried: 40 / 48 / 56 / 72 registers give 299 / 279 / 250 / 273 us at 4096^2; the compiler's own
//  choice without a minimum-blocks hint, 60 registers = 4 CTAs per SM, is the best at 241 us)
// PLAIN: the reference's default dynamics (periodic diffusion, identity food flow) known at compile time.
// FT: element type of the field arrays in HBM (medium, consumed_field).  float64 as in the reference, or float32 (the
// env's float32 field mode): values are widened on load, every operation stays the float64 one, and
// results are rounded once on store -- 24 instead of 48 B per cell-update.  The published gradient is formed from the
// ROUNDED chem1 values, so it equals np.gradient of the stored field exactly (what a forward kernel that samples chem1
// itself computes).
template <int R, int TH, int TW, int NT, bool GRAD, bool SLAB, bool PLAIN = false, typename FT = double>
__global__ void __launch_bounds__(NT)
field_step_kernel(const FieldArgs a) {
    static_assert(!SLAB || sizeof(FT) == 8, "the slab decomposition runs float64 fields");
    constexpr int G = GRAD ? 1 : 0;
    constexpr int OH = TH + 2 * G, OW = TW + 2 * G;     // blurred region
    constexpr int LW = OW + 2 * R;                      // staged row length
    constexpr int LH = OH + 2 * R;
    constexpr int NSTAGE = (LH * LW + NT - 1) / NT;
    extern __shared__ double smem[];
    double* s_in = smem;                     // [LH][LW]
    double* s_v = smem + LH * LW;            // [OH][LW]
    double* s_out = smem;                    // [OH][OW] blurred * keep (aliases s_in, GRAD only)

    const int H = a.H, W = a.W;              // GLOBAL field (== the local one unless SLAB)
    const int HL = SLAB ? a.sg.rows_per : H; // rows this launch produces
    const int row0 = SLAB ? a.sg.rank * a.sg.rows_per : 0;
    const int64_t C = (int64_t)HL * W;
    const int tiles = a.tiles_i * a.tiles_j;
    const int64_t b = blockIdx.x / (unsigned)tiles;
    const int t = blockIdx.x - (int)b * tiles;
    const int ti = t / a.tiles_j, tj = t - ti * a.tiles_j;
    const int i0 = ti * TH, j0 = tj * TW;    // LOCAL row / column of the tile

    const FT* min_l = SLAB ? (const FT*)a.st.medium_in[a.sg.rank] : (const FT*)a.medium_in + b * 3 * C;
    FT* mout_l = SLAB ? (FT*)a.st.medium_out[a.sg.rank] : (FT*)a.medium_out + b * 3 * C;
    const FT* food_in = min_l + C;
    const FT* chem_in = min_l + 2 * C;
    FT* occ_out = mout_l;
    FT* food_out = mout_l + C;
    FT* chem_out = mout_l + 2 * C;
    const int32_t* win = SLAB ? a.st.claim[a.sg.rank] : a.winner + b * C;
    const double* dep = SLAB ? nullptr : a.action + (b * 3 + 2) * a.M;
    FT* cons = SLAB ? (FT*)a.st.consumed[a.sg.rank] : (FT*)a.consumed + b * C;

    // food of the output tile is only needed by the last phase: pull its lines towards L2 now, so that
    // those loads do not start a fresh DRAM round trip after the blur (one 128-byte line per thread)
    if (a.prefetch_food) {
        constexpr int LINES_PER_ROW = TW * (int)sizeof(FT) / 128;
        for (int l = threadIdx.x; l < TH * LINES_PER_ROW; l += NT) {
            const int r = l / LINES_PER_ROW, c = (l - r * LINES_PER_ROW) * (128 / (int)sizeof(FT));
            if (i0 + r < HL && j0 + c < W)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(food_in + (int64_t)(i0 + r) * W + j0 + c));
        }
    }

    // ---- stage the periodic halo tile, deposit included -------------------------------------
    const bool near = H >= LH && W >= LW;    // every staged index is within one period of the field
    double v[NSTAGE];
    int w[NSTAGE];
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
        const int idx = threadIdx.x + s * NT;
        v[s] = 0.0;
        w[s] = -1;
        if (idx < LH * LW) {
            const int r = idx / LW, c = idx - r * LW;
            // global row / column this staged cell stands for (periodic over the WHOLE field by default)
            int gi, gj;
            if (SLAB || PLAIN) {
                if (near) {                   // (uniform: the field is at least as large as the halo tile)
                    gi = wrap_near(row0 + i0 - G - R + r, H);
                    gj = wrap_near(j0 - G - R + c, W);
                } else {
                    gi = wrap_index(row0 + i0 - G - R + r, H);
                    gj = wrap_index(j0 - G - R + c, W);
                }
            } else {
                gi = extend_index(i0 - G - R + r, H, a.diffuse_mode);
                gj = extend_index(j0 - G - R + c, W, a.diffuse_mode);
            }
            const int g = gi * W + gj;
            if (SLAB) {                       // rows outside this rank's slab come from the neighbours over NVLink
                v[s] = __ldg(slab_chan(a.st.medium_in, a.sg, 2, g));
                w[s] = __ldg(slab_cell(a.st.claim, a.sg, g));
            } else if (PLAIN || (gi >= 0 && gj >= 0)) {  // ('constant' extension: 0, no deposit)
                v[s] = (double)chem_in[g];
                w[s] = win[g];
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
        const int idx = threadIdx.x + s * NT;
        if (idx < LH * LW) {
            double x = v[s];
            if (w[s] >= 0) {
                if (SLAB) {                   // the winner is a GLOBAL slot id: its deposit lives on its owner
                    int64_t local;
                    const int q = slab_slot_owner(a.sg, w[s], local);
                    x = x + __ldg(a.st.action[q] + 2 * slab_slots_of(a.sg, q) + local);
                } else {
                    x = x + dep[w[s]];
                }
            }
            s_in[idx] = x;
        }
    }
    __syncthreads();

    // ---- axis-0 pass --------------------------------------------------------------------
    for (int idx = threadIdx.x; idx < OH * LW; idx += NT) {
        const int r = idx / LW, c = idx - r * LW;
        const double* p = s_in + (r + R) * LW + c;
        double acc = p[0] * a.bw.w[R];
#pragma unroll
        for (int k = R; k >= 1; --k) acc += (p[-k * LW] + p[k * LW]) * a.bw.w[R - k];
        s_v[idx] = acc;
    }
    __syncthreads();

    if constexpr (!GRAD) {
        // ---- axis-1 pass + elementwise channels ---------------------------------------------
        for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
            const int r = idx / TW, c = idx - r * TW;
            const int li = i0 + r, gj = j0 + c;
            const int g = li * W + gj;
            if (li < HL && gj < W) {
                const double* p = s_v + r * LW + c + R;
                double acc = p[0] * a.bw.w[R];
#pragma unroll
                for (int k = R; k >= 1; --k) acc += (p[-k] + p[k]) * a.bw.w[R - k];
                chem_out[g] = (FT)(acc * a.keep);
                const double occ = (win[g] >= 0) ? 1.0 : 0.0;
                const double f = (double)food_in[g];
                const double cf = (a.rate_feed * f) * occ;      // consumed_field, core/env.py:224
                const double fnew = next_food<SLAB || PLAIN>(a, f, cf, li, gj, g);
                food_out[g] = (FT)fnew;
                occ_out[g] = (FT)occ;
                if (!SLAB && sizeof(FT) == 8 && a.cell_pairs != nullptr) a.cell_pairs[b * C + g] = make_double2(cf, fnew);
                else cons[g] = (FT)cf;
            }
        }
    } else {
    // ---- axis-1 pass over the ring tile into shared memory ------------------------------------
    for (int idx = threadIdx.x; idx < OH * OW; idx += NT) {
        const int r = idx / OW, c = idx - r * OW;
        const double* p = s_v + r * LW + c + R;
        double acc = p[0] * a.bw.w[R];
#pragma unroll
        for (int k = R; k >= 1; --k) acc += (p[-k] + p[k]) * a.bw.w[R - k];
        s_out[idx] = (double)(FT)(acc * a.keep);         // (float32 fields: the value as it will be stored)
    }
    __syncthreads();

    // ---- outputs: new chem, its np.gradient, elementwise channels ------------------------------
    double2* grad = SLAB ? a.st.grad[a.sg.rank] : (a.grad != nullptr ? a.grad + b * C : nullptr);
    float2* grad32 = (!SLAB && a.grad32 != nullptr) ? a.grad32 + b * C : nullptr;
    for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
        const int r = idx / TW, c = idx - r * TW;
        const int li = i0 + r, gj = j0 + c;
        const int gi = row0 + li;            // global row: np.gradient is one-sided on the GLOBAL border only
        const int g = li * W + gj;
        if (li < HL && gj < W) {
            const double* q = s_out + (r + 1) * OW + (c + 1);
            chem_out[g] = (FT)q[0];
            // np.gradient: (f[i+1] - f[i-1]) / 2 inside, f[1] - f[0] / f[n-1] - f[n-2] at the edges
            const int um = (gi > 0) ? -OW : 0, up = (gi < H - 1) ? OW : 0;
            const int lm = (gj > 0) ? -1 : 0, lp = (gj < W - 1) ? 1 : 0;
            double gx = q[up] - q[um];
            double gy = q[lp] - q[lm];
            if (up - um == 2 * OW) gx *= 0.5;
            if (lp - lm == 2) gy *= 0.5;
            if (grad32 != nullptr) grad32[g] = make_float2((float)gx, (float)gy);
            else grad[g] = make_double2(gx, gy);

            const double occ = (win[g] >= 0) ? 1.0 : 0.0;
            const double f = (double)food_in[g];
            const double cf = (a.rate_feed * f) * occ;          // consumed_field, core/env.py:224
            const double fnew = next_food<SLAB || PLAIN>(a, f, cf, li, gj, g);
            food_out[g] = (FT)fnew;
            occ_out[g] = (FT)occ;
            if (!SLAB && sizeof(FT) == 8 && a.cell_pairs != nullptr) a.cell_pairs[b * C + g] = make_double2(cf, fnew);
            else cons[g] = (FT)cf;
        }
    }
    }   // GRAD
}

// ---------------------------------------------------------------------------------------------
// The same pass with 128-bit global accesses, for the reference's default dynamics (periodic diffusion, identity food
// flow: PLAIN) on fields with an even row length.  Same tile, same shared-memory arithmetic, same results; what changes
// is how the tile gets in and out:
//   staging   the halo tile starts at an EVEN column (one pad column when the halo width is odd), so every staged row
//             is a run of aligned cell pairs: one LDG.E.128 per two chem1 values and one LDG.E.64 per two claims
//             (a pair never straddles the periodic wrap: W is even);
//   outputs   a thread owns two adjacent cells: LDG.E.128 for the food, one STG.E.128 each for chem1, food, occupancy
//             and consumed_field, and one for the two float32 gradient pairs.
// Half as many memory instructions per cell as field_step_kernel -- and, measured on a B200, no faster (the pass is
// DRAM-bound: profiles/r02l_field_vec_ab.txt), so this is an opt-in (die_set_tuning("field_vec", 1)) and the scalar
// version stays the default.
// ---------------------------------------------------------------------------------------------
template <int R, int TH, int TW, int NT, bool GRAD>
__global__ void __launch_bounds__(NT)
field_step_vec_kernel(const FieldArgs a) {
    constexpr int G = GRAD ? 1 : 0;
    constexpr int HALO = R + G;
    constexpr int PADL = HALO & 1;                      // staged column 0 = field column j0 - HALO - PADL (even)
    constexpr int OH = TH + 2 * G, OW = TW + 2 * G;     // blurred region
    constexpr int LW = OW + 2 * R;                      // columns the blur reads
    constexpr int LH = OH + 2 * R;
    constexpr int SW = (PADL + LW + 1) & ~1;            // staged row length (even)
    constexpr int PW = SW / 2;                          // pairs per staged row
    constexpr int NPAIR = (LH * PW + NT - 1) / NT;
    static_assert(TW % 2 == 0 && (TH * TW / 2) % NT == 0, "whole pairs, whole passes");
    extern __shared__ double smem[];
    double* s_in = smem;                     // [LH][SW]
    double* s_v = smem + LH * SW;            // [OH][LW]
    double* s_out = smem;                    // [OH][OW] blurred * keep (aliases s_in, GRAD only)

    const int H = a.H, W = a.W;
    const int64_t C = (int64_t)H * W;
    const int tiles = a.tiles_i * a.tiles_j;
    const int64_t b = blockIdx.x / (unsigned)tiles;
    const int t = blockIdx.x - (int)b * tiles;
    const int ti = t / a.tiles_j, tj = t - ti * a.tiles_j;
    const int i0 = ti * TH, j0 = tj * TW;

    const double* __restrict__ min_l = a.medium_in + b * 3 * C;
    double* __restrict__ mout_l = a.medium_out + b * 3 * C;
    const double* __restrict__ food_in = min_l + C;
    const double* __restrict__ chem_in = min_l + 2 * C;
    double* __restrict__ occ_out = mout_l;
    double* __restrict__ food_out = mout_l + C;
    double* __restrict__ chem_out = mout_l + 2 * C;
    const int32_t* __restrict__ win = a.winner + b * C;
    const double* __restrict__ dep = a.action + (b * 3 + 2) * a.M;
    double* __restrict__ cons = a.consumed + b * C;

    if (a.prefetch_food) {
        constexpr int LINES_PER_ROW = TW * 8 / 128;
        for (int l = threadIdx.x; l < TH * LINES_PER_ROW; l += NT) {
            const int r = l / LINES_PER_ROW, c = (l - r * LINES_PER_ROW) * 16;
            if (i0 + r < H && j0 + c < W)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(food_in + (int64_t)(i0 + r) * W + j0 + c));
        }
    }

    // ---- stage the periodic halo tile pair by pair, deposit included --------------------------------
    double2 v[NPAIR];
    int2 w[NPAIR];
#pragma unroll
    for (int s = 0; s < NPAIR; ++s) {
        const int idx = threadIdx.x + s * NT;
        v[s] = make_double2(0.0, 0.0);
        w[s] = make_int2(-1, -1);
        if (idx < LH * PW) {
            const int r = idx / PW, c = (idx - r * PW) * 2;
            const int gi = wrap_index(i0 - HALO + r, H);
            const int gj = wrap_index(j0 - HALO - PADL + c, W);          // even; gj + 1 < W
            const int g = gi * W + gj;
            v[s] = *reinterpret_cast<const double2*>(chem_in + g);
            w[s] = *reinterpret_cast<const int2*>(win + g);
        }
    }
#pragma unroll
    for (int s = 0; s < NPAIR; ++s) {
        const int idx = threadIdx.x + s * NT;
        if (idx < LH * PW) {
            double2 x = v[s];
            if (w[s].x >= 0) x.x = x.x + dep[w[s].x];
            if (w[s].y >= 0) x.y = x.y + dep[w[s].y];
            *reinterpret_cast<double2*>(s_in + 2 * idx) = x;             // [r][c], c even: idx * 2 = r * SW + c
        }
    }
    __syncthreads();

    // ---- axis-0 pass (columns PADL .. PADL + LW of the staged rows) -----------------------------------
    for (int idx = threadIdx.x; idx < OH * LW; idx += NT) {
        const int r = idx / LW, c = idx - r * LW;
        const double* p = s_in + (r + R) * SW + PADL + c;
        double acc = p[0] * a.bw.w[R];
#pragma unroll
        for (int k = R; k >= 1; --k) acc += (p[-k * SW] + p[k * SW]) * a.bw.w[R - k];
        s_v[idx] = acc;
    }
    __syncthreads();

    if constexpr (GRAD) {
        // ---- axis-1 pass over the ring tile into shared memory ---------------------------------------
        for (int idx = threadIdx.x; idx < OH * OW; idx += NT) {
            const int r = idx / OW, c = idx - r * OW;
            const double* p = s_v + r * LW + c + R;
            double acc = p[0] * a.bw.w[R];
#pragma unroll
            for (int k = R; k >= 1; --k) acc += (p[-k] + p[k]) * a.bw.w[R - k];
            s_out[idx] = acc * a.keep;
        }
        __syncthreads();
    }

    // ---- outputs, two adjacent cells per thread -------------------------------------------------------
    float2* grad32 = (GRAD && a.grad32 != nullptr) ? a.grad32 + b * C : nullptr;
    double2* grad = (GRAD && a.grad != nullptr) ? a.grad + b * C : nullptr;
    for (int idx = threadIdx.x; idx < TH * TW / 2; idx += NT) {
        const int r = idx / (TW / 2), c = (idx - r * (TW / 2)) * 2;
        const int li = i0 + r, gj = j0 + c;
        if (li < H && gj < W) {                          // (gj even and W even: gj + 1 < W as well)
            const int g = li * W + gj;
            double o0, o1;
            if constexpr (GRAD) {
                const double* q = s_out + (r + 1) * OW + (c + 1);
                o0 = q[0];
                o1 = q[1];
                // np.gradient: (f[i+1] - f[i-1]) / 2 inside, f[1] - f[0] / f[n-1] - f[n-2] at the edges
                const int um = (li > 0) ? -OW : 0, up = (li < H - 1) ? OW : 0;
                double gx0 = q[up] - q[um], gx1 = q[1 + up] - q[1 + um];
                if (up - um == 2 * OW) { gx0 *= 0.5; gx1 *= 0.5; }
                const int lm0 = (gj > 0) ? -1 : 0;                       // cell gj: right neighbour gj + 1 always exists
                double gy0 = q[1] - q[lm0];
                if (lm0 != 0) gy0 *= 0.5;
                const int lp1 = (gj + 1 < W - 1) ? 1 : 0;                // cell gj + 1: left neighbour gj always exists
                double gy1 = q[1 + lp1] - q[0];
                if (lp1 != 0) gy1 *= 0.5;
                if (grad32 != nullptr) {
                    *reinterpret_cast<float4*>(grad32 + g) = make_float4((float)gx0, (float)gy0, (float)gx1, (float)gy1);
                } else {
                    grad[g] = make_double2(gx0, gy0);
                    grad[g + 1] = make_double2(gx1, gy1);
                }
            } else {
                const double* p = s_v + r * LW + c + R;
                double acc0 = p[0] * a.bw.w[R], acc1 = p[1] * a.bw.w[R];
#pragma unroll
                for (int k = R; k >= 1; --k) {
                    acc0 += (p[-k] + p[k]) * a.bw.w[R - k];
                    acc1 += (p[1 - k] + p[1 + k]) * a.bw.w[R - k];
                }
                o0 = acc0 * a.keep;
                o1 = acc1 * a.keep;
            }
            *reinterpret_cast<double2*>(chem_out + g) = make_double2(o0, o1);
            const int2 wn = *reinterpret_cast<const int2*>(win + g);
            const double2 f = *reinterpret_cast<const double2*>(food_in + g);
            const double occ0 = (wn.x >= 0) ? 1.0 : 0.0, occ1 = (wn.y >= 0) ? 1.0 : 0.0;
            const double cf0 = (a.rate_feed * f.x) * occ0, cf1 = (a.rate_feed * f.y) * occ1;   // consumed_field, core/env.py:224
            *reinterpret_cast<double2*>(food_out + g) = make_double2(next_food<true>(a, f.x, cf0, li, gj, g),
                                                                     next_food<true>(a, f.y, cf1, li, gj + 1, g + 1));
            *reinterpret_cast<double2*>(occ_out + g) = make_double2(occ0, occ1);
            *reinterpret_cast<double2*>(cons + g) = make_double2(cf0, cf1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Env._get_sensed_medium / _get_sense_mask (core/env.py:275-294), Dynamics.apply_sense_mask:
//   mask = ceil(round(gaussian(occupancy, sigma=2.0), 3));   obs_medium = medium.where(mask, other=0.)
// skimage's default mode is 'nearest' and truncate 4.0, i.e. a radius-8 blur with clamped borders, in
// scipy's operation order (as the diffusion pass above); round(v, 3) = rint(v * 1000) / 1000.
// One CTA per TH x TW tile: stage the clamped halo tile of the occupancy channel, two 1-D passes,
// then the three channels of the tile are copied or zeroed.
// ---------------------------------------------------------------------------------------------
template <int R, int TH, int TW, int NT>
__global__ void __launch_bounds__(NT)
sense_mask_kernel(const double* __restrict__ medium, double* __restrict__ obs, int H, int W,
                  int tiles_i, int tiles_j, const BlurWeights bw) {
    constexpr int LW = TW + 2 * R, LH = TH + 2 * R;
    extern __shared__ double smem[];
    double* s_in = smem;                     // [LH][LW]
    double* s_v = smem + LH * LW;            // [TH][LW]
    const int64_t C = (int64_t)H * W;
    const int tiles = tiles_i * tiles_j;
    const int64_t b = blockIdx.x / (unsigned)tiles;
    const int t = blockIdx.x - (int)b * tiles;
    const int ti = t / tiles_j, tj = t - ti * tiles_j;
    const int i0 = ti * TH, j0 = tj * TW;
    const double* med = medium + b * 3 * C;
    double* out = obs + b * 3 * C;

    for (int idx = threadIdx.x; idx < LH * LW; idx += NT) {
        const int r = idx / LW, c = idx - r * LW;
        const int gi = min(max(i0 - R + r, 0), H - 1), gj = min(max(j0 - R + c, 0), W - 1);     // 'nearest'
        s_in[idx] = med[(int64_t)gi * W + gj];                                                 // channel 0: occupancy
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < TH * LW; idx += NT) {
        const int r = idx / LW, c = idx - r * LW;
        const double* p = s_in + (r + R) * LW + c;
        double acc = p[0] * bw.w[R];
#pragma unroll
        for (int k = R; k >= 1; --k) acc += (p[-k * LW] + p[k * LW]) * bw.w[R - k];
        s_v[idx] = acc;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
        const int r = idx / TW, c = idx - r * TW;
        const int gi = i0 + r, gj = j0 + c;
        if (gi < H && gj < W) {
            const double* p = s_v + r * LW + c + R;
            double acc = p[0] * bw.w[R];
#pragma unroll
            for (int k = R; k >= 1; --k) acc += (p[-k] + p[k]) * bw.w[R - k];
            const bool seen = ceil(rint(acc * 1000.0) / 1000.0) != 0.0;
            const int64_t g = (int64_t)gi * W + gj;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) out[ch * C + g] = seen ? med[ch * C + g] : 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// EnvRenderer.render (core/render.py:76-132) + FieldTrace.update (:9-29): the three frames of one step in
// one pass, for plotting / GIF export (the step either side of the hot path).
//   medium frame  [H][W][3]   (agents, env_food, chem1) interleaved; optional colour rotation
//                             rgb -> cross(color, rgb) (RendererBase._set_colors, :47-58)
//   trace         [H][W]      trace = trace * decay + occupancy  (in place; the caller colour-maps it)
//   agents frame  [Wd][Ht][4] (0, agent_food, 0, alive != 0) with the reference's reshape
//                             (2, height, -1) -> transpose(1, 2, 0), where width, height = field_size
// One thread per cell (and per agent slot: the agents frame needs M == H*W, as in the reference).
// ---------------------------------------------------------------------------------------------
struct RenderArgs {
    const double* medium;     // [B][3][H][W]
    const double* agents;     // [B][4][M]
    double* trace;            // [B][H][W] in/out
    double* img_medium;       // [B][H][W][3]
    double* img_agents;       // [B][M][4] == [B][height = W][M / W][4], or null
    int H, W;
    int64_t M;
    double decay;
    int use_color;
    double color[3];          // normalised
};

__global__ void __launch_bounds__(256)
render_frames_kernel(const RenderArgs a, int64_t total) {
    const int64_t C = (int64_t)a.H * a.W;
    for (int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x; gid < total; gid += (int64_t)gridDim.x * 256) {
        const int64_t b = gid / C, g = gid - b * C;
        const double* med = a.medium + b * 3 * C;
        const double r = med[g], gr = med[C + g], bl = med[2 * C + g];
        double* px = a.img_medium + (b * C + g) * 3;
        if (a.use_color) {            // np.cross(color, rgb): (c1 b2 - c2 b1, c2 b0 - c0 b2, c0 b1 - c1 b0)
            px[0] = a.color[1] * bl - a.color[2] * gr;
            px[1] = a.color[2] * r - a.color[0] * bl;
            px[2] = a.color[0] * gr - a.color[1] * r;
        } else {
            px[0] = r;
            px[1] = gr;
            px[2] = bl;
        }
        a.trace[gid] = a.trace[gid] * a.decay + r;
        if (a.img_agents != nullptr) {          // slot g of the (2, M) block -> pixel g of the [W][M/W] image
            const double* ag = a.agents + b * 4 * a.M;
            double* q = a.img_agents + (b * a.M + g) * 4;
            q[0] = 0.0;
            q[1] = ag[3 * a.M + g];
            q[2] = 0.0;
            q[3] = (ag[2 * a.M + g] != 0.0) ? 1.0 : 0.0;
        }
    }
}

// No diffusion (blur_radius == 0): gaussian with radius 0 is the identity (w = [1]).
template <int NT, typename FT = double>
__global__ void __launch_bounds__(NT)
field_step_noblur_kernel(const FieldArgs a, int64_t total) {
    const int64_t C = (int64_t)a.H * a.W;
    for (int64_t gid = (int64_t)blockIdx.x * NT + threadIdx.x; gid < total; gid += (int64_t)gridDim.x * NT) {
        const int64_t b = gid / C, g = gid - b * C;
        const FT* min = (const FT*)a.medium_in + b * 3 * C;
        FT* mout = (FT*)a.medium_out + b * 3 * C;
        const int w = a.winner[gid];
        const double occ = (w >= 0) ? 1.0 : 0.0;
        const double f = (double)min[C + g];
        const double cf = (a.rate_feed * f) * occ;
        double chem = (double)min[2 * C + g];
        if (w >= 0) chem = chem + a.action[(b * 3 + 2) * a.M + w];
        mout[g] = (FT)occ;
        mout[C + g] = (FT)next_food(a, f, cf, (int)(g / a.W), (int)(g % a.W), g);
        mout[2 * C + g] = (FT)((chem * a.bw.w[0]) * a.bw.w[0] * a.keep);
        ((FT*)a.consumed)[gid] = (FT)cf;
    }
}

}  // namespace die
