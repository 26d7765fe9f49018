// die_field_bulk.cuh -- the field pass (die_field_kernels.cuh, same arithmetic, same outputs) as a PERSISTENT kernel
// whose halo tiles arrive by asynchronous bulk copies (the TMA unit, cp.async.bulk global -> shared, completing on an
// mbarrier) into a two-stage ring: while a CTA blurs tile k, the rows of tile k+1 -- raw chem1 and the claim table --
// are already landing in the other stage, without passing through registers.  The deposit fix-up (chem += deposit of
// the cell's winning slot) then runs over the tile in shared memory, and the occupancy of the output cells
// is read from the staged claims instead of a second global read.
//
// STAGED at the end of round 1 (tuning "field_impl" = 2): verified bit-identical to the tile kernel on the CPU
// emulator (tests/test_hostsim_kernels.py), compiled for sm_100a (UBLKCP in the SASS), NOT yet run on a GPU.
//
// Geometry.  A staged row holds LWA = roundup4(PAD + TW + 2 HALO) consecutive cells of one field row, starting at
// column j0 - HALO - PAD (HALO = R + G; PAD makes the start a multiple of 4 columns), so that with W % 4 == 0 every
// global source address is 32-byte (float64 chem) / 16-byte (int32 claims) aligned, as cp.async.bulk requires.
// Periodic wrap in the column direction splits a row into at most two copies (W >= LWA); rows wrap for free (every
// staged row is its own copy).  Per stage: chem [LH][LWA] f64 + claims [LH][LWA] i32; shared: s_v [OH][LWA].
// R = 2, GRAD: 2 x (21.9 + 10.9) + 19.6 KB = 85 KB per CTA, two CTAs per SM.
#pragma once
#include "die_async.cuh"
#include "die_field_kernels.cuh"

namespace die {

template <int R, int TH, int TW, bool GRAD>
struct BulkGeom {
    static constexpr int G = GRAD ? 1 : 0;
    static constexpr int HALO = R + G;
    static constexpr int PAD = (4 - HALO % 4) % 4;
    static constexpr int OH = TH + 2 * G, OW = TW + 2 * G;
    static constexpr int LW = OW + 2 * R;                          // columns the blur reads
    static constexpr int LH = OH + 2 * R;
    static constexpr int LWA = (PAD + LW + 3) / 4 * 4;             // staged row length
    static constexpr uint32_t kStageBytes = (uint32_t)LH * LWA * (8 + 4);
    static constexpr size_t kSmemBytes = (size_t)2 * LH * LWA * 8 + (size_t)2 * LH * LWA * 4 + (size_t)OH * LWA * 8 + 2 * 8 + 128;
};

template <int R, int TH, int TW, int NT, bool GRAD, bool PLAIN>
__global__ void __launch_bounds__(NT, 2)
field_step_bulk_kernel(const FieldArgs a, int total_tiles) {
    using GEO = BulkGeom<R, TH, TW, GRAD>;
    constexpr int HALO = GEO::HALO, PAD = GEO::PAD, OH = GEO::OH, OW = GEO::OW, LW = GEO::LW, LH = GEO::LH,
                  LWA = GEO::LWA;
    extern __shared__ double smem[];
    // carve the dynamic shared memory at a 128-byte boundary (bulk copies need 16)
    char* base = (char*)(((uintptr_t)smem + 127) & ~(uintptr_t)127);
    double* s_chem = (double*)base;                                  // [2][LH][LWA]
    double* s_v = s_chem + 2 * LH * LWA;                             // [OH][LWA]
    int32_t* s_claim = (int32_t*)(s_v + OH * LWA);                   // [2][LH][LWA]
    mbar_t* bars = (mbar_t*)(s_claim + 2 * LH * LWA);                // [2]

    const int H = a.H, W = a.W;
    const int64_t C = (int64_t)H * W;
    const int tiles = a.tiles_i * a.tiles_j;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // warp 0: announce the bytes of one stage, then every lane issues the copies of its rows
    auto issue = [&](int tile, int stage) {
        const int64_t b = tile / tiles;
        const int t = tile - (int)b * tiles;
        const int ti = t / a.tiles_j, tj = t - ti * a.tiles_j;
        const int i0 = ti * TH, j0 = tj * TW;
        const double* chem_in = a.medium_in + (b * 3 + 2) * C;
        const int32_t* win = a.winner + b * C;
        double* dc = s_chem + stage * LH * LWA;
        int32_t* dw = s_claim + stage * LH * LWA;
        if (lane == 0) mbar_arrive_expect_tx(&bars[stage], GEO::kStageBytes);
        __syncwarp();
        const int c0 = wrap_index(j0 - HALO - PAD, W);               // first staged column (a multiple of 4)
        const int n1 = min(LWA, W - c0);                             // cells before the row wraps (a multiple of 4)
        for (int r = lane; r < LH; r += 32) {
            const int64_t row = (int64_t)wrap_index(i0 - HALO + r, H) * W;
            bulk_g2s(dc + r * LWA, chem_in + row + c0, (uint32_t)n1 * 8, &bars[stage]);
            bulk_g2s(dw + r * LWA, win + row + c0, (uint32_t)n1 * 4, &bars[stage]);
            if (n1 < LWA) {
                bulk_g2s(dc + r * LWA + n1, chem_in + row, (uint32_t)(LWA - n1) * 8, &bars[stage]);
                bulk_g2s(dw + r * LWA + n1, win + row, (uint32_t)(LWA - n1) * 4, &bars[stage]);
            }
        }
    };

    uint32_t parity0 = 0, parity1 = 0;                               // phase parity of each stage's barrier
    int tile = blockIdx.x;
    if (warp == 0 && tile < total_tiles) issue(tile, 0);

    for (int k = 0; tile < total_tiles; ++k, tile += gridDim.x) {
        const int cur = k & 1;
        const int next = tile + gridDim.x;
        if (warp == 0 && next < total_tiles) {
            // stage cur^1 was last read (and written: the deposit fix-up, the blurred ring tile) before the
            // proxy fence + barrier that ended the previous iteration
            issue(next, cur ^ 1);
        }
        if (cur == 0) { mbar_wait(&bars[0], parity0); parity0 ^= 1; }
        else          { mbar_wait(&bars[1], parity1); parity1 ^= 1; }

        const int64_t b = tile / tiles;
        const int t = tile - (int)b * tiles;
        const int ti = t / a.tiles_j, tj = t - ti * a.tiles_j;
        const int i0 = ti * TH, j0 = tj * TW;
        double* s_in = s_chem + cur * LH * LWA;
        const int32_t* s_w = s_claim + cur * LH * LWA;
        double* s_out = s_in;                                        // [OH][OW] blurred * keep (GRAD; aliases the stage)
        const double* dep = a.action + (b * 3 + 2) * a.M;

        // ---- deposit of the winning slot, in shared memory (core/env.py:211) ---------------------------
        for (int idx = threadIdx.x; idx < LH * LWA; idx += NT) {
            const int w = s_w[idx];
            if (w >= 0) s_in[idx] = s_in[idx] + dep[w];
        }
        __syncthreads();

        // ---- axis-0 pass over columns [PAD, PAD + LW) ----------------------------------------------------
        for (int idx = threadIdx.x; idx < OH * LW; idx += NT) {
            const int r = idx / LW, c = idx - r * LW + PAD;
            const double* p = s_in + (r + R) * LWA + c;
            double acc = p[0] * a.bw.w[R];
#pragma unroll
            for (int q = R; q >= 1; --q) acc += (p[-q * LWA] + p[q * LWA]) * a.bw.w[R - q];
            s_v[r * LWA + c] = acc;
        }
        __syncthreads();

        const double* food_in = a.medium_in + (b * 3 + 1) * C;
        double* occ_out = a.medium_out + b * 3 * C;
        double* food_out = occ_out + C;
        double* chem_out = occ_out + 2 * C;
        double* cons = a.consumed + b * C;

        if constexpr (!GRAD) {
            for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
                const int r = idx / TW, c = idx - r * TW;
                const int gi = i0 + r, gj = j0 + c;
                if (gi < H && gj < W) {
                    const double* p = s_v + r * LWA + PAD + c + R;
                    double acc = p[0] * a.bw.w[R];
#pragma unroll
                    for (int q = R; q >= 1; --q) acc += (p[-q] + p[q]) * a.bw.w[R - q];
                    const int g = gi * W + gj;
                    chem_out[g] = acc * a.keep;
                    const double occ = (s_w[(r + HALO) * LWA + PAD + HALO + c] >= 0) ? 1.0 : 0.0;
                    const double f = food_in[g];
                    const double cf = (a.rate_feed * f) * occ;      // consumed_field, core/env.py:224
                    food_out[g] = next_food<PLAIN>(a, f, cf, gi, gj, g);
                    occ_out[g] = occ;
                    cons[g] = cf;
                }
            }
        } else {
            // ---- axis-1 pass over the ring tile into shared memory ---------------------------------------
            for (int idx = threadIdx.x; idx < OH * OW; idx += NT) {
                const int r = idx / OW, c = idx - r * OW;
                const double* p = s_v + r * LWA + PAD + c + R;
                double acc = p[0] * a.bw.w[R];
#pragma unroll
                for (int q = R; q >= 1; --q) acc += (p[-q] + p[q]) * a.bw.w[R - q];
                s_out[idx] = acc * a.keep;
            }
            __syncthreads();

            double2* grad = a.grad != nullptr ? a.grad + b * C : nullptr;
            float2* grad32 = a.grad32 != nullptr ? a.grad32 + b * C : nullptr;
            for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
                const int r = idx / TW, c = idx - r * TW;
                const int gi = i0 + r, gj = j0 + c;
                if (gi < H && gj < W) {
                    const double* q = s_out + (r + 1) * OW + (c + 1);
                    const int g = gi * W + gj;
                    chem_out[g] = q[0];
                    // np.gradient: (f[i+1] - f[i-1]) / 2 inside, f[1] - f[0] / f[n-1] - f[n-2] at the edges
                    const int um = (gi > 0) ? -OW : 0, up = (gi < H - 1) ? OW : 0;
                    const int lm = (gj > 0) ? -1 : 0, lp = (gj < W - 1) ? 1 : 0;
                    double gx = q[up] - q[um];
                    double gy = q[lp] - q[lm];
                    if (up - um == 2 * OW) gx *= 0.5;
                    if (lp - lm == 2) gy *= 0.5;
                    if (grad32 != nullptr) grad32[g] = make_float2((float)gx, (float)gy);
                    else grad[g] = make_double2(gx, gy);

                    const double occ = (s_w[(r + HALO) * LWA + PAD + HALO + c] >= 0) ? 1.0 : 0.0;
                    const double f = food_in[g];
                    const double cf = (a.rate_feed * f) * occ;      // consumed_field, core/env.py:224
                    food_out[g] = next_food<PLAIN>(a, f, cf, gi, gj, g);
                    occ_out[g] = occ;
                    cons[g] = cf;
                }
            }
        }
        proxy_fence_async();      // this thread's writes to the stage (fix-up, ring tile) before the next bulk copy into it
        __syncthreads();          // stage `cur` and s_v are free again
    }
}

}  // namespace die
