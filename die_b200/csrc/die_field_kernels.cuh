// die_field_kernels.cuh -- the dense half of Env.step: one pass over every cell that fuses
//   occupancy layout      medium['agents'] = 0; [cells] = 1        core/env.py:214-215
//   food consumption      food -= rate_feed * food * (occ > 0)     core/env.py:224-228
//   food flow             identity                                 core/env.py:147-150
//   diffusion * decay     gaussian(chem, sigma, 'wrap') * (1-d)    core/env.py:136-145
//   claim-table reset     (winner = -1 for the next step)
// reading medium_in (+ the claim table) and writing medium_out: 6 doubles of algorithmic
// traffic per cell (chem R+W, food R+W, occupancy W, + 2x4 B of claim table).
//
// The blur is scipy.ndimage's separable correlate1d in its exact operation order (axis 0
// then axis 1, each  x0*w0 + (x[-r]+x[+r])*w[-r] + ... + (x[-1]+x[+1])*w[-1], periodic), so
// the result is bit-identical to skimage.filters.gaussian on the same input.
#pragma once
#include "die_device.cuh"
#include "../../include/die_b200.h"

namespace die {

struct BlurWeights {
    double w[2 * DIE_MAX_RADIUS + 1];
};

struct FieldArgs {
    const double* medium_in;
    double* medium_out;
    int32_t* winner;
    int H, W;
    int tiles_i, tiles_j;
    double rate_feed;
    double keep;             // 1. - rate_decay_chem
    int food_infinite;
    BlurWeights bw;          // centre at [R]
};

__device__ __forceinline__ int wrap_index(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}

// Tile TH x TW outputs per CTA; (TH+2R) x (TW+2R) chem halo tile staged in shared memory,
// vertical (axis 0) pass into a second shared tile, horizontal (axis 1) pass to global.
template <int R, int TH, int TW, int NT>
__global__ void __launch_bounds__(NT)
field_step_kernel(const FieldArgs a) {
    constexpr int LW = TW + 2 * R;           // staged row length
    constexpr int LH = TH + 2 * R;
    extern __shared__ double smem[];
    double* s_in = smem;                     // [LH][LW]
    double* s_v = smem + LH * LW;            // [TH][LW]

    const int H = a.H, W = a.W;
    const int64_t C = (int64_t)H * W;
    const int tiles = a.tiles_i * a.tiles_j;
    const int64_t b = blockIdx.x / tiles;
    const int t = blockIdx.x - (int)b * tiles;
    const int ti = t / a.tiles_j, tj = t - ti * a.tiles_j;
    const int i0 = ti * TH, j0 = tj * TW;

    const double* food_in = a.medium_in + (b * 3 + 1) * C;
    const double* chem_in = a.medium_in + (b * 3 + 2) * C;
    double* occ_out = a.medium_out + (b * 3 + 0) * C;
    double* food_out = a.medium_out + (b * 3 + 1) * C;
    double* chem_out = a.medium_out + (b * 3 + 2) * C;
    int32_t* win = a.winner + b * C;

    // ---- stage the periodic halo tile ---------------------------------------------------
    for (int idx = threadIdx.x; idx < LH * LW; idx += NT) {
        const int r = idx / LW, c = idx - r * LW;
        const int gi = wrap_index(i0 - R + r, H);
        const int gj = wrap_index(j0 - R + c, W);
        s_in[idx] = chem_in[(int64_t)gi * W + gj];
    }
    __syncthreads();

    // ---- axis-0 pass --------------------------------------------------------------------
    for (int idx = threadIdx.x; idx < TH * LW; idx += NT) {
        const int r = idx / LW, c = idx - r * LW;
        const double* p = s_in + (r + R) * LW + c;
        double acc = p[0] * a.bw.w[R];
#pragma unroll
        for (int k = R; k >= 1; --k) acc += (p[-k * LW] + p[k * LW]) * a.bw.w[R - k];
        s_v[idx] = acc;
    }
    __syncthreads();

    // ---- axis-1 pass + elementwise channels ------------------------------------------------
    for (int idx = threadIdx.x; idx < TH * TW; idx += NT) {
        const int r = idx / TW, c = idx - r * TW;
        const int gi = i0 + r, gj = j0 + c;
        if (gi < H && gj < W) {
            const double* p = s_v + r * LW + c + R;
            double acc = p[0] * a.bw.w[R];
#pragma unroll
            for (int k = R; k >= 1; --k) acc += (p[-k] + p[k]) * a.bw.w[R - k];
            const int64_t g = (int64_t)gi * W + gj;
            chem_out[g] = acc * a.keep;

            const int w = win[g];
            const double occ = (w >= 0) ? 1.0 : 0.0;
            const double f = food_in[g];
            const double cf = (a.rate_feed * f) * occ;
            food_out[g] = a.food_infinite ? f : f - cf;
            occ_out[g] = occ;
            if (w >= 0) win[g] = -1;
        }
    }
}

// No diffusion (blur_radius == 0): gaussian with radius 0 is the identity (w = [1]).
template <int NT>
__global__ void __launch_bounds__(NT)
field_step_noblur_kernel(const FieldArgs a, int64_t total) {
    const int64_t C = (int64_t)a.H * a.W;
    for (int64_t gid = (int64_t)blockIdx.x * NT + threadIdx.x; gid < total; gid += (int64_t)gridDim.x * NT) {
        const int64_t b = gid / C, g = gid - b * C;
        const double* min = a.medium_in + b * 3 * C;
        double* mout = a.medium_out + b * 3 * C;
        const int w = a.winner[gid];
        const double occ = (w >= 0) ? 1.0 : 0.0;
        const double f = min[C + g];
        mout[g] = occ;
        mout[C + g] = a.food_infinite ? f : f - (a.rate_feed * f) * occ;
        mout[2 * C + g] = (min[2 * C + g] * a.bw.w[0]) * a.bw.w[0] * a.keep;
        if (w >= 0) a.winner[gid] = -1;
    }
}

}  // namespace die
