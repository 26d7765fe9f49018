// die_conv_kernels.cuh -- NeuralAutomataAgent.forward (core/agent/evo.py:117-209): the agent's perception model is a
// stack of small circular convolutions over the medium (ConvolutionModel, core/agent/evo.py:45-118: Conv2d, padding
// 'same', padding_mode 'circular', no bias, obs channels -> obs channels -> ... -> action channels, one Tanh at the end),
// evaluated in float32, and every slot's action is the model output at the agent's cell times (scale, scale, deposit)
// (tensor_by_agents with only_alive=False, core/utils.py:56-65; _rescale, core/agent/evo.py:183-186).
//
// At most 4 channels in and out and kernels up to 7 x 7: nothing here is a dense contraction worth a tensor core -- a
// layer is a shared-memory tiled stencil (81 .. 441 float32 FMAs per cell), bound by its FMA rate for 5 x 5 and larger,
// by HBM for 3 x 3.  One launch per layer (the last one applies tanh), one launch that gathers per slot.
//   conv_layer_kernel    TILE x TILE outputs per CTA, the input halo tile of every input channel in shared memory
//                        (periodic indices), weights in shared memory; input float64 / float32 medium channels (first
//                        layer, cast to float32 as th.as_tensor(..., dtype=float32) does) or the previous layer's float32
//   conv_gather_kernel   action[c][slot] = (double)(out[c][cell(slot)] * coef[c]); cell from the env's cache or from
//                        nearest_cell of the slot's position (clamped, as AgentIndexer does)
#pragma once
#include "die_device.cuh"

namespace die {

constexpr int kConvMaxCh = 4;
constexpr int kConvMaxK = 7;
constexpr int kConvTile = 32;

struct ConvLayerArgs {
    const void* in;        // [B][cin_total][H][W], element type TIN; channel c of the layer = in channel (c + in_ch0)
    float* out;            // [B][cout][H][W]
    const float* weight;   // [cout][cin][k][k] (torch Conv2d layout)
    int64_t weight_env_stride;  // floats between the weight sets of consecutive environments (0: one model for the batch;
                                // a POPULATION of models, one per environment, otherwise)
    int H, W, cin, cout, k, cin_total, in_ch0;
    int tiles_i, tiles_j;
    int apply_tanh;
};

template <typename TIN>
__global__ void __launch_bounds__(256)
conv_layer_kernel(const ConvLayerArgs a) {
    extern __shared__ float csmem[];
    const int r = a.k / 2;
    const int LW = kConvTile + 2 * r, LH = kConvTile + 2 * r;
    float* s_in = csmem;                                   // [cin][LH][LW]
    float* s_w = csmem + a.cin * LH * LW;                  // [cout][cin][k][k]
    const int H = a.H, W = a.W;
    const int64_t C = (int64_t)H * W;
    const int tiles = a.tiles_i * a.tiles_j;
    const int64_t b = blockIdx.x / (unsigned)tiles;
    const int t = blockIdx.x - (int)b * tiles;
    const int ti = t / a.tiles_j, tj = t - ti * a.tiles_j;
    const int i0 = ti * kConvTile, j0 = tj * kConvTile;
    const TIN* in = (const TIN*)a.in + (b * a.cin_total + a.in_ch0) * C;

    const float* wsrc = a.weight + b * a.weight_env_stride;
    for (int i = threadIdx.x; i < a.cout * a.cin * a.k * a.k; i += 256) s_w[i] = wsrc[i];
    for (int idx = threadIdx.x; idx < a.cin * LH * LW; idx += 256) {
        const int c = idx / (LH * LW), rem = idx - c * (LH * LW);
        const int rr = rem / LW, cc = rem - rr * LW;
        int gi = (i0 - r + rr) % H, gj = (j0 - r + cc) % W;        // padding_mode = 'circular'
        if (gi < 0) gi += H;
        if (gj < 0) gj += W;
        s_in[idx] = (float)in[c * C + (int64_t)gi * W + gj];
    }
    __syncthreads();

    float* out = a.out + b * a.cout * C;
    for (int idx = threadIdx.x; idx < kConvTile * kConvTile; idx += 256) {
        const int rr = idx / kConvTile, cc = idx - rr * kConvTile;
        const int gi = i0 + rr, gj = j0 + cc;
        if (gi >= H || gj >= W) continue;
        float acc[kConvMaxCh] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < a.cin; ++c) {
            const float* p = s_in + c * LH * LW + rr * LW + cc;
            for (int u = 0; u < a.k; ++u) {
                for (int v = 0; v < a.k; ++v) {
                    const float x = p[u * LW + v];                  // cross-correlation, as torch's conv2d
#pragma unroll
                    for (int o = 0; o < kConvMaxCh; ++o)
                        if (o < a.cout) acc[o] = fmaf(x, s_w[((o * a.cin + c) * a.k + u) * a.k + v], acc[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < kConvMaxCh; ++o)
            if (o < a.cout) out[o * C + (int64_t)gi * W + gj] = a.apply_tanh ? tanhf(acc[o]) : acc[o];
    }
}

// per-slot gather of the model output + rescale (core/agent/evo.py:160-166, 183-186)
__global__ void __launch_bounds__(256)
conv_gather_kernel(const float* __restrict__ sense, const double* __restrict__ agents, const int32_t* __restrict__ cells,
                   double* __restrict__ action, const Axis ax, const Axis ay, int64_t M, int cout, int B,
                   float coef0, float coef1, float coef2) {
    const int64_t C = (int64_t)ax.n * ay.n;
    const int64_t total = (int64_t)B * M;
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < total; g += (int64_t)gridDim.x * 256) {
        const int64_t b = g / M, i = g - b * M;
        int cell;
        if (cells != nullptr) {
            cell = cells[g];
        } else {
            const double* ag = agents + b * 4 * M;
            cell = nearest_cell(ag[i], ax) * ay.n + nearest_cell(ag[M + i], ay);
        }
        const float* s = sense + b * cout * C + cell;
        double* act = action + b * 3 * M + i;
        // per_agent_output *= action_coefs: float32 products, stored into the float64 action array
        act[0] = (double)(s[0] * coef0);
        act[M] = (double)((cout > 1 ? s[C] : 0.f) * coef1);
        act[2 * M] = (double)((cout > 2 ? s[2 * C] : 0.f) * coef2);
    }
}

}  // namespace die
