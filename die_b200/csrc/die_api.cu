// die_api.cu -- the C ABI of libdie_sm100a.so (declared in include/die_b200.h): argument
// checking, workspace ownership, launch geometry.  No torch, no C++ types in signatures.
#include <cstdio>
#include <cstring>
#include <new>

#include "die_agent_kernels.cuh"
#include "die_field_kernels.cuh"
#include "die_env_fused.cuh"
#include "die_conv_kernels.cuh"

using namespace die;

// ------------------------------------------------------------------------------------------
// error reporting
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define DIE_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t e_ = (expr);                                                    \
        if (e_ != cudaSuccess) return fail(DIE_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

#define DIE_REQUIRE(cond)                                                           \
    do {                                                                            \
        if (!(cond)) return fail(DIE_E_INVALID, "invalid argument: %s%s", #cond);   \
    } while (0)

struct die_env {
    int32_t H, W, B;
    int64_t M;
    die_dynamics_t dyn;
    int32_t* winner;       // [B][H*W]  claim table, -1 = empty
    int32_t* cells2[2];    // [B][M] x 2 linear cell of every slot after the move; [cur] is the valid one, the
    int cur;               //            other receives the cells of a speculative move (die_env_forward_gradient)
    uint32_t* alive_bits;  // [B][Mw]   (alive > 0) per slot, one bit each (die_env_refresh_alive)
    int64_t Mw;
    int alive_valid;
    double* burned;        // [B][M] lazily: linear_action_cost per slot of the action the last forward wrote (cost hint)
    int cost_pending;      // that array is valid for the action of the forward that just ran (DIE_FWD_WRITE_COST)
    int pending_move;      // 1: a speculative move (cells2[1-cur] + claims) waits for a step with DIE_STEP_ADOPT_MOVE;
                           // 2: a COMMITTED one (DIE_FWD_COMMIT_MOVE: the positions are already stored)
    // op_food_flow = WaveSequence flow operator (die_env_set_food_flow); borrowed device tables, host copy of ts
    const double* flow_rwave;
    const double* flow_col;    // [T][W]
    const double* flow_row;    // [T][H]
    double* flow_ts;           // host [T]
    const double* flow_frames; // [T][H*W] tabulated sequence (die_env_set_food_frames); borrowed device memory
    int64_t flow_T, flow_k;
    double flow_scale, flow_keep;
    double* consumed;      // [B][H*W]  consumed_field = rate_feed * food * occ of the current step
    // "pair mode" (large fields, see pair_mode_for): the per-cell scratch is {consumed_field, new env_food}; the feed kernel's
    // one gather per slot brings both and hands the food under the agent to the next forward pass per SLOT
    double2* cell_pairs;   // [B][H*W]  (lazy)
    double* food_here;     // [B][M]    env_food under every slot after the last step (lazy)
    int food_here_valid;   // the last step filled food_here for the cells in cells2[cur]
    double2* grad;         // [B][H*W]  np.gradient of the current chem1 (lazy; see die_env_publish_gradient)
    float2* grad32;        // [B][H*W]  the same rounded to float32 (lazy; tuning "grad_f32")
    int grad_kind;         // which of the two the LAST field pass wrote: 0 none, 1 grad, 2 grad32
    int grad_f32_ok;       // the last agent that acted through this env only needs the gradient for the guard-banded
                           // turn decision (normalised Physarum, die_turn_plan enabled): float32 is enough for it
    int publish_grad;
    double* part_gain;     // [B][nblk]
    int32_t* part_alive;   // [B][nblk]
    int nblk;
    double* action_stage;  // [B][3][M] device staging for die_env_step_host (lazy)
    cudaStream_t host_streams[2];      // die_env_step_host: chunks of the batch alternate between two streams (lazy)
    cudaEvent_t host_events[3];
    double* reward_dev;    // [B] (host path)
    int64_t* alive_dev;    // [B] (host path)
    int num_sms;
    int field_f32;                     // the field arrays (medium A/B, consumed_field) hold float32 (die_env_set_field_dtype)
    int profiling;                     // record events between the step's kernels
    int prof_steps;                    // steps recorded so far
    cudaEvent_t* prof_events;          // [DIE_MAX_PROFILED_STEPS][DIE_NUM_STEP_KERNELS + 1]
};

static inline void prof_mark(die_env* e, int k, cudaStream_t st) {
    if (e->profiling && e->prof_steps < DIE_MAX_PROFILED_STEPS)
        cudaEventRecord(e->prof_events[(size_t)e->prof_steps * (DIE_NUM_STEP_KERNELS + 1) + k], st);
}

extern "C" const char* die_version(void) { return "die_b200 0.1 (sm_100a)"; }

// launch counters (diagnostics: tests assert that the variant they mean to exercise is the one that ran)
static int64_t g_count_fwd_food_here = 0, g_count_field_tile = 0, g_count_field_vec = 0, g_count_step_fused = 0, g_count_fwd_lean = 0, g_count_fwd_lean_f32 = 0, g_count_fwd_general = 0,
               g_count_step_committed = 0, g_count_fwd_lean_move = 0, g_count_feed_cost = 0;

extern "C" int64_t die_get_counter(const char* key) {
    if (key == nullptr) return -1;
    if (strcmp(key, "field_tile") == 0) return g_count_field_tile;
    if (strcmp(key, "field_vec") == 0) return g_count_field_vec;
    if (strcmp(key, "step_fused") == 0) return g_count_step_fused;
    if (strcmp(key, "forward_lean") == 0) return g_count_fwd_lean;
    if (strcmp(key, "forward_lean_f32") == 0) return g_count_fwd_lean_f32;
    if (strcmp(key, "forward_general") == 0) return g_count_fwd_general;
    if (strcmp(key, "forward_food_here") == 0) return g_count_fwd_food_here;
    if (strcmp(key, "forward_lean_move") == 0) return g_count_fwd_lean_move;
    if (strcmp(key, "step_committed") == 0) return g_count_step_committed;
    if (strcmp(key, "feed_cost") == 0) return g_count_feed_cost;
    return -1;
}
extern "C" const char* die_last_error(void) { return g_err; }

static int check_dynamics(const die_dynamics_t* d) {
    DIE_REQUIRE(d != nullptr);
    DIE_REQUIRE(d->blur_radius >= 0 && d->blur_radius <= DIE_MAX_RADIUS);
    DIE_REQUIRE(d->boundary >= DIE_BOUNDARY_WRAP && d->boundary <= DIE_BOUNDARY_NONE);
    DIE_REQUIRE(d->diffuse_mode >= DIE_DIFFUSE_WRAP && d->diffuse_mode <= DIE_DIFFUSE_CONSTANT);
    return DIE_OK;
}

static inline int grid_for(int64_t total, int threads, int num_sms) {
    // enough CTAs for full occupancy, grid-stride beyond that: a multiple of the SM count
    const int64_t want = (total + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms * 32;
    int64_t g = want < cap ? want : cap;
    return (int)(g < 1 ? 1 : g);
}

extern "C" int die_env_create(int32_t H, int32_t W, int64_t M, int32_t B,
                              const die_dynamics_t* dyn, die_env_t** out) {
    DIE_REQUIRE(out != nullptr);
    *out = nullptr;
    DIE_REQUIRE(H >= 2 && W >= 2);
    DIE_REQUIRE(M >= 1 && M <= 0x7fffffffLL);
    DIE_REQUIRE(B >= 1);
    DIE_REQUIRE((int64_t)H * W <= 0x7fffffffLL);
    if (int rc = check_dynamics(dyn)) return rc;

    die_env* e = new (std::nothrow) die_env();
    if (e == nullptr) return fail(DIE_E_NOMEM, "out of host memory");
    memset(e, 0, sizeof(*e));
    e->H = H; e->W = W; e->B = B; e->M = M;
    e->dyn = *dyn;
    e->nblk = (int)((M + (int64_t)kAgentThreads * kFeedItems - 1) / ((int64_t)kAgentThreads * kFeedItems));
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t C = (size_t)H * W;
    if (err == cudaSuccess) err = cudaMalloc(&e->winner, sizeof(int32_t) * C * B);
    e->Mw = (M + 31) / 32;
    if (err == cudaSuccess) err = cudaMalloc(&e->cells2[0], sizeof(int32_t) * (size_t)M * B);
    if (err == cudaSuccess) err = cudaMalloc(&e->cells2[1], sizeof(int32_t) * (size_t)M * B);
    if (err == cudaSuccess) err = cudaMalloc(&e->alive_bits, sizeof(uint32_t) * (size_t)e->Mw * B);
    if (err == cudaSuccess) err = cudaMalloc(&e->consumed, sizeof(double) * C * B);
    if (err == cudaSuccess) err = cudaMalloc(&e->part_gain, sizeof(double) * (size_t)e->nblk * B);
    if (err == cudaSuccess) err = cudaMalloc(&e->part_alive, sizeof(int32_t) * (size_t)e->nblk * B);
    if (err == cudaSuccess) err = cudaMalloc(&e->reward_dev, sizeof(double) * B);
    if (err == cudaSuccess) err = cudaMalloc(&e->alive_dev, sizeof(int64_t) * B);
    if (err == cudaSuccess) err = cudaMemset(e->winner, 0xFF, sizeof(int32_t) * C * B);
    if (err == cudaSuccess) err = cudaMemset(e->cells2[0], 0, sizeof(int32_t) * (size_t)M * B);
    if (err == cudaSuccess) err = cudaMemset(e->cells2[1], 0, sizeof(int32_t) * (size_t)M * B);
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        die_env_destroy(e);
        return fail(DIE_E_CUDA, "die_env_create: %s", cudaGetErrorString(err));
    }
    *out = e;
    return DIE_OK;
}

extern "C" int die_env_destroy(die_env_t* e) {
    if (e == nullptr) return DIE_OK;
    cudaFree(e->winner);
    cudaFree(e->cells2[0]);
    cudaFree(e->cells2[1]);
    cudaFree(e->alive_bits);
    delete[] e->flow_ts;
    cudaFree(e->consumed);
    cudaFree(e->cell_pairs);
    cudaFree(e->food_here);
    cudaFree(e->burned);
    cudaFree(e->grad);
    cudaFree(e->grad32);
    cudaFree(e->part_gain);
    cudaFree(e->part_alive);
    cudaFree(e->action_stage);
    for (int k = 0; k < 2; ++k) if (e->host_streams[k]) cudaStreamDestroy(e->host_streams[k]);
    for (int k = 0; k < 3; ++k) if (e->host_events[k]) cudaEventDestroy(e->host_events[k]);
    cudaFree(e->reward_dev);
    cudaFree(e->alive_dev);
    if (e->prof_events != nullptr) {
        for (size_t k = 0; k < (size_t)DIE_MAX_PROFILED_STEPS * (DIE_NUM_STEP_KERNELS + 1); ++k)
            if (e->prof_events[k]) cudaEventDestroy(e->prof_events[k]);
        delete[] e->prof_events;
    }
    delete e;
    return DIE_OK;
}

extern "C" int die_env_set_field_dtype(die_env_t* e, int32_t dtype) {
    DIE_REQUIRE(e != nullptr && (dtype == DIE_FIELD_F64 || dtype == DIE_FIELD_F32));
    e->field_f32 = dtype == DIE_FIELD_F32;
    e->grad_kind = 0;
    return DIE_OK;
}

extern "C" int die_env_field_dtype(const die_env_t* e) { return (e && e->field_f32) ? DIE_FIELD_F32 : DIE_FIELD_F64; }

// element `elems` of a field array whose element type is the env's field dtype
static inline double* field_off(const die_env* e, const double* p, size_t elems) {
    return e->field_f32 ? (double*)((float*)p + elems) : (double*)p + elems;
}

extern "C" int die_env_set_dynamics(die_env_t* e, const die_dynamics_t* dyn) {
    DIE_REQUIRE(e != nullptr);
    if (int rc = check_dynamics(dyn)) return rc;
    e->dyn = *dyn;
    e->cost_pending = 0;              // (the cost weights may have changed)
    return DIE_OK;
}

extern "C" int die_env_set_food_flow(die_env_t* e, const double* rwave_dev, const double* col_dev, const double* row_dev,
                                     const double* ts_host, int64_t T, int64_t k0, double scale, double decay) {
    DIE_REQUIRE(e != nullptr);
    delete[] e->flow_ts;
    e->flow_ts = nullptr;
    e->flow_rwave = e->flow_col = e->flow_row = nullptr;
    e->flow_frames = nullptr;
    if (rwave_dev == nullptr) return DIE_OK;                // back to the identity flow
    DIE_REQUIRE(col_dev != nullptr && row_dev != nullptr && ts_host != nullptr && T >= 1 && k0 >= 0);
    e->flow_ts = new (std::nothrow) double[(size_t)T];
    if (e->flow_ts == nullptr) return fail(DIE_E_NOMEM, "out of host memory");
    memcpy(e->flow_ts, ts_host, sizeof(double) * (size_t)T);
    e->flow_rwave = rwave_dev;
    e->flow_col = col_dev;
    e->flow_row = row_dev;
    e->flow_T = T;
    e->flow_k = k0;
    e->flow_scale = scale;
    e->flow_keep = 1.0 - decay;
    return DIE_OK;
}

extern "C" int die_env_set_food_frames(die_env_t* e, const double* frames_dev, int64_t T, int64_t k0, double scale, double decay) {
    DIE_REQUIRE(e != nullptr);
    if (int rc = die_env_set_food_flow(e, nullptr, nullptr, nullptr, nullptr, 0, 0, 0.0, 0.0)) return rc;   // drop any flow
    if (frames_dev == nullptr) return DIE_OK;
    DIE_REQUIRE(T >= 1 && k0 >= 0);
    e->flow_frames = frames_dev;
    e->flow_T = T;
    e->flow_k = k0;
    e->flow_scale = scale;
    e->flow_keep = 1.0 - decay;
    return DIE_OK;
}

extern "C" const int32_t* die_env_cells(const die_env_t* e) { return e ? e->cells2[e->cur] : nullptr; }

static int g_grad_f32 = 1;         // the field pass may publish np.gradient(chem1) as float32 pairs (see die_env_gradient_kind)

// float32 pairs iff the switch is on AND the env's current consumer only thresholds the gradient
static inline bool publish_as_f32(const die_env* e) { return g_grad_f32 && e->grad_f32_ok; }

// the buffer the next field pass publishes into (allocated on first use)
static int ensure_gradient_buffer(die_env* e) {
    const size_t n = (size_t)e->H * e->W * e->B;
    if (publish_as_f32(e)) {
        if (e->grad32 == nullptr) DIE_CUDA(cudaMalloc(&e->grad32, sizeof(float2) * n));
    } else if (e->grad == nullptr) {
        DIE_CUDA(cudaMalloc(&e->grad, sizeof(double2) * n));
    }
    return DIE_OK;
}

extern "C" int die_env_publish_gradient(die_env_t* e, int32_t on) {
    DIE_REQUIRE(e != nullptr);
    if (!on) e->grad_kind = 0;
    e->publish_grad = on ? 1 : 0;        // (the buffer is allocated by the first field pass that publishes)
    return DIE_OK;
}

extern "C" const double* die_env_gradient(const die_env_t* e) {
    return (e && e->publish_grad && e->dyn.blur_radius > 0 && e->grad_kind != 2) ? (const double*)e->grad : nullptr;
}

extern "C" int die_env_gradient_kind(const die_env_t* e) {
    return (e && e->publish_grad && e->dyn.blur_radius > 0) ? e->grad_kind : 0;
}

extern "C" int die_env_set_profiling(die_env_t* e, int32_t on) {
    DIE_REQUIRE(e != nullptr);
    if (on && e->prof_events == nullptr) {
        const size_t n = (size_t)DIE_MAX_PROFILED_STEPS * (DIE_NUM_STEP_KERNELS + 1);
        e->prof_events = new (std::nothrow) cudaEvent_t[n]();
        if (e->prof_events == nullptr) return fail(DIE_E_NOMEM, "out of host memory");
        for (size_t k = 0; k < n; ++k) DIE_CUDA(cudaEventCreate(&e->prof_events[k]));
    }
    e->profiling = on ? 1 : 0;
    e->prof_steps = 0;
    return DIE_OK;
}

extern "C" int die_env_kernel_times(die_env_t* e, double* ms_out, int64_t* steps_out) {
    DIE_REQUIRE(e != nullptr && ms_out != nullptr && steps_out != nullptr);
    for (int k = 0; k < DIE_NUM_STEP_KERNELS; ++k) ms_out[k] = 0.0;
    *steps_out = e->prof_steps;
    for (int s = 0; s < e->prof_steps; ++s) {
        cudaEvent_t* ev = e->prof_events + (size_t)s * (DIE_NUM_STEP_KERNELS + 1);
        DIE_CUDA(cudaEventSynchronize(ev[DIE_NUM_STEP_KERNELS]));
        for (int k = 0; k < DIE_NUM_STEP_KERNELS; ++k) {
            float ms = 0.f;
            DIE_CUDA(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
            ms_out[k] += (double)ms;
        }
    }
    return DIE_OK;
}

// ------------------------------------------------------------------------------------------
// field pass launch
// ------------------------------------------------------------------------------------------
static int g_field_vec = 0;        // 1: the 128-bit field pass wherever it applies.  Measured on a B200 (round 2,
                                   // profiles/r02l_field_vec_ab.txt): half the memory instructions, 52 instead of 60 registers, and
                                   // NO faster (3.445 vs 3.409 ms batched, 0.230 vs 0.224 ms at 4096^2) -- the pass is bound by DRAM
                                   // (a 65 % write mix at 80 % of the copy bandwidth), not by its LSU instructions.  Opt-in.

static int g_field_tile = 0;       // threads per 32 x 64-cell tile.  0 = 512 (default: 4 cells and 6 staged values per thread, 32 registers,
                                   // 4 CTAs = 2048 threads per SM); 1 = 256 (the shape until the end of round 2: 56 registers, 1024 threads per
                                   // SM; 3.34 vs 3.05 ms batched, 0.2415 vs 0.2166 ms at 4096^2, profiles/r02zs_field_tile_ab.txt); 2 = 1024.
                                   // 1 / 2 exist for radius 2 only.  Larger tiles (32 x 128, 64 x 64 at 512 threads: two CTAs per SM by shared
                                   // memory) lose 7-9 %, a 16 x 64 tile of 256 threads gains less (halo 1.50x): profiles/r02zr_*.

template <int R, bool GRAD, int TH, int TW, int NT>
static cudaError_t launch_field_shape(const FieldArgs& fa, int B, cudaStream_t st, bool f32);

template <int R, bool GRAD>
static cudaError_t launch_field_g(const FieldArgs& fa, int B, cudaStream_t st, bool f32) {
    if constexpr (R == 2) {
        if (g_field_tile == 1 && !f32) return launch_field_shape<R, GRAD, 32, 64, 256>(fa, B, st, f32);
        if (g_field_tile == 2 && !f32) return launch_field_shape<R, GRAD, 32, 64, 1024>(fa, B, st, f32);
    }
    return launch_field_shape<R, GRAD, 32, 64, 512>(fa, B, st, f32);
}

template <int R, bool GRAD, int TH, int TW, int NT>
static cudaError_t launch_field_shape(const FieldArgs& fa, int B, cudaStream_t st, bool f32) {
    constexpr int G = GRAD ? 1 : 0;
    FieldArgs a = fa;
    a.tiles_i = (a.H + TH - 1) / TH;
    a.tiles_j = (a.W + TW - 1) / TW;
    const size_t smem = sizeof(double) * (size_t)((TH + 2 * G + 2 * R) * (TW + 2 * G + 2 * R) +
                                                  (TH + 2 * G) * (TW + 2 * G + 2 * R));
    const bool plain = a.diffuse_mode == DIE_DIFFUSE_WRAP && a.flow_rwave == nullptr && a.flow_frame == nullptr;
    if (f32) {                     // float32 field arrays: the scalar tile kernel on float (blur radius 1..4)
        if constexpr (R <= 4 && TH == 32 && TW == 64 && NT == 512) {
            auto kern = plain ? field_step_kernel<R, TH, TW, NT, GRAD, false, true, float>
                              : field_step_kernel<R, TH, TW, NT, GRAD, false, false, float>;
            cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return err;
            kern<<<(unsigned)((int64_t)a.tiles_i * a.tiles_j * B), NT, smem, st>>>(a);
            ++g_count_field_tile;
            return cudaGetLastError();
        } else {
            return cudaErrorInvalidValue;        // (float32 fields: blur radius 1..4; the Python layer refuses by name)
        }
    }
    if constexpr (TH == 32 && TW == 64 && NT == 512)
    if (plain && g_field_vec && a.cell_pairs == nullptr && (a.W & 1) == 0 && ((uintptr_t)a.medium_in & 15) == 0 && ((uintptr_t)a.medium_out & 15) == 0 &&
        ((uintptr_t)a.consumed & 15) == 0 && ((uintptr_t)a.winner & 7) == 0 && ((uintptr_t)a.grad32 & 15) == 0) {
        // the default dynamics on an even row length: the 128-bit version (same tile, same arithmetic, same results)
        constexpr int PADL = (R + G) & 1, SW = (PADL + TW + 2 * G + 2 * R + 1) & ~1;
        const size_t vsmem = sizeof(double) * (size_t)((TH + 2 * G + 2 * R) * SW + (TH + 2 * G) * (TW + 2 * G + 2 * R));
        auto vkern = field_step_vec_kernel<R, TH, TW, 256, GRAD>;          // (the opt-in keeps its 256 threads per tile)
        cudaError_t verr = cudaFuncSetAttribute(vkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem);
        if (verr != cudaSuccess) return verr;
        vkern<<<(unsigned)((int64_t)a.tiles_i * a.tiles_j * B), 256, vsmem, st>>>(a);
        ++g_count_field_vec;
        return cudaGetLastError();
    }
    auto kern = plain ? field_step_kernel<R, TH, TW, NT, GRAD, false, true> : field_step_kernel<R, TH, TW, NT, GRAD, false, false>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const int64_t grid = (int64_t)a.tiles_i * a.tiles_j * B;
    kern<<<(unsigned)grid, NT, smem, st>>>(a);
    ++g_count_field_tile;
    return cudaGetLastError();
}

static int g_field_prefetch = 1;   // tile kernel: prefetch the output tile's food lines to L2 while staging

template <int R>
static cudaError_t launch_field(const FieldArgs& fa, int B, int num_sms, cudaStream_t st, bool f32) {
    const bool want_grad = fa.grad != nullptr || fa.grad32 != nullptr;
    return want_grad ? launch_field_g<R, true>(fa, B, st, f32) : launch_field_g<R, false>(fa, B, st, f32);
}

static int g_step_impl = 0;        // 0 (default) = the three kernels move_claim / field_step / agent_feed;
                                   // 1 = the cluster-fused environment step wherever it applies (die_env_fused.cuh):
                                   //     bit-identical, 26 % less DRAM traffic, but 2.2x SLOWER on a B200 (19.0 vs 8.7 ms
                                   //     for 4096 x 256^2: latency / issue bound at two CTAs per SM) -- kept as an opt-in

extern "C" int die_set_step_impl(int32_t impl) {
    DIE_REQUIRE(impl == 0 || impl == 1);
    g_step_impl = impl;
    return DIE_OK;
}

// Pair mode: one field whose per-cell tables are far larger than L2 -- there every random 8-byte gather of a slot costs
// a DRAM sector plus change (77-93 B measured), so the step keeps {consumed_field, new food} in ONE 16-byte
// pair per cell and hands the food per slot to the next forward pass: three random accesses per slot and step become
// two.  Small fields (the batched workload) gather out of L2 and keep the 8-byte table.
static int g_cost_sqrt_near = 1;               // 0: the cost hint always evaluates sqrt() (A-B timing; same bits)
static int g_feed_min_blocks = 6;              // resident CTAs per SM asked of the COST feed kernel (0 = the compiler's choice, 48 registers)
static int g_cost_hint = 1;                    // 0: the feed kernel always re-reads dx, dy, deposit (A-B timing)
static int g_pair_mode = 1;                    // 0 never, 1 by size (pair_min_cells), 2 always (tests)
static int64_t g_pair_min_cells = 1 << 23;     // cells per environment from which it pays.  Measured on a B200, step time at
                                               // step 40 / step 3000: 2048^2 +6 % / +7 % (the tables still sit in L2),
                                               // 3072^2 -4 % / -6 %, 4096^2 +-0 / -11 %, 6144^2 -1 % / -15 %, 8192^2 -5 % / -16 %

static inline bool pair_mode_for(const die_env* e, bool speculative) {
    if (g_pair_mode == 0 || speculative || e->dyn.agents_die || e->field_f32 || e->dyn.blur_radius == 0) return false;
    return g_pair_mode == 2 || (int64_t)e->H * e->W >= g_pair_min_cells;
}

// Field pass over environments [b0, b0 + nb) of the batch; min / mout / action already point at environment b0.
static cudaError_t launch_field_any(die_env* e, int b0, int nb, const double* min, double* mout,
                                    const double* action, cudaStream_t st, bool pair = false) {
    const size_t C = (size_t)e->H * e->W;
    FieldArgs a;
    memset(&a, 0, sizeof(a));
    a.medium_in = min;
    a.medium_out = mout;
    a.winner = e->winner + b0 * C;
    a.action = action;
    a.consumed = field_off(e, e->consumed, (size_t)b0 * C);
    if (pair) a.cell_pairs = e->cell_pairs + (size_t)b0 * C;
    const bool f32 = e->field_f32 != 0;
    if (e->publish_grad && e->dyn.blur_radius > 0) {
        if (ensure_gradient_buffer(e) != DIE_OK) return cudaErrorMemoryAllocation;
        if (publish_as_f32(e)) a.grad32 = e->grad32 + b0 * C;
        else a.grad = e->grad + b0 * C;
        e->grad_kind = publish_as_f32(e) ? 2 : 1;
    } else {
        e->grad_kind = 0;
    }
    a.M = e->M;
    a.H = e->H;
    a.W = e->W;
    a.rate_feed = e->dyn.rate_feed;
    a.keep = 1.0 - e->dyn.rate_decay_chem;
    a.food_infinite = e->dyn.food_infinite;
    a.prefetch_food = g_field_prefetch;
    a.diffuse_mode = e->dyn.diffuse_mode;
    if (e->flow_rwave != nullptr) {
        const int64_t k = e->flow_k % e->flow_T;           // `for t in cycle(self._ts)`, core/data_init.py:40-42
        a.flow_rwave = e->flow_rwave;
        a.flow_col = e->flow_col + k * e->W;
        a.flow_row = e->flow_row + k * e->H;
        a.flow_t = e->flow_ts[k];
        a.flow_scale = e->flow_scale;
        a.flow_keep = e->flow_keep;
    } else if (e->flow_frames != nullptr) {
        a.flow_frame = e->flow_frames + (e->flow_k % e->flow_T) * (int64_t)C;
        a.flow_scale = e->flow_scale;
        a.flow_keep = e->flow_keep;
    }
    for (int k = 0; k < 2 * DIE_MAX_RADIUS + 1; ++k) a.bw.w[k] = e->dyn.blur_w[k];
    cudaError_t err = cudaErrorInvalidValue;
    switch (e->dyn.blur_radius) {
        case 0: {
            const int64_t total = (int64_t)e->H * e->W * nb;
            if (f32) field_step_noblur_kernel<256, float><<<grid_for(total, 256, e->num_sms), 256, 0, st>>>(a, total);
            else field_step_noblur_kernel<256><<<grid_for(total, 256, e->num_sms), 256, 0, st>>>(a, total);
            err = cudaGetLastError();
            break;
        }
        case 1: err = launch_field<1>(a, nb, e->num_sms, st, f32); break;
        case 2: err = launch_field<2>(a, nb, e->num_sms, st, f32); break;
        case 3: err = launch_field<3>(a, nb, e->num_sms, st, f32); break;
        case 4: err = launch_field<4>(a, nb, e->num_sms, st, f32); break;
        case 5: err = launch_field<5>(a, nb, e->num_sms, st, f32); break;
        case 6: err = launch_field<6>(a, nb, e->num_sms, st, f32); break;
        case 7: err = launch_field<7>(a, nb, e->num_sms, st, f32); break;
        case 8: err = launch_field<8>(a, nb, e->num_sms, st, f32); break;
    }
    return err;
}

// ------------------------------------------------------------------------------------------
// the cluster-fused environment step (die_env_fused.cuh)
// ------------------------------------------------------------------------------------------
constexpr size_t kFusedSmemMax = 227u << 10;

typedef void (*fused_kernel_t)(const FusedArgs);

static int g_fused_threads = 512;  // CTA size of the fused step (256 threads / 128 registers measured 45 % slower: removed)

template <int R, int NT>
static fused_kernel_t pick_fused_nt(bool grad, bool plain) {
    if (grad) return plain ? env_step_fused_kernel<R, NT, true, true> : env_step_fused_kernel<R, NT, true, false>;
    return plain ? env_step_fused_kernel<R, NT, false, true> : env_step_fused_kernel<R, NT, false, false>;
}

template <int R>
static fused_kernel_t pick_fused(bool grad, bool plain, int nt) {
    (void)nt;
    return pick_fused_nt<R, 512>(grad, plain);
}

// Per kernel instantiation and cluster size: has the shared-memory limit been raised, can the shape be scheduled?
struct FusedShape { fused_kernel_t kern; int S; size_t smem; int ok; int nt; };
static FusedShape g_fused_shapes[64];
static int g_fused_nshapes = 0;

static bool fused_shape_ok(fused_kernel_t kern, int S, size_t smem, int nt) {
    size_t kernel_max = 0;          // the dynamic shared-memory limit this kernel has been given so far: only ever raised
    for (int k = 0; k < g_fused_nshapes; ++k) {
        if (g_fused_shapes[k].kern != kern) continue;
        if (g_fused_shapes[k].smem > kernel_max) kernel_max = g_fused_shapes[k].smem;
        if (g_fused_shapes[k].S == S && g_fused_shapes[k].smem >= smem) return g_fused_shapes[k].ok != 0;
    }
    int ok = 1;
#if !defined(DIE_HOSTSIM)
    if (smem > kernel_max &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) ok = 0;
    if (ok && S > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ok = 0;
    if (ok && max_active_clusters(kern, (unsigned)nt, (unsigned)S, smem) < 1) ok = 0;
    if (!ok) cudaGetLastError();
#endif
    if (g_fused_nshapes < 64) g_fused_shapes[g_fused_nshapes++] = FusedShape{kern, S, smem, ok, nt};
    else return false;               // (table full: take the three kernels rather than launch an unchecked shape)
    return ok != 0;
}

// Runs Env.step for environments [b0, b0 + nb) as one cluster-fused launch if the configuration allows it.
// *done = 0: not applicable, the caller takes the three kernels.  Pointers already refer to environment b0.
static int try_fused_step(die_env* e, int b0, int nb, const double* medium_in, double* medium_out, double* agents,
                          const double* action, double* reward_dev, int64_t* alive_dev, const uint32_t* alive_bits,
                          cudaStream_t st, int* done) {
    *done = 0;
    const int R = e->dyn.blur_radius;
    if (e->dyn.diffuse_mode != DIE_DIFFUSE_WRAP || R < 1 || R > 4 || (e->W & 1) != 0 || e->W < R) return DIE_OK;
    if (((uintptr_t)medium_in & 15) != 0 || alive_bits == nullptr || e->field_f32) return DIE_OK;   // (a caller's odd view of a tensor)
    const int nt = g_fused_threads;
    const bool want_grad = e->publish_grad != 0;
    const bool plain = e->flow_rwave == nullptr && e->flow_frames == nullptr;
    fused_kernel_t kern = nullptr;
    switch (R) {
        case 1: kern = pick_fused<1>(want_grad, plain, nt); break;
        case 2: kern = pick_fused<2>(want_grad, plain, nt); break;
        case 3: kern = pick_fused<3>(want_grad, plain, nt); break;
        case 4: kern = pick_fused<4>(want_grad, plain, nt); break;
    }
    // the largest cluster whose CTAs get whole rows and whose shared memory fits (two CTAs per SM if possible)
    int S = 0;
    size_t smem = 0;
    for (int pass = 0; pass < 2 && S == 0; ++pass) {
        const size_t limit = pass == 0 ? (size_t)(113u << 10) : kFusedSmemMax;
        for (int cand = 16; cand >= 1 && S == 0; cand >>= 1) {
            if (e->H % cand != 0) continue;
            // a thread group works through at most kFusedMaxRounds feed blocks (their cells stay in registers)
            if ((e->nblk + cand - 1) / cand > kFusedMaxRounds * (nt / kAgentThreads)) continue;
            const size_t need = fused_smem_bytes(e->H / cand, e->W, R, want_grad, nt);
            if (need <= limit && fused_shape_ok(kern, cand, need, nt)) { S = cand; smem = need; }
        }
    }
    if (S == 0) return DIE_OK;

    const size_t C = (size_t)e->H * e->W;
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.agents = agents; a.action = action;
    a.cells = e->cells2[e->cur] + (size_t)b0 * e->M;
    a.part_gain = e->part_gain + (size_t)b0 * e->nblk;
    a.part_alive = e->part_alive + (size_t)b0 * e->nblk;
    a.reward = reward_dev + b0; a.alive_out = alive_dev + b0;
    a.alive_bits = alive_bits; a.Mw = e->Mw;
    a.ax = make_axis(e->H); a.ay = make_axis(e->W);
    a.boundary = e->dyn.boundary;
    a.w_dep = e->dyn.cost_w_deposit; a.w_dist = e->dyn.cost_w_dist;
    a.M = e->M; a.nblk = e->nblk; a.bpc = (e->nblk + S - 1) / S;
    a.medium_in = medium_in; a.medium_out = medium_out;
    if (want_grad) {
        if (int rc = ensure_gradient_buffer(e)) return rc;
        if (publish_as_f32(e)) a.grad32 = e->grad32 + (size_t)b0 * C;
        else a.grad = e->grad + (size_t)b0 * C;
        e->grad_kind = publish_as_f32(e) ? 2 : 1;
    } else {
        e->grad_kind = 0;
    }
    a.H = e->H; a.W = e->W; a.S = S; a.rows_per = e->H / S;
    a.slab_shift = -1;
    for (int sft = 0; sft < 31; ++sft) if ((1 << sft) == a.rows_per * e->W) a.slab_shift = sft;
    a.rate_feed = e->dyn.rate_feed;
    a.keep = 1.0 - e->dyn.rate_decay_chem;
    a.food_infinite = e->dyn.food_infinite;
    if (e->flow_rwave != nullptr) {
        const int64_t k = e->flow_k % e->flow_T;
        a.flow_rwave = e->flow_rwave;
        a.flow_col = e->flow_col + k * e->W;
        a.flow_row = e->flow_row + k * e->H;
        a.flow_t = e->flow_ts[k];
        a.flow_scale = e->flow_scale;
        a.flow_keep = e->flow_keep;
    } else if (e->flow_frames != nullptr) {
        a.flow_frame = e->flow_frames + (e->flow_k % e->flow_T) * (int64_t)C;
        a.flow_scale = e->flow_scale;
        a.flow_keep = e->flow_keep;
    }
    for (int k = 0; k < 2 * DIE_MAX_RADIUS + 1; ++k) a.bw.w[k] = e->dyn.blur_w[k];
    DIE_CUDA(launch_cluster(kern, (unsigned)((int64_t)nb * S), (unsigned)nt, (unsigned)S, smem, st, a));
    ++g_count_step_fused;
    *done = 1;
    return DIE_OK;
}

// ------------------------------------------------------------------------------------------
// Env.step
// ------------------------------------------------------------------------------------------
static int g_feed_bits = 1;        // feed kernel reads alive-ness from the bitmask (when valid) instead of the float64 channel

// The four launches of Env.step for environments [b0, b0 + nb) of the batch, on stream `st`.  Every pointer argument
// refers to the WHOLE batch; the range is resolved here (all per-env arrays are contiguous per environment).
static int env_step_range(die_env_t* e, int b0, int nb, double* medium_in, double* medium_out,
                          double* agents, const double* action, double* reward_dev, int64_t* alive_dev,
                          bool fused, const uint32_t* alive_bits, bool profile, cudaStream_t st,
                          bool committed = false, bool use_cost = false) {
    // fused: cells + claims of this action are in place (the forward kernel's MOVE instantiation); committed: so are the
    // positions, and the feed kernel is the plain one
    const size_t C = (size_t)e->H * e->W, M = (size_t)e->M;
    medium_in = field_off(e, medium_in, (size_t)b0 * 3 * C);
    medium_out = field_off(e, medium_out, (size_t)b0 * 3 * C);
    agents += (size_t)b0 * 4 * M;
    action += (size_t)b0 * 3 * M;
    int32_t* winner = e->winner + (size_t)b0 * C;
    int32_t* cells = e->cells2[e->cur] + (size_t)b0 * M;
    double* part_gain = e->part_gain + (size_t)b0 * e->nblk;
    int32_t* part_alive = e->part_alive + (size_t)b0 * e->nblk;
    if (alive_bits != nullptr) alive_bits += (size_t)b0 * e->Mw;

    if (e->dyn.agents_die) {
        // the lifecycle changes the alive channel every step: the cached bitmask is never valid, and a move evaluated
        // speculatively by the forward kernel used the bitmask
        if (fused) return fail(DIE_E_INVALID, "agents_die and a speculative move do not combine%s%s");
        alive_bits = nullptr;
        e->alive_valid = 0;
    }
    const bool pair = pair_mode_for(e, fused && !committed);
    if (pair) {
        if (e->cell_pairs == nullptr) DIE_CUDA(cudaMalloc(&e->cell_pairs, sizeof(double2) * C * e->B));
        if (e->food_here == nullptr) DIE_CUDA(cudaMalloc(&e->food_here, sizeof(double) * M * e->B));
    }
    e->food_here_valid = 0;
    if (profile) prof_mark(e, 0, st);
    if (!fused && !pair && g_step_impl == 1) {
        int done = 0;
        if (int rc = try_fused_step(e, b0, nb, medium_in, medium_out, agents, action, reward_dev, alive_dev, alive_bits, st, &done))
            return rc;
        if (done) {                  // (the whole step is the first interval of the per-kernel breakdown)
            if (profile) for (int k = 1; k <= DIE_NUM_STEP_KERNELS; ++k) prof_mark(e, k, st);
            return DIE_OK;
        }
    }
    if (!fused) {
        const int mchunk = chunks_for(e->M, kMoveItems);
        auto move = (alive_bits != nullptr) ? move_claim_kernel<false, true> : move_claim_kernel<false, false>;
        move<<<(unsigned)((int64_t)mchunk * nb), kAgentThreads, 0, st>>>(
            agents, action, winner, cells, make_axis(e->H), make_axis(e->W), e->M, mchunk,
            e->dyn.boundary, alive_bits, e->Mw, SlabGeom(), SlabTables());
        DIE_CUDA(cudaGetLastError());
    }
    if (profile) prof_mark(e, 1, st);

    DIE_CUDA(launch_field_any(e, b0, nb, medium_in, medium_out, action, st, pair));
    if (profile) prof_mark(e, 2, st);

    const unsigned fgrid = (unsigned)((int64_t)e->nblk * nb);
    const bool feed_bits = alive_bits != nullptr && (fused || g_feed_bits);
    auto feed = (fused && !committed) ? agent_feed_kernel<false, true, true>
                      : (feed_bits ? agent_feed_kernel<false, false, true> : agent_feed_kernel<false, false, false>);
    if (e->dyn.agents_die) feed = agent_feed_kernel<false, false, false, true>;
    if (e->field_f32) {
        if (fused) return fail(DIE_E_INVALID, "the speculative move is not available with float32 fields%s%s");
        feed = e->dyn.agents_die ? agent_feed_kernel<false, false, false, true, float>
                                 : (feed_bits ? agent_feed_kernel<false, false, true, false, float>
                                              : agent_feed_kernel<false, false, false, false, float>);
    }
    if (pair) feed = feed_bits ? agent_feed_kernel<false, false, true, false, double, true>
                               : agent_feed_kernel<false, false, false, false, double, true>;
    // cost hint: the forward kernel left linear_action_cost of exactly this action in e->burned (die_env_step_flags checked)
    const bool cost = use_cost && feed_bits && (!fused || committed) && !e->dyn.agents_die && !e->field_f32 &&
                      e->burned != nullptr;
    if (cost) {
        feed = pair ? agent_feed_kernel<false, false, true, false, double, true, true>
                    : agent_feed_kernel<false, false, true, false, double, false, true>;
        // six resident CTAs per SM (40 instead of 48 registers, no spills): more gathers in flight, feed 1.93 -> 1.86 ms batched;
        // eight (32 registers, 40 bytes of spills): 2.28 ms -- profiles/r02zu_feed_min_blocks_ab.txt
        if (!pair && g_feed_min_blocks == 6) feed = agent_feed_kernel<false, false, true, false, double, false, true, 6>;
        ++g_count_feed_cost;
    }
    FeedArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.agents = agents; fa.action = action;
    fa.consumed_field = field_off(e, e->consumed, (size_t)b0 * C);
    if (pair) {
        fa.cell_pairs = e->cell_pairs + (size_t)b0 * C;
        fa.food_here = e->food_here + (size_t)b0 * M;
    }
    fa.winner = winner; fa.cells = cells; fa.part_gain = part_gain; fa.part_alive = part_alive;
    fa.C = (int64_t)C; fa.M = e->M; fa.nblk = e->nblk;
    fa.w_dep = e->dyn.cost_w_deposit; fa.w_dist = e->dyn.cost_w_dist;
    fa.alive_bits = alive_bits; fa.Mw = e->Mw; fa.boundary = e->dyn.boundary;
    if (cost) fa.burned = e->burned + (size_t)b0 * M;
    feed<<<fgrid, kAgentThreads, 0, st>>>(fa, SlabGeom(), SlabTables());
    DIE_CUDA(cudaGetLastError());
    if (profile) prof_mark(e, 3, st);

    finalize_stats_kernel<<<nb, kFinalThreads, 0, st>>>(part_gain, part_alive, e->nblk, reward_dev + b0, alive_dev + b0);
    DIE_CUDA(cudaGetLastError());
    if (profile) prof_mark(e, 4, st);
    e->food_here_valid = pair ? 1 : 0;
    return DIE_OK;
}

extern "C" int die_env_step_flags(die_env_t* e, double* medium_in, double* medium_out,
                                  double* agents, const double* action,
                                  double* reward_dev, int64_t* alive_dev, int32_t flags, void* stream) {
    DIE_REQUIRE(e != nullptr);
    DIE_REQUIRE(medium_in != nullptr && medium_out != nullptr && medium_in != medium_out);
    DIE_REQUIRE(agents != nullptr && action != nullptr);
    DIE_REQUIRE(reward_dev != nullptr && alive_dev != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    const bool fused = (flags & DIE_STEP_ADOPT_MOVE) != 0;
    const bool bits = (flags & DIE_STEP_ALIVE_BITS) != 0;
    if ((fused || bits) && !e->alive_valid)
        return fail(DIE_E_INVALID, "die_env_step_flags: call die_env_refresh_alive first%s%s");
    const uint32_t* alive_bits = (fused || bits) ? e->alive_bits : nullptr;
    // DIE_STEP_USE_COST is a permission: honoured only while the forward's cost array is still the pending one
    const bool use_cost = (flags & DIE_STEP_USE_COST) != 0 && e->cost_pending && g_cost_hint;
    e->cost_pending = 0;
    bool committed = false;
    if (fused) {
        // the forward kernel already resolved cells and claims for exactly this action
        if (!e->pending_move) return fail(DIE_E_INVALID, "die_env_step_flags: no speculative move is pending%s%s");
        committed = e->pending_move == 2;
        e->cur ^= 1;
        e->pending_move = 0;
        if (committed) ++g_count_step_committed;
    } else if (e->pending_move == 2) {
        return fail(DIE_E_INVALID, "die_env_step_flags: a committed move is pending (DIE_FWD_COMMIT_MOVE): the next step "
                                   "must adopt it (DIE_STEP_ADOPT_MOVE, the very action of that forward)%s%s");
    } else if (e->pending_move) {       // abandoned speculation: its claims must not leak into this step
        if (int rc = die_env_discard_move(e, stream)) return rc;
    }
    if (int rc = env_step_range(e, 0, e->B, medium_in, medium_out, agents, action, reward_dev, alive_dev,
                                fused, alive_bits, true, st, committed, use_cost))
        return rc;
    if (e->flow_rwave != nullptr || e->flow_frames != nullptr) ++e->flow_k;
    if (e->profiling && e->prof_steps < DIE_MAX_PROFILED_STEPS) ++e->prof_steps;
    return DIE_OK;
}

extern "C" int die_env_step(die_env_t* e, double* medium_in, double* medium_out,
                            double* agents, const double* action,
                            double* reward_dev, int64_t* alive_dev, void* stream) {
    return die_env_step_flags(e, medium_in, medium_out, agents, action, reward_dev, alive_dev, 0, stream);
}

extern "C" int die_env_discard_move(die_env_t* e, void* stream) {
    DIE_REQUIRE(e != nullptr);
    if (!e->pending_move) return DIE_OK;
    DIE_CUDA(cudaMemsetAsync(e->winner, 0xFF, sizeof(int32_t) * (size_t)e->H * e->W * e->B, (cudaStream_t)stream));
    e->pending_move = 0;
    return DIE_OK;
}

extern "C" int die_env_refresh_alive(die_env_t* e, const double* agents, void* stream) {
    DIE_REQUIRE(e != nullptr && agents != nullptr);
    const int64_t total = (int64_t)e->B * e->Mw * 32;
    alive_bits_kernel<<<grid_for(total, 256, e->num_sms), 256, 0, (cudaStream_t)stream>>>(
        agents, e->alive_bits, e->M, e->Mw, e->B);
    DIE_CUDA(cudaGetLastError());
    e->alive_valid = 1;
    return DIE_OK;
}

extern "C" int die_env_pending_move(const die_env_t* e) { return e ? e->pending_move : 0; }

extern "C" int die_env_read_stats(die_env_t* e, const double* reward_dev, const int64_t* alive_dev,
                                  double* reward_host, int64_t* alive_host, void* stream) {
    DIE_REQUIRE(e != nullptr && reward_dev != nullptr && alive_dev != nullptr);
    DIE_REQUIRE(reward_host != nullptr && alive_host != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    DIE_CUDA(cudaMemcpyAsync(reward_host, reward_dev, sizeof(double) * e->B, cudaMemcpyDeviceToHost, st));
    DIE_CUDA(cudaMemcpyAsync(alive_host, alive_dev, sizeof(int64_t) * e->B, cudaMemcpyDeviceToHost, st));
    DIE_CUDA(cudaStreamSynchronize(st));
    return DIE_OK;
}

// Env.step through host buffers.  The environments of a batch are independent, so the batch is cut into chunks that
// run on two internal streams: while chunk k's observation travels device -> host, chunk k+1's action travels host ->
// device and its kernels run -- both PCIe directions stay busy instead of taking turns.
static int g_host_chunks = 4;
static size_t g_host_chunk_min_bytes = (size_t)32 << 20;   // below this a call is not worth chunking

static int env_step_host_impl(die_env_t* e, double* medium_in, double* medium_out,
                              double* agents, const double* action_host, const double* action_dev,
                              double* agents_host, double* medium_host,
                              double* reward_host, int64_t* alive_host, void* stream, int32_t flags = 0) {
    if ((flags & DIE_HOST_KEEP_ALIVE_CHANNEL) && e != nullptr && e->dyn.agents_die)
        return fail(DIE_E_INVALID, "DIE_HOST_KEEP_ALIVE_CHANNEL: with agents_die the alive channel changes every step%s%s");
    DIE_REQUIRE(e != nullptr && (action_host != nullptr) != (action_dev != nullptr));
    if (e->field_f32) return fail(DIE_E_INVALID, "the host-buffer step runs float64 fields only%s%s");
    DIE_REQUIRE(medium_in != nullptr && medium_out != nullptr && medium_in != medium_out && agents != nullptr);
    DIE_REQUIRE(reward_host != nullptr && alive_host != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t C = (size_t)e->H * e->W, M = (size_t)e->M;
    if (action_dev == nullptr && e->action_stage == nullptr)
        DIE_CUDA(cudaMalloc(&e->action_stage, sizeof(double) * 3 * M * e->B));
    const double* action = action_dev != nullptr ? action_dev : e->action_stage;
    e->cost_pending = 0;
    if (e->pending_move == 2)
        return fail(DIE_E_INVALID, "the host-buffer step cannot follow a committed move (DIE_FWD_COMMIT_MOVE)%s%s");
    if (e->pending_move) {
        if (int rc = die_env_discard_move(e, stream)) return rc;
    }
    // chunking pays once a chunk's copies are a few MB; small batches take the single-stream path
    int nchunks = (g_host_chunks > 1 && e->B >= 2 * g_host_chunks && sizeof(double) * 3 * M * e->B >= g_host_chunk_min_bytes)
                      ? g_host_chunks : 1;
    if (nchunks > 1 && e->host_streams[0] == nullptr) {
        for (int k = 0; k < 2; ++k) {
            DIE_CUDA(cudaStreamCreateWithFlags(&e->host_streams[k], cudaStreamNonBlocking));
            DIE_CUDA(cudaEventCreateWithFlags(&e->host_events[k], cudaEventDisableTiming));
        }
        DIE_CUDA(cudaEventCreateWithFlags(&e->host_events[2], cudaEventDisableTiming));
    }
    if (nchunks > 1) DIE_CUDA(cudaEventRecord(e->host_events[2], st));       // the chunks start after the caller's work
    for (int k = 0; k < nchunks; ++k) {
        const int b0 = (int)((int64_t)e->B * k / nchunks), b1 = (int)((int64_t)e->B * (k + 1) / nchunks);
        const int nb = b1 - b0;
        cudaStream_t s = (nchunks > 1) ? e->host_streams[k & 1] : st;
        if (nchunks > 1 && k < 2) DIE_CUDA(cudaStreamWaitEvent(s, e->host_events[2], 0));
        if (action_dev == nullptr)
            DIE_CUDA(cudaMemcpyAsync(e->action_stage + (size_t)b0 * 3 * M, action_host + (size_t)b0 * 3 * M,
                                     sizeof(double) * 3 * M * nb, cudaMemcpyHostToDevice, s));
        if (int rc = env_step_range(e, b0, nb, medium_in, medium_out, agents, action, e->reward_dev, e->alive_dev,
                                    false, nullptr, false, s))
            return rc;
        if (agents_host != nullptr && (flags & DIE_HOST_KEEP_ALIVE_CHANNEL)) {
            // the step never changes the alive channel (no lifecycle): the caller's host copy of it is still right, so only
            // x, y (channels 0-1, contiguous per env) and agent_food (channel 3) travel -- 24 instead of 32 B per slot
            DIE_CUDA(cudaMemcpy2DAsync(agents_host + (size_t)b0 * 4 * M, sizeof(double) * 4 * M,
                                       agents + (size_t)b0 * 4 * M, sizeof(double) * 4 * M,
                                       sizeof(double) * 2 * M, (size_t)nb, cudaMemcpyDeviceToHost, s));
            DIE_CUDA(cudaMemcpy2DAsync(agents_host + (size_t)b0 * 4 * M + 3 * M, sizeof(double) * 4 * M,
                                       agents + (size_t)b0 * 4 * M + 3 * M, sizeof(double) * 4 * M,
                                       sizeof(double) * M, (size_t)nb, cudaMemcpyDeviceToHost, s));
        } else if (agents_host != nullptr)
            DIE_CUDA(cudaMemcpyAsync(agents_host + (size_t)b0 * 4 * M, agents + (size_t)b0 * 4 * M,
                                     sizeof(double) * 4 * M * nb, cudaMemcpyDeviceToHost, s));
        if (medium_host != nullptr)
            DIE_CUDA(cudaMemcpyAsync(medium_host + (size_t)b0 * 3 * C, medium_out + (size_t)b0 * 3 * C,
                                     sizeof(double) * 3 * C * nb, cudaMemcpyDeviceToHost, s));
        DIE_CUDA(cudaMemcpyAsync(reward_host + b0, e->reward_dev + b0, sizeof(double) * nb, cudaMemcpyDeviceToHost, s));
        DIE_CUDA(cudaMemcpyAsync(alive_host + b0, e->alive_dev + b0, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, s));
    }
    if (nchunks > 1) {
        for (int k = 0; k < 2; ++k) {                                         // the caller's stream continues after both
            DIE_CUDA(cudaEventRecord(e->host_events[k], e->host_streams[k]));
            DIE_CUDA(cudaStreamWaitEvent(st, e->host_events[k], 0));
        }
    }
    if (e->flow_rwave != nullptr || e->flow_frames != nullptr) ++e->flow_k;
    DIE_CUDA(cudaStreamSynchronize(st));
    return DIE_OK;
}

extern "C" int die_env_step_host(die_env_t* e, double* medium_in, double* medium_out,
                                 double* agents, const double* action_host,
                                 double* agents_host, double* medium_host,
                                 double* reward_host, int64_t* alive_host, void* stream) {
    return env_step_host_impl(e, medium_in, medium_out, agents, action_host, nullptr, agents_host, medium_host,
                              reward_host, alive_host, stream);
}

extern "C" int die_env_step_host_dev(die_env_t* e, double* medium_in, double* medium_out,
                                     double* agents, const double* action_dev,
                                     double* agents_host, double* medium_host,
                                     double* reward_host, int64_t* alive_host, void* stream) {
    return env_step_host_impl(e, medium_in, medium_out, agents, nullptr, action_dev, agents_host, medium_host,
                              reward_host, alive_host, stream);
}

extern "C" int die_env_step_host_flags(die_env_t* e, double* medium_in, double* medium_out,
                                       double* agents, const double* action_host, const double* action_dev,
                                       double* agents_host, double* medium_host,
                                       double* reward_host, int64_t* alive_host, int32_t flags, void* stream) {
    return env_step_host_impl(e, medium_in, medium_out, agents, action_host, action_dev, agents_host, medium_host,
                              reward_host, alive_host, stream, flags);
}

// ------------------------------------------------------------------------------------------
// Env._get_sensed_medium with Dynamics.apply_sense_mask (core/env.py:275-294)
// ------------------------------------------------------------------------------------------
template <int R>
static cudaError_t launch_sense_mask(const double* medium, double* obs, int H, int W, int B,
                                     const BlurWeights& bw, cudaStream_t st) {
    constexpr int TH = 32, TW = 64, NT = 256;
    const int tiles_i = (H + TH - 1) / TH, tiles_j = (W + TW - 1) / TW;
    const size_t smem = sizeof(double) * (size_t)((TH + 2 * R) * (TW + 2 * R) + TH * (TW + 2 * R));
    auto kern = sense_mask_kernel<R, TH, TW, NT>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<(unsigned)((int64_t)tiles_i * tiles_j * B), NT, smem, st>>>(medium, obs, H, W, tiles_i, tiles_j, bw);
    return cudaGetLastError();
}

extern "C" int die_sense_mask(int32_t H, int32_t W, int32_t B, const double* weights, int32_t radius,
                              const double* medium, double* obs, void* stream) {
    DIE_REQUIRE(H >= 1 && W >= 1 && B >= 1 && (int64_t)H * W <= 0x7fffffffLL);
    DIE_REQUIRE(weights != nullptr && medium != nullptr && obs != nullptr && medium != obs);
    DIE_REQUIRE(radius >= 1 && radius <= DIE_MAX_RADIUS);
    BlurWeights bw;
    memset(&bw, 0, sizeof(bw));
    for (int k = 0; k < 2 * radius + 1; ++k) bw.w[k] = weights[k];
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t err = cudaErrorInvalidValue;
    switch (radius) {
        case 1: err = launch_sense_mask<1>(medium, obs, H, W, B, bw, st); break;
        case 2: err = launch_sense_mask<2>(medium, obs, H, W, B, bw, st); break;
        case 3: err = launch_sense_mask<3>(medium, obs, H, W, B, bw, st); break;
        case 4: err = launch_sense_mask<4>(medium, obs, H, W, B, bw, st); break;
        case 5: err = launch_sense_mask<5>(medium, obs, H, W, B, bw, st); break;
        case 6: err = launch_sense_mask<6>(medium, obs, H, W, B, bw, st); break;
        case 7: err = launch_sense_mask<7>(medium, obs, H, W, B, bw, st); break;
        case 8: err = launch_sense_mask<8>(medium, obs, H, W, B, bw, st); break;
    }
    DIE_CUDA(err);
    return DIE_OK;
}

// ------------------------------------------------------------------------------------------
// EnvRenderer.render (core/render.py:76-132)
// ------------------------------------------------------------------------------------------
extern "C" int die_render_frames(int32_t H, int32_t W, int64_t M, int32_t B,
                                 const double* medium, const double* agents, double* trace, double decay,
                                 const double* color_host, double* img_medium, double* img_agents, void* stream) {
    DIE_REQUIRE(H >= 1 && W >= 1 && B >= 1 && (int64_t)H * W <= 0x7fffffffLL);
    DIE_REQUIRE(medium != nullptr && trace != nullptr && img_medium != nullptr);
    DIE_REQUIRE(img_agents == nullptr || (agents != nullptr && M == (int64_t)H * W));
    RenderArgs a;
    memset(&a, 0, sizeof(a));
    a.medium = medium; a.agents = agents; a.trace = trace; a.img_medium = img_medium; a.img_agents = img_agents;
    a.H = H; a.W = W; a.M = M; a.decay = decay;
    if (color_host != nullptr) {
        a.use_color = 1;
        for (int k = 0; k < 3; ++k) a.color[k] = color_host[k];
    }
    const int64_t total = (int64_t)H * W * B;
    render_frames_kernel<<<grid_for(total, 256, 148), 256, 0, (cudaStream_t)stream>>>(a, total);
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}

// ------------------------------------------------------------------------------------------
// Agent.forward
// ------------------------------------------------------------------------------------------
extern "C" int die_brownian_forward(const double* agents, double* action, int64_t M, int32_t B,
                                    double move_scale, double deposit_scale,
                                    const double* u, uint64_t seed, uint64_t step, void* stream) {
    DIE_REQUIRE(agents != nullptr && action != nullptr);
    DIE_REQUIRE(M >= 1 && B >= 1);
    DIE_REQUIRE(M <= 0x7fffffffLL);
    const int nchunk = chunks_for(M, kBrownItems);
    brownian_forward_kernel<<<(unsigned)((int64_t)nchunk * B), kAgentThreads, 0, (cudaStream_t)stream>>>(
        agents, action, M, nchunk, move_scale, deposit_scale, u, seed, step, nullptr);
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}

extern "C" int die_brownian_forward_dev(const double* agents, double* action, int64_t M, int32_t B,
                                        double move_scale, double deposit_scale,
                                        uint64_t seed, const uint64_t* step_dev, void* stream) {
    DIE_REQUIRE(agents != nullptr && action != nullptr && step_dev != nullptr);
    DIE_REQUIRE(M >= 1 && B >= 1);
    DIE_REQUIRE(M <= 0x7fffffffLL);
    const int nchunk = chunks_for(M, kBrownItems);
    brownian_forward_kernel<<<(unsigned)((int64_t)nchunk * B), kAgentThreads, 0, (cudaStream_t)stream>>>(
        agents, action, M, nchunk, move_scale, deposit_scale, nullptr, seed, 0, step_dev);
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}

extern "C" int die_const_forward(double* action, int64_t M, int32_t B,
                                 double dx, double dy, double deposit, void* stream) {
    DIE_REQUIRE(action != nullptr);
    DIE_REQUIRE(M >= 1 && B >= 1);
    DIE_REQUIRE(M <= 0x7fffffffLL);
    const int nchunk = chunks_for(M, kBrownItems);
    const_forward_kernel<<<(unsigned)((int64_t)nchunk * B), kAgentThreads, 0, (cudaStream_t)stream>>>(
        action, M, nchunk, dx, dy, deposit);
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}

// JonesAgent.forward (the classic three-sensor particle; specification: oracle/die_ref.py:JonesAgent)
extern "C" int die_jones_forward(const die_jones_params_t* p, int32_t H, int32_t W, int64_t M, int32_t B,
                                 const double* agents, const void* medium, int32_t medium_f32,
                                 double* theta, double* action, const uint8_t* coin,
                                 uint64_t seed, uint64_t step, int32_t step_on_device, void* stream) {
    DIE_REQUIRE(p != nullptr);
    DIE_REQUIRE(H >= 2 && W >= 2 && M >= 1 && B >= 1);
    DIE_REQUIRE((int64_t)H * W <= 0x7fffffffLL && M <= 0x7fffffffLL);
    DIE_REQUIRE(agents != nullptr && medium != nullptr && theta != nullptr && action != nullptr);
    DIE_REQUIRE(p->sense_radians >= 0.0 && p->sense_radians <= DIE_PI && p->turn_radians >= 0.0 && p->turn_radians <= DIE_PI);
    JonesArgs a;
    memset(&a, 0, sizeof(a));
    a.ax = make_axis(H);
    a.ay = make_axis(W);
    a.W = W; a.M = M; a.C = (int64_t)H * W;
    a.nchunk = chunks_for(M, kJonesItems);
    a.agents = agents; a.medium = (const double*)medium; a.theta = theta; a.action = action; a.coin = coin;
    a.scale = p->scale; a.deposit = p->deposit; a.sense_offset = p->sense_offset;
    a.sense_radians = p->sense_radians; a.turn_radians = p->turn_radians;
    a.seed = seed;
    if (step_on_device) a.step_dev = (const uint64_t*)(uintptr_t)step;
    else a.step = step;
    const unsigned grid = (unsigned)((int64_t)a.nchunk * B);
    if (medium_f32) jones_forward_kernel<float><<<grid, kAgentThreads, 0, (cudaStream_t)stream>>>(a);
    else jones_forward_kernel<double><<<grid, kAgentThreads, 0, (cudaStream_t)stream>>>(a);
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}

static int g_turn_quick = 1;       // 0: every slot runs die_turn_exact (diagnosis / A-B tests; same results)
static int g_sense_quick = 1;      // 0: the sensed cell always comes from the float64 die_sincos (A-B tests; same results)
static int g_fwd_lean = 1;         // use the LEAN instantiation of the forward kernel when its preconditions hold
static int g_fwd_min_blocks = 4;   // register cap of the forward kernel (3 / 4 / 5 resident CTAs per SM)

extern "C" int die_set_turn_quick(int32_t on) {
    g_turn_quick = on ? 1 : 0;
    return DIE_OK;
}

// cudaLimitMaxL2FetchGranularity of the current device (32 / 64 / 128 bytes; a hint to the L2: how much a miss fetches
// from DRAM).  Random 8-byte gathers over tables larger than L2 -- the spread-out ghost slots of one large field, DESIGN.md
// 5.1 -- pay the full granule per gather.  Device-wide and sticky for the process: returns the previous value in *previous.
extern "C" int die_device_l2_fetch_granularity(int32_t bytes, int32_t* previous) {
    size_t old = 0;
    DIE_CUDA(cudaDeviceGetLimit(&old, cudaLimitMaxL2FetchGranularity));
    if (previous != nullptr) *previous = (int32_t)old;
    if (bytes > 0) {
        DIE_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128);
        DIE_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
    }
    return DIE_OK;
}

extern "C" int die_set_tuning(const char* key, int32_t value) {
    DIE_REQUIRE(key != nullptr);
    if (strcmp(key, "turn_quick") == 0) g_turn_quick = value ? 1 : 0;
    else if (strcmp(key, "sense_quick") == 0) g_sense_quick = value ? 1 : 0;
    else if (strcmp(key, "fwd_min_blocks") == 0) { DIE_REQUIRE(value >= 3 && value <= 5); g_fwd_min_blocks = value; }
    else if (strcmp(key, "feed_bits") == 0) g_feed_bits = value ? 1 : 0;
    else if (strcmp(key, "host_chunks") == 0) { DIE_REQUIRE(value >= 1 && value <= 64); g_host_chunks = value; }
    else if (strcmp(key, "host_chunk_min_kb") == 0) { DIE_REQUIRE(value >= 0); g_host_chunk_min_bytes = (size_t)value << 10; }
    else if (strcmp(key, "fwd_lean") == 0) g_fwd_lean = value;      // 0 off, 1 on (4 CTAs/SM), 5: 48-register cap
    else if (strcmp(key, "field_prefetch") == 0) g_field_prefetch = value ? 1 : 0;
    else if (strcmp(key, "grad_f32") == 0) g_grad_f32 = value ? 1 : 0;
    else if (strcmp(key, "step_impl") == 0) return die_set_step_impl(value);
    else if (strcmp(key, "field_vec") == 0) g_field_vec = value ? 1 : 0;
    else if (strcmp(key, "field_tile") == 0) { DIE_REQUIRE(value >= 0 && value <= 2); g_field_tile = value; }
    else if (strcmp(key, "feed_min_blocks") == 0) { DIE_REQUIRE(value == 0 || value == 6); g_feed_min_blocks = value; }
    else if (strcmp(key, "cost_hint") == 0) g_cost_hint = value ? 1 : 0;
    else if (strcmp(key, "cost_sqrt_near") == 0) g_cost_sqrt_near = value ? 1 : 0;
    else if (strcmp(key, "pair_mode") == 0) { DIE_REQUIRE(value >= 0 && value <= 2); g_pair_mode = value; }
    else if (strcmp(key, "pair_min_cells_log2") == 0) { DIE_REQUIRE(value >= 2 && value <= 31); g_pair_min_cells = (int64_t)1 << value; }
    else if (strcmp(key, "fused_threads") == 0) { DIE_REQUIRE(value == 512); g_fused_threads = value; }
    else return fail(DIE_E_INVALID, "die_set_tuning: unknown key %s%s", key);
    return DIE_OK;
}

static die_turn_plan_t plan_for(const die_gradient_params_t* p) {
    die_turn_plan_t plan;
    memset(&plan, 0, sizeof(plan));
    if (g_turn_quick && p->discrete_turn)
        plan = die_turn_plan(p->normalized_grad, p->use_grad_clip, p->grad_clip,
                             p->turn_radians * p->turn_tolerance, p->sense_radians);
    return plan;
}

static int gradient_forward_impl(die_env_t* env, int move_mode, const die_gradient_params_t* p,
                                 int32_t H, int32_t W, int64_t M, int32_t B,
                                 const double* agents, const double* medium,
                                 double* theta, double* prev_grad, double* action,
                                 const uint8_t* coin, const double* noise, int32_t* sense_cells,
                                 const double* grad_hint, const int32_t* cells_hint,
                                 uint64_t seed, uint64_t step, void* stream, int b0 = 0,
                                 const float2* grad32_hint = nullptr, const uint64_t* step_dev = nullptr,
                                 bool field_f32 = false, bool write_cost = false) {
    DIE_REQUIRE(p != nullptr);
    DIE_REQUIRE(H >= 2 && W >= 2 && M >= 1 && B >= 1);
    DIE_REQUIRE((int64_t)H * W <= 0x7fffffffLL);
    DIE_REQUIRE(agents != nullptr && medium != nullptr && theta != nullptr && action != nullptr);
    // move_mode: 0 the action only; 1 + the move, speculatively (DIE_FWD_SPECULATE_MOVE); 2 + the move, committed
    // (DIE_FWD_COMMIT_MOVE: x, y stored into `agents`)
    const bool speculate = move_mode != 0;
    // prev_grad may only be omitted when it cannot influence any output
    DIE_REQUIRE(prev_grad != nullptr || (p->inertia == 0.0 && p->noise_scale == 0.0));
    GradientArgs a;
    memset(&a, 0, sizeof(a));
    a.p = *p;
    a.plan = plan_for(p);
    DIE_REQUIRE(M <= 0x7fffffffLL);
    a.H = H; a.W = W; a.M = M; a.nchunk = chunks_for(M, kFwdItems);
    a.ax = make_axis(H);
    a.ay = make_axis(W);
    a.agents = agents; a.medium = medium; a.theta = theta; a.prev_grad = prev_grad;
    a.action = action; a.coin = coin; a.noise = noise; a.sense_cells = sense_cells;
    a.grad = (const double2*)grad_hint; a.cells = cells_hint;
    // the float32 gradient serves the guard-banded turn decision only (die_turn.h); anything that needs the VALUE of
    // the gradient (GradientAgent, unnormalised gradients, the exact turn path for every slot) samples chem1 instead
    if (grad_hint == nullptr && grad32_hint != nullptr && p->discrete_turn && a.plan.enabled && p->normalized_grad)
        a.grad32 = grad32_hint;
    a.seed = seed; a.step = step;
    a.step_dev = step_dev;
    if (g_sense_quick) {           // float32 sin / cos for the sensed cell, guarded (die_device.cuh: nearest_cell_guarded)
        a.sense_guard_x = fabs(p->sense_offset) * DIE_SINCOSF_ERR * a.ax.nm1 + 1e-9;
        a.sense_guard_y = fabs(p->sense_offset) * DIE_SINCOSF_ERR * a.ay.nm1 + 1e-9;
    }
    a.b0 = b0;
    const unsigned grid = (unsigned)((int64_t)a.nchunk * B);
    cudaStream_t st = (cudaStream_t)stream;
    if (speculate) {
        a.winner = env->winner;
        a.cells_out = env->cells2[1 - env->cur];
        a.alive_bits = env->alive_bits;
        a.Mw = env->Mw;
        a.boundary = env->dyn.boundary;
        if (move_mode == 2) a.commit_xy = const_cast<double*>(agents);
    }
    if (write_cost && env != nullptr && g_cost_hint && !field_f32) {
        if (env->burned == nullptr) DIE_CUDA(cudaMalloc(&env->burned, sizeof(double) * (size_t)M * B));
        a.burned_out = env->burned;
        a.cost_w_dep = env->dyn.cost_w_deposit;
        a.cost_w_dist = env->dyn.cost_w_dist;
        // a Physarum turn on a normalised gradient with an identity momentum step writes scale * (cos d, sin d)
        if (g_cost_sqrt_near && p->discrete_turn && p->normalized_grad && p->inertia == 0.0 && p->noise_scale == 0.0)
            a.cost_sqrt = die_sqrt_near_plan(p->scale);
    }
    void (*kern)(const GradientArgs) = nullptr;
#define DIE_PICK_FWD(MINB)                                                                               \
    kern = speculate ? (p->discrete_turn ? gradient_forward_kernel<true, false, true, MINB>              \
                                         : gradient_forward_kernel<false, false, true, MINB>)            \
                     : (p->discrete_turn ? gradient_forward_kernel<true, false, false, MINB>             \
                                         : gradient_forward_kernel<false, false, false, MINB>)
    if (g_fwd_min_blocks == 3) { DIE_PICK_FWD(3); }
    else if (g_fwd_min_blocks == 5) { DIE_PICK_FWD(5); }
    else { DIE_PICK_FWD(4); }
#undef DIE_PICK_FWD
    // the steady-state Physarum configuration has its own instantiation (see LEAN in die_agent_kernels.cuh)
    const bool lean = g_fwd_lean && (!speculate || move_mode == 2) && p->discrete_turn && a.plan.enabled && p->normalized_grad &&
                      prev_grad == nullptr && coin == nullptr && noise == nullptr && sense_cells == nullptr &&
                      (a.grad != nullptr || a.grad32 != nullptr) && a.cells != nullptr && g_fwd_min_blocks == 4;
    // the food under every slot, handed over by the last step's feed kernel (pair mode): valid exactly when the cell
    // cache is (the caller proved the observation is the env's own state of that step)
    const bool fh = lean && !field_f32 && env != nullptr && env->food_here_valid && env->food_here != nullptr &&
                    cells_hint == env->cells2[env->cur] && g_fwd_lean != 5;
    if (fh) a.food_here = env->food_here;
    if (lean && speculate) {
        // the committed move of the run loop: the steady-state instantiation + Env._agent_move in one launch (3 CTAs per
        // SM, 73-76 registers: the move needs x, y, the axes and the alive word alive to the end of the item; capped at
        // 64 registers it spills 32 bytes and is 6 % slower, profiles/r02zj_committed_move_ab.txt)
        if (fh) kern = (a.grad32 != nullptr) ? gradient_forward_kernel<true, false, true, 3, true, true, double, true>
                                             : gradient_forward_kernel<true, false, true, 3, true, false, double, true>;
        else kern = (a.grad32 != nullptr) ? gradient_forward_kernel<true, false, true, 3, true, true>
                                          : gradient_forward_kernel<true, false, true, 3, true, false>;
        if (fh) ++g_count_fwd_food_here;
        ++g_count_fwd_lean_move;
    } else if (fh) {
        kern = (a.grad32 != nullptr) ? gradient_forward_kernel<true, false, false, 4, true, true, double, true>
                                     : gradient_forward_kernel<true, false, false, 4, true, false, double, true>;
        ++g_count_fwd_food_here;
    } else if (lean) {
        if (a.grad32 != nullptr) {
            if (g_fwd_lean == 5) kern = gradient_forward_kernel<true, false, false, 5, true, true>;
            else kern = gradient_forward_kernel<true, false, false, 4, true, true>;
        } else {
            if (g_fwd_lean == 5) kern = gradient_forward_kernel<true, false, false, 5, true>;
            else kern = gradient_forward_kernel<true, false, false, 4, true>;
        }
    }
    if (field_f32) {               // the medium holds float32: the same kernels reading float (64-register variants only)
        if (speculate) return fail(DIE_E_INVALID, "the speculative move is not available with float32 fields%s%s");
        if (lean) kern = (a.grad32 != nullptr) ? gradient_forward_kernel<true, false, false, 4, true, true, float>
                                               : gradient_forward_kernel<true, false, false, 4, true, false, float>;
        else kern = p->discrete_turn ? gradient_forward_kernel<true, false, false, 4, false, false, float>
                                     : gradient_forward_kernel<false, false, false, 4, false, false, float>;
    }
    kern<<<grid, kAgentThreads, 0, st>>>(a);
    DIE_CUDA(cudaGetLastError());
    ++(lean ? (a.grad32 != nullptr ? g_count_fwd_lean_f32 : g_count_fwd_lean) : g_count_fwd_general);
    if (speculate) env->pending_move = move_mode;
    if (env != nullptr) env->cost_pending = (a.burned_out != nullptr) ? 1 : 0;
    return DIE_OK;
}

extern "C" int die_gradient_forward(const die_gradient_params_t* p,
                                    int32_t H, int32_t W, int64_t M, int32_t B,
                                    const double* agents, const double* medium,
                                    double* theta, double* prev_grad, double* action,
                                    const uint8_t* coin, const double* noise,
                                    int32_t* sense_cells,
                                    const double* grad_hint, const int32_t* cells_hint,
                                    uint64_t seed, uint64_t step, void* stream) {
    return gradient_forward_impl(nullptr, 0, p, H, W, M, B, agents, medium, theta, prev_grad, action,
                                 coin, noise, sense_cells, grad_hint, cells_hint, seed, step, stream);
}

extern "C" int die_gradient_forward_f32(const die_gradient_params_t* p,
                                        int32_t H, int32_t W, int64_t M, int32_t B,
                                        const double* agents, const float* medium,
                                        double* theta, double* prev_grad, double* action,
                                        const uint8_t* coin, const double* noise,
                                        int32_t* sense_cells, uint64_t seed, uint64_t step, void* stream) {
    return gradient_forward_impl(nullptr, 0, p, H, W, M, B, agents, (const double*)medium, theta, prev_grad, action,
                                 coin, noise, sense_cells, nullptr, nullptr, seed, step, stream, 0, nullptr, nullptr, true);
}

extern "C" int die_env_forward_gradient(die_env_t* e, const die_gradient_params_t* p,
                                        const double* agents, const double* medium,
                                        double* theta, double* prev_grad, double* action,
                                        const uint8_t* coin, const double* noise, int32_t* sense_cells,
                                        int32_t flags, uint64_t seed, uint64_t step, void* stream) {
    DIE_REQUIRE(e != nullptr);
    const bool speculate = (flags & (DIE_FWD_SPECULATE_MOVE | DIE_FWD_COMMIT_MOVE)) != 0;
    const int move_mode = (flags & DIE_FWD_COMMIT_MOVE) ? 2 : (speculate ? 1 : 0);
    if (e->pending_move == 2)
        return fail(DIE_E_INVALID, "die_env_forward_gradient: a committed move is pending (DIE_FWD_COMMIT_MOVE): the env's "
                                   "positions are already those of the next step, which must adopt it first%s%s");
    if (move_mode == 2 && (e->dyn.agents_die || e->field_f32))
        return fail(DIE_E_INVALID, "the committed move is not available with agents_die or float32 fields%s%s");
    if (speculate) {
        if (!e->alive_valid) return fail(DIE_E_INVALID, "die_env_forward_gradient: call die_env_refresh_alive first%s%s");
        if (e->pending_move) {          // a previous speculation was never adopted: drop its claims
            if (int rc = die_env_discard_move(e, stream)) return rc;
        }
    }
    // what the NEXT field pass publishes follows this consumer: float32 pairs serve the guard-banded turn decision only
    e->grad_f32_ok = (p != nullptr && p->discrete_turn && p->normalized_grad && plan_for(p).enabled) ? 1 : 0;
    const double* grad_hint = (flags & DIE_FWD_USE_GRADIENT) ? die_env_gradient(e) : nullptr;
    const float2* grad32_hint = ((flags & DIE_FWD_USE_GRADIENT) && die_env_gradient_kind(e) == 2) ? e->grad32 : nullptr;
    const int32_t* cells_hint = (flags & DIE_FWD_USE_CELLS) ? e->cells2[e->cur] : nullptr;
    // DIE_FWD_STEP_ON_DEVICE: `step` is the device address of a uint64 call counter (CUDA-graph replays)
    const uint64_t* step_dev = (flags & DIE_FWD_STEP_ON_DEVICE) ? (const uint64_t*)(uintptr_t)step : nullptr;
    return gradient_forward_impl(e, move_mode, p, e->H, e->W, e->M, e->B, agents, medium, theta, prev_grad, action,
                                 coin, noise, sense_cells, grad_hint, cells_hint, seed, step_dev ? 0 : step, stream, 0,
                                 grad32_hint, step_dev, e->field_f32 != 0, (flags & DIE_FWD_WRITE_COST) != 0);
}

// ------------------------------------------------------------------------------------------
// Agent.forward through host buffers, chunked over the environments of the batch
// ------------------------------------------------------------------------------------------
struct die_host_ctx {
    cudaStream_t streams[2];
    cudaEvent_t events[3];
};

extern "C" int die_host_ctx_create(die_host_ctx_t** out) {
    DIE_REQUIRE(out != nullptr);
    *out = nullptr;
    die_host_ctx* c = new (std::nothrow) die_host_ctx();
    if (c == nullptr) return fail(DIE_E_NOMEM, "out of host memory");
    memset(c, 0, sizeof(*c));
    cudaError_t err = cudaSuccess;
    for (int k = 0; k < 2 && err == cudaSuccess; ++k) err = cudaStreamCreateWithFlags(&c->streams[k], cudaStreamNonBlocking);
    for (int k = 0; k < 3 && err == cudaSuccess; ++k) err = cudaEventCreateWithFlags(&c->events[k], cudaEventDisableTiming);
    if (err != cudaSuccess) {
        die_host_ctx_destroy(c);
        return fail(DIE_E_CUDA, "die_host_ctx_create: %s", cudaGetErrorString(err));
    }
    *out = c;
    return DIE_OK;
}

extern "C" int die_host_ctx_destroy(die_host_ctx_t* c) {
    if (c == nullptr) return DIE_OK;
    for (int k = 0; k < 2; ++k) if (c->streams[k]) cudaStreamDestroy(c->streams[k]);
    for (int k = 0; k < 3; ++k) if (c->events[k]) cudaEventDestroy(c->events[k]);
    delete c;
    return DIE_OK;
}

extern "C" int die_gradient_forward_host(die_host_ctx_t* ctx, const die_gradient_params_t* p,
                                         int32_t H, int32_t W, int64_t M, int32_t B,
                                         const double* agents_host, const double* medium_host,
                                         double* agents_stage, double* medium_stage,
                                         double* theta, double* prev_grad, double* action, double* action_host,
                                         const uint8_t* coin, const double* noise, int32_t* sense_cells,
                                         uint64_t seed, uint64_t step, void* stream) {
    DIE_REQUIRE(ctx != nullptr && p != nullptr && B >= 1 && M >= 1 && H >= 2 && W >= 2);
    DIE_REQUIRE(agents_host != nullptr && medium_host != nullptr && agents_stage != nullptr && medium_stage != nullptr);
    DIE_REQUIRE(theta != nullptr && action != nullptr && action_host != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t C = (size_t)H * W, Ms = (size_t)M;
    const size_t obs_bytes = sizeof(double) * (4 * Ms + 3 * C) * B;
    const int nchunks = (g_host_chunks > 1 && B >= 2 * g_host_chunks && obs_bytes >= g_host_chunk_min_bytes) ? g_host_chunks : 1;
    if (nchunks > 1) DIE_CUDA(cudaEventRecord(ctx->events[2], st));
    for (int k = 0; k < nchunks; ++k) {
        const int b0 = (int)((int64_t)B * k / nchunks), b1 = (int)((int64_t)B * (k + 1) / nchunks);
        const int nb = b1 - b0;
        cudaStream_t s = (nchunks > 1) ? ctx->streams[k & 1] : st;
        if (nchunks > 1 && k < 2) DIE_CUDA(cudaStreamWaitEvent(s, ctx->events[2], 0));
        // only what the policy reads travels: x, y (agents channels 0-1) and env_food, chem1 (medium channels 1-2), each
        // contiguous per environment -- 32 instead of 56 B per slot; the other channels of the staging buffers (alive,
        // agent_food, the occupancy) are never read by the forward kernel (core/agent/gradient.py:96-124 does not either)
        DIE_CUDA(cudaMemcpy2DAsync(agents_stage + b0 * 4 * Ms, sizeof(double) * 4 * Ms, agents_host + b0 * 4 * Ms,
                                   sizeof(double) * 4 * Ms, sizeof(double) * 2 * Ms, (size_t)nb, cudaMemcpyHostToDevice, s));
        DIE_CUDA(cudaMemcpy2DAsync(medium_stage + b0 * 3 * C + C, sizeof(double) * 3 * C, medium_host + b0 * 3 * C + C,
                                   sizeof(double) * 3 * C, sizeof(double) * 2 * C, (size_t)nb, cudaMemcpyHostToDevice, s));
        if (int rc = gradient_forward_impl(nullptr, false, p, H, W, M, nb, agents_stage + b0 * 4 * Ms, medium_stage + b0 * 3 * C,
                                           theta + b0 * Ms, prev_grad ? prev_grad + b0 * 2 * Ms : nullptr, action + b0 * 3 * Ms,
                                           coin ? coin + b0 * Ms : nullptr, noise ? noise + b0 * 2 * Ms : nullptr,
                                           sense_cells ? sense_cells + b0 * Ms : nullptr, nullptr, nullptr,
                                           seed, step, (void*)s, b0))
            return rc;
        DIE_CUDA(cudaMemcpyAsync(action_host + b0 * 3 * Ms, action + b0 * 3 * Ms, sizeof(double) * 3 * Ms * nb,
                                 cudaMemcpyDeviceToHost, s));
    }
    if (nchunks > 1) {
        for (int k = 0; k < 2; ++k) {
            DIE_CUDA(cudaEventRecord(ctx->events[k], ctx->streams[k]));
            DIE_CUDA(cudaStreamWaitEvent(st, ctx->events[k], 0));
        }
    }
    DIE_CUDA(cudaStreamSynchronize(st));
    return DIE_OK;
}

// ------------------------------------------------------------------------------------------
// NeuralAutomataAgent.forward (core/agent/evo.py:117-209)
// ------------------------------------------------------------------------------------------
extern "C" int die_conv_policy_forward_population(int32_t H, int32_t W, int64_t M, int32_t B, int32_t field_dtype,
                                       const void* medium, int32_t in_ch_total, int32_t in_ch0, int32_t cin, int32_t cout_last,
                                       int32_t n_layers, const int32_t* kernel_sizes, const float* weights,
                                       int64_t weight_env_stride,
                                       float* scratch_a, float* scratch_b,
                                       const double* agents, const int32_t* cells_hint, const float* coefs,
                                       double* action, int32_t* final_scratch, void* stream) {
    DIE_REQUIRE(weight_env_stride >= 0);
    DIE_REQUIRE(H >= 1 && W >= 1 && M >= 1 && B >= 1 && (int64_t)H * W <= 0x7fffffffLL);
    DIE_REQUIRE(field_dtype == DIE_FIELD_F64 || field_dtype == DIE_FIELD_F32);
    DIE_REQUIRE(medium != nullptr && weights != nullptr && scratch_a != nullptr && scratch_b != nullptr);
    DIE_REQUIRE(kernel_sizes != nullptr && (action == nullptr || (agents != nullptr && coefs != nullptr && cout_last <= 3)));
    DIE_REQUIRE(cin >= 1 && cin <= kConvMaxCh && cout_last >= 1 && cout_last <= kConvMaxCh);
    DIE_REQUIRE(in_ch0 >= 0 && in_ch0 + cin <= in_ch_total);
    DIE_REQUIRE(n_layers >= 1 && n_layers <= 16);
    cudaStream_t st = (cudaStream_t)stream;
    const float* w = weights;
    const void* in = medium;
    for (int l = 0; l < n_layers; ++l) {
        const int k = kernel_sizes[l];
        DIE_REQUIRE(k >= 1 && k <= kConvMaxK && (k & 1) == 1);
        ConvLayerArgs a;
        memset(&a, 0, sizeof(a));
        a.in = in;
        a.out = (l & 1) ? scratch_b : scratch_a;
        a.weight = w;
        a.weight_env_stride = weight_env_stride;
        a.H = H; a.W = W; a.k = k;
        a.cin = cin;
        a.cout = (l == n_layers - 1) ? cout_last : cin;
        a.cin_total = (l == 0) ? in_ch_total : cin;
        a.in_ch0 = (l == 0) ? in_ch0 : 0;
        a.tiles_i = (H + kConvTile - 1) / kConvTile;
        a.tiles_j = (W + kConvTile - 1) / kConvTile;
        a.apply_tanh = (l == n_layers - 1) ? 1 : 0;
        const size_t smem = sizeof(float) * ((size_t)cin * (kConvTile + k - 1) * (kConvTile + k - 1) + (size_t)a.cout * cin * k * k);
        const unsigned grid = (unsigned)((int64_t)a.tiles_i * a.tiles_j * B);
        if (l == 0 && field_dtype == DIE_FIELD_F64) conv_layer_kernel<double><<<grid, 256, smem, st>>>(a);
        else conv_layer_kernel<float><<<grid, 256, smem, st>>>(a);
        DIE_CUDA(cudaGetLastError());
        w += (size_t)a.cout * cin * k * k;
        in = a.out;
    }
    const float* sense = (const float*)in;
    if (final_scratch != nullptr) *final_scratch = (n_layers & 1) ? 0 : 1;
    if (action == nullptr) return DIE_OK;                  // the model alone (ConvolutionModel.forward)
    const int64_t total = (int64_t)B * M;
    conv_gather_kernel<<<grid_for(total, 256, 148), 256, 0, st>>>(sense, agents, cells_hint, action, make_axis(H), make_axis(W),
                                                               M, cout_last, B, coefs[0], coefs[1], coefs[2]);
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}

extern "C" int die_conv_policy_forward(int32_t H, int32_t W, int64_t M, int32_t B, int32_t field_dtype,
                                       const void* medium, int32_t in_ch_total, int32_t in_ch0, int32_t cin, int32_t cout_last,
                                       int32_t n_layers, const int32_t* kernel_sizes, const float* weights,
                                       float* scratch_a, float* scratch_b,
                                       const double* agents, const int32_t* cells_hint, const float* coefs,
                                       double* action, int32_t* final_scratch, void* stream) {
    return die_conv_policy_forward_population(H, W, M, B, field_dtype, medium, in_ch_total, in_ch0, cin, cout_last, n_layers,
                                              kernel_sizes, weights, 0, scratch_a, scratch_b, agents, cells_hint, coefs,
                                              action, final_scratch, stream);
}

// ------------------------------------------------------------------------------------------
// diagnostics: die_math.h on device arrays
// ------------------------------------------------------------------------------------------
__global__ void math_sincos_kernel(const double* x, double* s, double* c, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        die_sincos(x[i], s + i, c + i);
}

__global__ void math_atan2_kernel(const double* y, const double* x, double* out, int64_t n, int fast) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = fast ? die_atan2_fast(y[i], x[i]) : die_atan2(y[i], x[i]);
}

extern "C" int die_math_sincos(const double* x, double* s, double* c, int64_t n, void* stream) {
    DIE_REQUIRE(x != nullptr && s != nullptr && c != nullptr && n >= 0);
    if (n == 0) return DIE_OK;
    math_sincos_kernel<<<grid_for(n, 256, 148), 256, 0, (cudaStream_t)stream>>>(x, s, c, n);
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}

extern "C" int die_math_atan2(const double* y, const double* x, double* out, int64_t n, int32_t fast, void* stream) {
    DIE_REQUIRE(x != nullptr && y != nullptr && out != nullptr && n >= 0);
    if (n == 0) return DIE_OK;
    math_atan2_kernel<<<grid_for(n, 256, 148), 256, 0, (cudaStream_t)stream>>>(y, x, out, n, fast);
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}

// ------------------------------------------------------------------------------------------
// one field over G GPUs: row slabs + NVLink peer access (die_slab.cuh)
// ------------------------------------------------------------------------------------------
struct die_slab {
    SlabGeom g;
    die_dynamics_t dyn;
    SlabTables tbl[2];         // [cur]: medium_in = medium[cur], medium_out = medium[1-cur]
    int64_t Ml;                // local slots
    int32_t* cells;            // [Ml] global linear cell after the move
    double* part_gain;
    int32_t* part_alive;
    double* reward_dev;        // [1]
    int64_t* alive_dev;        // [1]
    int nblk;
    // corner mirror (die_slab.cuh): local copies of the four corner_r x corner_r corner patches
    int corner_r;
    double2* corner_grad;
    double* corner_food;
    double* corner_cons;
    int corner_valid;          // food + consumed_field mirrored by a refresh since the last field pass
    int corner_grad_valid;
};

// the tables a kernel of this rank gets: mirror pointers only while the mirror is valid
static SlabTables slab_tables(const die_slab* e, int cur, bool want_mirror) {
    SlabTables t = e->tbl[cur];
    t.corner_grad = nullptr;
    t.corner_food = nullptr;
    t.corner_cons = nullptr;
    t.corner_r = e->corner_r;
    if (e->corner_r > 0 && want_mirror && e->corner_valid) {
        t.corner_food = e->corner_food;
        t.corner_cons = e->corner_cons;
        if (e->corner_grad_valid) t.corner_grad = e->corner_grad;
    }
    return t;
}

__global__ void slab_pack_stats_kernel(const double* reward, const int64_t* alive, double* out) {
    out[0] = reward[0];
    out[1] = (double)alive[0];
}

extern "C" int die_slab_create(const die_slab_geom_t* geom, const die_dynamics_t* dyn, die_slab_t** out) {
    DIE_REQUIRE(out != nullptr && geom != nullptr);
    *out = nullptr;
    DIE_REQUIRE(geom->G >= 1 && geom->G <= DIE_MAX_RANKS && geom->rank >= 0 && geom->rank < geom->G);
    DIE_REQUIRE(geom->H >= 2 && geom->W >= 2 && geom->H % geom->G == 0);
    DIE_REQUIRE((int64_t)geom->H * geom->W <= 0x7fffffffLL && geom->M >= 1 && geom->M <= 0x7fffffffLL);
    if (int rc = check_dynamics(dyn)) return rc;
    DIE_REQUIRE(dyn->blur_radius >= 1);
    DIE_REQUIRE(dyn->diffuse_mode == DIE_DIFFUSE_WRAP);
    die_slab* e = new (std::nothrow) die_slab();
    if (e == nullptr) return fail(DIE_E_NOMEM, "out of host memory");
    memset(e, 0, sizeof(*e));
    SlabGeom& g = e->g;
    g.G = geom->G; g.rank = geom->rank; g.H = geom->H; g.W = geom->W;
    g.rows_per = geom->H / geom->G;
    g.slab_cells = g.rows_per * g.W;
    g.slab_shift = -1;
    for (int sft = 0; sft < 31; ++sft) if ((1 << sft) == g.slab_cells) g.slab_shift = sft;
    g.M = geom->M;
    int64_t total = 0;
    for (int q = 0; q < geom->G; ++q) {
        g.s0[q] = geom->s0[q]; g.n0[q] = geom->n0[q]; g.s1[q] = geom->s1[q]; g.n1[q] = geom->n1[q];
        total += g.n0[q] + g.n1[q];
    }
    if (total != g.M) { delete e; return fail(DIE_E_INVALID, "slot ranges do not cover M%s%s"); }
    e->dyn = *dyn;
    e->Ml = g.n0[g.rank] + g.n1[g.rank];
    e->nblk = (int)((e->Ml + (int64_t)kAgentThreads * kFeedItems - 1) / ((int64_t)kAgentThreads * kFeedItems));
    if (e->nblk < 1) e->nblk = 1;
    cudaError_t err = cudaMalloc(&e->cells, sizeof(int32_t) * (size_t)(e->Ml > 0 ? e->Ml : 1));
    if (err == cudaSuccess) err = cudaMalloc(&e->part_gain, sizeof(double) * (size_t)e->nblk);
    if (err == cudaSuccess) err = cudaMalloc(&e->part_alive, sizeof(int32_t) * (size_t)e->nblk);
    if (err == cudaSuccess) err = cudaMalloc(&e->reward_dev, sizeof(double));
    if (err == cudaSuccess) err = cudaMalloc(&e->alive_dev, sizeof(int64_t));
    if (err != cudaSuccess) {
        die_slab_destroy(e);
        return fail(DIE_E_CUDA, "die_slab_create: %s", cudaGetErrorString(err));
    }
    *out = e;
    return DIE_OK;
}

extern "C" int die_slab_destroy(die_slab_t* e) {
    if (e == nullptr) return DIE_OK;
    cudaFree(e->cells);
    cudaFree(e->part_gain);
    cudaFree(e->part_alive);
    cudaFree(e->reward_dev);
    cudaFree(e->alive_dev);
    cudaFree(e->corner_grad);
    cudaFree(e->corner_food);
    cudaFree(e->corner_cons);
    delete e;
    return DIE_OK;
}

extern "C" int die_slab_set_corner_mirror(die_slab_t* e, int32_t r) {
    DIE_REQUIRE(e != nullptr && r >= 0 && 2 * r <= e->g.H && 2 * r <= e->g.W);
    cudaFree(e->corner_grad);
    cudaFree(e->corner_food);
    cudaFree(e->corner_cons);
    e->corner_grad = nullptr;
    e->corner_food = e->corner_cons = nullptr;
    e->corner_r = 0;
    e->corner_valid = e->corner_grad_valid = 0;
    if (r == 0) return DIE_OK;
    const size_t n = (size_t)4 * r * r;
    DIE_CUDA(cudaMalloc(&e->corner_grad, sizeof(double2) * n));
    DIE_CUDA(cudaMalloc(&e->corner_food, sizeof(double) * n));
    DIE_CUDA(cudaMalloc(&e->corner_cons, sizeof(double) * n));
    e->corner_r = r;
    return DIE_OK;
}

extern "C" int die_slab_corner_refresh(die_slab_t* e, int32_t cur, int32_t with_grad, void* stream) {
    DIE_REQUIRE(e != nullptr && (cur == 0 || cur == 1));
    if (e->corner_r == 0) return DIE_OK;
    const SlabTables t = slab_tables(e, cur, false);
    const int64_t total = (int64_t)4 * e->corner_r * e->corner_r;
    slab_corner_copy_kernel<<<grid_for(total, 256, 148), 256, 0, (cudaStream_t)stream>>>(
        e->g, t, e->corner_grad, e->corner_food, e->corner_cons, with_grad ? 1 : 0);
    DIE_CUDA(cudaGetLastError());
    e->corner_valid = 1;
    e->corner_grad_valid = with_grad ? 1 : 0;
    return DIE_OK;
}

extern "C" int die_slab_bind(die_slab_t* e, const void* med_a, const void* med_b, const void* claim,
                             const void* consumed, const void* grad, const void* action) {
    DIE_REQUIRE(e != nullptr && med_a != nullptr && med_b != nullptr && claim != nullptr);
    DIE_REQUIRE(consumed != nullptr && grad != nullptr && action != nullptr);
    for (int cur = 0; cur < 2; ++cur) {
        SlabTables& t = e->tbl[cur];
        t.medium_in = (double* const*)(cur == 0 ? med_a : med_b);
        t.medium_out = (double* const*)(cur == 0 ? med_b : med_a);
        t.claim = (int32_t* const*)claim;
        t.consumed = (double* const*)consumed;
        t.grad = (double2* const*)grad;
        t.action = (double* const*)action;
    }
    return DIE_OK;
}

extern "C" const int32_t* die_slab_cells(const die_slab_t* e) { return e ? e->cells : nullptr; }

extern "C" int die_slab_forward(die_slab_t* e, const die_gradient_params_t* p, int32_t cur,
                                const double* agents, double* theta, double* action,
                                const uint8_t* coin, int32_t hints,
                                uint64_t seed, uint64_t step, void* stream) {
    DIE_REQUIRE(e != nullptr && p != nullptr && (cur == 0 || cur == 1));
    DIE_REQUIRE(agents != nullptr && theta != nullptr && action != nullptr);
    DIE_REQUIRE(p->inertia == 0.0 && p->noise_scale == 0.0);       // no prev_grad in slab mode (Physarum defaults)
    if (e->Ml == 0) return DIE_OK;
    GradientArgs a;
    memset(&a, 0, sizeof(a));
    a.p = *p;
    a.H = e->g.H; a.W = e->g.W; a.M = e->Ml; a.nchunk = chunks_for(e->Ml, kFwdItems);
    a.ax = make_axis(e->g.H);
    a.ay = make_axis(e->g.W);
    a.plan = plan_for(p);
    a.agents = agents; a.theta = theta; a.action = action; a.coin = coin;
    a.seed = seed; a.step = step;
    if (g_sense_quick) {           // as the single-GPU forward: float32 sin / cos for the sensed cell, guarded
        a.sense_guard_x = fabs(p->sense_offset) * DIE_SINCOSF_ERR * a.ax.nm1 + 1e-9;
        a.sense_guard_y = fabs(p->sense_offset) * DIE_SINCOSF_ERR * a.ay.nm1 + 1e-9;
    }
    a.sg = e->g;
    a.st = slab_tables(e, cur, true);
    if (!(hints & 1)) {                          // bit 0: the gradient published by the last die_slab_field
        a.st.grad = nullptr;
        a.st.corner_grad = nullptr;
    }
    if (hints & 2) a.cells = e->cells;           // bit 1: the cell cache of the last die_slab_move_claim
    const unsigned grid = (unsigned)a.nchunk;
    if (p->discrete_turn)
        gradient_forward_kernel<true, true, false, 4><<<grid, kAgentThreads, 0, (cudaStream_t)stream>>>(a);
    else
        gradient_forward_kernel<false, true, false, 4><<<grid, kAgentThreads, 0, (cudaStream_t)stream>>>(a);
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}

extern "C" int die_slab_move_claim(die_slab_t* e, double* agents, const double* action, void* stream) {
    DIE_REQUIRE(e != nullptr && agents != nullptr && action != nullptr);
    if (e->Ml == 0) return DIE_OK;
    const int mchunk = chunks_for(e->Ml, kMoveItems);
    move_claim_kernel<true, false><<<(unsigned)mchunk, kAgentThreads, 0, (cudaStream_t)stream>>>(
        agents, action, nullptr, e->cells, make_axis(e->g.H), make_axis(e->g.W), e->Ml, mchunk,
        e->dyn.boundary, nullptr, 0, e->g, slab_tables(e, 0, false));
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}

template <int R>
static cudaError_t launch_field_slab(const FieldArgs& fa, bool grad, cudaStream_t st) {
    constexpr int TH = 32, TW = 64, NT = 256;
    FieldArgs a = fa;
    a.tiles_i = (a.sg.rows_per + TH - 1) / TH;
    a.tiles_j = (a.W + TW - 1) / TW;
    const int G = grad ? 1 : 0;
    const size_t smem = sizeof(double) * (size_t)((TH + 2 * G + 2 * R) * (TW + 2 * G + 2 * R) +
                                                  (TH + 2 * G) * (TW + 2 * G + 2 * R));
    cudaError_t err;
    const unsigned grid = (unsigned)(a.tiles_i * a.tiles_j);
    if (grad) {
        auto kern = field_step_kernel<R, TH, TW, NT, true, true>;
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kern<<<grid, NT, smem, st>>>(a);
    } else {
        auto kern = field_step_kernel<R, TH, TW, NT, false, true>;
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        kern<<<grid, NT, smem, st>>>(a);
    }
    return cudaGetLastError();
}

extern "C" int die_slab_field(die_slab_t* e, int32_t cur, int32_t publish_grad, void* stream) {
    DIE_REQUIRE(e != nullptr && (cur == 0 || cur == 1));
    FieldArgs a;
    memset(&a, 0, sizeof(a));
    a.H = e->g.H;
    a.W = e->g.W;
    a.rate_feed = e->dyn.rate_feed;
    a.keep = 1.0 - e->dyn.rate_decay_chem;
    a.food_infinite = e->dyn.food_infinite;
    for (int k = 0; k < 2 * DIE_MAX_RADIUS + 1; ++k) a.bw.w[k] = e->dyn.blur_w[k];
    a.sg = e->g;
    a.st = slab_tables(e, cur, false);
    e->corner_valid = e->corner_grad_valid = 0;  // the mirror describes the previous step until die_slab_corner_refresh
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t err = cudaErrorInvalidValue;
    switch (e->dyn.blur_radius) {
        case 1: err = launch_field_slab<1>(a, publish_grad != 0, st); break;
        case 2: err = launch_field_slab<2>(a, publish_grad != 0, st); break;
        case 3: err = launch_field_slab<3>(a, publish_grad != 0, st); break;
        case 4: err = launch_field_slab<4>(a, publish_grad != 0, st); break;
        default: return fail(DIE_E_INVALID, "slab mode supports blur radius 1..4%s%s");
    }
    DIE_CUDA(err);
    return DIE_OK;
}

extern "C" int die_slab_feed(die_slab_t* e, double* agents, const double* action, double* stats, void* stream) {
    DIE_REQUIRE(e != nullptr && agents != nullptr && action != nullptr && stats != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    if (e->Ml > 0) {
        FeedArgs fa;
        memset(&fa, 0, sizeof(fa));
        fa.agents = agents; fa.action = action; fa.cells = e->cells;
        fa.part_gain = e->part_gain; fa.part_alive = e->part_alive;
        fa.C = (int64_t)e->g.slab_cells; fa.M = e->Ml; fa.nblk = e->nblk;
        fa.w_dep = e->dyn.cost_w_deposit; fa.w_dist = e->dyn.cost_w_dist;
        fa.boundary = e->dyn.boundary;
        agent_feed_kernel<true, false, false><<<(unsigned)e->nblk, kAgentThreads, 0, st>>>(fa, e->g, slab_tables(e, 0, true));
        DIE_CUDA(cudaGetLastError());
        finalize_stats_kernel<<<1, kFinalThreads, 0, st>>>(e->part_gain, e->part_alive, e->nblk, e->reward_dev, e->alive_dev);
    } else {
        DIE_CUDA(cudaMemsetAsync(e->reward_dev, 0, sizeof(double), st));
        DIE_CUDA(cudaMemsetAsync(e->alive_dev, 0, sizeof(int64_t), st));
    }
    DIE_CUDA(cudaGetLastError());
    slab_pack_stats_kernel<<<1, 1, 0, st>>>(e->reward_dev, e->alive_dev, stats);
    DIE_CUDA(cudaGetLastError());
    return DIE_OK;
}
