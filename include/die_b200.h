/* die_b200.h -- C ABI of libdie_sm100a.so: the per-step hot path of gkirgizov/die
 * (Env.step, BrownianAgent.forward, GradientAgent/PhysarumAgent.forward) as hand-written
 * sm_100a CUDA kernels.
 *
 * The reference has no FFI: its "operator API" for this path is the Python protocol
 *     Env.step(action)            /root/reference/core/env.py:101-131
 *     Agent.forward(obs)          /root/reference/core/agent/base.py:13-16
 * so each entry point below names the reference method(s) it replaces.  The binding a
 * maintainer of the reference would add (a ctypes stub) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every function returns DIE_OK (0) or a DIE_E_* code and never throws;
 *     die_last_error() gives the message of the calling thread's last failure;
 *   - "dev" pointers are device memory owned by the caller (borrowed for the call, stream-
 *     ordered); "host" pointers are host memory (pinned for full copy speed);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - everything is float64, channel-major, exactly the reference's layouts
 *     (core/base_types.py:31-36, core/data_init.py:94-130):
 *         medium [B][3][H][W]  channels (agents, env_food, chem1)
 *         agents [B][4][M]     channels (x, y, alive, agent_food)
 *         action [B][3][M]     channels (dx, dy, deposit1)
 *     B = number of independent environments in the batch (1 = the reference's case).
 */
#ifndef DIE_B200_H
#define DIE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIE_OK          0
#define DIE_E_INVALID   1   /* bad argument */
#define DIE_E_CUDA      2   /* a CUDA runtime call failed; see die_last_error() */
#define DIE_E_NOMEM     3

#define DIE_MAX_RADIUS  8   /* blur radius int(4*sigma+.5) <= 8  <=>  sigma <= 2.1 */

/* Dynamics.diffuse_mode: the `mode` skimage.filters.gaussian hands to scipy.ndimage (core/env.py:140-143) */
#define DIE_DIFFUSE_WRAP      0   /* the reference's default */
#define DIE_DIFFUSE_REFLECT   1
#define DIE_DIFFUSE_NEAREST   2
#define DIE_DIFFUSE_MIRROR    3
#define DIE_DIFFUSE_CONSTANT  4   /* cval = 0 */

#define DIE_BOUNDARY_WRAP   0   /* BoundaryCondition.wrap  core/env.py:154-155 */
#define DIE_BOUNDARY_LIMIT  1   /* BoundaryCondition.limit core/env.py:156-157 */
#define DIE_BOUNDARY_NONE   2   /* unknown enum: warn + leave as is, core/env.py:158-161 */

/* Dynamics, core/env.py:42-61, as a POD.  op_action_cost is restricted to the reference's
 * two operators: linear_action_cost (weights 0.02, 0.01; core/env.py:29-35) and zero_cost
 * (both weights 0; core/env.py:38-39).  op_food_flow is identity (core/env.py:45).
 * blur_w holds the 2*blur_radius+1 weights scipy.ndimage.gaussian_filter1d builds for
 * diffuse_sigma (w[k] = exp(-.5/sigma^2 k^2)/sum, k=-r..r). */
typedef struct die_dynamics {
    double  rate_feed;
    double  rate_decay_chem;
    double  cost_w_deposit;
    double  cost_w_dist;
    double  blur_w[2 * DIE_MAX_RADIUS + 1];
    int32_t blur_radius;
    int32_t boundary;
    int32_t food_infinite;
    int32_t diffuse_mode;     /* DIE_DIFFUSE_*; 0 = 'wrap' */
    int32_t agents_die;       /* Dynamics.agents_die, core/env.py:245-250: after feeding, every channel of a slot whose
                               * agent_food is not above 1e-4 becomes 0 (the lifecycle the reference intends: its own
                               * version loses track of the array it rebinds, core/utils.py:22) */
} die_dynamics_t;

/* GradientAgent / PhysarumAgent constructor state, core/agent/gradient.py:19-45,139-163. */
typedef struct die_gradient_params {
    double  scale;
    double  deposit;
    double  inertia;
    double  sense_offset;
    double  noise_scale;
    double  grad_clip;        /* used iff use_grad_clip */
    double  turn_radians;     /* np.radians(turn_angle)   (Physarum) */
    double  sense_radians;    /* np.radians(sense_angle)  (Physarum) */
    double  turn_tolerance;   /* rtol                     (Physarum) */
    int32_t normalized_grad;
    int32_t use_grad_clip;    /* grad_clip is not None */
    int32_t discrete_turn;    /* 1 = PhysarumAgent._process_gradient, 0 = GradientAgent */
    int32_t reserved;
} die_gradient_params_t;

typedef struct die_env die_env_t;

const char* die_version(void);
const char* die_last_error(void);

/* Field dtype.  The reference computes in float64 throughout (numpy defaults) and that is the default here; in the
 * float32 field mode the arrays that hold FIELDS in HBM -- medium A/B and the consumed_field scratch -- are float32
 * (24 instead of 48 B per cell-update), agents / actions / headings stay float64 (cell indices stay bit-exact for a given
 * action), values are widened on load and every operation is the float64 one, rounded once on store.  Call right after
 * die_env_create, before the first step; every `medium` pointer of the env's entry points is then a float array.
 * Not available in this mode: the host-buffer step, the speculative move, the cluster-fused step, blur radius > 4. */
#define DIE_FIELD_F64 0
#define DIE_FIELD_F32 1
int die_env_set_field_dtype(die_env_t* env, int32_t dtype);
int die_env_field_dtype(const die_env_t* env);

/* Env.__init__ workspace (core/env.py:65-72): per-cell claim table, per-slot cell cache,
 * reduction partials.  Must be called with the target device current. */
int die_env_create(int32_t H, int32_t W, int64_t M, int32_t B,
                   const die_dynamics_t* dyn, die_env_t** out);
int die_env_destroy(die_env_t* env);
int die_env_set_dynamics(die_env_t* env, const die_dynamics_t* dyn);

/* Dynamics.op_food_flow = WaveSequence(field_size, dt, t_bounds).get_flow_operator(scale, decay)
 * (core/data_init.py:16-47, 71-89; used by examples/simple_agents.py:95-100): every step
 *     food = scale * F_t + (1 - decay) * food,   t cycling over np.arange(*t_bounds, dt),
 *     F_t = 0.75 cos(pi (rwave + t)) + 0.25 (sin(pi x 3 + t) + cos(pi y 3 + t)),  rwave = r + cos(pi x) + sin(0.4 pi y)
 * evaluated inside the field pass.  The caller tabulates (with numpy, i.e. the reference's own arithmetic)
 * rwave_dev[H*W], col_dev[T][W] = sin(pi x 3 + t_k), row_dev[T][H] = cos(pi y 3 + t_k) and ts_host[T]; the
 * device tables are borrowed until the flow is reset or the env destroyed.  The per-cell cosine is
 * die_math.h's.  k0 = index of the first time step to use.  rwave_dev == NULL restores the identity flow. */
int die_env_set_food_flow(die_env_t* env, const double* rwave_dev, const double* col_dev, const double* row_dev,
                          const double* ts_host, int64_t T, int64_t k0, double scale, double decay);
/* The same operator for ANY FieldSequence (core/data_init.py:16-51: the reference's PerlinNoiseSequence :54-68, a
 * user's own subclass): frames_dev [T][H*W] holds sequence[t_k] for every time step, tabulated by the host with the
 * sequence's own code; the field pass reads one value per cell,  food = scale * F_k[cell] + (1 - decay) * food,  k
 * advancing by one per step and cycling, starting at k0.  Shared by all environments of a batch; borrowed until
 * replaced (either setter with a null table restores the identity flow). */
int die_env_set_food_frames(die_env_t* env, const double* frames_dev, int64_t T, int64_t k0, double scale, double decay);

/* Env.step, core/env.py:101-131: move -> deposit + layout -> feed -> food flow ->
 * diffuse*decay -> reward / num_agents.  Reads medium_in (never written: the previous
 * observation stays valid), writes the next medium into medium_out (a different buffer),
 * updates agents in place.  reward_dev[B] = sum over ALL M slots of gained; alive_dev[B] =
 * #(alive > 0).  die_env_cells() exposes the int32 [B][M] linear cell index ix*W+iy of
 * every slot after the move (validation aid). */
int die_env_step(die_env_t* env, double* medium_in_dev, double* medium_out_dev,
                 double* agents_dev, const double* action_dev,
                 double* reward_dev, int64_t* alive_dev, void* stream);
const int32_t* die_env_cells(const die_env_t* env);   /* device ptr, int32 [B][M], valid after a step */

/* Derived fields the environment can publish for the agents' forward pass (optional fast path).
 * die_env_publish_gradient(env, 1): from the next step on, the field pass also writes
 * np.gradient of the NEW chem1 channel (raw d/dx, d/dy pairs, core/agent/gradient.py:57) into an
 * internal [B][H*W][2] buffer; die_env_gradient() returns it (NULL when not published).  Passing it
 * (and die_env_cells()) to die_gradient_forward replaces four 8-byte gathers and two coordinate
 * look-ups per slot by one 16-byte gather and one 4-byte read.  The hints are valid only for the
 * medium / agents written by the last die_env_step and until they are modified by the caller. */
int die_env_publish_gradient(die_env_t* env, int32_t on);
const double* die_env_gradient(const die_env_t* env);
/* When the agent that last acted through this env (die_env_forward_gradient) only thresholds the gradient -- a
 * PhysarumAgent on normalised gradients whose turn rule admits the guard-banded float32 decision -- the pairs are
 * published rounded to float32 instead (8 bytes per cell written and gathered; die_set_tuning("grad_f32", 0) turns
 * this off): that decision rounds the gradient to float32 as its first operation anyway and re-samples chem1 in
 * float64 whenever it defers to the reference arithmetic (die_b200/csrc/die_turn.h), so results do not change.  A
 * policy that uses the gradient's VALUE (GradientAgent, unnormalised gradients) gets float64 pairs.
 * die_env_gradient_kind: what the last step published -- 0 nothing, 1 float64 (die_env_gradient()), 2 float32
 * (internal; reached through die_env_forward_gradient's DIE_FWD_USE_GRADIENT). */
int die_env_gradient_kind(const die_env_t* env);

/* Per-kernel timing of Env.step with CUDA events recorded on the launching stream between
 * the step's kernels (measurement aid for bench.py's roofline; off by default).
 * die_env_kernel_times synchronises the recorded events and returns the accumulated
 * milliseconds of {move_claim, field_step, agent_feed, finalize_stats} and the number of
 * profiled steps since profiling was (re-)enabled; at most DIE_MAX_PROFILED_STEPS are kept. */
#define DIE_NUM_STEP_KERNELS     4
#define DIE_MAX_PROFILED_STEPS   2048
int die_env_set_profiling(die_env_t* env, int32_t on);
int die_env_kernel_times(die_env_t* env, double* ms_out /*[DIE_NUM_STEP_KERNELS]*/, int64_t* steps_out);

/* The read-back that makes Env.step's return value (core/env.py:117-131): reward_dev[B] / alive_dev[B] of the last
 * die_env_step -> host (pinned for full speed), then synchronises `stream`.  16 bytes per environment. */
int die_env_read_stats(die_env_t* env, const double* reward_dev, const int64_t* alive_dev,
                       double* reward_host, int64_t* alive_host, void* stream);

/* Env.step through HOST buffers: H2D of action_host[B][3][M], the step, D2H of the new
 * observation (agents_host[B][4][M], medium_host[B][3][H][W]; either may be NULL to skip)
 * and of reward_host[B] / alive_host[B]; synchronises `stream` before returning.  Large batches are
 * processed in chunks of environments on two internal streams, so that one chunk's D2H overlaps the next
 * chunk's H2D and kernels (die_set_tuning("host_chunks", n); 1 = one stream). */
int die_env_step_host(die_env_t* env, double* medium_in_dev, double* medium_out_dev,
                      double* agents_dev, const double* action_host,
                      double* agents_host, double* medium_host,
                      double* reward_host, int64_t* alive_host, void* stream);
/* The same step when the action is ALREADY on the device -- the host array handed to Env.step is the very (read-only)
 * array Agent.forward's host path returned and its device copy is still valid: no H2D of the action. */
int die_env_step_host_dev(die_env_t* env, double* medium_in_dev, double* medium_out_dev,
                          double* agents_dev, const double* action_dev,
                          double* agents_host, double* medium_host,
                          double* reward_host, int64_t* alive_host, void* stream);

/* Both of the above with options (exactly one of action_host / action_dev is non-NULL):
 *   DIE_HOST_KEEP_ALIVE_CHANNEL  agents_host already holds the current `alive` channel (a full download of this env's
 *                                agents went into that very buffer and neither the caller nor a lifecycle changed the
 *                                channel since -- Env.step never does, core/env.py:152-243): only x, y and agent_food
 *                                are copied, 24 instead of 32 B per slot.  Refused with Dynamics.agents_die. */
#define DIE_HOST_KEEP_ALIVE_CHANNEL 1
int die_env_step_host_flags(die_env_t* env, double* medium_in_dev, double* medium_out_dev,
                            double* agents_dev, const double* action_host, const double* action_dev,
                            double* agents_host, double* medium_host,
                            double* reward_host, int64_t* alive_host, int32_t flags, void* stream);

/* Env._get_sensed_medium with Dynamics.apply_sense_mask (core/env.py:275-294): the observation's medium is
 *     medium.where(ceil(round(gaussian(medium['agents'], sigma=2.0), 3)), other=0.)
 * i.e. every channel zeroed outside the blurred neighbourhood of the agents.  weights_host[2*radius+1] = the
 * scipy.ndimage weights of that gaussian (sigma 2 -> radius 8); borders are clamped (skimage's default
 * mode='nearest').  medium_dev [B][3][H][W] -> obs_dev [B][3][H][W] (a different buffer). */
int die_sense_mask(int32_t H, int32_t W, int32_t B, const double* weights_host, int32_t radius,
                   const double* medium_dev, double* obs_dev, void* stream);

/* EnvRenderer.render + FieldTrace.update (core/render.py:9-29, 76-132): the frames of one step, on the device.
 *   img_medium_dev [B][H][W][3] = (agents, env_food, chem1) per pixel, or cross(color, rgb) when color_host[3]
 *                                 (normalised) is given (RendererBase._set_colors);
 *   trace_dev      [B][H][W]    = trace * decay + occupancy, in place (decay = 1 - 1/trace_steps); the caller
 *                                 colour-maps it (matplotlib's LUT is not part of this library);
 *   img_agents_dev [B][M][4]    = (0, agent_food, 0, alive != 0) per slot -- with M == H*W this IS the reference's
 *                                 (2, height, -1) -> transpose(1, 2, 0) image [W][H][4]; NULL to skip. */
int die_render_frames(int32_t H, int32_t W, int64_t M, int32_t B,
                      const double* medium_dev, const double* agents_dev, double* trace_dev, double decay,
                      const double* color_host, double* img_medium_dev, double* img_agents_dev, void* stream);

/* BrownianAgent.forward, core/agent/static.py:40-50 (+ core/data_init.py:159-169,218-220,
 * 248-253).  u_dev[B][3][M] = the three uniform draws in the reference's order
 * (dx, dy, deposit1); NULL = draw in-kernel (Philox4x32-10 keyed on seed, step, slot). */
int die_brownian_forward(const double* agents_dev, double* action_dev, int64_t M, int32_t B,
                         double move_scale, double deposit_scale,
                         const double* u_dev, uint64_t seed, uint64_t step, void* stream);
/* The same with in-kernel draws only and the call counter read from device memory (see DIE_FWD_STEP_ON_DEVICE). */
int die_brownian_forward_dev(const double* agents_dev, double* action_dev, int64_t M, int32_t B,
                             double move_scale, double deposit_scale, uint64_t seed,
                             const uint64_t* step_dev, void* stream);

/* ConstAgent.forward, core/agent/static.py:19-28 (not alive-masked). */
int die_const_forward(double* action_dev, int64_t M, int32_t B,
                      double dx, double dy, double deposit, void* stream);

/* JonesAgent.forward: the classic three-sensor Physarum particle (Jones 2010, "Characteristics of pattern formation and
 * evolution in approximations of Physarum transport networks") on the reference's Env protocol -- SURVEY.md 8(f) rank 4,
 * "optional ... not parity-checkable": the reference has no such class, so the specification is oracle/die_ref.py:JonesAgent
 * (sensors FL / F / FR at heading +sense, 0, -sense, `sense_offset` away, reading chem1 at the nearest cell, clamped as
 * core/utils.py:39-54 does; F largest: straight on; F smallest: +-turn by a coin; else towards the larger side;
 * heading' = renormalize_radians(heading + turn), core/utils.py:177-179; action = (scale cos, scale sin, deposit * food
 * under the agent), every slot, as core/agent/gradient.py:113-124).
 * medium_f32: the medium holds float32 elements (the env's float32 field mode).  coin_dev: [B][M] in {0,1}, or NULL for
 * in-kernel Philox draws keyed on (seed, step, env, slot).  step_on_device: `step` is the DEVICE ADDRESS of the uint64 call
 * counter (CUDA-graph replays, as DIE_FWD_STEP_ON_DEVICE). */
typedef struct die_jones_params {
    double scale, deposit, sense_offset, sense_radians, turn_radians;
} die_jones_params_t;
int die_jones_forward(const die_jones_params_t* p, int32_t H, int32_t W, int64_t M, int32_t B,
                      const double* agents_dev, const void* medium_dev, int32_t medium_f32,
                      double* theta_dev, double* action_dev, const uint8_t* coin_dev,
                      uint64_t seed, uint64_t step, int32_t step_on_device, void* stream);

/* GradientAgent.forward / PhysarumAgent.forward, core/agent/gradient.py:96-124 (+ :55-91,
 * :168-219).  theta_dev[B][M] (in/out) = _direction_rads.  prev_grad_dev[B][2][M] (in/out)
 * = _prev_grad; may be NULL when inertia == 0 and noise_scale == 0 (its value then cannot
 * reach any output).  coin_dev[B][M] uint8 in {0,1} = np.random.randint(0, 2, M); NULL =
 * Philox.  noise_dev[B][2][M] = rng.normal(0, .4, (2, M)); NULL = Philox Box-Muller
 * (skipped entirely when noise_scale == 0).  sense_cells_dev (optional) int32 [B][M]:
 * linear index of the cell each slot sensed (validation aid).  grad_hint_dev / cells_hint_dev
 * (optional): die_env_gradient() / die_env_cells() of the environment that produced the
 * observation -- same results bit for bit, fewer gathers. */
int die_gradient_forward(const die_gradient_params_t* p,
                         int32_t H, int32_t W, int64_t M, int32_t B,
                         const double* agents_dev, const double* medium_dev,
                         double* theta_dev, double* prev_grad_dev, double* action_dev,
                         const uint8_t* coin_dev, const double* noise_dev,
                         int32_t* sense_cells_dev,
                         const double* grad_hint_dev, const int32_t* cells_hint_dev,
                         uint64_t seed, uint64_t step, void* stream);

/* GradientAgent.forward / PhysarumAgent.forward through HOST buffers: H2D of the observation (agents_host[B][4][M],
 * medium_host[B][3][H][W]) into the caller's device staging buffers -- of the channels the policy reads, x, y and
 * env_food, chem1 (32 of the observation's 56 B per slot; the rest of the staging buffers is not written) --, the forward
 * kernel, D2H of the action into action_host[B][3][M]; synchronises `stream` before returning.  Large batches are cut into chunks of environments that
 * alternate between the two streams of a die_host_ctx_t, so one chunk's action download overlaps the next chunk's
 * observation upload (both PCIe directions busy).  In-kernel random draws do not depend on the chunking. */
typedef struct die_host_ctx die_host_ctx_t;
int die_host_ctx_create(die_host_ctx_t** out);
int die_host_ctx_destroy(die_host_ctx_t* ctx);
int die_gradient_forward_host(die_host_ctx_t* ctx, const die_gradient_params_t* p,
                              int32_t H, int32_t W, int64_t M, int32_t B,
                              const double* agents_host, const double* medium_host,
                              double* agents_stage_dev, double* medium_stage_dev,
                              double* theta_dev, double* prev_grad_dev, double* action_dev, double* action_host,
                              const uint8_t* coin_dev, const double* noise_dev, int32_t* sense_cells_dev,
                              uint64_t seed, uint64_t step, void* stream);

/* The same forward pass bound to the environment that produced the observation (what
 * die_b200.PhysarumAgent.forward calls when `obs` is provably that Env's own, die_b200/_hints.py).
 * flags:
 *   DIE_FWD_USE_GRADIENT    the gradient published by the env's last step is valid for `medium_dev`
 *   DIE_FWD_USE_CELLS       the env's cell cache is valid for `agents_dev`
 *   DIE_FWD_SPECULATE_MOVE  additionally evaluate Env._agent_move + cell resolution + the claim
 *                           (core/env.py:152-172, :211) for the action being written: post-move cells go
 *                           to the env's second cell buffer, claims into its claim table; positions are NOT
 *                           touched.  If the very next step receives this action unmodified
 *                           (the caller's responsibility: die_b200.Env checks the tensor identity and torch's
 *                           version counters of action and agents), die_env_step_flags(DIE_STEP_ADOPT_MOVE)
 *                           skips the move+claim kernel and
 *                           the feed kernel commits the positions -- same results bit for bit, one launch
 *                           and ~32 B per slot less.  Any other continuation (die_env_step, another
 *                           forward) first calls die_env_discard_move, which empties the claim table.
 *   DIE_FWD_COMMIT_MOVE     (with DIE_FWD_SPECULATE_MOVE) the run-loop contract: the caller PROMISES that the next call on
 *                           this env is die_env_step_flags(DIE_STEP_ADOPT_MOVE) with exactly this action
 *                           (`for ...: action = agent.forward(obs); obs, ... = env.step(action)`,
 *                           examples/minimal_run.py:21-25).  The forward launch then also stores the moved
 *                           positions into `agents_dev` (which it therefore WRITES: x, y of every slot), the
 *                           adopting step runs the field pass and the plain feed kernel only: no move+claim launch
 *                           and none of its 56 B per slot.  Same results bit for bit.  A committed move cannot be
 *                           withdrawn: until the adopting step any other step / forward on the env fails with
 *                           DIE_E_INVALID; die_env_discard_move (reset paths) only forgets it.
 * Needs die_env_refresh_alive() after every change of the `alive` channel (the speculative claim and
 * the fused feed read alive-ness from a bitmask). */
#define DIE_FWD_USE_GRADIENT    1
#define DIE_FWD_USE_CELLS       2
#define DIE_FWD_SPECULATE_MOVE  4
/* die_gradient_forward on a float32 medium (an observation that did not come from an env, or whose hints are off). */
int die_gradient_forward_f32(const die_gradient_params_t* p, int32_t H, int32_t W, int64_t M, int32_t B,
                             const double* agents_dev, const float* medium_dev,
                             double* theta_dev, double* prev_grad_dev, double* action_dev,
                             const uint8_t* coin_dev, const double* noise_dev, int32_t* sense_cells_dev,
                             uint64_t seed, uint64_t step, void* stream);

/*   DIE_FWD_STEP_ON_DEVICE  `step` is not the call counter but the DEVICE ADDRESS of a uint64 holding it: a CUDA graph
 *                           that captured this call can be replayed while something else (a captured increment) advances
 *                           the counter, so every replay draws fresh in-kernel random numbers (die_b200/graph.py) */
#define DIE_FWD_STEP_ON_DEVICE  8
#define DIE_FWD_COMMIT_MOVE     16
/*   DIE_FWD_WRITE_COST      cost hint: the launch also stores linear_action_cost (core/env.py:29-35, with the env's weights) of
 *                           the action it writes, 8 B per slot, into a buffer of the env.  If the very next step receives
 *                           this action unmodified (the caller's responsibility, as above) and passes DIE_STEP_USE_COST,
 *                           the feed kernel reads that value instead of dx, dy, deposit (24 B per slot): same bits.  Any
 *                           other step, a host-buffer step or die_env_set_dynamics drops the hint. */
#define DIE_FWD_WRITE_COST      32
int die_env_forward_gradient(die_env_t* env, const die_gradient_params_t* p,
                             const double* agents_dev, const double* medium_dev,
                             double* theta_dev, double* prev_grad_dev, double* action_dev,
                             const uint8_t* coin_dev, const double* noise_dev, int32_t* sense_cells_dev,
                             int32_t flags, uint64_t seed, uint64_t step, void* stream);
/* Env.step (as die_env_step) with result-neutral options:
 *   DIE_STEP_ADOPT_MOVE  adopt the pending speculative move (DIE_E_INVALID if none is pending);
 *   DIE_STEP_ALIVE_BITS  the bitmask built by die_env_refresh_alive is valid for `agents_dev`: the move
 *                        and feed kernels read alive-ness from it (1 bit instead of 8 bytes per slot). */
#define DIE_STEP_ADOPT_MOVE  1
#define DIE_STEP_ALIVE_BITS  2
/*   DIE_STEP_USE_COST    `action_dev` is exactly what the last die_env_forward_gradient(DIE_FWD_WRITE_COST) on this env wrote:
 *                        the feed kernel may take the action cost from the hint (ignored when no hint is pending). */
#define DIE_STEP_USE_COST    4
int die_env_step_flags(die_env_t* env, double* medium_in_dev, double* medium_out_dev,
                       double* agents_dev, const double* action_dev,
                       double* reward_dev, int64_t* alive_dev, int32_t flags, void* stream);
int die_env_discard_move(die_env_t* env, void* stream);
int die_env_pending_move(const die_env_t* env);          /* 1 while a speculative move waits for its step, 2: a committed one */
int die_env_refresh_alive(die_env_t* env, const double* agents_dev, void* stream);

/* PhysarumAgent._choose_turn has two implementations (die_b200/csrc/die_turn.h): the reference's own
 * arithmetic, and a guard-banded float32 shortcut that defers to it whenever a threshold is close.
 * die_set_turn_quick(0) forces the former for every slot (A-B tests; results are identical). */
int die_set_turn_quick(int32_t on);

/* Performance switches that never change results (A-B timing, bench.py --tune): "turn_quick" 0/1,
 * "fwd_min_blocks" 3/4/5 (register cap of the forward kernel: resident CTAs per SM), "host_chunks" n (see die_env_step_host), "fwd_lean" 0/1 (compile-time specialised forward kernel for the
 * steady-state Physarum configuration), "feed_bits" 0/1 (feed kernel takes
 * alive-ness from the bitmask), "field_prefetch" 0/1, "field_vec" 0/1 (the 128-bit field pass where it applies), "pair_mode" 0/1/2 (never / by size / always) and "pair_min_cells_log2" (default 23): {consumed_field, food} pairs + per-slot food hand-over for large fields, "step_impl" 0/1 (see die_set_step_impl), "grad_f32" 0/1 (see die_env_gradient_kind). */
int die_set_tuning(const char* key, int32_t value);
/* cudaLimitMaxL2FetchGranularity of the current device: bytes = 32 / 64 / 128 sets it (0 = only query); *previous gets the
 * old value.  A device-wide hint: how many bytes an L2 miss fetches from DRAM.  Matters for the random 8-byte gathers of
 * one LARGE field once the ghost slots have spread over it; result-neutral. */
int die_device_l2_fetch_granularity(int32_t bytes, int32_t* previous);
/* How often a kernel variant has been launched by this process (diagnostics for tests: "the variant I selected is the
 * one that ran"): "field_tile", "field_vec", "step_fused", "forward_lean", "forward_lean_f32", "forward_general", "forward_food_here";
 * -1 for an unknown key. */
int64_t die_get_counter(const char* key);

/* ---------------------------------------------------------------------------------------------
 * One field split into row slabs over G GPUs (BASELINE configs[4]; SURVEY section 8e, mode 2).
 * Rank r owns rows [r*H/G, (r+1)*H/G) of every per-cell array and the agent slots given by two
 * global ranges (s0/n0: its slab's alive agents in the reference's row-major order; s1/n1: its share of
 * the ghost slots).  Per-cell arrays and the action array are allocated by the caller as SYMMETRIC
 * memory (e.g. torch.distributed._symmetric_memory) and passed as device tables of G peer base
 * pointers; the kernels read / atomically update remote cells directly over NVLink.  The caller
 * places a cross-rank barrier after die_slab_move_claim, after die_slab_field and after die_slab_feed.
 * Results are identical to a single-GPU Env on the concatenated state.
 *   medium tables: each peer buffer [3][H/G][W];  claim [H/G*W] int32 (init -1);  consumed [H/G*W];
 *   grad [H/G*W][2];  action [3][n0[q]+n1[q]];  agents_local [4][Ml], theta_local [Ml]. */
#define DIE_MAX_RANKS 8
typedef struct die_slab_geom {
    int32_t G, rank, H, W;
    int64_t M;                                  /* global slot count */
    int64_t s0[DIE_MAX_RANKS], n0[DIE_MAX_RANKS], s1[DIE_MAX_RANKS], n1[DIE_MAX_RANKS];
} die_slab_geom_t;
typedef struct die_slab die_slab_t;

int die_slab_create(const die_slab_geom_t* geom, const die_dynamics_t* dyn, die_slab_t** out);
int die_slab_destroy(die_slab_t* slab);
int die_slab_bind(die_slab_t* slab, const void* medium_a_tbl_dev, const void* medium_b_tbl_dev,
                  const void* claim_tbl_dev, const void* consumed_tbl_dev, const void* grad_tbl_dev,
                  const void* action_tbl_dev);
/* GradientAgent/PhysarumAgent.forward for this rank's slots; cur = index (0/1) of the current medium. */
int die_slab_forward(die_slab_t* slab, const die_gradient_params_t* p, int32_t cur,
                     const double* agents_local_dev, double* theta_local_dev, double* action_local_dev,
                     const uint8_t* coin_local_dev, int32_t hints /* bit0: published grad, bit1: cell cache */,
                     uint64_t seed, uint64_t step, void* stream);
int die_slab_move_claim(die_slab_t* slab, double* agents_local_dev, const double* action_local_dev, void* stream);
int die_slab_field(die_slab_t* slab, int32_t cur, int32_t publish_grad, void* stream);
/* stats_dev[0] = this rank's sum of gained, stats_dev[1] = its alive count (as double); the caller
 * all-reduces them (the only collective of the step). */
int die_slab_feed(die_slab_t* slab, double* agents_local_dev, const double* action_local_dev,
                  double* stats_dev, void* stream);
const int32_t* die_slab_cells(const die_slab_t* slab);    /* int32 [Ml] GLOBAL linear cell of every local slot */
/* Corner mirror: local copies of the four r x r corner patches of the published gradient, the current env_food
 * and consumed_field.  The reference creates ~90 % of its slots as ghosts at (0, 0); positions wrap, so for
 * thousands of steps they stay within a few hundred cells of the four corners and every rank's ghost gathers
 * would otherwise cross NVLink into ranks 0 and G-1.  die_slab_corner_refresh (same `cur` as the die_slab_field
 * before it) pulls the patches from their owners; call it after the barrier that follows die_slab_field and
 * before die_slab_feed.  Gathers outside the patches still go to the owner, so `r` is a performance parameter
 * only; results do not change. */
int die_slab_set_corner_mirror(die_slab_t* slab, int32_t r);
int die_slab_corner_refresh(die_slab_t* slab, int32_t cur, int32_t with_grad, void* stream);

/* Env.step implementation switch (tests / A-B timing): 0 (default) = the three kernels move_claim / field_step /
 * agent_feed; 1 = the cluster-fused environment step wherever it applies (small periodic environments: one thread-block
 * cluster per environment, claim table in distributed shared memory, field rows by TMA bulk copies;
 * die_b200/csrc/die_env_fused.cuh).  Both give bit-identical results.  Measured on a B200 (round 2): the fused step
 * moves 26 % fewer DRAM bytes but is 2.2x slower (latency / issue bound), so it is an opt-in.
 * (CTAs of 512 threads; 256 threads with twice the registers measured 45 % slower.) */
int die_set_step_impl(int32_t impl);

/* NeuralAutomataAgent.forward, core/agent/evo.py:117-209 (+ ConvolutionModel, :45-118): a stack of n_layers small circular
 * convolutions (Conv2d, padding 'same', padding_mode 'circular', no bias) over `cin` channels of the medium starting at channel
 * in_ch0 (all three, or env_food + chem1), in float32 as the reference casts them, Tanh after the last layer; every slot's
 * action = model output at the agent's nearest cell (all M slots, core/utils.py:56-65) times coefs_host[3] = (scale, scale,
 * deposit).  kernel_sizes_host[n_layers] (odd, <= 7); weights_dev = the layers' weights back to back, each [cout][cin][k][k]
 * float32 (torch's layout; cout = cin except cout_last for the last layer); scratch_a / scratch_b: float [B][cin or cout_last]
 * [H][W] work buffers -- *final_scratch tells which of them (0 / 1) holds the model output afterwards (what render() shows);
 * cells_hint_dev: die_env_cells of the env that produced the observation, or NULL to resolve cells from agents_dev.
 * medium_dev has in_ch_total channels per environment (3 for an env's medium); action_dev == NULL runs the model alone
 * (ConvolutionModel.forward on any [B][in_ch_total][H][W] input; cout_last <= 4 then). */
int die_conv_policy_forward(int32_t H, int32_t W, int64_t M, int32_t B, int32_t field_dtype,
                            const void* medium_dev, int32_t in_ch_total, int32_t in_ch0, int32_t cin, int32_t cout_last,
                            int32_t n_layers, const int32_t* kernel_sizes_host, const float* weights_dev,
                            float* scratch_a_dev, float* scratch_b_dev,
                            const double* agents_dev, const int32_t* cells_hint_dev, const float* coefs_host,
                            double* action_dev, int32_t* final_scratch, void* stream);

/* The same with a POPULATION of models, one per environment of the batch (what a neuro-evolution search evaluates: the
 * reference's examples/learning_agents.py scores one candidate after the other on one env): environment b uses the
 * weights at weights_dev + b * weight_env_stride (floats; 0 = one model for every environment, as above). */
int die_conv_policy_forward_population(int32_t H, int32_t W, int64_t M, int32_t B, int32_t field_dtype,
                            const void* medium_dev, int32_t in_ch_total, int32_t in_ch0, int32_t cin, int32_t cout_last,
                            int32_t n_layers, const int32_t* kernel_sizes_host, const float* weights_dev,
                            int64_t weight_env_stride,
                            float* scratch_a_dev, float* scratch_b_dev,
                            const double* agents_dev, const int32_t* cells_hint_dev, const float* coefs_host,
                            double* action_dev, int32_t* final_scratch, void* stream);

/* Diagnostics: the kernels' bit-reproducible sin/cos/atan2 (die_b200/csrc/die_math.h) applied
 * to device arrays, so tests can check the device results equal the host build of the same
 * source bit-for-bit.  fast != 0 selects die_atan2_fast. */
int die_math_sincos(const double* x_dev, double* sin_dev, double* cos_dev, int64_t n, void* stream);
int die_math_atan2(const double* y_dev, const double* x_dev, double* out_dev, int64_t n,
                   int32_t fast, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIE_B200_H */
