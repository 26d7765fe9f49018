"""The kernel SOURCES of die_b200/csrc, executed thread by thread on the CPU (tests/hostsim: one fiber per CUDA
thread, __syncthreads / shuffles / ballots as fiber barriers), through the same C ABI and in the same call order
as die_b200.Env / die_b200.PhysarumAgent, against the numpy oracle.

What this buys: the logic of every kernel (indexing, tiling, halos, the claim protocol, the reductions, every
template instantiation the dispatcher can pick) is checked on every CPU run, bit for bit -- the arithmetic is the
same IEEE float64 sequence on both machines (-fmad=false / -ffp-contract=off, die_math.h for sin / cos / atan2).
What it does not: speed, and races between CTAs (blocks run one after the other).  The `-m gpu` tests stay the
parity tests proper; the product never loads the emulated library.
"""
import numpy as np
import pytest

import die_b200 as D
from die_b200 import _lib as L
from oracle import die_ref as R
from tests._parity import lattice_theta, ref_cells_linear, assert_state_equal
from tests.hostsim import sim as S

PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)     # README.md:45-48


def _rel(a, b):
    return abs(a - b) / max(abs(a), abs(b), 1e-300)


def make_pair(field, seed=1, ratio=0.1, dynamics_kw=None, ref_dynamics_kw=None, batch=None):
    dynamics_kw = dynamics_kw or {}
    ref_dynamics_kw = ref_dynamics_kw if ref_dynamics_kw is not None else dict(dynamics_kw)
    refs = []
    for b in range(batch or 1):
        np.random.seed(seed + b)
        refs.append(R.Env(field, R.Dynamics(init_agent_ratio=ratio, **ref_dynamics_kw), noise_seed=seed + b))
    env = S.SimEnv(field, np.stack([r.medium for r in refs]), np.stack([r.agents for r in refs]),
                   D.Dynamics(init_agent_ratio=ratio, **dynamics_kw), batch=batch)
    return refs, env


@pytest.fixture
def portable_math():
    R.set_math_backend('portable')
    yield
    R.set_math_backend('numpy')


@pytest.fixture
def tuning():
    """set_tuning(key, value) with every switch restored afterwards."""
    defaults = dict(turn_quick=1, fwd_min_blocks=4, feed_bits=1, fwd_lean=1, field_prefetch=1, grad_f32=1, step_impl=0, fused_threads=512, field_vec=0, sense_quick=1, pair_mode=1, pair_min_cells_log2=23, cost_hint=1, cost_sqrt_near=1, field_tile=0, feed_min_blocks=6)
    yield S.set_tuning
    for k, v in defaults.items():
        S.set_tuning(k, v)


# ------------------------------------------------------------------------------------------
# config 1: BrownianAgent free running, bit-exact (brownian_forward, move_claim, field_step, agent_feed, finalize)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("field,iters,flags", [((37, 53), 25, L.STEP_ALIVE_BITS), ((64, 64), 10, 0), ((24, 200), 10, L.STEP_ALIVE_BITS)])
def test_brownian_free_run_bit_exact(field, iters, flags):
    (ref,), env = make_pair(field, seed=1)
    ra = R.BrownianAgent(0.01)
    m = ref.agents.shape[-1]
    rng = np.random.default_rng(7)
    robs = ref._get_current_obs
    for it in range(iters):
        u = rng.random((3, m))
        ract = ra.forward(robs, u=u)
        gact = S.brownian_forward(env.agents[0], u=u)
        assert np.array_equal(ract, gact), f"action differs at step {it}"
        robs, rr, rterm, _, rinfo = ref.step(ract)
        r, alive = env.step(gact, flags=flags)
        assert np.array_equal(ref_cells_linear(ref), env.cells()[0]), f"cells differ at step {it}"
        assert rinfo['num_agents'] == alive[0]
        assert _rel(rr, r[0]) < 1e-11
        assert_state_equal(ref, env.medium[0], env.agents[0], float_exact=True)


@pytest.mark.parametrize("field,rate_feed", [((28, 36), 0.1), ((24, 64), 0.0)])
def test_agents_die_brownian(field, rate_feed):
    """Dynamics(agents_die=True): Env._agent_lifecycle (core/env.py:245-250) folded into the feed kernel -- slots whose
    stock is not above 1e-4 after feeding lose every channel (dead, back at (0, 0)); agents starve over the run, ghosts
    are reset every step; num_agents counts the survivors.  Semantics pinned against the reference in
    tests/test_golden_oracle.py::test_agents_die_equals_reference_with_its_indexer_rebound."""
    dyn = dict(agents_die=True, rate_feed=rate_feed)
    (ref,), env = make_pair(field, seed=31, ratio=0.3, dynamics_kw=dyn)
    ra = R.BrownianAgent(0.03, 2.0)
    m = ref.agents.shape[-1]
    rng = np.random.default_rng(7)
    robs = ref._get_current_obs
    alive0 = ref.num_alive
    for it in range(30):
        u = rng.random((3, m))
        ract = ra.forward(robs, u=u)
        gact = S.brownian_forward(env.agents[0], move_scale=0.03, deposit_scale=2.0, u=u)
        assert np.array_equal(ract, gact), f"action differs at step {it}"
        robs, rr, rterm, _, rinfo = ref.step(ract)
        r, alive = env.step(gact)
        assert rinfo['num_agents'] == alive[0]
        assert _rel(rr, r[0]) < 1e-11
        assert_state_equal(ref, env.medium[0], env.agents[0], float_exact=True)
    assert ref.num_alive < alive0


def test_agents_die_physarum(portable_math):
    """The same with PhysarumAgent acting through the env's hints: the cell cache of a slot put back at (0, 0) is reset."""
    dyn = dict(agents_die=True, rate_feed=0.02)
    _physarum_free_run((28, 40), 25, dict(PHYS, scale=0.05, deposit=8.0), seed=5, dynamics_kw=dyn, ref_dynamics_kw=dyn)


def test_brownian_philox_is_launch_invariant_and_masked():
    """In-kernel Philox draws: ghosts get exactly zero (Q8), values lie on the 3-decimal lattice (Q9), and a batch
    of two environments draws different numbers per environment."""
    (ref,), _ = make_pair((32, 32), seed=3)
    ag = np.stack([ref.agents, ref.agents])
    act = S.brownian_forward(ag, move_scale=0.01, seed=5, step=9)
    ghosts = ag[0, 2] == 0
    assert (act[:, :, ghosts] == 0).all()
    q = (act[0, 0, ~ghosts] + 0.01) / 0.02 * 1000
    assert np.allclose(q, np.rint(q), atol=1e-9)
    assert not np.array_equal(act[0], act[1])
    assert np.array_equal(act, S.brownian_forward(ag, move_scale=0.01, seed=5, step=9))


# ------------------------------------------------------------------------------------------
# config 2: PhysarumAgent, oracle in its portable math backend -> everything bit-exact, free running
# ------------------------------------------------------------------------------------------
def _physarum_free_run(field, iters, agent_kw, seed=2, dynamics_kw=None, ref_dynamics_kw=None, use_hints=True,
                       record=True, fuse=False, exact=True):
    """exact=False: the policy involves a libm function die_math.h does not replace (hypot of an unnormalised
    gradient), so headings / moves are compared to 1e-13 and the oracle continues from the kernel's values."""
    (ref,), env = make_pair(field, seed=seed, dynamics_kw=dynamics_kw, ref_dynamics_kw=ref_dynamics_kw)
    m = ref.agents.shape[-1]
    theta0, prev = lattice_theta(m, agent_kw.get('turn_angle', 30), seed)
    ra = R.PhysarumAgent(max_agents=m, prev_grad=prev, **agent_kw)
    ga = S.SimGradientAgent(m, **agent_kw)
    ga.theta[0] = theta0
    if ga.prev_grad is not None:
        ga.prev_grad[0] = prev
    ga.record_sense_cells = record
    ga.fuse_move = fuse
    rng = np.random.default_rng(seed)
    robs = ref._get_current_obs
    w = field[1]
    seen_flags = set()
    for it in range(iters):
        coin = rng.integers(0, 2, m)
        # the momentum step's noise draws matter even at noise_scale = 0 wherever g' can hold a -0.0 (see _needs_prev)
        noise = rng.normal(0., 0.4, size=(2, m)) if ga.prev_grad is not None else None
        ract = ra.forward(robs, coin=coin.copy(), noise=None if noise is None else noise.copy())
        gact = ga.forward(env, coin=coin, noise=noise, use_hints=use_hints)[0]
        seen_flags.add(ga.last_flags)
        if record:
            sx, sy = ra.last_sense_cells
            assert np.array_equal((sx * w + sy).astype(np.int32), ga.sense_cells[0]), f"sense cells, step {it}"
        if exact:
            assert np.array_equal(ga.theta[0], ra._direction_rads), f"theta differs at step {it}"
            assert np.array_equal(gact, ract), f"action differs at step {it}"
        else:
            d = np.abs(R.renormalize_radians(ga.theta[0] - ra._direction_rads))
            assert np.minimum(d, 2 * np.pi - d).max() < 1e-13, f"theta differs at step {it}"
            np.testing.assert_allclose(gact, ract, rtol=1e-13, atol=1e-17)
            ra._direction_rads, ra._prev_grad, ract = ga.theta[0].copy(), ga.prev_grad[0].copy(), gact.copy()
        robs, rr, _, _, rinfo = ref.step(ract)
        flags = L.STEP_ALIVE_BITS | (L.STEP_ADOPT_MOVE if fuse else 0)
        r, alive = env.step(gact, flags=flags)
        expect_cells = ref_cells_linear(ref)
        if ref.dynamics.agents_die:           # the cache serves the NEXT forward: a slot put back at (0, 0) is in cell 0
            expect_cells = np.where((ref.agents == 0).all(axis=0), 0, expect_cells)
        assert np.array_equal(expect_cells, env.cells()[0]), f"cells differ at step {it}"
        assert rinfo['num_agents'] == alive[0]
        assert _rel(rr, r[0]) < 1e-11
        assert_state_equal(ref, env.medium[0], env.agents[0], float_exact=True)
    return env, ga, seen_flags


def test_physarum_free_run_with_env_hints(portable_math):
    """Published gradient + cell cache (the steady-state path of the drop-in loop)."""
    env, _, flags = _physarum_free_run((48, 80), 30, PHYS)
    assert L.FWD_USE_GRADIENT | L.FWD_USE_CELLS in flags
    assert S.lib().die_env_gradient_kind(env.handle) == 2        # a decision-only consumer: float32 pairs by default


def test_physarum_free_run_without_hints(portable_math):
    """die_gradient_forward: cells resolved and np.gradient evaluated per sample (4 chem gathers)."""
    _physarum_free_run((45, 131), 20, dict(scale=0.02, turn_angle=35, sense_angle=120, sense_offset=0.06,
                                           turn_tolerance=0.05), seed=5, use_hints=False)


def test_physarum_free_run_fused_move(portable_math):
    """The forward kernel evaluates the move + claims speculatively, the step adopts them (feed commits positions)."""
    _physarum_free_run((40, 72), 20, dict(scale=0.02, turn_angle=30, sense_offset=0.06), seed=8, fuse=True)


def test_physarum_limit_boundary_sigma08_infinite_food_zero_cost(portable_math):
    _physarum_free_run((33, 70), 15, PHYS, seed=4,
                       dynamics_kw=dict(boundary=D.BoundaryCondition.limit, diffuse_sigma=0.8, food_infinite=True,
                                        op_action_cost=D.zero_cost),
                       ref_dynamics_kw=dict(boundary='limit', diffuse_sigma=0.8, food_infinite=True,
                                            op_action_cost=R.zero_cost))


@pytest.mark.parametrize("lean", [1, 0])
def test_physarum_free_run_float32_gradient_cache(portable_math, tuning, lean):
    """grad_f32: the field pass publishes np.gradient(chem1) as float32 pairs; the quick turn
    decision reads those, deferred slots re-sample chem1 in float64.  Against the oracle, and with the structured
    fields of a long run where gradients underflow float32 / sit on the clip threshold."""
    tuning("grad_f32", 1)
    tuning("fwd_lean", lean)
    env, ga, flags = _physarum_free_run((48, 80), 30, PHYS)
    assert S.lib().die_env_gradient_kind(env.handle) == 2 and not S.lib().die_env_gradient(env.handle)
    assert L.FWD_USE_GRADIENT | L.FWD_USE_CELLS in flags
    # GradientAgent needs the gradient's value: the float32 cache must be ignored, results still exact
    test_gradient_agent_inertia_noise(None, 0.9, 0.025)


@pytest.mark.parametrize("mode", ['reflect', 'nearest', 'mirror', 'constant'])
def test_diffuse_modes(portable_math, mode):
    _physarum_free_run((20, 37), 8, PHYS, seed=6, dynamics_kw=dict(diffuse_mode=mode, diffuse_sigma=0.8))


def test_no_diffusion(portable_math):
    """blur radius 0 (sigma 0.1): the no-blur field kernel; no gradient is published then."""
    _physarum_free_run((20, 37), 8, PHYS, seed=6, dynamics_kw=dict(diffuse_sigma=0.1))


# ------------------------------------------------------------------------------------------
# A-B: every instantiation the dispatcher can pick gives the same bits
# ------------------------------------------------------------------------------------------
def _philox_run(field, iters, seed=11, batch=None, agent_kw=PHYS, record=False):
    """Free run with in-kernel Philox coins (the LEAN forward's precondition) -> final state."""
    refs, env = make_pair(field, seed=seed, batch=batch)
    m = env.M
    ga = S.SimGradientAgent(m, B=env.B, seed=3, **agent_kw)
    for b in range(env.B):
        ga.theta[b] = lattice_theta(m, 30, seed + b)[0]
    ga.record_sense_cells = record
    rewards = []
    for it in range(iters):
        act = ga.forward(env)
        r, _ = env.step(act)
        rewards.append(r)
    return env.medium.copy(), env.agents.copy(), ga.theta.copy(), np.array(rewards)


@pytest.mark.parametrize("key,values", [("fwd_lean", [0, 5]), ("turn_quick", [0]), ("fwd_min_blocks", [3, 5]),
                                        ("feed_bits", [0]), ("field_prefetch", [0]), ("grad_f32", [0]), ("step_impl", [1]), ("field_vec", [1]), ("sense_quick", [0]), ("pair_mode", [2]),
                                        ("field_tile", [1, 2]), ("cost_hint", [0]), ("cost_sqrt_near", [0]), ("feed_min_blocks", [0])])
def test_tuning_switches_do_not_change_results(tuning, key, values):
    lean0 = S.lib().die_get_counter(b"forward_lean_f32")
    base = _philox_run((40, 72), 12)
    # the default configuration's steady state IS the LEAN forward on the float32 gradient cache (every step but the
    # first, which has no hints yet)
    assert S.lib().die_get_counter(b"forward_lean_f32") == lean0 + 11
    for v in values:
        tuning(key, v)
        out = _philox_run((40, 72), 12)
        for a, b, what in zip(base, out, ("medium", "agents", "theta", "reward")):
            assert np.array_equal(a, b), f"{key}={v}: {what} differs"


@pytest.mark.parametrize("shape,sigma,batch,kw", [
    ((40, 72), 0.5, None, {}), ((33, 47), 1.0, 3, {}), ((64, 64), 0.5, 2, dict(food_infinite=True)),
    ((24, 50), 0.8, None, dict(diffuse_mode='reflect', boundary=D.BoundaryCondition.limit)), ((6, 90), 0.3, 2, {})])
@pytest.mark.parametrize("grad_f32", [1, 0])
def test_pair_mode_equals_the_separate_tables(tuning, shape, sigma, batch, kw, grad_f32):
    """pair_mode (large single fields by default, forced here): the field pass writes {consumed_field, new food} per
    cell, the feed kernel's gather brings both and stores the food under every slot, the next LEAN forward reads it per
    slot instead of gathering -- every output equals the three separate tables' bit for bit; a Brownian agent (no
    gradient, no forward consumer) and an observation that is not the env's own (no hints: the general forward) too."""
    outs = []
    fh0 = S.lib().die_get_counter(b"forward_food_here")
    tuning("grad_f32", grad_f32)
    for mode in (0, 2):
        tuning("pair_mode", mode)
        refs, env = make_pair(shape, seed=5, dynamics_kw=dict(diffuse_sigma=sigma, **kw), batch=batch)
        B = env.B
        ga = S.SimGradientAgent(env.M, B=B, seed=1, **PHYS)
        for b in range(B):
            ga.theta[b] = lattice_theta(env.M, 30, 5 + b)[0]
        rewards = []
        for it in range(9):
            if it == 4:
                act = np.stack([S.brownian_forward(env.agents[b], move_scale=0.02, seed=4 + b, step=it) for b in range(B)])
            elif it == 6:
                act = ga.forward(env, use_hints=False)
            else:
                act = ga.forward(env)
            rewards.append(env.step(act)[0].copy())
        outs.append((env.medium.copy(), env.agents.copy(), ga.theta.copy(), np.array(rewards), env.cells().copy()))
    # steps 1-3, 5, 7, 8 of the pair-mode run had valid hints + a pair-mode step before them
    assert S.lib().die_get_counter(b"forward_food_here") == fh0 + 6
    for a, b, what in zip(outs[0], outs[1], ("medium", "agents", "theta", "reward", "cells")):
        assert np.array_equal(a, b), f"{what} differs"


def test_pair_mode_is_chosen_by_field_size_and_declines_where_it_does_not_apply(tuning):
    """Automatic mode: fields of pair_min_cells and more; never with agents_die, float32 fields, no diffusion, a
    speculative move."""
    fh = lambda: S.lib().die_get_counter(b"forward_food_here")
    def run(shape, dynamics_kw=None, fuse=False):
        refs, env = make_pair(shape, seed=3, dynamics_kw=dynamics_kw)
        ga = S.SimGradientAgent(env.M, seed=1, **PHYS)
        ga.fuse_move = fuse
        n0 = fh()
        for it in range(3):
            env.step(ga.forward(env), flags=L.STEP_ALIVE_BITS | (L.STEP_ADOPT_MOVE if fuse else 0))
        return fh() - n0
    tuning("pair_min_cells_log2", 11)
    assert run((32, 64)) == 2 and run((32, 63)) == 0
    assert run((32, 64), dict(agents_die=True)) == 0
    assert run((32, 64), dict(diffuse_sigma=0.1)) == 0
    assert run((32, 64), fuse=True) == 0
    tuning("pair_mode", 0)
    assert run((32, 64)) == 0


@pytest.mark.parametrize("shape,sigma,batch", [((64, 64), 0.5, None), ((40, 72), 0.5, 3), ((6, 76), 0.5, None),
                                               ((48, 96), 1.0, 2), ((33, 132), 0.3, 2), ((16, 34), 0.8, 5)])
@pytest.mark.parametrize("grad,threads", [(True, 512), (False, 512)])
def test_fused_step_equals_three_kernels(tuning, shape, sigma, batch, grad, threads):
    """step_impl = 1 (die_env_fused.cuh): one thread-block cluster per environment -- claims by atomicMax into the
    distributed shared memory of the CTA that owns the row, chem rows by bulk copies on an mbarrier, the feed phase
    reading remote claims, block partials and the final sum in the order of the three-kernel path.  Under the emulator
    the CTAs of a cluster run concurrently as fibers with separate shared memories.  Cluster sizes 16 / 8 / 2 / 1 (by the
    divisors of H), CTAs without slots (few feed blocks), radii 1..4, batches, with (Physarum) and without (Brownian)
    the published gradient: medium, agents, headings, rewards and the cell cache must not differ in a bit."""
    outs = []
    fused0 = S.lib().die_get_counter(b"step_fused")
    tuning("fused_threads", threads)
    for impl in (0, 1):
        tuning("step_impl", impl)
        refs, env = make_pair(shape, seed=13, dynamics_kw=dict(diffuse_sigma=sigma), batch=batch)
        B = env.B
        ga = S.SimGradientAgent(env.M, B=B, seed=1, **PHYS)
        for b in range(B):
            ga.theta[b] = lattice_theta(env.M, 30, 13 + b)[0]
        for it in range(5):
            if grad:
                act = ga.forward(env)
            else:
                act = S.brownian_forward(env.agents, move_scale=0.02, seed=4, step=it)
            env.step(act)
        outs.append((env.medium.copy(), env.agents.copy(), ga.theta.copy(), env.reward.copy(), env.cells().copy(),
                     env.gradient() if grad else None))
    # (33 rows only admit a cluster of one CTA: five rounds of feed blocks for 256 threads, one more than the kernel holds)
    expect = 0 if (threads == 256 and shape == (33, 132)) else 5
    assert S.lib().die_get_counter(b"step_fused") == fused0 + expect, "the fused step must be the one that ran"
    for a, b, what in zip(outs[0], outs[1], ("medium", "agents", "theta", "reward", "cells", "gradient")):
        if a is None:
            continue
        assert np.array_equal(a, b), f"{what} differs"


@pytest.mark.parametrize("shape,sigma,batch", [((64, 128), 0.5, None), ((40, 72), 0.5, None), ((6, 76), 0.5, None),
                                               ((70, 200), 0.8, None), ((33, 132), 0.3, 3), ((48, 96), 1.0, 2),
                                               ((96, 80), 0.5, 5), ((37, 2), 0.5, None)])
@pytest.mark.parametrize("grad", [True, False])
def test_vectorised_field_pass_equals_the_scalar_one(tuning, shape, sigma, batch, grad):
    """field_vec = 1 (field_step_vec_kernel): the field pass with 128-bit global accesses -- halo tile staged in aligned
    cell pairs from an even column, two adjacent output cells per thread.  Even widths that are not multiples of the
    tile (the pair at the right edge), fields lower than a tile, radii 1..4 (odd and even halo widths: with / without
    the pad column), batches, with (Physarum, float64 and float32 gradient cache) and without (Brownian) the gradient."""
    outs = []
    vec0 = S.lib().die_get_counter(b"field_vec")
    for vec, f32 in ((0, 0), (1, 0), (1, 1)):
        tuning("field_vec", vec)
        tuning("grad_f32", f32)
        refs, env = make_pair(shape, seed=13, dynamics_kw=dict(diffuse_sigma=sigma), batch=batch)
        B = env.B
        ga = S.SimGradientAgent(env.M, B=B, seed=1, **PHYS)
        for b in range(B):
            ga.theta[b] = lattice_theta(env.M, 30, 13 + b)[0]
        for it in range(5):
            if grad:
                act = ga.forward(env)
            else:
                act = S.brownian_forward(env.agents, move_scale=0.02, seed=4, step=it)
            env.step(act)
        outs.append((env.medium.copy(), env.agents.copy(), ga.theta.copy(), env.reward.copy(),
                     None if f32 or not grad else env.gradient()))
    assert S.lib().die_get_counter(b"field_vec") == vec0 + 10, "the 128-bit kernel must be the one that ran"
    for k in (1, 2):
        for a, b, what in zip(outs[0], outs[k], ("medium", "agents", "theta", "reward", "gradient")):
            if a is None or b is None:
                continue
            assert np.array_equal(a, b), f"variant {k}: {what} differs"


def test_vectorised_field_pass_falls_back_where_it_does_not_apply(tuning):
    """Odd widths (a pair would straddle the row end) and the other scipy boundary modes (time-varying food flows: the
    wave / tabulated flow tests run with field_vec = 1 and stay bit-exact against the oracle)."""
    vec0 = S.lib().die_get_counter(b"field_vec")
    for shape, kw in (((40, 71), {}), ((40, 64), dict(diffuse_mode='reflect'))):
        refs, env = make_pair(shape, seed=3, dynamics_kw=kw)
        ga = S.SimGradientAgent(env.M, seed=1, **PHYS)
        for it in range(2):
            env.step(ga.forward(env))
    assert S.lib().die_get_counter(b"field_vec") == vec0


def test_fused_step_falls_back_where_it_does_not_apply(tuning):
    """Odd row lengths (the bulk copies need 16-byte rows), non-periodic diffusion, no diffusion: the three kernels run."""
    fused0 = S.lib().die_get_counter(b"step_fused")
    for shape, kw in (((40, 71), {}), ((40, 64), dict(diffuse_mode='reflect')), ((40, 64), dict(diffuse_sigma=0.1))):
        refs, env = make_pair(shape, seed=3, dynamics_kw=kw)
        ga = S.SimGradientAgent(env.M, seed=1, **PHYS)
        for it in range(2):
            env.step(ga.forward(env))
    assert S.lib().die_get_counter(b"step_fused") == fused0


def test_sensed_cells_next_to_cell_boundaries(portable_math, tuning):
    """sense_quick: the sense position pos + sense_offset (cos, sin)(theta) is formed with a float32 sin / cos (+-4e-7)
    and accepted only if the sensed cell is not within the guard of a cell boundary; otherwise the float64 die_sincos
    decides.  Adversarial slots: positions placed so that the EXACT sense coordinate lies on a boundary between two
    cells, and 1e-12 ... 1e-6 cells to either side of it, on both axes, for headings all around the lattice."""
    field = (40, 72)
    np.random.seed(3)
    ref = R.Env(field, R.Dynamics(init_agent_ratio=0.1), noise_seed=3)
    m = ref.agents.shape[-1]
    rng = np.random.default_rng(0)
    so = 0.04
    theta = (rng.integers(-6, 7, m) * np.radians(30)).astype(np.float64)
    theta[::7] = rng.uniform(-np.pi, np.pi, theta[::7].size)
    sn, cs = R.math_sincos(theta) if hasattr(R, "math_sincos") else (np.sin(theta), np.cos(theta))
    deltas = np.array([0.0, 1e-12, -1e-12, 1e-9, -1e-9, 3e-8, -3e-8, 2e-7, -2e-7, 4.5e-7, -4.5e-7, 1e-6, -1e-6])
    for axis, (n, off) in enumerate(((field[0], so * cs), (field[1], so * sn))):
        k = rng.integers(1, n - 2, m)
        target = (k + 0.5 + deltas[rng.integers(0, deltas.size, m)]) / (n - 1)       # a boundary between cells k and k + 1
        ref.agents[axis] = np.clip(target - off, 0.0, 1.0)
    env = S.SimEnv(field, ref.medium[None], ref.agents[None], D.Dynamics(init_agent_ratio=0.1))
    outs = []
    for quick in (1, 0):
        tuning("sense_quick", quick)
        ga = S.SimGradientAgent(m, **PHYS)
        ga.theta[0] = theta
        ga.record_sense_cells = True
        ga.forward(env, coin=np.zeros(m, dtype=np.int64), use_hints=False)
        outs.append((ga.sense_cells[0].copy(), ga.action.copy(), ga.theta.copy()))
    ra = R.PhysarumAgent(max_agents=m, prev_grad=np.ones((2, m)), **PHYS)
    ra._direction_rads = theta.copy()
    ra.forward(ref._get_current_obs, coin=np.zeros(m, dtype=np.int64))
    sx, sy = ra.last_sense_cells
    assert np.array_equal((sx * field[1] + sy).astype(np.int32), outs[0][0]), "sensed cells differ from the oracle's"
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_batched_envs_match_single_envs():
    """B = 3 environments in one set of launches == three single environments (same Philox coins: the in-kernel
    RNG is keyed on (environment, slot), not on the launch geometry)."""
    med, ag, th, rw = _philox_run((32, 48), 8, batch=3)
    # single environments see the same coins only through injected draws, so compare against the oracle instead
    refs, env = make_pair((32, 48), seed=11, batch=3)
    m = env.M
    ga = S.SimGradientAgent(m, B=3, seed=3, **PHYS)
    singles = []
    for b in range(3):
        ga.theta[b] = lattice_theta(m, 30, 11 + b)[0]
        e1 = S.SimEnv((32, 48), refs[b].medium, refs[b].agents)
        a1 = S.SimGradientAgent(m, seed=3, **PHYS)
        a1.theta[0] = ga.theta[b]
        singles.append((e1, a1))
    rng = np.random.default_rng(0)
    for it in range(8):
        coin = rng.integers(0, 2, (3, m))
        env.step(ga.forward(env, coin=coin))
        for b, (e1, a1) in enumerate(singles):
            e1.step(a1.forward(e1, coin=coin[b]))
    for b, (e1, a1) in enumerate(singles):
        assert np.array_equal(env.medium[b], e1.medium[0])
        assert np.array_equal(env.agents[b], e1.agents[0])
        assert np.array_equal(ga.theta[b], a1.theta[0])
        assert env.reward[b] == e1.reward[0] and env.alive[b] == e1.alive[0]


@pytest.mark.parametrize("inertia,noise_scale", [(0.9, 0.025), (0.0, 0.0), (0.0, 0.025)])
def test_gradient_agent_inertia_noise(portable_math, inertia, noise_scale):
    """GradientAgent (no discrete turn): momentum with injected noise, prev_grad state.  With inertia = noise_scale = 0
    the momentum step only matters for the signs of zeros (clipped gradients) -- which decide heading 0 vs pi."""
    (ref,), env = make_pair((40, 56), seed=21)
    m = env.M
    kw = dict(scale=0.01, deposit=4.0, inertia=inertia, sense_offset=0.02, noise_scale=noise_scale)
    rng = np.random.default_rng(21)
    prev = rng.normal(0., 0.4, size=(2, m))
    ra = R.GradientAgent(max_agents=m, prev_grad=prev.copy(), **kw)
    ga = S.SimGradientAgent(m, discrete_turn=False, **kw)
    ga.theta[0] = R.get_radians(prev)
    ga.prev_grad[0] = prev
    robs = ref._get_current_obs
    for it in range(10):
        noise = rng.normal(0., 0.4, size=(2, m))
        ract = ra.forward(robs, noise=noise.copy())
        gact = ga.forward(env, noise=noise)[0]
        assert np.array_equal(gact, ract), f"action differs at step {it}"
        assert np.array_equal(ga.theta[0], ra._direction_rads)
        robs, rr, *_ = ref.step(ract)
        r, _ = env.step(gact)
        assert _rel(rr, r[0]) < 1e-11
        assert_state_equal(ref, env.medium[0], env.agents[0], float_exact=True)
    assert S.lib().die_env_gradient_kind(env.handle) == 1        # GradientAgent uses the gradient's value: float64 pairs


def test_gradient_cache_follows_its_consumer(portable_math):
    """One env, a PhysarumAgent and a GradientAgent taking turns: the field pass publishes float32 pairs after a
    decision-only forward and float64 pairs after a value-using one; a forward that finds the wrong kind samples chem1
    itself.  Everything stays bit-exact against the oracle."""
    (ref,), env = make_pair((40, 56), seed=23)
    m = env.M
    theta0, prev = lattice_theta(m, 30, 23)
    rp = R.PhysarumAgent(max_agents=m, prev_grad=prev, **PHYS)
    gp = S.SimGradientAgent(m, **PHYS)
    gp.theta[0] = theta0
    kw = dict(scale=0.01, deposit=4.0, inertia=0.9, sense_offset=0.02, noise_scale=0.025)
    rg = R.GradientAgent(max_agents=m, prev_grad=prev.copy(), **kw)
    gg = S.SimGradientAgent(m, discrete_turn=False, **kw)
    gg.theta[0], gg.prev_grad[0] = R.get_radians(prev), prev
    rng = np.random.default_rng(3)
    kinds = []
    for it in range(12):
        physarum = (it // 2) % 2 == 0
        if physarum:
            coin = rng.integers(0, 2, m)
            ract = rp.forward(ref._get_current_obs, coin=coin.copy())
            gact = gp.forward(env, coin=coin)[0]
        else:
            noise = rng.normal(0., 0.4, size=(2, m))
            ract = rg.forward(ref._get_current_obs, noise=noise.copy())
            gact = gg.forward(env, noise=noise)[0]
        assert np.array_equal(gact, ract), f"action differs at step {it}"
        ref.step(ract)
        env.step(gact)
        assert_state_equal(ref, env.medium[0], env.agents[0], float_exact=True)
        kinds.append(S.lib().die_env_gradient_kind(env.handle))
    assert kinds == [2, 2, 1, 1] * 3


def test_host_buffer_step_is_chunked_and_identical(tuning):
    """die_env_step_host cuts the batch into chunks on two streams; same results as the device-pointer step."""
    tuning("host_chunk_min_kb", 0)
    try:
        outs = []
        for host in (False, True):
            refs, env = make_pair((24, 32), seed=17, batch=9)
            ga = S.SimGradientAgent(env.M, B=9, seed=2, **PHYS)
            for b in range(9):
                ga.theta[b] = lattice_theta(env.M, 30, b)[0]
            for it in range(4):
                act = ga.forward(env, use_hints=False)
                if host:
                    (ag_h, med_h), _, _ = env.step_host(act)
                    assert np.array_equal(ag_h, env.agents) and np.array_equal(med_h, env.medium)
                else:
                    env.step(act)
            outs.append((env.medium.copy(), env.agents.copy(), env.reward.copy(), env.alive.copy()))
        for a, b in zip(*outs):
            assert np.array_equal(a, b)
    finally:
        tuning("host_chunk_min_kb", 32 << 10)


# ------------------------------------------------------------------------------------------
# edge cases of the reference's semantics, on the emulated kernels
# ------------------------------------------------------------------------------------------
def test_everybody_on_one_cell_and_full_occupancy():
    h, w = 8, 8
    m = h * w
    medium = np.zeros((3, h, w))
    medium[1] = 0.5
    agents = np.zeros((4, m))
    agents[2] = 1.0                     # all alive, all at (0, 0)
    agents[3] = 0.3
    env = S.SimEnv((h, w), medium, agents)
    action = np.zeros((3, m))
    action[2] = np.arange(m) + 1.0      # deposit of slot i = i + 1: the LAST slot's deposit must land (Q2)
    env.step(action)
    # chem before blur had one cell = m (last writer); blur * decay conserves mass * 0.9
    assert abs(env.medium[0, 2].sum() - 0.9 * m) < 1e-9
    assert env.medium[0, 0].sum() == 1.0 and env.medium[0, 0, 0, 0] == 1.0
    # every slot eats the full rate_feed * food of the cell (Q7), the cell loses it once
    assert np.allclose(env.agents[0, 3], 0.3 + 0.1 * 0.5 - 0.02 * (np.arange(m) + 1.0))
    assert env.medium[0, 1, 0, 0] == 0.5 - 0.1 * 0.5
    assert env.alive[0] == m


def test_fewer_and_more_slots_than_cells(portable_math):
    for m in (50, 700):
        h, w = 16, 24
        rng = np.random.default_rng(m)
        medium = np.zeros((3, h, w))
        medium[1] = rng.random((h, w)).round(3)
        agents = np.zeros((4, m))
        n = m // 3
        agents[0, :n] = rng.integers(0, h, n) / (h - 1)
        agents[1, :n] = rng.integers(0, w, n) / (w - 1)
        agents[2, :n] = 1.0
        agents[3, :n] = 0.5
        env = S.SimEnv((h, w), medium, agents)
        ga = S.SimGradientAgent(m, seed=1, **PHYS)
        for it in range(5):
            env.step(ga.forward(env))
        assert env.alive[0] == n
        assert np.isfinite(env.medium).all() and np.isfinite(env.agents).all()
        assert env.medium[0, 0].sum() <= n


def test_smallest_field():
    """2 x 2: coordinate 1.0 wraps to 0.0 on the move (Q3), so both agents end on cell (0, 0); the oracle agrees."""
    medium = np.zeros((3, 2, 2))
    medium[1] = [[0.2, 0.4], [0.6, 0.8]]
    agents = np.array([[0., 1., 0., 0.], [1., 0., 0., 0.], [1., 1., 0., 0.], [.5, .5, 0., 0.]])
    ref = R.Env((2, 2), R.Dynamics(), medium=medium, agents=agents)
    env = S.SimEnv((2, 2), medium, agents)
    act = np.zeros((3, 4))
    act[2, :2] = [1.0, 2.0]
    for it in range(3):
        _, rr, _, _, info = ref.step(act)
        r, alive = env.step(act)
        assert alive[0] == 2 == info['num_agents']
        assert _rel(rr, r[0]) < 1e-12
        assert_state_equal(ref, env.medium[0], env.agents[0], float_exact=True)
    assert env.medium[0, 0].tolist() == [[1., 0.], [0., 0.]]


# ------------------------------------------------------------------------------------------
# the library's other kernels
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("field,ratio", [((24, 32), 0.1), ((37, 70), 0.02)])
def test_sense_mask_kernel(field, ratio):
    from die_b200.env import gaussian_kernel1d
    (ref,), env = make_pair(field, seed=9, ratio=ratio, ref_dynamics_kw=dict(apply_sense_mask=True))
    w = np.ascontiguousarray(gaussian_kernel1d(2.0))
    obs = np.full_like(env.medium, np.nan)
    S.check(S.lib().die_sense_mask(field[0], field[1], 1, S.ptr(w), (len(w) - 1) // 2, S.ptr(env.medium), S.ptr(obs), None))
    assert np.array_equal(obs[0], ref._get_current_obs[1])


@pytest.mark.parametrize("field,colors", [((24, 32), 'rgb'), ((37, 53), 'one')])
def test_render_kernel(field, colors):
    (ref,), env = make_pair(field, seed=9, ratio=0.2)
    rr = R.EnvRenderer(field, field_colors_id=colors)
    h, w = field
    trace = np.zeros((1, h, w))
    img_m = np.full((1, h, w, 3), np.nan)
    img_a = np.full((1, env.M, 4), np.nan)
    color = None if rr._color is None else np.ascontiguousarray(rr._color)
    ra = R.BrownianAgent(0.02)
    rng = np.random.default_rng(1)
    for it in range(3):
        frames = rr.render(ref.medium, ref.agents)
        S.check(S.lib().die_render_frames(h, w, env.M, 1, S.ptr(env.medium), S.ptr(env.agents), S.ptr(trace), rr._decay,
                                          S.ptr(color), S.ptr(img_m), S.ptr(img_a), None))
        assert np.array_equal(img_m[0], frames[0])
        assert np.array_equal(trace[0], frames[1])
        assert np.array_equal(img_a[0].reshape(frames[2].shape), frames[2])
        act = ra.forward(ref._get_current_obs, u=rng.random((3, env.M)))
        ref.step(act)
        env.step(act)


def test_math_kernels_equal_the_host_build():
    from oracle import portable_math
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-10, 10, 5000), [0.0, -0.0, np.pi, -np.pi, 1e-300, 1e6]])
    s, c = np.empty_like(x), np.empty_like(x)
    S.check(S.lib().die_math_sincos(S.ptr(x), S.ptr(s), S.ptr(c), len(x), None))
    ps, pc = portable_math.sincos(x)
    assert np.array_equal(s, ps) and np.array_equal(c, pc)
    y = rng.normal(size=len(x))
    for fast in (0, 1):
        out = np.empty_like(x)
        S.check(S.lib().die_math_atan2(S.ptr(y), S.ptr(x), S.ptr(out), len(x), fast, None))
        assert np.array_equal(out, portable_math.atan2(y, x, fast=bool(fast)))


def test_emulator_reports_a_deadlock_instead_of_hanging(tmp_path):
    """The scheduler's own safety net: a barrier that part of a block never reaches is an error, not a hang."""
    import ctypes, subprocess, textwrap
    from tests.hostsim import build as B
    src = tmp_path / "dead.cpp"
    src.write_text(textwrap.dedent('''
        #include "cuda_runtime.h"
        static void bad_kernel(int* out) { if (threadIdx.x < 16) __syncthreads(); else { for (;;) { __syncthreads(); } } out[0] = 1; }
        static void good_kernel(int* out) { __shared__ int s[64]; s[threadIdx.x] = threadIdx.x; __syncthreads();
                                            int v = __shfl_sync(0xffffffffu, s[63 - threadIdx.x], (threadIdx.x + 1) & 31);
                                            out[blockIdx.x * 64 + threadIdx.x] = v + (int)__ballot_sync(0xffffffffu, threadIdx.x & 1); }
        extern "C" int run_bad(int* out) { hostsim::launch(dim3(1), dim3(64), 0, nullptr, bad_kernel, out); return cudaGetLastError(); }
        extern "C" int run_good(int* out) { hostsim::launch(dim3(2), dim3(64), 0, nullptr, good_kernel, out); return cudaGetLastError(); }
    '''))
    so = tmp_path / "dead.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", B.HERE, "-o", str(so), str(src),
                    B.os.path.join(B.HERE, "hostsim.cpp")], check=True)
    lib = ctypes.CDLL(str(so))
    out = np.zeros(128, dtype=np.int32)
    assert lib.run_good(out.ctypes.data_as(ctypes.c_void_p)) == 0
    t = np.arange(64)
    src_lane = (t & ~31) | ((t + 1) & 31)
    assert np.array_equal(out[:64], (63 - src_lane) + 0xAAAAAAAA - (1 << 32))
    assert np.array_equal(out[64:], out[:64])


# ------------------------------------------------------------------------------------------
# one field over G ranks (row slabs, peer-pointer kernels): the emulated multi-rank world == the single env
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,G,band", [((32, 40), 2, 0), ((48, 36), 3, 0), ((64, 32), 8, 0), ((32, 40), 2, 5),
                                          ((48, 36), 3, 2), ((64, 32), 4, 16)])
def test_slab_world_equals_single_env(shape, G, band):
    phys = dict(scale=0.02, turn_angle=30, sense_offset=0.08)      # large moves: agents cross slab seams quickly
    (ref,), env = make_pair(shape, seed=31, ratio=0.15)
    m = env.M
    theta0, _ = lattice_theta(m, 30, 31)
    ga = S.SimGradientAgent(m, **phys)
    ga.theta[0] = theta0
    world = S.SimSlabWorld(env.medium[0], env.agents[0], theta0, G, corner_r=band, **phys)
    rng = np.random.default_rng(1)
    crossed = 0
    for it in range(12):
        coin = rng.integers(0, 2, m)
        act = ga.forward(env, coin=coin)[0]
        world.forward(coin)
        r, alive = env.step(act)
        wr, walive = world.step()
        wmed, wag, wth, wact, wcells = world.gather()
        assert np.array_equal(wact, act), f"action differs at step {it}"
        assert np.array_equal(wcells, env.cells()[0]), f"cells differ at step {it}"
        assert np.array_equal(wmed, env.medium[0]), f"medium differs at step {it}"
        assert np.array_equal(wag, env.agents[0]), f"agents differ at step {it}"
        assert np.array_equal(wth, ga.theta[0]), f"theta differs at step {it}"
        assert walive == alive[0]
        assert abs(wr - r[0]) <= 1e-11 * max(1.0, abs(r[0]))
        owner_of_cell = (wcells // shape[1]) // (shape[0] // G)
        slot_owner = np.concatenate([np.full(world.layout.local_slots(q), q) for q in range(G)])
        ids = np.concatenate([world.layout.global_ids(q) for q in range(G)])
        crossed = max(crossed, int((owner_of_cell[ids] != slot_owner).sum()))
    assert crossed > 0, "the test must exercise agents standing on another rank's slab"


# ------------------------------------------------------------------------------------------
# the committed golden vectors (tests/golden: recorded by executing the reference's own source,
# oracle/make_golden_from_reference.py) replayed through the emulated kernels
# ------------------------------------------------------------------------------------------
import os      # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_brownian_golden_bit_exact():
    g = np.load(os.path.join(GOLD, "brownian_24x32.npz"))
    size = tuple(int(v) for v in g["size"])
    env = S.SimEnv(size, g["medium0"], g["agents0"])
    np.random.seed(int(g["loop_seed"]))
    m = env.M
    for k in range(len(g["rewards"])):
        u = np.stack([np.random.random_sample(m) for _ in range(3)])      # the reference's draw order: dx, dy, deposit1
        act = S.brownian_forward(env.agents[0], float(g["move_scale"]), float(g["deposit_scale"]), u=u)
        assert np.array_equal(act, g["actions"][k]), k
        r, alive = env.step(act)
        assert abs(r[0] - g["rewards"][k]) <= 1e-12 * max(1.0, abs(g["rewards"][k]))
        assert alive[0] == g["num_agents"][k]
    assert np.array_equal(env.medium[0], g["medium_final"]) and np.array_equal(env.agents[0], g["agents_final"])


GOLDEN_CASES = {
    "physarum_24x32.npz": (dict(), dict(scale=0.007, turn_angle=30, sense_offset=0.04)),
    "physarum_limit_sigma08_20x20.npz": (
        dict(boundary=D.BoundaryCondition.limit, diffuse_sigma=0.8, food_infinite=True, op_action_cost=D.zero_cost),
        dict(scale=0.03, turn_angle=35, sense_angle=120, sense_offset=0.06, turn_tolerance=0.05)),
    "physarum_waveflow_24x32.npz": (dict(waveflow=(0.5, 0.5)), dict(scale=0.007, turn_angle=30, sense_offset=0.04)),
}


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_physarum_golden_every_step(name):
    """Same bar as tests/test_gpu_golden.py: every recorded step from its recorded pre-state; decisions, cells,
    occupancy and the deposit channel exact; headings / moves to 1e-13 (die_math.h vs the recording host's libm);
    Env.step on the recorded action bit-exact in every field (the wave flow's per-cell cosine: 1e-15)."""
    g = np.load(os.path.join(GOLD, name))
    dyn_kw, agent_kw = GOLDEN_CASES[name]
    dyn_kw = dict(dyn_kw)
    size = tuple(int(v) for v in g["size"])
    wf = dyn_kw.pop("waveflow", None)
    flow = D.WaveSequence(size, dt=0.01).get_flow_operator(scale=wf[0], decay=wf[1]) if wf is not None else None
    m = g["agents_pre"].shape[-1]
    for k in range(len(g["reward"])):
        env = S.SimEnv(size, g["medium_pre"][k], g["agents_pre"][k], D.Dynamics(**dyn_kw))
        if flow is not None:
            flow.calls = k
            env.set_food_flow(flow)
        agent = S.SimGradientAgent(m, **agent_kw)
        agent.theta[0] = g["theta_pre"][k]
        act = agent.forward(env, coin=g["coin"][k])[0]
        assert np.array_equal(act[2], g["action"][k][2]), f"deposit differs at step {k}"
        np.testing.assert_allclose(act[:2], g["action"][k][:2], rtol=0, atol=1e-15)
        d = np.abs((agent.theta[0] - g["theta_post"][k] + np.pi) % (2 * np.pi) - np.pi)
        assert d.max() < 1e-13, f"heading differs at step {k}"
        r, alive = env.step(g["action"][k])
        med, ag = env.medium[0].copy(), env.agents[0]
        if flow is not None:
            np.testing.assert_allclose(med[1], g["medium_post"][k][1], rtol=0, atol=1e-15)
            med[1] = g["medium_post"][k][1]
        assert np.array_equal(med, g["medium_post"][k]) and np.array_equal(ag, g["agents_post"][k]), k
        assert abs(r[0] - g["reward"][k]) <= 1e-12 * max(1.0, abs(g["reward"][k]))
        assert alive[0] == g["num_agents"][k]


@pytest.mark.parametrize("field,sigma", [((24, 40), 0.5), ((37, 53), 0.5), ((20, 20), 0.1), ((24, 64), 0.5)])
def test_wave_flow_bit_exact_with_portable_math(portable_math, tuning, field, sigma):
    rflow = R.WaveSequence(field, dt=0.01).get_flow_operator(scale=0.5, decay=0.5)
    gflow = D.WaveSequence(field, dt=0.01).get_flow_operator(scale=0.5, decay=0.5)
    (ref,), env = make_pair(field, seed=5, dynamics_kw=dict(diffuse_sigma=sigma),
                            ref_dynamics_kw=dict(op_food_flow=rflow, diffuse_sigma=sigma))
    env.set_food_flow(gflow)
    ra = R.BrownianAgent(0.01)
    rng = np.random.default_rng(3)
    for it in range(10):
        act = ra.forward(ref._get_current_obs, u=rng.random((3, env.M)))
        _, rr, *_ = ref.step(act)
        r, _ = env.step(act)
        assert_state_equal(ref, env.medium[0], env.agents[0], float_exact=True)
        assert abs(rr - r[0]) <= 1e-10 * max(1.0, abs(rr))


def test_float32_gradient_cache_on_adversarial_fields(tuning):
    """Gradients that underflow / overflow float32, sit on the clip threshold, are exactly symmetric or exactly zero:
    with the float32 cache every such slot must be deferred (and re-sampled in float64) exactly when it matters.
    The state after each of 6 steps must equal the float64-cache run bit for bit, LEAN and general kernel."""
    shape = (64, 64)
    yy, xx = np.meshgrid(np.arange(64), np.arange(64))
    blobs = np.zeros(shape)
    for cx, cy in [(16, 16), (48, 16), (16, 48), (48, 48), (32, 32)]:
        blobs += np.maximum(0, 6 - np.maximum(abs(xx - cx), abs(yy - cy))) * 0.25
    rng = np.random.default_rng(5)
    noise = rng.random(shape)
    scales = [1e-300, 1e-50, 1e-46, 1e-44, 1e-40, 1e-10, 3e-6, 1e-5, 1.0, 1e20, 1e37, 1e39, 1e200, 0.0, 1e-5, 2e-5]
    chem = blobs.copy()
    for k, sc in enumerate(scales):                       # 4-row bands of noise at wildly different magnitudes
        chem[4 * k:4 * k + 4, 32:] = noise[4 * k:4 * k + 4, 32:] * sc
    chem[40:, :8] = 3e-6
    outs, counts = {}, {}
    f32_0 = S.lib().die_get_counter(b"forward_lean_f32")
    for f32 in (0, 1):
        for lean in (0, 1):
            tuning("grad_f32", f32)
            tuning("fwd_lean", lean)
            (ref,), env = make_pair(shape, seed=3)
            env.medium[0, 2] = chem
            m = env.M
            ga = S.SimGradientAgent(m, seed=2, **PHYS)
            ga.theta[0] = np.random.default_rng(0).integers(-6, 7, m) * np.radians(30)
            trace = []
            for it in range(6):
                act = ga.forward(env).copy()
                env.step(act)
                trace.append((act, ga.theta.copy(), env.medium.copy(), env.agents.copy()))
            outs[(f32, lean)] = trace
            assert S.lib().die_env_gradient_kind(env.handle) == (2 if f32 else 1)
            counts[(f32, lean)] = S.lib().die_get_counter(b"forward_lean_f32")
    assert counts[(1, 0)] == f32_0 and counts[(1, 1)] == f32_0 + 5       # LEAN + float32 cache: steps 2..6
    base = outs[(0, 0)]
    for key, trace in outs.items():
        for it, (a, b) in enumerate(zip(base, trace)):
            for x, y, what in zip(a, b, ("action", "theta", "medium", "agents")):
                assert np.array_equal(x, y, equal_nan=True), f"grad_f32, fwd_lean = {key}: {what} differs at step {it}"


def test_fenced_memory_faults_on_an_overrun():
    """The emulator's memcheck: arrays end at an inaccessible page, so a kernel told to process one element more
    than the array holds dies with SIGSEGV (in a child process) instead of silently reading a neighbour."""
    import subprocess, sys, textwrap
    code = textwrap.dedent('''
        import sys
        import numpy as np
        from tests.hostsim import sim as S
        n = int(sys.argv[1])
        x = S.fenced((512,), fill=0.5); s = S.fenced((512,)); c = S.fenced((512,))
        S.check(S.lib().die_math_sincos(S.ptr(x), S.ptr(s), S.ptr(c), n, None))
        print("ok", float(s[0]))
    ''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ok = subprocess.run([sys.executable, "-c", code, "512"], cwd=root, capture_output=True, text=True)
    assert ok.returncode == 0 and ok.stdout.startswith("ok"), ok.stderr
    bad = subprocess.run([sys.executable, "-c", code, "513"], cwd=root, capture_output=True, text=True)
    assert bad.returncode == -11, (bad.returncode, bad.stdout, bad.stderr[-300:])


# ------------------------------------------------------------------------------------------
# the remaining branches (line coverage of the kernel sources under the emulator: tools/REPRODUCE.md)
# ------------------------------------------------------------------------------------------
def test_const_agent_and_moves_longer_than_the_field():
    """ConstAgent.forward (core/agent/static.py:19-28) with a step of 2.3 field lengths: `% 1.` goes through the
    generic fmod path of die_np_remainder; the oracle agrees bit for bit."""
    (ref,), env = make_pair((24, 32), seed=4)
    m = env.M
    action = S.fenced((1, 3, m), fill=np.nan)
    S.check(S.lib().die_const_forward(S.ptr(action), m, 1, 2.3, -3.7, 0.25, None))
    ra = R.ConstAgent((2.3, -3.7), 0.25)
    ract = ra.forward(ref._get_current_obs)
    assert np.array_equal(ract, action[0])
    for it in range(4):
        _, rr, *_ = ref.step(ract)
        r, _ = env.step(action)
        assert np.array_equal(ref_cells_linear(ref), env.cells()[0])
        assert_state_equal(ref, env.medium[0], env.agents[0], float_exact=True)
        assert _rel(rr, r[0]) < 1e-11


@pytest.mark.parametrize("kw", [dict(grad_clip=None), dict(normalized_grad=False), dict(normalized_grad=False, grad_clip=None)])
def test_gradient_processing_variants(portable_math, kw):
    """grad_clip=None (nan_to_num of 0/0) and unnormalised gradients: the quick turn decision is not offered there,
    every slot takes the reference arithmetic."""
    _physarum_free_run((28, 44), 8, dict(PHYS, **kw), seed=12, exact=kw.get('normalized_grad', True))


def test_unnormalised_zero_gradient_headings_match_numpy_itself():
    """The oracle in its DEFAULT math backend (numpy's own complex arithmetic, as the reference): a zero gradient with
    normalized_grad=False gives g' = 0 * exp(1j d) = (0 c - 0 s) + 1j (0 s + 0 c) -- numpy promotes the real radius to a
    complex number -- and the signs of those zeros, of 0 * prev_grad and of 0 * noise decide whether the next heading
    is 0 or pi.  No libm function is involved at step 0 (chem1 == 0), so the comparison is exact, signs included."""
    (ref,), env = make_pair((32, 48), seed=12)
    m = env.M
    theta0, prev = lattice_theta(m, 30, 12)
    kw = dict(PHYS, normalized_grad=False)
    ra = R.PhysarumAgent(max_agents=m, prev_grad=prev, **kw)
    ga = S.SimGradientAgent(m, **kw)
    ga.theta[0], ga.prev_grad[0] = theta0, prev
    rng = np.random.default_rng(0)
    for variant in range(4):                               # all sign patterns of prev / noise
        coin = rng.integers(0, 2, m)
        noise = rng.normal(0., 0.4, size=(2, m)) * (1 if variant % 2 == 0 else -1)
        ra._prev_grad = np.abs(ra._prev_grad) * (1 if variant < 2 else -1) * np.where(rng.random((2, m)) < 0.5, 1, -1)
        ga.prev_grad[0] = ra._prev_grad
        ra._direction_rads = theta0.copy()
        ga.theta[0] = theta0
        ract = ra.forward(ref._get_current_obs, coin=coin.copy(), noise=noise.copy())
        gact = ga.forward(env, coin=coin, noise=noise, use_hints=False)[0]
        assert np.array_equal(ga.theta[0], ra._direction_rads)
        assert np.array_equal(gact, ract) and np.array_equal(np.signbit(gact), np.signbit(ract))
        assert np.array_equal(np.signbit(ga.prev_grad[0]), np.signbit(ra._prev_grad))


def test_gradient_agent_in_kernel_noise_is_reproducible():
    """GradientAgent with rng='philox': Box-Muller noise drawn in the kernel, keyed on (seed, step, env, slot)."""
    outs = []
    for rep in range(2):
        (ref,), env = make_pair((24, 32), seed=6)
        ga = S.SimGradientAgent(env.M, seed=9, discrete_turn=False, scale=0.01, inertia=0.9, sense_offset=0.02, noise_scale=0.025)
        ga.prev_grad[...] = np.random.default_rng(1).normal(0, 0.4, ga.prev_grad.shape)
        for it in range(4):
            env.step(ga.forward(env))
        outs.append((ga.prev_grad.copy(), ga.theta.copy(), env.agents.copy()))
    assert all(np.array_equal(a, b) for a, b in zip(*outs))
    assert np.isfinite(outs[0][0]).all() and np.std(outs[0][0]) > 0.01


def test_abandoned_speculation_leaves_no_trace():
    """The forward kernel evaluated move + claims speculatively (DIE_FWD_SPECULATE_MOVE) but the step does not adopt
    them (another action arrives): die_env_step_flags must drop the pending claims and run the plain step; a second
    speculation replaces the first; die_env_pending_move tracks the state."""
    phys = dict(scale=0.02, turn_angle=30, sense_offset=0.06)
    outs = []
    for abandon in (False, True):
        (ref,), env = make_pair((40, 72), seed=8)
        ga = S.SimGradientAgent(env.M, seed=3, **phys)
        ga.theta[0] = lattice_theta(env.M, 30, 8)[0]
        other = S.SimGradientAgent(env.M, seed=3, **phys)
        other.theta[0] = ga.theta[0]
        for it in range(6):
            if abandon:
                ga.fuse_move = True
                spec = ga.forward(env).copy()                       # speculation pending ...
                assert S.lib().die_env_pending_move(env.handle) == 1
                if it % 2 == 0:
                    ga.theta[...] = other.theta                     # ... repeated forward: the first speculation is replaced
                    ga.step_no -= 1
                    spec2 = ga.forward(env).copy()
                    assert np.array_equal(spec, spec2)
                other.theta[...] = ga.theta
                env.step(spec, flags=L.STEP_ALIVE_BITS)             # ... and abandoned: plain step on a copy of the action
                assert S.lib().die_env_pending_move(env.handle) == 0
            else:
                ga.fuse_move = False
                act = ga.forward(env).copy()
                other.theta[...] = ga.theta
                env.step(act)
        outs.append((env.medium.copy(), env.agents.copy(), ga.theta.copy(), env.cells()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_physarum_free_run_committed_move(portable_math):
    """DIE_FWD_COMMIT_MOVE (the run-loop contract): the forward launch also stores the moved positions, the adopting step
    runs the field pass + the plain feed kernel only.  Against the oracle, injected coins (the general MOVE kernel)."""
    n0 = S.lib().die_get_counter(b"step_committed")
    _physarum_free_run((40, 72), 20, dict(scale=0.02, turn_angle=30, sense_offset=0.06), seed=8, fuse='commit')
    assert S.lib().die_get_counter(b"step_committed") == n0 + 20


@pytest.mark.parametrize("shape,batch,kw,pair", [
    ((40, 72), None, {}, 0), ((33, 47), 3, dict(diffuse_sigma=1.0), 0), ((64, 64), 2, {}, 2),
    ((24, 50), None, dict(boundary=D.BoundaryCondition.limit), 0), ((6, 90), 2, dict(food_infinite=True), 2)])
@pytest.mark.parametrize("grad_f32", [1, 0])
def test_committed_move_equals_the_plain_loop(tuning, shape, batch, kw, pair, grad_f32):
    """In-kernel Philox coins, the env's hints valid: the steady-state (LEAN) instantiation with the move folded in
    (float64 / float32 gradient cache, with and without the food handed over by pair mode) -- the whole loop equals
    forward + move_claim + field + feed bit for bit, every step."""
    tuning("grad_f32", grad_f32)
    tuning("pair_mode", pair)
    outs = []
    so = S.lib()
    lm0, sc0 = so.die_get_counter(b"forward_lean_move"), so.die_get_counter(b"step_committed")
    for fuse in (False, 'commit'):
        refs, env = make_pair(shape, seed=5, dynamics_kw=kw, batch=batch)
        B = env.B
        ga = S.SimGradientAgent(env.M, B=B, seed=1, **PHYS)
        for b in range(B):
            ga.theta[b] = lattice_theta(env.M, 30, 5 + b)[0]
        ga.fuse_move = fuse
        trace = []
        for it in range(10):
            act = ga.forward(env).copy()
            if fuse:
                assert so.die_env_pending_move(env.handle) == 2
            r, alive = env.step(act, flags=L.STEP_ALIVE_BITS | (L.STEP_ADOPT_MOVE if fuse else 0))
            trace.append((act, r.copy(), alive.copy(), env.agents.copy(), env.cells().copy()))
        outs.append((env.medium.copy(), ga.theta.copy(), trace))
    assert so.die_get_counter(b"step_committed") == sc0 + 10
    assert so.die_get_counter(b"forward_lean_move") == lm0 + 9       # the first forward has no hints yet: general kernel
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    for it, (a, b) in enumerate(zip(outs[0][2], outs[1][2])):
        for x, y, what in zip(a, b, ("action", "reward", "num_agents", "agents", "cells")):
            assert np.array_equal(x, y), f"step {it}: {what} differs"


@pytest.mark.parametrize("shape,batch,kw,pair,fuse", [
    ((40, 72), None, {}, 0, False), ((33, 47), 3, dict(op_action_cost=D.zero_cost), 0, False), ((64, 64), 2, {}, 2, False),
    ((24, 50), None, dict(boundary=D.BoundaryCondition.limit), 0, 'commit'), ((6, 90), 2, dict(food_infinite=True), 2, 'commit')])
def test_cost_hint_equals_the_plain_feed(tuning, shape, batch, kw, pair, fuse):
    """DIE_FWD_WRITE_COST / DIE_STEP_USE_COST: the forward kernel leaves linear_action_cost of its action, the feed kernel
    reads it instead of dx, dy, deposit -- every output equal bit for bit, with pair mode and the committed move too; a step
    with the permission but no pending hint (a Brownian action) runs the plain kernel."""
    tuning("pair_mode", pair)
    outs = []
    so = S.lib()
    n0 = so.die_get_counter(b"feed_cost")
    for cost in (False, True):
        refs, env = make_pair(shape, seed=5, dynamics_kw=kw, batch=batch)
        B = env.B
        ga = S.SimGradientAgent(env.M, B=B, seed=1, **PHYS)
        for b in range(B):
            ga.theta[b] = lattice_theta(env.M, 30, 5 + b)[0]
        ga.fuse_move, ga.write_cost = fuse, cost
        trace = []
        for it in range(8):
            if it == 5:
                act = np.stack([S.brownian_forward(env.agents[b], move_scale=0.02, seed=4 + b, step=it) for b in range(B)])
                adopt = 0
            else:
                act = ga.forward(env).copy()
                adopt = L.STEP_ADOPT_MOVE if (fuse and so.die_env_pending_move(env.handle)) else 0
            r, alive = env.step(act, flags=L.STEP_ALIVE_BITS | adopt | (L.STEP_USE_COST if cost else 0))
            trace.append((r.copy(), alive.copy(), env.agents.copy()))
        outs.append((env.medium.copy(), ga.theta.copy(), trace))
    assert so.die_get_counter(b"feed_cost") == n0 + 7          # every Physarum step of the second run, not the Brownian one
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    for it, (a, b) in enumerate(zip(outs[0][2], outs[1][2])):
        for x, y, what in zip(a, b, ("reward", "num_agents", "agents")):
            assert np.array_equal(x, y), f"step {it}: {what} differs"


def test_committed_move_must_be_adopted():
    """A committed move cannot be withdrawn: a plain step, a host-buffer step or another forward on the env fail by name
    until the adopting step has run; die_env_discard_move (the reset paths) forgets it."""
    so = S.lib()
    (ref,), env = make_pair((24, 32), seed=2)
    ga = S.SimGradientAgent(env.M, seed=3, **PHYS)
    ga.theta[0] = lattice_theta(env.M, 30, 8)[0]
    env.step(ga.forward(env).copy())
    ga.fuse_move = 'commit'
    before = env.agents.copy()
    act = ga.forward(env).copy()
    assert so.die_env_pending_move(env.handle) == 2
    assert not np.array_equal(before[0, :2], env.agents[0, :2])           # the positions are already those of the next step
    assert np.array_equal(before[0, 2:], env.agents[0, 2:])
    with pytest.raises(RuntimeError, match="committed move is pending"):
        env.step(act)                                                       # plain step
    with pytest.raises(RuntimeError, match="committed move is pending"):
        ga.forward(env)                                                     # another forward
    with pytest.raises(RuntimeError, match="committed move"):
        env.step_host(act)
    env.step(act, flags=L.STEP_ALIVE_BITS | L.STEP_ADOPT_MOVE)
    assert so.die_env_pending_move(env.handle) == 0
    ga.forward(env)
    assert so.die_env_pending_move(env.handle) == 2
    S.check(so.die_env_discard_move(env.handle, None))
    assert so.die_env_pending_move(env.handle) == 0


# ------------------------------------------------------------------------------------------
# JonesAgent: the classic three-sensor particle (optional); specification = oracle/die_ref.py:JonesAgent
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("field,kw,dyn", [
    ((40, 72), dict(scale=0.02, sense_offset=0.06), {}),
    ((33, 47), dict(scale=0.015, sense_offset=0.09, turn_angle=22.5, sense_angle=45), {}),
    ((24, 50), dict(scale=0.03, sense_offset=0.05, turn_angle=60, sense_angle=30, deposit=2.0),
     dict(boundary=D.BoundaryCondition.limit, diffuse_sigma=0.8))])
def test_jones_agent_free_run(portable_math, field, kw, dyn):
    """Free run against the oracle in its portable math backend: heading, action, and the whole env state bit-exact
    every step; all five branches of the sensor rule occur."""
    ref_dyn = dict(dyn)
    if 'boundary' in ref_dyn:
        ref_dyn['boundary'] = 'limit'
    (ref,), env = make_pair(field, seed=6, ratio=0.25, dynamics_kw=dyn, ref_dynamics_kw=ref_dyn)
    m = ref.agents.shape[-1]
    tr = np.radians(kw.get('turn_angle', 45))
    theta0 = (lattice_theta(m, 30, 6)[0] // tr) * tr
    ra = R.JonesAgent(max_agents=m, theta0=theta0, **kw)
    ga = S.SimJonesAgent(m, **kw)
    ga.theta[0] = theta0
    rng = np.random.default_rng(3)
    robs = ref._get_current_obs
    turns = set()
    for it in range(25):
        coin = rng.integers(0, 2, m)
        before = ra._direction_rads.copy()
        ract = ra.forward(robs, coin=coin.copy())
        gact = ga.forward(env, coin=coin)[0]
        assert np.array_equal(ga.theta[0], ra._direction_rads), f"heading differs at step {it}"
        assert np.array_equal(gact, ract), f"action differs at step {it}"
        turns |= set(np.unique(np.round(R.renormalize_radians(ra._direction_rads - before) / tr)).astype(int))
        robs, rr, _, _, rinfo = ref.step(ract)
        r, alive = env.step(gact)
        assert rinfo['num_agents'] == alive[0] and _rel(rr, r[0]) < 1e-11
        assert_state_equal(ref, env.medium[0], env.agents[0], float_exact=True)
    assert {-1, 0, 1} <= turns


def test_jones_agent_philox_batch_and_float32_medium():
    """In-kernel coins: reproducible, different per environment and per call; a float32 medium is read as such."""
    refs, env = make_pair((32, 48), seed=9, ratio=0.3, batch=2)
    kw = dict(scale=0.02, sense_offset=0.07)
    outs = []
    for rep in range(2):
        ga = S.SimJonesAgent(env.M, B=2, seed=11, **kw)
        acts = [ga.forward(env).copy(), ga.forward(env).copy()]
        outs.append((acts, ga.theta.copy()))
    assert all(np.array_equal(a, b) for a, b in zip(outs[0][0], outs[1][0])) and np.array_equal(outs[0][1], outs[1][1])
    # a float32 medium (the env's float32 field mode) is read as float32: same result as a float64 medium holding the
    # float32-rounded values
    med = np.stack([r.medium for r in refs]).copy()
    med[:, 2] = np.random.default_rng(1).random(med[:, 2].shape)          # some chem to sense
    med32 = med.astype(np.float32)
    ag = np.stack([r.agents for r in refs])
    e32 = S.SimEnv((32, 48), med32, ag, D.Dynamics(), batch=2, field_dtype=np.float32)
    e64 = S.SimEnv((32, 48), med32.astype(np.float64), ag, D.Dynamics(), batch=2)
    g32, g64 = S.SimJonesAgent(env.M, B=2, seed=11, **kw), S.SimJonesAgent(env.M, B=2, seed=11, **kw)
    assert np.array_equal(g32.forward(e32), g64.forward(e64)) and np.array_equal(g32.theta, g64.theta)
    assert len(np.unique(np.round(g32.theta, 6))) >= 3
    # two identical environments draw different coins (where the front sensor reads the smallest value)
    med[1], ag[1] = med[0], ag[0]
    twin = S.SimEnv((32, 48), med, ag, D.Dynamics(), batch=2)
    gt = S.SimJonesAgent(env.M, B=2, seed=11, **kw)
    a = gt.forward(twin).copy()
    assert not np.array_equal(a[0], a[1]) and np.array_equal(a[:, 2], a[::-1, 2])


def test_host_step_keeps_the_alive_channel_where_it_is(tuning):
    """DIE_HOST_KEEP_ALIVE_CHANNEL: x, y and agent_food are copied into the caller's buffer, the alive channel is left
    alone (the poison written over it survives), everything else equals the full download; refused with agents_die."""
    tuning("host_chunk_min_kb", 0)
    try:
        refs, env = make_pair((24, 32), seed=17, batch=9)
        refs2, env2 = make_pair((24, 32), seed=17, batch=9)
        rng = np.random.default_rng(0)
        keep = None
        for it in range(4):
            act = np.stack([S.brownian_forward(env.agents[b], move_scale=0.02, seed=4 + b, step=it) for b in range(9)])
            (full_agents, full_medium), r_full, _ = env2.step_host(act)
            if keep is None:
                (keep, med), r, _ = env.step_host(act)
            else:
                keep[:, 2] = -7.0                                   # poison: must not be overwritten
                (_, med), r, _ = env.step_host(act, agents_host=keep, flags=L.HOST_KEEP_ALIVE_CHANNEL)
                assert (keep[:, 2] == -7.0).all()
                keep[:, 2] = full_agents[:, 2]
            assert np.array_equal(keep, full_agents) and np.array_equal(med, full_medium) and np.array_equal(r, r_full)
        (_,), dying = make_pair((24, 32), seed=2, dynamics_kw=dict(agents_die=True))
        with pytest.raises(RuntimeError, match="KEEP_ALIVE"):
            dying.step_host(np.zeros((1, 3, dying.M)), agents_host=np.empty_like(dying.agents), flags=L.HOST_KEEP_ALIVE_CHANNEL)
    finally:
        tuning("host_chunk_min_kb", 32 << 10)


def test_forward_through_host_buffers_is_chunked_and_identical(tuning):
    """die_gradient_forward_host: observation uploaded / action downloaded chunk by chunk on two streams; the in-kernel
    random draws are keyed on the GLOBAL environment index, so chunking does not change them."""
    tuning("host_chunk_min_kb", 0)
    try:
        refs, env = make_pair((24, 32), seed=17, batch=9)
        B, M = 9, env.M
        p = S.gradient_params(**PHYS)
        theta0 = np.stack([lattice_theta(M, 30, b)[0] for b in range(B)])
        so = S.lib()
        ctx = S.C.c_void_p()
        S.check(so.die_host_ctx_create(S.C.byref(ctx)))
        res = []
        for host in (False, True):
            theta, action = S.fenced_copy(theta0), S.fenced((B, 3, M), fill=np.nan)
            if host:
                ag_stage, med_stage = S.fenced(env.agents.shape, fill=np.nan), S.fenced(env.medium.shape, fill=np.nan)
                action_host = S.fenced((B, 3, M), fill=np.nan)
                S.check(so.die_gradient_forward_host(ctx, S.C.byref(p), 24, 32, M, B, S.ptr(env.agents), S.ptr(env.medium),
                                                     S.ptr(ag_stage), S.ptr(med_stage), S.ptr(theta), None, S.ptr(action),
                                                     S.ptr(action_host), None, None, None, 7, 3, None))
                assert np.array_equal(action_host, action)
            else:
                S.check(so.die_gradient_forward(S.C.byref(p), 24, 32, M, B, S.ptr(env.agents), S.ptr(env.medium), S.ptr(theta),
                                                None, S.ptr(action), None, None, None, None, None, 7, 3, None))
            res.append((action.copy(), theta.copy()))
        S.check(so.die_host_ctx_destroy(ctx))
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    finally:
        tuning("host_chunk_min_kb", 32 << 10)


# ------------------------------------------------------------------------------------------
# seeded random sweep over configurations: emulated kernels vs the oracle, bit for bit
# ------------------------------------------------------------------------------------------
def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    h = int(rng.choice([2, 3, 5, 8, 17, 31, 32, 33, 40, 64, 65, 70]))
    w = int(rng.choice([2, 3, 7, 16, 63, 64, 65, 72, 76, 100, 129, 200]))
    sigma = float(rng.choice([0.1, 0.3, 0.5, 0.5, 0.8, 1.0, 1.3]))
    mode = str(rng.choice(['wrap', 'wrap', 'wrap', 'reflect', 'nearest', 'mirror', 'constant']))
    limit = bool(rng.random() < 0.3)
    dyn = dict(diffuse_sigma=sigma, diffuse_mode=mode, food_infinite=bool(rng.random() < 0.2),
               rate_feed=float(rng.choice([0.1, 0.25])), rate_decay_chem=float(rng.choice([0.1, 0.02])))
    rdyn = dict(dyn)
    if limit:
        dyn['boundary'], rdyn['boundary'] = D.BoundaryCondition.limit, 'limit'
    agent = dict(scale=float(rng.choice([0.007, 0.03, 0.2, 1.7])), sense_offset=float(rng.choice([0.0, 0.04, 0.3, 1.5])),
                 turn_angle=float(rng.choice([30, 35, 45, 90])), sense_angle=float(rng.choice([60, 90, 120, 170])),
                 turn_tolerance=float(rng.choice([0.05, 0.1, 0.3])), deposit=float(rng.choice([4.0, 0.5])))
    tune = dict(field_vec=int(rng.random() < 0.7), step_impl=int(rng.random() < 0.3), grad_f32=int(rng.random() < 0.5), fwd_lean=int(rng.random() < 0.7),
                feed_bits=int(rng.random() < 0.7))
    return (h, w), float(rng.choice([0.05, 0.1, 0.5, 1.0])), dyn, rdyn, agent, tune, bool(rng.random() < 0.5)


@pytest.mark.parametrize("seed", range(int(os.environ.get("DIE_SWEEP_SEEDS", "80"))))
def test_random_configurations_against_the_oracle(portable_math, tuning, seed):
    field, ratio, dyn, rdyn, agent_kw, tune, hints = _random_case(seed)
    if agent_kw['sense_angle'] <= agent_kw['turn_angle'] * agent_kw['turn_tolerance'] / 0.99 + 1:
        agent_kw['sense_angle'] = 170.0
    for k, v in tune.items():
        tuning(k, v)
    (ref,), env = make_pair(field, seed=seed, ratio=ratio, dynamics_kw=dyn, ref_dynamics_kw=rdyn)
    m = env.M
    theta0, prev = lattice_theta(m, agent_kw['turn_angle'], seed)
    ra = R.PhysarumAgent(max_agents=m, prev_grad=prev, **agent_kw)
    ga = S.SimGradientAgent(m, **agent_kw)
    ga.theta[0] = theta0
    # (a second generator, so that the cases of the seeds above stay what they were) the round-2 variants: cost hint,
    # committed move, threads per field tile, the feed kernel's register cap
    extra = np.random.default_rng(9000 + seed)
    ga.write_cost = bool(extra.random() < 0.6)
    ga.fuse_move = 'commit' if (hints and tune['feed_bits'] and extra.random() < 0.4) else False
    tuning("field_tile", int(extra.integers(0, 3)))
    tuning("feed_min_blocks", int(extra.choice([0, 6])))
    rng = np.random.default_rng(seed)
    robs = ref._get_current_obs
    for it in range(4):
        coin = rng.integers(0, 2, m)
        ract = ra.forward(robs, coin=coin.copy())
        gact = ga.forward(env, coin=coin, use_hints=hints)[0]
        assert np.array_equal(ga.theta[0], ra._direction_rads), f"theta differs at step {it}"
        assert np.array_equal(gact, ract), f"action differs at step {it}"
        robs, rr, _, _, rinfo = ref.step(ract)
        adopt = L.STEP_ADOPT_MOVE if S.lib().die_env_pending_move(env.handle) else 0
        r, alive = env.step(gact, flags=(L.STEP_ALIVE_BITS if tune['feed_bits'] else 0) | adopt
                            | (L.STEP_USE_COST if ga.write_cost else 0))
        assert np.array_equal(ref_cells_linear(ref), env.cells()[0]), f"cells differ at step {it}"
        assert rinfo['num_agents'] == alive[0]
        assert _rel(rr, r[0]) < 1e-10 or abs(rr - r[0]) < 1e-9
        assert_state_equal(ref, env.medium[0], env.agents[0], float_exact=True)


@pytest.mark.parametrize("seed", range(int(os.environ.get("DIE_SWEEP_SEEDS", "40"))))
def test_random_batches_slot_counts_and_policies(portable_math, tuning, seed):
    """Second sweep: batches of environments, slot counts below / above the cell count (M != H*W), Brownian and
    Physarum steps interleaved on the same environments, every kernel variant -- against one oracle per environment."""
    rng = np.random.default_rng(5000 + seed)
    h, w = int(rng.choice([3, 8, 20, 33, 40])), int(rng.choice([4, 16, 65, 72, 100]))
    B = int(rng.choice([1, 2, 3, 5]))
    C = h * w
    m = int(rng.choice([max(C // 3, 1), C, C, 2 * C + 7]))
    sigma = float(rng.choice([0.3, 0.5, 0.8]))
    for k, v in dict(field_vec=int(rng.random() < 0.7), step_impl=int(rng.random() < 0.3), grad_f32=int(rng.random() < 0.7), fwd_lean=int(rng.random() < 0.5),
                     feed_bits=int(rng.random() < 0.7)).items():
        tuning(k, v)
    refs, mediums, agentss = [], [], []
    for b in range(B):
        np.random.seed(seed * 10 + b)
        r0 = R.Env((h, w), R.Dynamics(init_agent_ratio=0.2, diffuse_sigma=sigma), noise_seed=seed * 10 + b)
        ag = np.zeros((4, m))
        n = min(m, C)
        ag[:, :n] = r0.agents[:, :n]
        refs.append(R.Env((h, w), R.Dynamics(init_agent_ratio=0.2, diffuse_sigma=sigma), medium=r0.medium, agents=ag))
        mediums.append(r0.medium)
        agentss.append(ag)
    env = S.SimEnv((h, w), np.stack(mediums), np.stack(agentss), D.Dynamics(init_agent_ratio=0.2, diffuse_sigma=sigma), batch=B)
    th = [lattice_theta(m, 30, seed + b) for b in range(B)]
    ras = [R.PhysarumAgent(max_agents=m, prev_grad=th[b][1], **PHYS) for b in range(B)]
    rb = R.BrownianAgent(0.03)
    ga = S.SimGradientAgent(m, B=B, **PHYS)
    for b in range(B):
        ga.theta[b] = th[b][0]
    for it in range(5):
        if rng.random() < 0.6:
            coin = rng.integers(0, 2, (B, m))
            racts = [ras[b].forward(refs[b]._get_current_obs, coin=coin[b].copy()) for b in range(B)]
            gact = ga.forward(env, coin=coin).copy()
            for b in range(B):
                assert np.array_equal(ga.theta[b], ras[b]._direction_rads), f"theta, env {b}, step {it}"
        else:
            u = rng.random((B, 3, m))
            racts = [rb.forward(refs[b]._get_current_obs, u=u[b]) for b in range(B)]
            gact = S.brownian_forward(env.agents, move_scale=0.03, u=u)
        for b in range(B):
            assert np.array_equal(gact[b], racts[b]), f"action, env {b}, step {it}"
        outs = [refs[b].step(racts[b]) for b in range(B)]
        r, alive = env.step(gact, flags=L.STEP_ALIVE_BITS if rng.random() < 0.7 else 0)
        for b in range(B):
            assert np.array_equal(ref_cells_linear(refs[b]), env.cells()[b]), f"cells, env {b}, step {it}"
            assert outs[b][4]['num_agents'] == alive[b]
            assert _rel(outs[b][1], r[b]) < 1e-10 or abs(outs[b][1] - r[b]) < 1e-9
            assert_state_equal(refs[b], env.medium[b], env.agents[b], float_exact=True)


@pytest.mark.parametrize("field,sigma,batch", [((24, 40), 0.5, None), ((37, 53), 0.5, None), ((20, 20), 0.1, None),
                                               ((40, 64), 0.5, 3), ((33, 70), 0.8, 2)])
def test_tabulated_food_flow(tuning, field, sigma, batch):
    """op_food_flow of ANY FieldSequence (the reference's PerlinNoiseSequence, a user's own): the host tabulates
    sequence[t] for every time step, the field pass reads one value per cell: scale * F_k + (1 - decay) * food, k cycling
    (die_env_set_food_frames).  T = 4 frames over 10 steps: the iterator wraps twice; a batch shares the sequence."""
    rng = np.random.default_rng(9)
    frames = rng.normal(0.0, 0.3, size=(4, *field)).round(3)
    gflow = D.TabulatedSequence(frames).get_flow_operator(scale=0.5, decay=0.25)
    gflow.calls = 2                                       # start in the middle of the sequence
    refs = []
    for b in range(batch or 1):
        np.random.seed(5 + b)
        rflow = R.FrameSequence(frames).get_flow_operator(scale=0.5, decay=0.25, k0=2)
        refs.append(R.Env(field, R.Dynamics(init_agent_ratio=0.1, op_food_flow=rflow, diffuse_sigma=sigma), noise_seed=5 + b))
    env = S.SimEnv(field, np.stack([r.medium for r in refs]), np.stack([r.agents for r in refs]),
                   D.Dynamics(init_agent_ratio=0.1, diffuse_sigma=sigma), batch=batch)
    env.set_food_frames(gflow)
    ra = R.BrownianAgent(0.01)
    for it in range(10):
        u = rng.random((env.B, 3, env.M))
        acts = np.stack([ra.forward(refs[b]._get_current_obs, u=u[b]) for b in range(env.B)])
        rr = [refs[b].step(acts[b])[1] for b in range(env.B)]
        r, _ = env.step(acts)
        for b in range(env.B):
            assert_state_equal(refs[b], env.medium[b], env.agents[b], float_exact=True)
            assert abs(rr[b] - r[b]) <= 1e-10 * max(1.0, abs(rr[b]))
    # back to the identity flow
    S.check(S.lib().die_env_set_food_frames(env.handle, None, 0, 0, 0.0, 0.0))
    food = env.medium[:, 1].copy()
    env.step(np.zeros((env.B, 3, env.M)))
    occ = env.medium[:, 0]
    assert np.array_equal(env.medium[:, 1], food - (0.1 * food) * occ)


def test_api_argument_checks_profiling_and_set_dynamics():
    """Host logic of the C ABI, the same C++ the CUDA build compiles: invalid arguments come back as DIE_E_INVALID with
    a message naming the violated condition (nothing is launched), the per-kernel profiling counts steps, and
    die_env_set_dynamics changes the dynamics of a live handle."""
    so = S.lib()
    (ref,), env = make_pair((24, 32), seed=2)
    m = env.M
    act = S.fenced((1, 3, m), fill=0.0)
    # same buffer as input and output medium
    rc = so.die_env_step(env.handle, S.ptr(env.medium), S.ptr(env.medium), S.ptr(env.agents), S.ptr(act),
                         S.ptr(env.reward), S.ptr(env.alive), None)
    assert rc != 0 and b"medium_in != medium_out" in so.die_last_error()
    # adopting a move nobody speculated
    S.check(so.die_env_refresh_alive(env.handle, S.ptr(env.agents), None))
    rc = so.die_env_step_flags(env.handle, S.ptr(env.medium_buf[0]), S.ptr(env.medium_buf[1]), S.ptr(env.agents), S.ptr(act),
                               S.ptr(env.reward), S.ptr(env.alive), L.STEP_ADOPT_MOVE, None)
    assert rc != 0 and b"no speculative move" in so.die_last_error()
    h = S.C.c_void_p()
    bad = _dyn = D.env._dynamics_to_c(D.Dynamics())
    assert so.die_env_create(1, 32, 10, 1, S.C.byref(bad), S.C.byref(h)) != 0 and not h.value      # H >= 2
    bad.blur_radius = 99
    assert so.die_env_create(8, 8, 10, 1, S.C.byref(bad), S.C.byref(h)) != 0 and b"blur_radius" in so.die_last_error()
    assert so.die_set_tuning(b"no_such_switch", 1) != 0 and b"unknown key" in so.die_last_error()
    assert so.die_get_counter(b"no_such_counter") == -1
    assert so.die_env_destroy(None) == 0
    # profiling: events are recorded between the step's kernels; the emulator's events carry no time
    S.check(so.die_env_set_profiling(env.handle, 1))
    for _ in range(3):
        env.step(act)
    ms = (S.C.c_double * 4)()
    n = S.C.c_int64()
    S.check(so.die_env_kernel_times(env.handle, ms, S.C.byref(n)))
    assert n.value == 3 and list(ms) == [0.0] * 4
    S.check(so.die_env_set_profiling(env.handle, 0))
    # set_dynamics on a live handle: no decay, no feeding -> food untouched, chem mass conserved by the blur
    dyn = D.env._dynamics_to_c(D.Dynamics(rate_feed=0.0, rate_decay_chem=0.0))
    S.check(so.die_env_set_dynamics(env.handle, S.C.byref(dyn)))
    env.medium[0, 2] = np.random.default_rng(0).random((24, 32))
    food, mass = env.medium[0, 1].copy(), env.medium[0, 2].sum()
    env.step(act)
    assert np.array_equal(env.medium[0, 1], food) and abs(env.medium[0, 2].sum() - mass) < 1e-9


@pytest.mark.parametrize("seed", range(int(os.environ.get("DIE_SWEEP_SEEDS", "24")) // 2))
def test_random_slab_worlds(seed):
    """Third sweep: the multi-rank slab world (peer-pointer kernels, owner look-ups, corner mirror) against the single
    env for random rank counts, field shapes, mirror sizes, agent geometry and blur radii."""
    rng = np.random.default_rng(9000 + seed)
    G = int(rng.choice([2, 3, 4, 8]))
    rows_per = int(rng.choice([4, 5, 8, 16]))
    shape = (G * rows_per, int(rng.choice([16, 24, 37, 64])))
    band = int(rng.choice([0, 0, 2, 5, 100]))
    sigma = float(rng.choice([0.3, 0.5, 0.8, 1.0]))
    phys = dict(scale=float(rng.choice([0.007, 0.02, 0.1])), turn_angle=float(rng.choice([30, 45])),
                sense_offset=float(rng.choice([0.0, 0.04, 0.2])))
    dyn = dict(diffuse_sigma=sigma)
    (ref,), env = make_pair(shape, seed=seed, ratio=float(rng.choice([0.05, 0.15, 0.5])), dynamics_kw=dyn)
    m = env.M
    theta0, _ = lattice_theta(m, phys['turn_angle'], seed)
    ga = S.SimGradientAgent(m, **phys)
    ga.theta[0] = theta0
    world = S.SimSlabWorld(env.medium[0], env.agents[0], theta0, G, dynamics=D.Dynamics(**dyn), corner_r=band, **phys)
    for it in range(8):
        coin = rng.integers(0, 2, m)
        act = ga.forward(env, coin=coin)[0]
        world.forward(coin)
        r, alive = env.step(act)
        wr, walive = world.step()
        wmed, wag, wth, wact, wcells = world.gather()
        assert np.array_equal(wact, act), f"action differs at step {it}"
        assert np.array_equal(wcells, env.cells()[0]), f"cells differ at step {it}"
        assert np.array_equal(wmed, env.medium[0]), f"medium differs at step {it}"
        assert np.array_equal(wag, env.agents[0]), f"agents differ at step {it}"
        assert np.array_equal(wth, ga.theta[0]), f"theta differs at step {it}"
        assert walive == alive[0] and abs(wr - r[0]) <= 1e-10 * max(1.0, abs(r[0]))


@pytest.mark.parametrize("shape", [(256, 256), (5, 7), (1031, 33), (1000, 3)])
def test_cells_match_pandas_on_adversarial_coordinates(shape):
    """nearest_cell (die_device.cuh: the add-magic rounding trick, no correction step) against
    pandas.Index.get_indexer(method='nearest') -- what xarray's .sel(method='nearest') runs -- on every grid
    coordinate, every midpoint, their +-3 ulp neighbours and random points, through move_claim."""
    h, w = shape
    rng = np.random.default_rng(h * 131 + w)

    def adversarial(n):
        g = R.grid_coords(n)
        mids = (g[:-1] + g[1:]) / 2
        pts = [g, mids]
        for base in (g, mids):
            up, dn = base.copy(), base.copy()
            for _ in range(3):
                up, dn = np.nextafter(up, 2), np.nextafter(dn, -2)
                pts += [up.copy(), dn.copy()]
        pts.append(rng.uniform(0, 1, 5000))
        return np.clip(np.concatenate(pts), 0.0, 1.0)

    xs, ys = adversarial(h), adversarial(w)
    m = max(len(xs), len(ys))
    xs = np.resize(xs, m)
    ys = np.resize(rng.permutation(ys), m)
    agents = np.zeros((4, m))
    agents[0], agents[1] = xs, ys
    agents[2, ::3] = 1.0
    env = S.SimEnv(shape, np.zeros((3, h, w)), agents, D.Dynamics(boundary=D.BoundaryCondition.limit))
    env.step(np.zeros((3, m)))            # 'limit' boundary + zero action: positions unchanged
    ref = R.nearest_index_pandas(xs, h) * w + R.nearest_index_pandas(ys, w)
    assert np.array_equal(env.cells()[0], ref.astype(np.int32))


def test_sense_cells_clamp_out_of_range():
    """pos + offset outside [0, 1] clamps to the edge cells (Q4), through the forward kernel."""
    from oracle import portable_math as P
    h, w = 64, 48
    rng = np.random.default_rng(0)
    m = 4096
    agents = np.zeros((4, m))
    agents[0], agents[1] = rng.uniform(0, 1, m), rng.uniform(0, 1, m)
    agents[0, :64] = np.linspace(0, 0.02, 64)
    agents[1, 64:128] = np.linspace(0.98, 1.0, 64)
    env = S.SimEnv((h, w), np.zeros((3, h, w)), agents)
    theta = rng.uniform(-np.pi, np.pi, m)
    ga = S.SimGradientAgent(m, sense_offset=0.3, scale=0.01)
    ga.theta[0] = theta
    ga.record_sense_cells = True
    ga.forward(env, coin=np.zeros(m, dtype=np.uint8))
    s, c = P.sincos(theta)
    px, py = agents[0] + 0.3 * c, agents[1] + 0.3 * s
    ref = R.nearest_index_pandas(px, h) * w + R.nearest_index_pandas(py, w)
    assert np.array_equal(ga.sense_cells[0], ref.astype(np.int32))
    assert (px < 0).any() and (px > 1).any() and (py < 0).any() and (py > 1).any()


@pytest.mark.parametrize("sigma,shape", [(1.2, (20, 37)), (1.4, (40, 40)), (1.6, (5, 7)), (1.9, (33, 70)), (2.0, (64, 64))])
def test_large_blur_radii(portable_math, sigma, shape):
    """Radii 5..8 of the tile kernel (the bulk and march kernels stop at 4 / 3), incl. a radius larger than the field."""
    _physarum_free_run(shape, 4, PHYS, seed=16, dynamics_kw=dict(diffuse_sigma=sigma))


def test_reward_reduction_with_thousands_of_partials():
    """finalize_stats_kernel's four-accumulator loop only runs with more than 3072 block partials per environment
    (M > 3.1 M slots: the 4096^2 field): here 3.3 M slots on a tiny field, against numpy's sum of the same terms."""
    h, w, m = 8, 8, 3_300_000
    rng = np.random.default_rng(0)
    medium = np.zeros((3, h, w))
    medium[1] = rng.random((h, w)).round(3)
    agents = np.zeros((4, m))
    n = 40_000
    agents[0, :n], agents[1, :n] = rng.integers(0, h, n) / (h - 1), rng.integers(0, w, n) / (w - 1)
    agents[2, :n], agents[3, :n] = 1.0, 0.5
    env = S.SimEnv((h, w), medium, agents)
    action = np.zeros((3, m))
    action[2] = rng.random(m).round(3)
    action[0] = rng.normal(0, 0.01, m)
    stock0 = env.agents[0, 3].copy()
    r, alive = env.step(action)
    assert alive[0] == n
    gained = env.agents[0, 3] - stock0
    assert abs(r[0] - gained.sum()) <= 1e-9 * abs(gained).sum()
    # die_env_read_stats: the two D2H copies + sync of the synchronous Env.step
    rh, ah = S.fenced((1,), fill=np.nan), S.fenced((1,), np.int64, fill=-1)
    S.check(S.lib().die_env_read_stats(env.handle, S.ptr(env.reward), S.ptr(env.alive), S.ptr(rh), S.ptr(ah), None))
    assert rh[0] == r[0] and ah[0] == n


def test_unknown_boundary_moves_without_wrapping_or_clipping():
    """core/env.py:158-161: an unfamiliar boundary condition logs a warning and does nothing -- positions leave [0, 1]
    and their cells clamp to the edge."""
    (ref,), env = make_pair((16, 16), seed=1)
    dyn = D.env._dynamics_to_c(D.Dynamics())
    dyn.boundary = L.BOUNDARY_NONE
    S.check(S.lib().die_env_set_dynamics(env.handle, S.C.byref(dyn)))
    pos0 = env.agents[0, :2].copy()
    action = np.zeros((3, env.M))
    action[0], action[1] = 0.7, -0.6
    env.step(action)
    assert np.array_equal(env.agents[0, :2], pos0 + np.array([[0.7], [-0.6]]))
    cells = env.cells()[0]
    assert ((cells // 16)[pos0[0] + 0.7 > 1.0] == 15).all() and ((cells % 16)[pos0[1] - 0.6 < 0.0] == 0).all()


@pytest.mark.parametrize("sigma", [0.3, 0.7, 1.0, 1.5])
def test_sense_mask_other_radii(sigma):
    """die_sense_mask is templated on the radius; the reference only uses sigma = 2 (radius 8)."""
    from die_b200.env import gaussian_kernel1d
    import scipy.ndimage as ndi
    (ref,), env = make_pair((30, 44), seed=9, ratio=0.05)
    wts = np.ascontiguousarray(gaussian_kernel1d(sigma))
    obs = S.fenced(env.medium.shape, fill=np.nan)
    S.check(S.lib().die_sense_mask(30, 44, 1, S.ptr(wts), (len(wts) - 1) // 2, S.ptr(env.medium), S.ptr(obs), None))
    mask = np.ceil(ndi.gaussian_filter(env.medium[0, 0], sigma, mode='nearest').round(3)) != 0
    assert np.array_equal(obs[0], np.where(mask[None], env.medium[0], 0.0))


def test_results_do_not_depend_on_block_or_thread_order():
    """The emulator runs blocks in ascending order and a block's threads from 0 up; the hardware promises neither.
    A subset of this file again, in a child process, with the blocks shuffled and the threads scheduled from the last
    one down (HOSTSIM_BLOCK_ORDER / HOSTSIM_THREAD_ORDER): claims by atomicMax, the persistent bulk kernel's tile
    assignment, the slab world's peer accesses and every producer / consumer phase must give the same bits."""
    import subprocess
    import sys
    if os.environ.get("HOSTSIM_BLOCK_ORDER") or os.environ.get("HOSTSIM_THREAD_ORDER"):
        pytest.skip("already running under an alternative order")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HOSTSIM_BLOCK_ORDER="shuffle", HOSTSIM_THREAD_ORDER="reverse", DIE_SWEEP_SEEDS="12")
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "tests/test_hostsim_kernels.py",
                          "-k", "brownian_free_run or with_env_hints or fused_move or bulk_field_kernel_equals or "
                                "slab_world_equals or golden or random_configurations or tuning_switches or "
                                "committed_move_equals or cost_hint_equals or jones_agent_free_run"],
                         cwd=root, env=env, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
