"""The cluster-fused environment step (die_b200/csrc/die_env_fused.cuh: one thread-block cluster per environment,
claim table in distributed shared memory, chem rows by TMA bulk copies) against the three-kernel path and the oracle.

Both paths implement core/env.py:101-131; the fused one must reproduce the other BIT FOR BIT, reward included
(its block partials and final sum follow agent_feed_kernel / finalize_stats_kernel)."""
import numpy as np
import pytest

from tests._parity import assert_state_equal, lattice_theta, make_pair

pytestmark = pytest.mark.gpu
PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)


def _free_run(shape, steps, impl, batch=None, dynamics_kw=None, agent="physarum", seed=13):
    import die_b200 as D
    from die_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.die_set_tuning(b"step_impl", impl))
    try:
        n0 = lib.die_get_counter(b"step_fused")
        refs, env = make_pair(shape, seed=seed, dynamics_kw=dynamics_kw or {}, batch=batch)
        m = env.max_agents
        if agent == "physarum":
            ag = D.PhysarumAgent(max_agents=m, seed=5, **PHYS)
        else:
            ag = D.BrownianAgent(move_scale=0.02, seed=7)
        obs = env._get_current_obs
        rewards = []
        for _ in range(steps):
            obs, r, *_ = env.step(ag.forward(obs))
            rewards.append(np.array(r, dtype=np.float64).copy())
        launched = lib.die_get_counter(b"step_fused") - n0
        theta = ag.get_state()[0] if agent == "physarum" else np.zeros(1)
        return (*env.get_state(), theta, np.array(rewards), env.last_cells().cpu().numpy()), launched
    finally:
        _lib.check(lib.die_set_tuning(b"step_impl", 0))


@pytest.mark.parametrize("shape,batch,sigma", [((256, 256), 5, 0.5), ((256, 256), None, 0.5), ((64, 64), 3, 0.5),
                                               ((48, 80), 7, 0.8), ((24, 40), None, 0.3), ((128, 96), 2, 1.0),
                                               ((96, 130), 2, 0.5), ((32, 512), 2, 0.5), ((50, 64), 3, 0.5)])
@pytest.mark.parametrize("agent", ["physarum", "brownian"])
def test_fused_step_equals_three_kernels(shape, batch, sigma, agent):
    base, n_base = _free_run(shape, 25, 0, batch=batch, dynamics_kw=dict(diffuse_sigma=sigma), agent=agent)
    out, n_fused = _free_run(shape, 25, 1, batch=batch, dynamics_kw=dict(diffuse_sigma=sigma), agent=agent)
    assert n_base == 0 and n_fused == 25, "step_impl must select the path"
    for a, b, what in zip(base, out, ("medium", "agents", "theta", "rewards", "cells")):
        assert np.array_equal(a, b), f"{what} differs between the fused step and the three kernels"


def test_fused_step_falls_back_where_it_does_not_apply():
    """Odd row lengths (rows not 16-byte multiples), non-periodic diffusion, no diffusion, fields too large for a
    cluster's shared memory: step_impl = 1 runs the three kernels."""
    for shape, kw in (((40, 71), {}), ((40, 64), dict(diffuse_mode='reflect')), ((40, 64), dict(diffuse_sigma=0.1)),
                      ((1024, 1024), {})):
        _, n = _free_run(shape, 3, 1, dynamics_kw=kw)
        assert n == 0


@pytest.mark.parametrize("dyn", [dict(), dict(food_infinite=True), dict(boundary="limit")])
def test_fused_step_against_the_oracle(dyn):
    """Brownian free run with injected draws, fused step, every step compared with the oracle bit for bit."""
    import die_b200 as D
    from die_b200 import _lib
    from oracle import die_ref as R
    lib = _lib.load()
    gkw = dict(dyn)
    if "boundary" in gkw:
        gkw["boundary"] = D.BoundaryCondition.limit
    n0 = lib.die_get_counter(b"step_fused")
    _lib.check(lib.die_set_tuning(b"step_impl", 1))
    (ref,), env = make_pair((64, 96), seed=3, dynamics_kw=gkw, ref_dynamics_kw=dyn)
    ra, ga = R.BrownianAgent(0.03), D.BrownianAgent(move_scale=0.03)
    m = env.max_agents
    rng = np.random.default_rng(1)
    robs, gobs = ref._get_current_obs, env._get_current_obs
    for _ in range(20):
        u = rng.random((3, m))
        ract, gact = ra.forward(robs, u=u), ga.forward(gobs, u=u)
        robs, rr, *_ = ref.step(ract)
        gobs, gr, *_ = env.step(gact)
        med, ag = env.get_state()
        assert_state_equal(ref, med, ag, float_exact=True)
        assert abs(rr - gr) <= 1e-10 * max(1.0, abs(rr))
    _lib.check(lib.die_set_tuning(b"step_impl", 0))
    assert lib.die_get_counter(b"step_fused") - n0 == 20


def test_fused_step_with_wave_food_flow():
    import die_b200 as D
    field = (64, 64)
    outs = []
    for impl in (0, 1):
        flow = D.WaveSequence(field, dt=0.01).get_flow_operator(scale=0.5, decay=0.5)
        out, n = _free_run(field, 12, impl, batch=None, dynamics_kw=dict(op_food_flow=flow))
        assert n == (12 if impl else 0)
        outs.append(out)
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
