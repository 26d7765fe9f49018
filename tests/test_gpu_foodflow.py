"""Dynamics.op_food_flow = WaveSequence(...).get_flow_operator(scale, decay) (core/data_init.py:16-47, 71-89;
examples/simple_agents.py:95-100) evaluated inside the CUDA field pass, against the oracle: bit for bit with the
oracle's 'portable' math backend (the per-cell cosine is die_math.h's on both sides), to 1e-13 with numpy's."""
import numpy as np
import pytest

from oracle import die_ref as R
from tests._parity import make_pair, assert_state_equal, ref_cells_linear

pytestmark = pytest.mark.gpu


def _run(field, steps, math, scale, decay, dyn_extra=None, impl=0, sigma=0.5):
    import die_b200 as D
    from die_b200 import _lib
    lib = _lib.load()
    dyn_extra = dyn_extra or {}
    R.set_math_backend(math)
    lib.die_set_step_impl(impl)
    try:
        rflow = R.WaveSequence(field, dt=0.01).get_flow_operator(scale=scale, decay=decay)
        gflow = D.WaveSequence(field, dt=0.01).get_flow_operator(scale=scale, decay=decay)
        (ref,), gpu = make_pair(field, seed=5,
                                dynamics_kw=dict(op_food_flow=gflow, diffuse_sigma=sigma, **dyn_extra),
                                ref_dynamics_kw=dict(op_food_flow=rflow, diffuse_sigma=sigma, **dyn_extra))
        ra, ga = R.BrownianAgent(0.01), D.BrownianAgent(move_scale=0.01)
        m = ref.agents.shape[-1]
        rng = np.random.default_rng(3)
        robs, gobs = ref._get_current_obs, gpu._get_current_obs
        worst = 0.0
        for it in range(steps):
            u = rng.random((3, m))
            ract, gact = ra.forward(robs, u=u), ga.forward(gobs, u=u)
            assert np.array_equal(ract, gact.cpu().numpy())
            robs, rr, *_ = ref.step(ract)
            gobs, gr, *_ = gpu.step(gact)
            med, ag = gpu.get_state()
            assert np.array_equal(ref_cells_linear(ref), gpu.last_cells().cpu().numpy())
            if math == 'portable':
                assert_state_equal(ref, med, ag, float_exact=True)
            else:
                worst = max(worst, np.abs(med[1] - ref.medium[1]).max())
                assert np.array_equal(med[0], ref.medium[0]) and np.array_equal(med[2], ref.medium[2])
                assert np.array_equal(ag[:3], ref.agents[:3])
            assert abs(rr - gr) <= 1e-10 * max(1.0, abs(rr))
        assert gflow.calls == steps
        return worst
    finally:
        R.set_math_backend('numpy')
        lib.die_set_step_impl(0)


@pytest.mark.parametrize("field,impl,sigma", [((48, 64), 0, 0.5), ((37, 53), 0, 0.5), ((64, 40), 1, 0.5),
                                              ((40, 40), 0, 0.1), ((96, 80), 0, 0.8)])
def test_wave_flow_bit_exact_with_portable_math(field, impl, sigma):
    _run(field, 40, 'portable', 0.5, 0.5, impl=impl, sigma=sigma)


def test_wave_flow_with_numpy_math_and_infinite_food():
    worst = _run((48, 64), 60, 'numpy', 1.0, 0.1, dyn_extra=dict(food_infinite=True))
    assert worst < 1e-13, worst


def test_wave_flow_cycles_and_survives_reset():
    """t cycles over np.arange(*t_bounds, dt) (3 values here); reset() keeps the operator's iterator position,
    as the reference's closure does."""
    import die_b200 as D
    field = (32, 32)
    R.set_math_backend('portable')
    try:
        rseq, gseq = R.WaveSequence(field, dt=0.5, t_bounds=(0, 1.5)), D.WaveSequence(field, dt=0.5, t_bounds=(0, 1.5))
        rflow, gflow = rseq.get_flow_operator(0.3, 0.2), gseq.get_flow_operator(0.3, 0.2)
        (ref,), gpu = make_pair(field, seed=8, dynamics_kw=dict(op_food_flow=gflow), ref_dynamics_kw=dict(op_food_flow=rflow))
        ra, ga = R.ConstAgent((0.01, -0.02), 0.3), D.ConstAgent((0.01, -0.02), 0.3)
        for it in range(7):
            ref.step(ra.forward(ref._get_current_obs))
            gpu.step(ga.forward(gpu._get_current_obs))
            assert_state_equal(ref, *gpu.get_state(), float_exact=True)
        assert gflow.calls == 7
        gpu.reset()
        med, ag = gpu.get_state()
        ref.medium[...] = med
        ref.agents[...] = ag
        for it in range(4):                    # continues at time step 7 % 3
            ref.step(ra.forward(ref._get_current_obs))
            gpu.step(ga.forward(gpu._get_current_obs))
            assert_state_equal(ref, *gpu.get_state(), float_exact=True)
    finally:
        R.set_math_backend('numpy')


def test_wave_flow_batched_envs_share_the_sequence():
    import die_b200 as D
    field = (40, 48)
    gflow = D.WaveSequence(field).get_flow_operator(0.5, 0.5)
    (r0, r1), gpu = make_pair(field, seed=2, batch=2, dynamics_kw=dict(op_food_flow=gflow),
                              ref_dynamics_kw={})
    single_flow = D.WaveSequence(field).get_flow_operator(0.5, 0.5)
    single = D.Env(field, D.Dynamics(op_food_flow=single_flow), init_state=(r1.medium, r1.agents))
    ga, gs = D.ConstAgent((0.004, 0.003), 0.2), D.ConstAgent((0.004, 0.003), 0.2)
    for it in range(10):
        gpu.step(ga.forward(gpu._get_current_obs))
        single.step(gs.forward(single._get_current_obs))
    med, ag = gpu.get_state()
    ms, as_ = single.get_state()
    assert np.array_equal(med[1], ms) and np.array_equal(ag[1], as_)


def test_unsupported_flows_are_refused():
    import die_b200 as D
    with pytest.raises(NotImplementedError):
        D.Env((16, 16), D.Dynamics(op_food_flow=lambda f: f * 0.5))
    seq = D.WaveSequence((8, 8))
    with pytest.raises(ValueError):
        D.Env((16, 16), D.Dynamics(op_food_flow=seq.get_flow_operator()))


# ------------------------------------------------------------------------------------------------------------
# Dynamics.diffuse_mode: every boundary extension skimage.filters.gaussian / scipy.ndimage offers
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ['reflect', 'nearest', 'mirror', 'constant', 'wrap'])
@pytest.mark.parametrize("field,sigma", [((40, 56), 0.5), ((33, 70), 0.8), ((5, 7), 0.5), ((64, 64), 1.6), ((5, 7), 1.6)])
def test_diffuse_modes_bit_exact(mode, field, sigma):
    """Brownian free run (only +,-,*,/ on the path): every field must equal scipy's blur with that mode bit for bit;
    a Physarum agent is then run for a few steps so that the published gradient is exercised under the mode too."""
    import die_b200 as D
    (ref,), gpu = make_pair(field, seed=6, ratio=0.3, dynamics_kw=dict(diffuse_mode=mode, diffuse_sigma=sigma))
    ra, ga = R.BrownianAgent(0.05, 1.0), D.BrownianAgent(move_scale=0.05, deposit_scale=1.0)
    m = ref.agents.shape[-1]
    rng = np.random.default_rng(4)
    robs, gobs = ref._get_current_obs, gpu._get_current_obs
    for it in range(15):
        u = rng.random((3, m))
        ract, gact = ra.forward(robs, u=u), ga.forward(gobs, u=u)
        robs, rr, *_ = ref.step(ract)
        gobs, gr, *_ = gpu.step(gact)
        assert_state_equal(ref, *gpu.get_state(), float_exact=True)
    R.set_math_backend('portable')
    try:
        kw = dict(scale=0.02, turn_angle=30, sense_offset=0.06)
        from tests._parity import lattice_theta
        theta0, prev = lattice_theta(m, 30, 6)
        rp, gp = R.PhysarumAgent(max_agents=m, prev_grad=prev, **kw), D.PhysarumAgent(max_agents=m, **kw)
        gp.set_state(theta=theta0)
        for it in range(8):
            coin = rng.integers(0, 2, m)
            ract = rp.forward(robs, coin=coin.copy())
            gact = gp.forward(gobs, coin=coin)
            assert np.array_equal(ract, gact.cpu().numpy()), (mode, it)
            robs, *_ = ref.step(ract)
            gobs, *_ = gpu.step(gact)
            assert_state_equal(ref, *gpu.get_state(), float_exact=True)
        assert gp.last_hints == (True, True)
    finally:
        R.set_math_backend('numpy')
