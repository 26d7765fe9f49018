"""Env(init='device'): the initial state generated on the GPU (die_b200/device_init.py), then stepped --
the oracle, started from a copy of that state, must follow bit for bit."""
import numpy as np
import pytest

from oracle import die_ref as R
from tests._parity import assert_state_equal, ref_cells_linear, lattice_theta

pytestmark = pytest.mark.gpu


def test_device_init_state_is_well_formed_and_steps_like_the_oracle():
    import die_b200 as D
    field = (96, 72)
    env = D.Env(field, D.Dynamics(init_agent_ratio=0.1), init='device', seed=4)
    med, ag = env.get_state()
    m = env.max_agents
    assert med.shape == (3, *field) and ag.shape == (4, m) and m == field[0] * field[1]
    n = int(ag[2].sum())
    assert n == int(med[0].sum()) and (ag[2, :n] == 1).all() and (ag[:, n:] == 0).all()
    ix, iy = np.nonzero(med[0])
    assert np.array_equal(ag[0, :n], np.linspace(0, 1, field[0])[ix]) and np.array_equal(ag[1, :n], np.linspace(0, 1, field[1])[iy])
    # the oracle takes over the same state
    np.random.seed(0)
    ref = R.Env(field, R.Dynamics(init_agent_ratio=0.1), noise_seed=0)
    ref.medium[...] = med
    ref.agents[...] = ag
    theta0, prev = lattice_theta(m, 30, 4)
    kw = dict(scale=0.007, turn_angle=30, sense_offset=0.04)
    ra, ga = R.PhysarumAgent(max_agents=m, prev_grad=prev, **kw), D.PhysarumAgent(max_agents=m, **kw)
    ga.set_state(theta=theta0)
    rng = np.random.default_rng(1)
    gobs = env._get_current_obs
    for it in range(10):
        ra._direction_rads = ga.get_state()[0].copy()
        coin = rng.integers(0, 2, m)
        ra.forward(ref._get_current_obs, coin=coin.copy())
        gact = ga.forward(gobs, coin=coin)
        ref.step(gact.cpu().numpy())
        gobs, *_ = env.step(gact)
        assert np.array_equal(ref_cells_linear(ref), env.last_cells().cpu().numpy())
        assert_state_equal(ref, *env.get_state(), float_exact=True)


def test_device_init_batched_and_reset():
    import die_b200 as D
    env = D.Env((48, 40), D.Dynamics(init_agent_ratio=0.2), batch=5, init='device', seed=9)
    med, ag = env.get_state()
    assert med.shape == (5, 3, 48, 40) and ag.shape == (5, 4, 48 * 40)
    assert abs(med[:, 0].mean() - 0.2) < 0.03
    for b in range(5):
        n = int(ag[b, 2].sum())
        assert n == int(med[b, 0].sum()) and (ag[b, :, n:] == 0).all()
    env.reset()
    med2, _ = env.get_state()
    assert med2.shape == med.shape and not np.array_equal(med, med2)     # a fresh state, as in the reference
    same = D.Env((48, 40), D.Dynamics(init_agent_ratio=0.2), batch=5, init='device', seed=9)
    assert np.array_equal(same.get_state()[0], med)


def test_device_init_large_field_is_fast():
    import time
    import torch
    import die_b200 as D
    t0 = time.time()
    env = D.Env((2048, 2048), D.Dynamics(init_agent_ratio=0.1), init='device', seed=2)
    torch.cuda.synchronize()
    dt = time.time() - t0
    n = int(env._num_alive_agents)
    assert abs(n / 2048 ** 2 - 0.1) < 0.005
    assert dt < 30, dt           # the reference's per-cell Python loop needs minutes here
