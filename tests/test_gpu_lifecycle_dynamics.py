"""Dynamics(agents_die=True) on the GPU (core/env.py:245-250 with the indexer following the rebound array: semantics
pinned in tests/test_golden_oracle.py), a Dynamics edited between steps (the reference reads self.dynamics every step),
Env.invalidate_caches after a write torch's version counters cannot see, and the food-frame size guard."""
import numpy as np
import pytest

from tests._parity import assert_state_equal, lattice_theta, make_pair

pytestmark = pytest.mark.gpu
PHYS = dict(scale=0.05, turn_angle=30, sense_offset=0.04, deposit=8.0)


@pytest.mark.parametrize("batch", [None, 3])
def test_agents_die_brownian(batch):
    import die_b200 as D
    from oracle import die_ref as R
    dyn = dict(agents_die=True, rate_feed=0.02)
    refs, env = make_pair((48, 64), ratio=0.3, seed=31, dynamics_kw=dyn, batch=batch)
    B, m = env.batch, env.max_agents
    ra, ga = R.BrownianAgent(0.03, 2.0), D.BrownianAgent(move_scale=0.03, deposit_scale=2.0)
    rng = np.random.default_rng(7)
    gobs = env._get_current_obs
    alive0 = sum(r.num_alive for r in refs)
    for it in range(40):
        u = rng.random((B, 3, m))
        gact = ga.forward(gobs, u=u if batch else u[0])
        infos = [refs[b].step(ra.forward(refs[b]._get_current_obs, u=u[b]))[4] for b in range(B)]
        gobs, gr, gterm, _, ginfo = env.step(gact)
        med, ag = env.get_state()
        med, ag = med.reshape(B, 3, 48, 64), ag.reshape(B, 4, m)
        for b in range(B):
            assert_state_equal(refs[b], med[b], ag[b], float_exact=True)
            assert np.atleast_1d(ginfo['num_agents'])[b] == infos[b]['num_agents']
    assert sum(r.num_alive for r in refs) < alive0, "the scenario must starve agents"


def test_agents_die_physarum_with_hints():
    """PhysarumAgent through the env's hints (cell cache, published gradient): a slot put back at (0, 0) must be looked
    up in cell 0 by the next forward."""
    import die_b200 as D
    from oracle import die_ref as R
    R.set_math_backend('portable')
    try:
        dyn = dict(agents_die=True, rate_feed=0.02)
        (ref,), env = make_pair((40, 56), ratio=0.3, seed=5, dynamics_kw=dyn)
        m = env.max_agents
        theta0, prev = lattice_theta(m, 30, 5)
        ra = R.PhysarumAgent(max_agents=m, prev_grad=prev, **PHYS)
        ga = D.PhysarumAgent(max_agents=m, **PHYS)
        ga.set_state(theta=theta0)
        rng = np.random.default_rng(3)
        gobs = env._get_current_obs
        for it in range(30):
            coin = rng.integers(0, 2, m)
            ract = ra.forward(ref._get_current_obs, coin=coin.copy())
            gact = ga.forward(gobs, coin=coin)
            assert np.array_equal(gact.cpu().numpy(), ract), it
            ref.step(ract)
            gobs, *_ = env.step(gact)
            assert_state_equal(ref, *env.get_state(), float_exact=True)
        assert ga.last_hints == (True, True)
    finally:
        R.set_math_backend('numpy')


def test_dynamics_edited_between_steps_take_effect():
    """env.dynamics.rate_feed = ..., food_infinite, diffuse_sigma, agents_die changed between steps: the reference reads
    self.dynamics on every step (core/env.py:136-150, 220-250), so must we (die_env_set_dynamics)."""
    import die_b200 as D
    from oracle import die_ref as R
    (ref,), env = make_pair((32, 48), seed=2)
    m = env.max_agents
    ra, ga = R.BrownianAgent(0.02), D.BrownianAgent(move_scale=0.02)
    rng = np.random.default_rng(0)
    gobs = env._get_current_obs
    edits = {3: dict(rate_feed=0.4), 6: dict(food_infinite=True), 9: dict(diffuse_sigma=0.8), 12: dict(rate_decay_chem=0.3),
             15: dict(agents_die=True)}
    for it in range(20):
        for k, v in edits.get(it, {}).items():
            setattr(env.dynamics, k, v)
            setattr(ref.dynamics, k, v)
        u = rng.random((3, m))
        ref.step(ra.forward(ref._get_current_obs, u=u))
        gobs, *_ = env.step(ga.forward(gobs, u=u))
        assert_state_equal(ref, *env.get_state(), float_exact=True)


def test_invalidate_caches_after_an_untracked_write():
    """A write through tensor.data does not bump torch's version counter: the alive bitmask goes stale until
    invalidate_caches() (documented contract; the reference re-reads its arrays every step)."""
    import torch
    import die_b200 as D
    from oracle import die_ref as R
    (ref,), env = make_pair((32, 32), seed=4)
    m = env.max_agents
    ra, ga = R.BrownianAgent(0.02), D.BrownianAgent(move_scale=0.02)
    rng = np.random.default_rng(0)
    u = rng.random((3, m))
    ref.step(ra.forward(ref._get_current_obs, u=u))
    obs, *_ = env.step(ga.forward(env._get_current_obs, u=u))
    alive = np.flatnonzero(ref.agents[2] > 0)[:10]
    ref.agents[2, alive] = 0.0                                   # kill ten agents behind torch's back
    env.agents.data[2, torch.as_tensor(alive, device=env.device)] = 0.0
    env.invalidate_caches()
    u = rng.random((3, m))
    _, _, _, _, rinfo = ref.step(ra.forward(ref._get_current_obs, u=u))
    _, _, _, _, ginfo = env.step(ga.forward(env._get_current_obs, u=u))
    assert ginfo['num_agents'] == rinfo['num_agents']
    assert_state_equal(ref, *env.get_state(), float_exact=True)


def test_food_frames_that_cannot_fit_are_refused_by_name():
    import die_b200 as D

    class Huge(D.FieldSequence):
        def __getitem__(self, t):
            raise AssertionError("must not be tabulated")

    seq = Huge((4096, 4096), dt=0.01, t_bounds=(0, 10))
    with pytest.raises(MemoryError, match="tabulating"):
        D.Env((4096, 4096), D.Dynamics(op_food_flow=seq.get_flow_operator(scale=0.5, decay=0.5)), init="device")


def test_verify_caches_catches_a_write_torch_does_not_see():
    """Env(verify_caches=True), a debug mode: every cache carries a checksum of the bytes it was built from; a write
    through tensor.data (no version bump) is reported instead of stepping on a stale alive mask / cell cache, and after
    invalidate_caches() the run continues exactly like one that was told about the write."""
    import torch
    import die_b200 as D
    outs = []
    for verify in (True, False):
        env = D.Env((32, 48), D.Dynamics(init_agent_ratio=0.2), init='device', seed=1, verify_caches=verify)
        ag = D.PhysarumAgent(max_agents=env.max_agents, seed=3, scale=0.02, turn_angle=30, sense_offset=0.05)
        obs = env._get_current_obs
        for _ in range(3):
            obs, *_ = env.step(ag.forward(obs))
        v0 = env.agents._version
        env.agents.data[2, :40] = 0.0                    # kills 40 agents behind torch's back
        env.agents.data[0, :40] = 0.9
        assert env.agents._version == v0
        if verify:
            with pytest.raises(RuntimeError, match="verify_caches"):
                ag.forward(obs)
            with pytest.raises(RuntimeError, match="verify_caches"):
                env.step(ag._action)
        env.invalidate_caches()
        for _ in range(3):
            obs, *_ = env.step(ag.forward(obs))
        outs.append((*env.get_state(), ag.get_state()[0]))
    assert all(np.array_equal(a, b) for a, b in zip(*outs))
    env.medium.data[2].mul_(0.5)                         # the same for the medium (verify=False here: nothing is raised)
    env2 = D.Env((32, 48), D.Dynamics(init_agent_ratio=0.2), init='device', seed=1, verify_caches=True)
    obs = env2._get_current_obs
    obs, *_ = env2.step(ag.forward(obs))
    env2.medium.data[2].mul_(0.5)
    with pytest.raises(RuntimeError, match="medium was written"):
        ag.forward(obs)
