"""The committed move on the GPU (``agent.fuse_move = 'commit'``, DIE_FWD_COMMIT_MOVE in include/die_b200.h).

The reference's run loop is ``action = agent.forward(obs); obs, ... = env.step(action)`` (examples/minimal_run.py:21-25):
under that contract the forward launch -- the steady-state (LEAN) instantiation, in-kernel Philox coins, the env's own
caches -- also evaluates Env._agent_move (core/env.py:152-172) + cell resolution + the claim AND stores the positions, so
the step is the field pass + the plain feed kernel.  Every bit of every step must equal the plain five-launch loop.
"""
import numpy as np
import pytest

from tests._parity import make_pair, lattice_theta

pytestmark = pytest.mark.gpu

PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)


def _setup(shape, seed, batch=None, dynamics_kw=None, agent_seed=9):
    import die_b200 as D
    envs, agents = [], []
    for mode in (False, 'commit'):
        _, env = make_pair(shape, seed=seed, batch=batch, dynamics_kw=dynamics_kw, ref_dynamics_kw={})
        m = env.max_agents
        ag = D.PhysarumAgent(max_agents=m, rng='philox', seed=agent_seed, **PHYS)
        B = batch or 1
        th = np.stack([lattice_theta(m, 30, seed + b)[0] for b in range(B)])
        ag.set_state(theta=th if batch else th[0])
        ag.fuse_move = mode
        envs.append(env)
        agents.append(ag)
    return envs, agents


def _same(env_a, env_b, ag_a, ag_b):
    ma, aa = env_a.get_state()
    mb, ab = env_b.get_state()
    assert np.array_equal(ma, mb) and np.array_equal(aa, ab)
    assert np.array_equal(env_a.last_cells().cpu().numpy(), env_b.last_cells().cpu().numpy())
    assert np.array_equal(ag_a.get_state()[0], ag_b.get_state()[0])


@pytest.mark.parametrize("shape,batch,dyn,pair", [
    ((96, 80), None, {}, 0), ((37, 53), 3, {}, 0), ((256, 256), 8, {}, 0), ((64, 48), None, dict(diffuse_sigma=0.8), 0),
    ((128, 64), 2, {}, 2), ((64, 48), None, dict(boundary='limit'), 2)])
def test_committed_move_matches_the_plain_loop(shape, batch, dyn, pair):
    import die_b200 as D
    from die_b200 import _lib
    dyn = dict(dyn)
    if dyn.get('boundary') == 'limit':
        dyn['boundary'] = D.BoundaryCondition.limit
    lib = _lib.load()
    _lib.check(lib.die_set_tuning(b"pair_mode", pair))
    try:
        (env_p, env_c), (ag_p, ag_c) = _setup(shape, 31, batch, dyn)
        lm0, sc0 = lib.die_get_counter(b"forward_lean_move"), lib.die_get_counter(b"step_committed")
        op, oc = env_p._get_current_obs, env_c._get_current_obs
        for it in range(40):
            ap, ac = ag_p.forward(op), ag_c.forward(oc)
            # the first forward sees no published gradient / cell cache yet: plain; from then on the move is committed
            assert ag_c.last_committed == (it > 0) and not ag_p.last_committed
            assert np.array_equal(ap.cpu().numpy(), ac.cpu().numpy()), it
            op, rp, _, _, ip = env_p.step(ap)
            oc, rc, _, _, ic = env_c.step(ac)
            assert env_c.last_step_fused == (it > 0) and not env_p.last_step_fused
            assert np.array_equal(rp, rc) and str(ip) == str(ic), it
            _same(env_p, env_c, ag_p, ag_c)
        assert lib.die_get_counter(b"step_committed") == sc0 + 39
        assert lib.die_get_counter(b"forward_lean_move") == lm0 + 39
    finally:
        _lib.check(lib.die_set_tuning(b"pair_mode", 1))


def test_committed_move_in_a_graphed_loop_and_async():
    """step_async loop and the CUDA-graph replay of it (die_b200.GraphedLoop): the same bits as the plain eager loop."""
    import torch
    import die_b200 as D
    (env_p, env_c), (ag_p, ag_c) = _setup((96, 128), 7, 4)
    (_, env_g), (_, ag_g) = _setup((96, 128), 7, 4)
    op, oc = env_p._get_current_obs, env_c._get_current_obs
    rsum_p = torch.zeros(4, dtype=torch.float64, device=env_p.device)
    rsum_c = torch.zeros_like(rsum_p)
    for it in range(24):
        op, rp, _ = env_p.step_async(ag_p.forward(op))
        oc, rc, _ = env_c.step_async(ag_c.forward(oc))
        if it >= 2:                       # (the GraphedLoop sums from its first iteration after the two warm-up ones)
            rsum_p += rp
            rsum_c += rc
    loop = D.GraphedLoop(env_g, ag_g, warmup=2)
    _, rsum_g = loop.run(22)
    assert ag_g.last_committed
    torch.cuda.synchronize()
    _same(env_p, env_c, ag_p, ag_c)
    _same(env_p, env_g, ag_p, ag_g)
    assert torch.equal(rsum_p, rsum_c)
    assert torch.equal(rsum_p, rsum_g)


def test_a_committed_move_must_be_stepped():
    import die_b200 as D
    (_, env), (_, ag) = _setup((48, 64), 3)
    obs = env._get_current_obs
    obs, *_ = env.step(ag.forward(obs))
    before = env.agents.clone()
    action = ag.forward(obs)
    assert ag.last_committed
    assert not np.array_equal(before[:2].cpu().numpy(), env.agents[:2].cpu().numpy())     # already moved
    with pytest.raises(RuntimeError, match="committed"):
        env.step(action.clone())                       # not the tensor forward returned
    with pytest.raises(RuntimeError, match="committed"):
        ag.forward(obs)                                # a second forward before the step
    with pytest.raises(RuntimeError, match="committed"):
        env.step(action.cpu().numpy())                 # the host-buffer step
    obs, r, *_ = env.step(action)                      # the right one still works
    assert env.last_step_fused
    # a Brownian agent in between never commits anything; the Physarum agent resumes afterwards
    br = D.BrownianAgent(0.01, seed=4)
    obs, *_ = env.step(br.forward(obs))
    assert not env.last_step_fused
    obs, *_ = env.step(ag.forward(obs))
    assert env.last_step_fused
    # reset() forgets a pending commitment
    ag.forward(obs)
    env.reset(seed=5)
    obs = env._get_current_obs
    obs, *_ = env.step(ag.forward(obs))
