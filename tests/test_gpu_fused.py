"""The two result-neutral fast paths of the Physarum step, on the GPU:

* the speculative move: ``PhysarumAgent.forward`` evaluates Env._agent_move + the claims of the action it
  writes, ``Env.step`` adopts them iff it receives that very action untouched (die_env_step_flags, DIE_STEP_ADOPT_MOVE) and
  otherwise discards them -- every continuation must give the bits of the plain four-kernel step;
* the guard-banded quick turn decision (die_turn.h) and the register cap of the forward kernel
  (die_set_tuning): switching them must not change a single bit.
"""
import numpy as np
import pytest

from tests._parity import make_pair, lattice_theta

pytestmark = pytest.mark.gpu

PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)


def _pair_of_envs(shape, seed, **kw):
    (_,), a = make_pair(shape, seed=seed, **kw)
    (_,), b = make_pair(shape, seed=seed, **kw)
    return a, b


def _agents(m, seed, **kw):
    import die_b200 as D
    theta0, _ = lattice_theta(m, kw.get('turn_angle', 30), seed)
    out = []
    for _ in range(2):
        ag = D.PhysarumAgent(max_agents=m, **kw)
        ag.set_state(theta=theta0)
        ag.fuse_move = True
        out.append(ag)
    return out


def _same(env_a, env_b, ag_a, ag_b):
    ma, aa = env_a.get_state()
    mb, ab = env_b.get_state()
    assert np.array_equal(ma, mb) and np.array_equal(aa, ab)
    assert np.array_equal(env_a.last_cells().cpu().numpy(), env_b.last_cells().cpu().numpy())
    assert np.array_equal(ag_a.get_state()[0], ag_b.get_state()[0])


@pytest.mark.parametrize("shape,dyn", [((96, 80), {}), ((37, 53), {}), ((256, 256), {}),
                                       ((64, 48), dict(boundary='limit', diffuse_sigma=0.8))])
def test_fused_move_matches_the_plain_step(shape, dyn):
    import die_b200 as D
    dyn = dict(dyn)
    if dyn.get('boundary') == 'limit':
        dyn['boundary'] = D.BoundaryCondition.limit
    env_f, env_p = _pair_of_envs(shape, 31, dynamics_kw=dyn, ref_dynamics_kw={})
    m = env_f.max_agents
    ag_f, ag_p = _agents(m, 31, **PHYS)
    ag_p.fuse_move = False
    rng = np.random.default_rng(5)
    of, op = env_f._get_current_obs, env_p._get_current_obs
    for it in range(40):
        coin = rng.integers(0, 2, m)
        af = ag_f.forward(of, coin=coin)
        ap = ag_p.forward(op, coin=coin)
        assert ag_f.last_speculated and not ag_p.last_speculated
        assert np.array_equal(af.cpu().numpy(), ap.cpu().numpy())
        of, rf, _, _, inf_f = env_f.step(af)
        op, rp, _, _, inf_p = env_p.step(ap)
        assert env_f.last_step_fused and not env_p.last_step_fused
        assert rf == rp and inf_f == inf_p, it
        _same(env_f, env_p, ag_f, ag_p)


def test_fused_move_batched_envs():
    import die_b200 as D
    (_, _, _), env_f = make_pair((48, 40), seed=4, batch=3)
    (_, _, _), env_p = make_pair((48, 40), seed=4, batch=3)
    m = env_f.max_agents
    th = np.stack([lattice_theta(m, 30, 40 + b)[0] for b in range(3)])
    ag_f, ag_p = D.PhysarumAgent(max_agents=m, **PHYS), D.PhysarumAgent(max_agents=m, **PHYS)
    ag_f.set_state(theta=th)
    ag_p.set_state(theta=th)
    ag_f.fuse_move, ag_p.fuse_move = True, False
    rng = np.random.default_rng(6)
    of, op = env_f._get_current_obs, env_p._get_current_obs
    for it in range(25):
        coin = rng.integers(0, 2, (3, m))
        of, rf, *_ = env_f.step(ag_f.forward(of, coin=coin))
        op, rp, *_ = env_p.step(ag_p.forward(op, coin=coin))
        assert env_f.last_step_fused and not env_p.last_step_fused
        assert np.array_equal(rf, rp)
        _same(env_f, env_p, ag_f, ag_p)


def test_abandoned_speculations_leave_no_trace():
    """Every way of NOT stepping the speculated action: an edited action, a cloned action, a second
    forward, another agent's action, edited agents, a host-path step.  The plain env does the same
    steps with speculation off; states must stay identical throughout."""
    import torch
    import die_b200 as D
    env_s, env_p = _pair_of_envs((72, 88), 17)
    m = env_s.max_agents
    ag_s, ag_p = _agents(m, 17, **PHYS)
    ag_p.fuse_move = False
    brown_s, brown_p = D.BrownianAgent(0.01), D.BrownianAgent(0.01)
    rng = np.random.default_rng(9)
    os_, op = env_s._get_current_obs, env_p._get_current_obs

    def both_forward():
        coin = rng.integers(0, 2, m)
        return ag_s.forward(os_, coin=coin), ag_p.forward(op, coin=coin)

    for it in range(6):
        kind = it % 6
        a_s, a_p = both_forward()
        if kind == 0:                       # in-place edit of the action (version counter moves)
            a_s[2] *= 0.5
            a_p[2] *= 0.5
        elif kind == 1:                     # a copy of the action
            a_s, a_p = a_s.clone(), a_p.clone()
        elif kind == 2:                     # forward twice, step the second
            ag_p.set_state(theta=ag_s.get_state()[0])          # keep both agents' headings aligned
            a_s, a_p = both_forward()
        elif kind == 3:                     # somebody else's action
            u = rng.random((3, m))
            a_s, a_p = brown_s.forward(os_, u=u), brown_p.forward(op, u=u)
        elif kind == 4:                     # agents edited between forward and step
            env_s.agents[0] += 0.001
            env_p.agents[0] += 0.001
        elif kind == 5:                     # host-path step
            a_s, a_p = a_s.cpu().numpy(), a_p.cpu().numpy()
        os_, r_s, *_ = env_s.step(a_s)
        op, r_p, *_ = env_p.step(a_p)
        if kind == 2:
            assert env_s.last_step_fused    # the SECOND forward's speculation is the live one
        else:
            assert not env_s.last_step_fused, kind
        assert r_s == r_p, (it, kind)
        if kind == 5:
            os_, op = env_s._get_current_obs, env_p._get_current_obs
        _same(env_s, env_p, ag_s, ag_p)
        # and a clean fused step right after
        a_s, a_p = both_forward()
        os_, r_s, *_ = env_s.step(a_s)
        op, r_p, *_ = env_p.step(a_p)
        assert env_s.last_step_fused and r_s == r_p
        _same(env_s, env_p, ag_s, ag_p)


def test_one_agent_two_envs_never_adopts_the_wrong_speculation():
    """The agent re-uses one action buffer: after forward(obs of env A) then forward(obs of env B), env A
    must not adopt what is now B's action as its own speculated one."""
    import die_b200 as D
    env_a, env_b = _pair_of_envs((40, 56), 23)
    ref_a, ref_b = _pair_of_envs((40, 56), 23)
    m = env_a.max_agents
    ag, ag_ref = _agents(m, 23, **PHYS)
    ag_ref.fuse_move = False
    ag_ref.use_env_hints = False
    coin1, coin2 = np.random.default_rng(1).integers(0, 2, (2, m))
    ag.forward(env_a._get_current_obs, coin=coin1)
    act = ag.forward(env_b._get_current_obs, coin=coin2)          # same buffer, now B's action
    env_a.step(act)
    assert not env_a.last_step_fused
    ag_ref.forward(ref_a._get_current_obs, coin=coin1)
    act_ref = ag_ref.forward(ref_b._get_current_obs, coin=coin2)
    assert np.array_equal(act.cpu().numpy(), act_ref.cpu().numpy())
    ref_a.step(act_ref)
    _same(env_a, ref_a, ag, ag_ref)
    env_b.step(act)
    ref_b.step(act_ref)
    assert env_b.last_step_fused
    _same(env_b, ref_b, ag, ag_ref)


@pytest.mark.parametrize("key,values", [("turn_quick", (1, 0)), ("fwd_min_blocks", (4, 3, 5)), ("feed_bits", (1, 0)), ("field_prefetch", (1, 0)), ("fwd_lean", (1, 0, 5)), ("sense_quick", (1, 0))])
def test_tuning_switches_do_not_change_results(key, values):
    """Free-running 60 steps (in-kernel Philox coins, identical seeds) under every setting of a switch."""
    import die_b200 as D
    from die_b200 import _lib
    lib = _lib.load()
    states = []
    try:
        for v in values:
            _lib.check(lib.die_set_tuning(key.encode(), v))
            (_,), env = make_pair((128, 96), seed=13)
            m = env.max_agents
            ag = D.PhysarumAgent(max_agents=m, seed=5, **PHYS)
            ag.set_state(theta=lattice_theta(m, 30, 13)[0])
            ag.fuse_move = key in ("fwd_min_blocks", "feed_bits", "field_prefetch")   # also covers the MOVE instantiations
            obs = env._get_current_obs
            total = 0.0
            for _ in range(60):
                obs, r, *_ = env.step(ag.forward(obs))
                total += r
            states.append((*env.get_state(), ag.get_state()[0], total))
    finally:
        _lib.check(lib.die_set_tuning(key.encode(), values[0]))
    for s in states[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(states[0][:3], s[:3])) and states[0][3] == s[3]


def test_quick_turn_on_structured_fields():
    """A field made of exactly symmetric blobs (axis-aligned and diagonal gradients, equal neighbours, zeros)
    with lattice headings: the knife-edge cases the quick decision must hand to the exact path."""
    import torch
    import die_b200 as D
    from die_b200 import _lib
    lib = _lib.load()
    shape = (64, 64)
    outs = []
    try:
        for quick in (1, 0):
            _lib.check(lib.die_set_tuning(b"turn_quick", quick))
            (ref,), env = make_pair(shape, seed=3)
            med, agn = env.get_state()
            yy, xx = np.meshgrid(np.arange(64), np.arange(64))
            chem = np.zeros(shape)
            for cx, cy in [(16, 16), (48, 16), (16, 48), (48, 48), (32, 32)]:
                chem += np.maximum(0, 6 - np.maximum(abs(xx - cx), abs(yy - cy))) * 0.25
            chem[40:, :8] = 3e-6                              # clipped plateau next to zeros
            med[2] = chem
            env.set_state(medium=med)
            m = env.max_agents
            ag = D.PhysarumAgent(max_agents=m, seed=2, **PHYS)
            k = np.random.default_rng(0).integers(-6, 7, m)
            ag.set_state(theta=k * np.radians(30))
            ag.use_env_hints = False
            obs = env._get_current_obs
            acts = []
            for _ in range(8):
                act = ag.forward(obs)
                acts.append(act.cpu().numpy().copy())
                obs, *_ = env.step(act)
            outs.append((np.array(acts), ag.get_state()[0], *env.get_state()))
    finally:
        _lib.check(lib.die_set_tuning(b"turn_quick", 1))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
