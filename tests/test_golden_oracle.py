"""The oracle against golden vectors produced by executing the reference's OWN source files over
stand-in packages (oracle/make_golden_from_reference.py, oracle/shims/README.md), and -- where
/root/reference is present -- against the reference executed live, bit-for-bit."""
import os

import numpy as np
import pytest

from oracle import die_ref as R
from oracle import run_reference

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


def test_brownian_golden_replay_bit_exact():
    g = _load("brownian_24x32.npz")
    env = R.Env(tuple(g["size"]), R.Dynamics(), medium=g["medium0"], agents=g["agents0"])
    agent = R.BrownianAgent(float(g["move_scale"]), float(g["deposit_scale"]))
    np.random.seed(int(g["loop_seed"]))            # the oracle draws from the global RNG in the reference's order
    obs = env._get_current_obs
    for k in range(len(g["rewards"])):
        act = agent.forward(obs)
        assert np.array_equal(act, g["actions"][k]), k
        obs, r, _, _, info = env.step(act)
        assert r == g["rewards"][k] and info["num_agents"] == g["num_agents"][k], k
    assert np.array_equal(env.medium, g["medium_final"]) and np.array_equal(env.agents, g["agents_final"])


PHYS_CASES = {
    "physarum_24x32.npz": (dict(), dict(scale=0.007, turn_angle=30, sense_offset=0.04)),
    "physarum_limit_sigma08_20x20.npz": (
        dict(boundary='limit', diffuse_sigma=0.8, food_infinite=True, op_action_cost=R.zero_cost),
        dict(scale=0.03, turn_angle=35, sense_angle=120, sense_offset=0.06, turn_tolerance=0.05)),
    # op_food_flow = WaveSequence flow operator; replayed step k starts the sequence at its k-th time step
    "physarum_waveflow_24x32.npz": (dict(waveflow=(0.5, 0.5)), dict(scale=0.007, turn_angle=30, sense_offset=0.04)),
}


def _dynamics(dyn_kw, size, k0=0):
    dyn_kw = dict(dyn_kw)
    wf = dyn_kw.pop("waveflow", None)
    if wf is not None:
        dyn_kw["op_food_flow"] = R.WaveSequence(size, dt=0.01).get_flow_operator(scale=wf[0], decay=wf[1], k0=k0)
    return R.Dynamics(**dyn_kw)


@pytest.mark.parametrize("name", sorted(PHYS_CASES))
def test_physarum_golden_every_step(name):
    """Every recorded step from its recorded pre-state.  Integer-valued results (occupancy, alive,
    num_agents) exact; floats to 1e-12 relative (numpy's sin/cos may differ by an ulp across hosts)
    -- on the host that generated the fixtures they are bit-exact."""
    g = _load(name)
    dyn_kw, agent_kw = PHYS_CASES[name]
    size = tuple(g["size"])
    m = g["agents_pre"].shape[-1]
    exact_all = True
    for k in range(len(g["reward"])):
        env = R.Env(size, _dynamics(dyn_kw, size, k0=k), medium=g["medium_pre"][k], agents=g["agents_pre"][k])
        agent = R.PhysarumAgent(max_agents=m, prev_grad=np.ones((2, m)), **agent_kw)
        agent._direction_rads = g["theta_pre"][k].copy()
        act = agent.forward(env._get_current_obs, coin=g["coin"][k].astype(np.int64))
        np.testing.assert_allclose(act, g["action"][k], rtol=1e-12, atol=1e-16)
        _, r, _, _, info = env.step(g["action"][k])                  # same action -> Env.step is exact
        assert np.array_equal(env.medium, g["medium_post"][k]) and np.array_equal(env.agents, g["agents_post"][k])
        assert r == g["reward"][k] and info["num_agents"] == g["num_agents"][k]
        np.testing.assert_allclose(np.cos(agent._direction_rads), np.cos(g["theta_post"][k]), atol=1e-12)
        exact_all &= np.array_equal(act, g["action"][k]) and np.array_equal(agent._direction_rads, g["theta_post"][k])
    if run_reference.available():
        assert exact_all, "on the generating host the oracle reproduces the reference bit-for-bit"


@pytest.mark.skipif(not run_reference.available(), reason="/root/reference not present on this box")
@pytest.mark.parametrize("kind,size,steps,seed,dyn,akw", [
    ("brownian", (32, 48), 30, 1, {}, dict(move_scale=0.01)),
    ("brownian", (40, 40), 30, 2, {}, dict(move_scale=0.3, deposit_scale=0.1)),
    ("physarum", (32, 48), 40, 3, {}, dict(scale=0.007, turn_angle=30, sense_offset=0.04)),
    ("physarum", (48, 48), 40, 4, {}, dict(scale=0.02, turn_angle=35, sense_angle=120, sense_offset=0.06,
                                            turn_tolerance=0.05)),
    ("physarum", (36, 36), 30, 5, dict(limit=True, diffuse_sigma=0.8, rate_feed=0.3, rate_decay_chem=0.2),
     dict(scale=0.05, sense_offset=0.1)),
    ("const", (24, 24), 20, 6, {}, dict(delta_xy=(-0.01, 0.005), deposit=0.1)),
    ("physarum", (32, 48), 30, 7, dict(waveflow=(0.5, 0.5)), dict(scale=0.007, turn_angle=30, sense_offset=0.04)),
    ("brownian", (28, 20), 30, 8, dict(waveflow=(1.0, 0.1), food_infinite=True), dict(move_scale=0.02)),
])
def test_oracle_equals_reference_executed_live(kind, size, steps, seed, dyn, akw):
    """The reference's own classes (over the stand-in packages) and the index-form restatement, side
    by side on the same seeded state and the same global RNG stream: actions, rewards, info, medium,
    agents and headings must be IDENTICAL at every step."""
    ref = run_reference.load()
    dyn = dict(dyn)
    limit = dyn.pop("limit", False)
    wf = dyn.pop("waveflow", None)
    rdyn, odyn = dict(dyn), dict(dyn)
    if wf is not None:
        rdyn["op_food_flow"] = ref.WaveSequence(size, dt=0.01).get_flow_operator(scale=wf[0], decay=wf[1])
        odyn["op_food_flow"] = R.WaveSequence(size, dt=0.01).get_flow_operator(scale=wf[0], decay=wf[1])
    np.random.seed(seed)
    renv = ref.Env(size, ref.Dynamics(init_agent_ratio=0.1, boundary=ref.BoundaryCondition.limit if limit
                                      else ref.BoundaryCondition.wrap, **rdyn))
    oenv = R.Env(size, R.Dynamics(boundary='limit' if limit else 'wrap', **odyn),
                 medium=renv.medium.values.copy(), agents=renv.agents.values.copy())
    m = renv.agents.shape[-1]
    if kind == "brownian":
        ra, oa = ref.BrownianAgent(**akw), R.BrownianAgent(**akw)
    elif kind == "const":
        ra, oa = ref.ConstAgent(**akw), R.ConstAgent(**akw)
    else:
        ra = ref.PhysarumAgent(max_agents=m, **akw)
        oa = R.PhysarumAgent(max_agents=m, prev_grad=ra._prev_grad.copy(), **akw)
        assert np.array_equal(oa._direction_rads, ra._direction_rads)
    robs, oobs = renv._get_current_obs, oenv._get_current_obs
    for it in range(steps):
        state = np.random.get_state()
        ract = ra.forward(robs)
        np.random.set_state(state)
        oact = oa.forward(oobs)
        assert np.array_equal(ract.values, oact), it
        robs, rr, rt, rtr, rinfo = renv.step(ract)
        oobs, orr, ot, otr, oinfo = oenv.step(oact)
        assert (rr, rt, rtr, rinfo) == (orr, ot, otr, oinfo), it
        assert np.array_equal(renv.medium.values, oenv.medium) and np.array_equal(renv.agents.values, oenv.agents), it
        assert np.array_equal(robs[1].values, oobs[1])
        if kind == "physarum":
            assert np.array_equal(ra._direction_rads, oa._direction_rads)


@pytest.mark.skipif(not run_reference.available(), reason="/root/reference not present on this box")
@pytest.mark.parametrize("kind,rate_feed,steps", [("brownian", 0.1, 40), ("physarum", 0.02, 40), ("brownian", 0.0, 25)])
def test_agents_die_equals_reference_with_its_indexer_rebound(kind, rate_feed, steps):
    """Dynamics(agents_die=True), core/env.py:245-261: the reference's lifecycle rebinds ``self.agents`` to a NEW array
    (``agents.where(have_food, 0)``) while the AgentIndexer it built in __init__ keeps the old one (core/utils.py:22), so
    from then on it moves one array and resolves cells from another.  The semantics of a WORKING agents_die are pinned
    here as: the reference's own code, with the indexer re-pointed at ``self.agents`` after every step (one assignment,
    nothing else touched).  Agents starve (low rate_feed), ghosts are put back at (0, 0) every step."""
    ref = run_reference.load()
    size = (28, 36)
    np.random.seed(31)
    renv = ref.Env(size, ref.Dynamics(init_agent_ratio=0.3, agents_die=True, rate_feed=rate_feed))
    oenv = R.Env(size, R.Dynamics(agents_die=True, rate_feed=rate_feed),
                 medium=renv.medium.values.copy(), agents=renv.agents.values.copy())
    m = renv.agents.shape[-1]
    if kind == "brownian":
        ra, oa = ref.BrownianAgent(move_scale=0.03, deposit_scale=2.0), R.BrownianAgent(move_scale=0.03, deposit_scale=2.0)
    else:
        akw = dict(scale=0.05, turn_angle=30, sense_offset=0.04, deposit=8.0)
        ra = ref.PhysarumAgent(max_agents=m, **akw)
        oa = R.PhysarumAgent(max_agents=m, prev_grad=ra._prev_grad.copy(), **akw)
    robs, oobs = renv._get_current_obs, oenv._get_current_obs
    alive0 = int((renv.agents.values[2] > 0).sum())
    for it in range(steps):
        state = np.random.get_state()
        ract = ra.forward(robs)
        np.random.set_state(state)
        oact = oa.forward(oobs)
        assert np.array_equal(ract.values, oact), it
        robs, rr, rt, rtr, rinfo = renv.step(ract)
        renv._agent_idx._AgentIndexer__agents = renv.agents          # THE fix: the indexer follows the rebound array
        rinfo = dict(rinfo, num_agents=int((renv.agents.values[2] > 0).sum()))
        oobs, orr, ot, otr, oinfo = oenv.step(oact)
        assert rr == orr and oinfo["num_agents"] == rinfo["num_agents"] and oinfo["reward"] == rinfo["reward"], it
        assert np.array_equal(renv.medium.values, oenv.medium) and np.array_equal(renv.agents.values, oenv.agents), it
    assert oinfo["num_agents"] < alive0, "the scenario must actually starve agents"


@pytest.mark.skipif(not run_reference.available(), reason="/root/reference not present on this box")
@pytest.mark.parametrize("kernel_sizes,with_agent_channel,size", [((3,), True, (24, 32)), ((3, 5), True, (20, 28)),
                                                                 ((5, 3, 3), False, (16, 40)), ((7,), True, (12, 12))])
def test_neural_automata_agent_equals_reference_executed_live(kernel_sizes, with_agent_channel, size):
    """core/agent/evo.py (NeuralAutomataAgent + ConvolutionModel, the reference's own source over the stand-in packages)
    vs the oracle's index-form restatement with the same weights: identical float32 actions and model output, step after
    step while the env evolves under those actions."""
    import torch as th
    ref = run_reference.load()
    import core.agent.evo as evo
    th.manual_seed(7)
    np.random.seed(7)
    renv = ref.Env(size, ref.Dynamics(init_agent_ratio=0.2))
    oenv = R.Env(size, R.Dynamics(), medium=renv.medium.values.copy(), agents=renv.agents.values.copy())
    ra = evo.NeuralAutomataAgent(scale=0.05, deposit=0.7, with_agent_channel=with_agent_channel, kernel_sizes=kernel_sizes)
    ra.model.init_weights()
    weights = [k.weight.detach().numpy().copy() for k in ra.model.kernels if hasattr(k, 'weight')]
    oa = R.NeuralAutomataAgent(weights, scale=0.05, deposit=0.7, with_agent_channel=with_agent_channel)
    robs, oobs = renv._get_current_obs, oenv._get_current_obs
    for it in range(6):
        ract = ra.forward(robs)
        oact = oa.forward(oobs)
        assert ract.values.dtype == np.float32 and oact.dtype == np.float32
        assert np.array_equal(ract.values, oact), it
        assert np.array_equal(ra._sense_output.detach().numpy(), oa.sense_output.numpy())
        robs, rr, *_ = renv.step(ract)
        oobs, orr, *_ = oenv.step(oact)                 # float32, as the reference hands it to Env.step
        assert rr == orr
        assert np.array_equal(renv.medium.values, oenv.medium) and np.array_equal(renv.agents.values, oenv.agents), it


@pytest.mark.skipif(not run_reference.available(), reason="/root/reference not present on this box")
@pytest.mark.parametrize("colors", ['rgb', 'one', 'two'])
def test_oracle_renderer_equals_reference_executed_live(colors):
    """core/render.py's EnvRenderer (own source, over the stand-in packages; its matplotlib colour map aside) vs the
    oracle's restatement: medium frame, agent trace and agents frame over several steps, non-square field."""
    ref = run_reference.load()
    import core.render as render
    size = (24, 40)
    np.random.seed(9)
    renv = ref.Env(size, ref.Dynamics(init_agent_ratio=0.2))
    rr = render.EnvRenderer(size, field_colors_id=colors)
    orr = R.EnvRenderer(size, field_colors_id=colors)
    agent = ref.BrownianAgent(move_scale=0.02)
    obs = renv._get_current_obs
    for it in range(6):
        frames = orr.render(renv.medium.values, renv.agents.values)
        assert np.array_equal(rr._upd_img_medium(renv.medium), frames[0])
        rr._agent_trace.update(renv.medium.sel(channel='agents'))
        assert np.array_equal(np.asarray(rr._agent_trace.as_mask()), frames[1])
        assert np.array_equal(rr._upd_img_agents(renv.agents), frames[2])
        obs, *_ = renv.step(agent.forward(obs))
