"""The CUDA path against the committed golden vectors (tests/golden, produced by executing the
reference's own source; see oracle/make_golden_from_reference.py).  Self-contained on the GPU box:
no oracle, no /root/reference."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_brownian_golden_300_percent_bit_exact():
    """Brownian + Env.step involve only +,-,*,/,sqrt,fmod: the GPU reproduces the reference's
    recorded actions, cells, fields and agent state BIT FOR BIT over the whole recorded run."""
    import die_b200 as D
    g = np.load(os.path.join(GOLD, "brownian_24x32.npz"))
    env = D.Env(tuple(g["size"]), D.Dynamics(), init_state=(g["medium0"], g["agents0"]))
    agent = D.BrownianAgent(float(g["move_scale"]), float(g["deposit_scale"]), rng='numpy')
    np.random.seed(int(g["loop_seed"]))            # validation mode: host draws in the reference's order
    obs = env._get_current_obs
    for k in range(len(g["rewards"])):
        act = agent.forward(obs)
        assert np.array_equal(act.cpu().numpy(), g["actions"][k]), k
        obs, r, _, _, info = env.step(act)
        assert abs(r - g["rewards"][k]) <= 1e-12 * max(1.0, abs(g["rewards"][k]))
        assert info["num_agents"] == g["num_agents"][k]
    med, ag = env.get_state()
    assert np.array_equal(med, g["medium_final"]) and np.array_equal(ag, g["agents_final"])


CASES = {
    "physarum_24x32.npz": (dict(), dict(scale=0.007, turn_angle=30, sense_offset=0.04)),
    "physarum_limit_sigma08_20x20.npz": (
        dict(boundary='limit', diffuse_sigma=0.8, food_infinite=True, op_action_cost='zero'),
        dict(scale=0.03, turn_angle=35, sense_angle=120, sense_offset=0.06, turn_tolerance=0.05)),
    # op_food_flow = WaveSequence flow (core/data_init.py:29-38, 71-89): the field kernel evaluates one cosine per
    # cell with die_math.h (<= 0.7 ulp), so the food channel is compared to 1e-15 instead of bit for bit
    "physarum_waveflow_24x32.npz": (dict(waveflow=(0.5, 0.5)), dict(scale=0.007, turn_angle=30, sense_offset=0.04)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_physarum_golden_every_step(name):
    """Every recorded step from its recorded pre-state: occupancy / alive / num_agents exact, the
    deposit channel exact (same cells, same decisions), headings and moves to 1e-13 (the kernels'
    sin/cos/atan2 and the reference's libm differ by <= 1 ulp), and Env.step on the recorded action
    bit-exact in every field."""
    import die_b200 as D
    g = np.load(os.path.join(GOLD, name))
    dyn_kw, agent_kw = CASES[name]
    dyn_kw = dict(dyn_kw)
    if dyn_kw.get("boundary") == 'limit':
        dyn_kw["boundary"] = D.BoundaryCondition.limit
    if dyn_kw.get("op_action_cost") == 'zero':
        dyn_kw["op_action_cost"] = D.zero_cost
    size = tuple(int(v) for v in g["size"])
    wf = dyn_kw.pop("waveflow", None)
    flow = D.WaveSequence(size, dt=0.01).get_flow_operator(scale=wf[0], decay=wf[1]) if wf is not None else None
    if flow is not None:
        dyn_kw["op_food_flow"] = flow
    m = g["agents_pre"].shape[-1]
    agent = D.PhysarumAgent(max_agents=m, **agent_kw)
    import torch
    for k in range(len(g["reward"])):
        if flow is not None:
            flow.calls = k                      # the recorded step k used the sequence's k-th time step
        env = D.Env(size, D.Dynamics(**dyn_kw), init_state=(g["medium_pre"][k], g["agents_pre"][k]))
        agent.set_state(theta=g["theta_pre"][k])
        act = agent.forward(env._get_current_obs, coin=g["coin"][k]).cpu().numpy()
        assert np.array_equal(act[2], g["action"][k][2]), f"deposit differs at step {k}"
        np.testing.assert_allclose(act[:2], g["action"][k][:2], rtol=0, atol=1e-15)
        th = agent.get_state()[0]
        d = np.abs((th - g["theta_post"][k] + np.pi) % (2 * np.pi) - np.pi)
        assert d.max() < 1e-13, f"heading differs at step {k}"
        gold_act = torch.from_numpy(g["action"][k]).cuda()
        _, r, _, _, info = env.step(gold_act)
        med, ag = env.get_state()
        if flow is not None:
            np.testing.assert_allclose(med[1], g["medium_post"][k][1], rtol=0, atol=1e-15)
            assert (med[1] == g["medium_post"][k][1]).mean() > 0.9
            med[1] = g["medium_post"][k][1]
        assert np.array_equal(med, g["medium_post"][k]) and np.array_equal(ag, g["agents_post"][k]), k
        assert abs(r - g["reward"][k]) <= 1e-12 * max(1.0, abs(g["reward"][k]))
        assert info["num_agents"] == g["num_agents"][k]
