"""die_b200.GraphedLoop: the reference's run loop (examples/minimal_run.py:21-25) replayed from a CUDA graph must give
the bits of the eager loop -- the in-kernel random draws included (the call counter lives on the device)."""
import numpy as np
import pytest

from tests._parity import lattice_theta, make_pair

pytestmark = pytest.mark.gpu
PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)


def _agent(D, kind, m):
    if kind == "brownian":
        return D.BrownianAgent(move_scale=0.01, seed=9)
    ag = D.PhysarumAgent(max_agents=m, seed=9, **PHYS)
    return ag


@pytest.mark.parametrize("kind", ["brownian", "physarum"])
@pytest.mark.parametrize("field,batch,iters", [((256, 256), None, 61), ((64, 96), 4, 40)])
def test_graphed_loop_equals_eager_loop(kind, field, batch, iters):
    import torch
    import die_b200 as D
    outs = []
    for graphed in (False, True):
        _, env = make_pair(field, seed=5, batch=batch)
        m = env.max_agents
        ag = _agent(D, kind, m)
        if kind == "physarum":
            ag.set_state(theta=np.stack([lattice_theta(m, 30, 5 + b)[0] for b in range(env.batch)]))
        if graphed:
            loop = D.GraphedLoop(env, ag, warmup=2)
            obs, rsum = loop.run(iters - 2)                 # the constructor ran two eager iterations
            obs, rsum = loop.run(7)                         # a second run, odd: eager head / tail handling
            total = None
        else:
            obs = env._get_current_obs
            total = torch.zeros(env.batch, dtype=torch.float64, device=env.device)
            for it in range(iters + 7):
                obs, r, alive = env.step_async(ag.forward(obs))
                if it >= 2:
                    total += r
        torch.cuda.synchronize()
        theta = ag.get_state()[0] if kind == "physarum" else np.zeros(1)
        outs.append((*env.get_state(), theta, (rsum if graphed else total).cpu().numpy(), np.array(ag._step)))
    for a, b, what in zip(outs[0], outs[1], ("medium", "agents", "theta", "reward sum", "call counter")):
        assert np.array_equal(a, b), f"{what} differs between the eager and the graphed loop"


def test_graphed_loop_notices_changed_dynamics_and_recreated_buffers():
    import die_b200 as D
    _, env = make_pair((32, 32), seed=1)
    loop = D.GraphedLoop(env, D.BrownianAgent())
    loop.run(4)
    env.dynamics.rate_feed = 0.3
    with pytest.raises(RuntimeError, match="dynamics changed"):
        loop.run(2)
    loop = D.GraphedLoop(env, D.BrownianAgent())
    loop.run(2)
    env.dynamics.rate_feed = 0.1
    env.reset()                      # under changed dynamics reset() re-creates the env (new handle, new buffers)
    with pytest.raises(RuntimeError, match="reset"):
        loop.run(2)


@pytest.mark.parametrize("agent_kind", ["physarum", "brownian"])
def test_graphed_loop_survives_an_in_place_reset_and_edited_tensors(agent_kind):
    """Env.reset() writes the new state into the existing tensors, so a captured loop stays bound to live buffers; its
    baked-in use of the env's caches (cells, gradient) is re-validated by eager iterations at the start of run().  The
    same after set_state / an in-place edit of the medium.  Twin: the eager loop on an identically seeded env."""
    import torch
    import die_b200 as D
    field = (48, 40)
    envs = [D.Env(field, D.Dynamics(init_agent_ratio=0.2), init='device', seed=9) for _ in range(2)]
    def make_agent():
        if agent_kind == "brownian":
            return D.BrownianAgent(0.02, seed=4)
        return D.PhysarumAgent(max_agents=envs[0].max_agents, seed=5, scale=0.02, turn_angle=30, sense_offset=0.05)
    agents = [make_agent() for _ in range(2)]
    loop = D.GraphedLoop(envs[0], agents[0])                 # 2 eager warm-up iterations
    loop.run(6)
    def eager(n):
        obs = envs[1]._get_current_obs
        for _ in range(n):
            obs, *_ = envs[1].step(agents[1].forward(obs))
    eager(8)
    ptrs = (envs[0]._medium_buf[0].data_ptr(), envs[0]._agents.data_ptr())
    for e in envs:
        e.reset()
    assert ptrs == (envs[0]._medium_buf[0].data_ptr(), envs[0]._agents.data_ptr())
    assert all(torch.equal(a, b) for a, b in zip(envs[0]._get_current_obs, envs[1]._get_current_obs))
    loop.run(7)
    eager(7)
    assert all(np.array_equal(a, b) for a, b in zip(envs[0].get_state(), envs[1].get_state()))
    for e in envs:                                           # an edit through the public tensors
        e.medium[2].mul_(0.5)
        e.agents[3].add_(0.25)
    loop.run(5)
    eager(5)
    assert all(np.array_equal(a, b) for a, b in zip(envs[0].get_state(), envs[1].get_state()))
    if agent_kind == "physarum":
        assert np.array_equal(agents[0].get_state()[0], agents[1].get_state()[0])


def test_graphed_loop_refuses_what_it_cannot_capture():
    import die_b200 as D
    field = (32, 32)
    flow = D.WaveSequence(field, dt=0.01).get_flow_operator(scale=0.5, decay=0.5)
    _, env = make_pair(field, seed=1, dynamics_kw=dict(op_food_flow=flow))
    with pytest.raises(NotImplementedError):
        D.GraphedLoop(env, D.BrownianAgent())
    _, env = make_pair(field, seed=1)
    with pytest.raises(NotImplementedError):
        D.GraphedLoop(env, D.BrownianAgent(rng='numpy'))
