#!/usr/bin/env python
"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE (/root/reference/core/*.py,
unmodified) over the stand-in packages of oracle/shims (see oracle/shims/README.md).

    python oracle/make_golden_from_reference.py

Fixtures (float64, seeded; legacy np.random streams are frozen across numpy versions):
  brownian_24x32.npz   initial state, seed, 25 steps: per-step action / reward / num_agents, final state.
                       The uniforms come from np.random.seed(seed) in the reference's own draw order.
  physarum_24x32.npz   25 steps, README agent parameters: for EVERY step the pre-step state (medium,
                       agents, theta), the coin flips, the action, the post-step state, reward, info.
  physarum_limit_sigma08_20x20.npz   as above with boundary='limit', diffuse_sigma=0.8 (radius-3 blur),
                       food_infinite, zero_cost, turn_angle=35, sense_angle=120.
  physarum_waveflow_24x32.npz   as the first with op_food_flow = WaveSequence((24, 32), dt=0.01)
                       .get_flow_operator(scale=0.5, decay=0.5); step k uses the sequence's k-th time step.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import run_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def brownian(ref, name, size, steps, seed, **akw):
    np.random.seed(seed)
    env = ref.Env(size, ref.Dynamics(init_agent_ratio=0.1))
    medium0, agents0 = env.medium.values.copy(), env.agents.values.copy()
    agent = ref.BrownianAgent(**akw)
    np.random.seed(seed + 1)                       # stream used by the step loop
    obs = env._get_current_obs
    actions, rewards, alive = [], [], []
    for _ in range(steps):
        act = agent.forward(obs)
        obs, r, _, _, info = env.step(act)
        actions.append(act.values.copy())
        rewards.append(r)
        alive.append(info['num_agents'])
    np.savez_compressed(os.path.join(OUT, name), size=np.array(size), loop_seed=seed + 1,
                        move_scale=akw.get('move_scale', 0.01), deposit_scale=akw.get('deposit_scale', 0.5),
                        medium0=medium0, agents0=agents0, actions=np.array(actions), rewards=np.array(rewards),
                        num_agents=np.array(alive), medium_final=env.medium.values, agents_final=env.agents.values)


def physarum(ref, name, size, steps, seed, dyn_kw, agent_kw):
    np.random.seed(seed)
    env = ref.Env(size, ref.Dynamics(init_agent_ratio=0.1, **dyn_kw))
    m = env.agents.shape[-1]
    agent = ref.PhysarumAgent(max_agents=m, **agent_kw)
    # the agent's private generator is unseeded in the reference: pin its initial state
    rng = np.random.default_rng(seed)
    agent._prev_grad = rng.normal(0., 0.4, (2, m))
    agent._direction_rads = agent._discretize_grad(agent._prev_grad)
    np.random.seed(seed + 1)
    obs = env._get_current_obs
    rec = {k: [] for k in ('medium_pre', 'agents_pre', 'theta_pre', 'coin', 'action', 'medium_post',
                           'agents_post', 'theta_post', 'reward', 'num_agents')}
    for _ in range(steps):
        rec['medium_pre'].append(env.medium.values.copy())
        rec['agents_pre'].append(env.agents.values.copy())
        rec['theta_pre'].append(np.array(agent._direction_rads).copy())
        state = np.random.get_state()
        coin = np.random.randint(0, 2, m)            # the draw _choose_turn is about to make (gradient.py:181)
        np.random.set_state(state)
        act = agent.forward(obs)
        obs, r, _, _, info = env.step(act)
        rec['coin'].append(coin.astype(np.uint8))
        rec['action'].append(act.values.copy())
        rec['medium_post'].append(env.medium.values.copy())
        rec['agents_post'].append(env.agents.values.copy())
        rec['theta_post'].append(np.array(agent._direction_rads).copy())
        rec['reward'].append(r)
        rec['num_agents'].append(info['num_agents'])
    np.savez_compressed(os.path.join(OUT, name), size=np.array(size), loop_seed=seed + 1,
                        **{k: np.array(v) for k, v in rec.items()})


def main():
    ref = run_reference.load()
    os.makedirs(OUT, exist_ok=True)
    brownian(ref, "brownian_24x32.npz", (24, 32), 25, 11, move_scale=0.01)
    physarum(ref, "physarum_24x32.npz", (24, 32), 25, 21, {},
             dict(scale=0.007, turn_angle=30, sense_offset=0.04))
    physarum(ref, "physarum_limit_sigma08_20x20.npz", (20, 20), 15, 31,
             dict(boundary=ref.BoundaryCondition.limit, diffuse_sigma=0.8, food_infinite=True,
                  op_action_cost=ref.zero_cost),
             dict(scale=0.03, turn_angle=35, sense_angle=120, sense_offset=0.06, turn_tolerance=0.05))
    # Dynamics.op_food_flow = WaveSequence flow (examples/simple_agents.py:95-100, 'dyn-pred')
    wave = ref.WaveSequence((24, 32), dt=0.01).get_flow_operator(scale=0.5, decay=0.5)
    physarum(ref, "physarum_waveflow_24x32.npz", (24, 32), 12, 41, dict(op_food_flow=wave),
             dict(scale=0.007, turn_angle=30, sense_offset=0.04))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
