#!/usr/bin/env python
"""The reference's examples/simple_agents.py on die_b200: the four hand-written policies (ConstAgent, BrownianAgent,
GradientAgent, PhysarumAgent with the example's own parameters, examples/simple_agents.py:41-73) on the two dynamics it
offers ('st-perlin': static food; 'dyn-pred': the WaveSequence food flow, :95-100).  Same calls as the reference, only
the imports differ; the interactive matplotlib plotter is replaced by the frames of `Env.render` (the arrays the
reference's EnvRenderer hands to matplotlib).  Needs a CUDA device (there is no CPU fallback).

    python examples/simple_agents.py [--agent const|rand|grad|physarum|jones] [--dynamics st-perlin|dyn-pred]
                                     [--field 156] [--iters 1000] [--ratio 0.1] [--frames-every 0]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from die_b200 import Env, Dynamics, ConstAgent, BrownianAgent, GradientAgent, PhysarumAgent, JonesAgent, WaveSequence


def try_const_agent(**kwargs):
    return ConstAgent(delta_xy=(-0.01, 0.005), deposit=0.1)


def try_random_agent(**kwargs):
    return BrownianAgent(move_scale=0.01, deposit_scale=0.1)


def try_gradient_agent(num_agents, **kwargs):
    return GradientAgent(num_agents, sense_offset=0.03, inertia=0.95, scale=0.01, deposit=4.5, noise_scale=0.025,
                         normalized_grad=True)


def try_physarum_agent(num_agents, **kwargs):
    return PhysarumAgent(num_agents, turn_angle=35, sense_angle=120, sense_offset=0.03, turn_tolerance=0.05,
                         inertia=0., scale=0.0075, deposit=4.5, noise_scale=0.0, normalized_grad=True)


def try_jones_agent(num_agents, **kwargs):
    """Not in the reference's example: the classic three-sensor particle (die_b200.JonesAgent) with comparable parameters."""
    return JonesAgent(num_agents, turn_angle=45, sense_angle=45, sense_offset=0.03, scale=0.0075, deposit=4.5)


def run_agent(env, agent, iters=1000, frames_every=0):
    total_reward = 0
    obs = env._get_current_obs
    t0 = time.perf_counter()
    for i in range(iters):
        action = agent.forward(obs)
        obs, reward, terminated, _, stats = env.step(action)
        total_reward += reward
        if i % 100 == 0 or i == iters - 1:
            print(f"iter {i:5d}  total_reward {np.round(total_reward, 3)}  {stats}")
        if frames_every and i % frames_every == 0:
            medium_rgb, trace, agents_rgba = env.render(host=True)      # what InteractivePlotter.draw() would show
        if terminated:
            break
    return total_reward, (time.perf_counter() - t0) / max(i + 1, 1)


def run_experiment(field_size=156, agent_id='rand', dynamics_id='st-perlin', iters=1000, agent_ratio=0.15, frames_every=0):
    max_agents = field_size * field_size
    field_size = (field_size, field_size)
    agents = {
        'const': try_const_agent,
        'rand': try_random_agent,
        'grad': lambda: try_gradient_agent(max_agents),
        'physarum': lambda: try_physarum_agent(max_agents),
        'jones': lambda: try_jones_agent(max_agents),
    }
    wave_flow = WaveSequence(field_size, dt=0.01).get_flow_operator(scale=0.5, decay=0.5)
    dynamics_choice = {
        'st-perlin': Dynamics(init_agent_ratio=agent_ratio, food_infinite=False),
        'dyn-pred': Dynamics(init_agent_ratio=agent_ratio, food_infinite=False, op_food_flow=wave_flow),
    }
    env = Env(field_size, dynamics_choice[dynamics_id])
    agent = agents[agent_id]()
    total, per_iter = run_agent(env, agent, iters=iters, frames_every=frames_every)
    print(f"{agent_id} on {dynamics_id} {field_size}: total reward {total:.3f}, {per_iter * 1e3:.3f} ms per iteration")
    frames = env.render(host=True)
    print("frames:", [f.shape for f in frames])
    return total


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument("--agent", default="grad", choices=["const", "rand", "grad", "physarum", "jones"])
    ap.add_argument("--dynamics", default="st-perlin", choices=["st-perlin", "dyn-pred"])
    ap.add_argument("--field", type=int, default=156)
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--ratio", type=float, default=0.1)
    ap.add_argument("--frames-every", type=int, default=0)
    a = ap.parse_args()
    run_experiment(field_size=a.field, agent_id=a.agent, dynamics_id=a.dynamics, iters=a.iters, agent_ratio=a.ratio,
                   frames_every=a.frames_every)
