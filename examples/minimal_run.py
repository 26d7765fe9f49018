#!/usr/bin/env python
"""The reference's examples/minimal_run.py (and the README loop, README.md:23-39) on die_b200: the same calls,
only the import differs.  Needs a CUDA device (there is no CPU fallback).

    python examples/minimal_run.py [--agent brownian|physarum] [--field 256] [--iters 300] [--waves]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from die_b200 import Env, Dynamics, BrownianAgent, PhysarumAgent, WaveSequence


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agent", default="physarum", choices=["brownian", "physarum"])
    ap.add_argument("--field", type=int, default=256)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--waves", action="store_true", help="the 'dyn-pred' dynamics of examples/simple_agents.py")
    args = ap.parse_args()

    field_size = (args.field, args.field)
    dynamics = Dynamics(init_agent_ratio=0.1)
    if args.waves:
        dynamics.op_food_flow = WaveSequence(field_size, dt=0.01).get_flow_operator(scale=0.5, decay=0.5)
    env = Env(field_size, dynamics, init='device' if args.field > 1024 else 'host')
    if args.agent == "brownian":
        agent = BrownianAgent(move_scale=0.01)
    else:
        agent = PhysarumAgent(max_agents=args.field ** 2, scale=0.007, turn_angle=30, sense_offset=0.04)

    total_reward = 0
    obs = env._get_current_obs
    t0 = time.perf_counter()
    for i in range(args.iters):
        action = agent.forward(obs)
        obs, reward, terminated, truncated, stats = env.step(action)
        total_reward += reward
        if i % 50 == 0 or i == args.iters - 1:
            print(f"iter {i:4d}  total_reward {np.round(total_reward, 3)}  {stats}")
        if terminated:
            break
    dt = time.perf_counter() - t0
    print(f"{args.iters} iterations of {args.agent} on {field_size}: {dt / args.iters * 1e3:.3f} ms per iteration "
          f"(including the per-step reward read-back)")
    frames = env.render(host=True)          # [medium (H, W, 3), agent trace (H, W), agents (W, H, 4)] as numpy arrays
    print("frames:", [f.shape for f in frames])


if __name__ == "__main__":
    main()
