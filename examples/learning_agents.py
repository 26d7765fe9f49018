#!/usr/bin/env python
"""The reference's examples/learning_agents.py on die_b200: evolve the weights of a NeuralAutomataAgent (a stack of small
circular convolutions over the medium) with PGPE, fitness = the reward summed over `epoch_iters` iterations.

The reference scores its population of 10 one candidate after the other on one env through evotorch; here every
generation is ONE batched run (one environment per candidate, one model per environment, the loop replayed from a CUDA
graph) and the searcher is the compact PGPE + ClipUp of die_b200/evolve.py (evotorch, mlflow and the plotting of the
reference's example are not needed).  Needs a CUDA device.

    python examples/learning_agents.py [--field 156] [--epochs 100] [--epoch-iters 50] [--popsize 10]
                                       [--dynamics st-perlin|st-perlin-wide] [--reset] [--out saved_models/agent.pt]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from die_b200 import Env, Dynamics, NeuralAutomataAgent, PopulationEvaluator, PGPE


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--field", type=int, default=156)
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--epoch-iters", type=int, default=50)
    ap.add_argument("--popsize", type=int, default=10)
    ap.add_argument("--dynamics", default="st-perlin", choices=["st-perlin", "st-perlin-wide"])
    ap.add_argument("--agent-ratio", type=float, default=0.15)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--reset", action="store_true", help="start every generation from a freshly initialised env "
                    "(less noisy scores; the reference's example lets the env run on)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    field_size = (args.field, args.field)
    dynamics = {
        'st-perlin': Dynamics(init_agent_ratio=args.agent_ratio, food_infinite=True),
        'st-perlin-wide': Dynamics(init_agent_ratio=args.agent_ratio, food_infinite=True,
                                   rate_decay_chem=0.025, diffuse_sigma=.8),
    }[args.dynamics]
    env = Env(field_size, dynamics, init='device', seed=args.seed, batch=args.popsize)
    agent = NeuralAutomataAgent(kernel_sizes=[3, 3], scale=0.01, deposit=2.0)
    n = agent.model.num_parameters
    print(f'Network has {n} parameters; population {args.popsize}, {args.epoch_iters} iterations per evaluation')

    radius_init = 1.5
    max_speed = radius_init / 15.
    searcher = PGPE(n, popsize=args.popsize, radius_init=radius_init, center_learning_rate=max_speed / 2.,
                    stdev_learning_rate=0.1, max_speed=max_speed, momentum=0.9, seed=args.seed)
    evaluator = PopulationEvaluator(env, agent)
    t0 = time.perf_counter()
    for epoch in range(args.epochs):
        candidates = searcher.ask()
        fitness = evaluator.evaluate(candidates, args.epoch_iters, reset=args.reset)
        searcher.tell(fitness)
        if epoch % 10 == 0 or epoch == args.epochs - 1:
            print(f'epoch {epoch:4d}  mean {fitness.mean():10.3f}  best {fitness.max():10.3f}  '
                  f'best so far {searcher.best[0]:10.3f}  stdev {searcher.stdev.mean():.4f}')
    dt = time.perf_counter() - t0
    steps = args.epochs * args.epoch_iters * args.popsize
    print(f'{steps} env iterations in {dt:.2f} s = {steps * args.field ** 2 / dt / 1e6:.1f} M cell-updates/s')

    agent.set_population(None)
    # the learned distribution's centre and the best single candidate seen (evotorch's "pop_best", what the reference's
    # example saves), each on a fresh single env as the reference's example does at its end
    single_dyn = dynamics
    results = {}
    for name, vec in (('center', searcher.center), ('pop_best', searcher.best[1])):
        agent.model.set_parameters_vector(vec)
        single = Env(field_size, single_dyn, init='device', seed=args.seed + 1)
        obs, total = single._get_current_obs, 0.
        for _ in range(args.epoch_iters):
            obs, reward, _, _, stats = single.step(agent.forward(obs))
            total += reward
        results[name] = total
        print(f'Reward of {name} over {args.epoch_iters} iterations on a fresh env: {np.round(total, 3)}  {stats}')
    agent.model.set_parameters_vector(searcher.best[1])
    if args.out:
        os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
        agent.save(args.out)                 # the reference's TorchAgent file format: loads in the reference, too
        print(f'saved the best agent to {args.out}')


if __name__ == '__main__':
    main()
