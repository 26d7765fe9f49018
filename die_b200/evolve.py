"""Population evaluation for a neuro-evolution search over ``NeuralAutomataAgent`` weights.

The reference's training example (examples/learning_agents.py:21-63) hands evotorch a fitness function that runs ONE
candidate for ``epoch_iters`` iterations of ``action = agent.forward(obs); obs, reward, ... = env.step(action)`` on ONE env
and returns the summed reward; a population of 10 is scored one candidate after the other.  Here the whole population is
scored at once: a batched ``Env`` with one environment per candidate, ``agent.set_population`` giving environment p the
weights of candidate p (die_conv_policy_forward_population), the loop replayed from a CUDA graph with the rewards summed on
the device -- one host synchronisation per generation.

``PGPE`` is a compact restatement of the searcher that example configures (PGPE with symmetric sampling, centred ranks,
the ClipUp optimiser; Sehnke et al. 2010, Toklu et al. 2020) so that the example runs where evotorch is absent; any other
ask / tell searcher can drive ``PopulationEvaluator.evaluate`` the same way.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .env import Env
from .graph import GraphedLoop


class PopulationEvaluator:
    """fitness[p] = sum over ``iters`` steps of the reward of environment p under candidate p.

    ``env``: a batched Env (``batch`` == population size); ``agent``: a NeuralAutomataAgent.  ``reset=True`` starts every
    evaluation from a freshly initialised env (``env.reset()``); the reference's example does not (each candidate continues
    where the previous one left the env), which is ``reset=False``.  ``graph=False`` runs the eager loop instead of
    the captured one (same results)."""

    def __init__(self, env: Env, agent, graph: bool = True):
        if not hasattr(agent, 'set_population'):
            raise TypeError(f"{type(agent).__name__} has no per-environment weights (set_population)")
        self.env, self.agent, self.graph = env, agent, graph
        self._loop: Optional[GraphedLoop] = None
        self._pop_ptr = None
        self.evaluations = 0

    @property
    def population_size(self) -> int:
        return self.env.batch

    def evaluate(self, vectors, iters: int, reset: bool = False) -> np.ndarray:
        v = torch.as_tensor(vectors, dtype=torch.float32)
        if v.dim() != 2 or v.shape[0] != self.env.batch:
            raise ValueError(f"expected [{self.env.batch}, num_parameters] candidates, got {tuple(v.shape)}")
        self.agent.set_population(v)
        if reset:
            self.env.reset()               # in place: a captured loop stays valid (GraphedLoop.run re-validates the caches)
            if self._loop is not None and self._loop._generation != self.env._generation:
                self._loop = None
        self.evaluations += 1
        if not self.graph:
            total = torch.zeros(self.env.batch, dtype=torch.float64, device=self.env.device)
            obs = self.env._get_current_obs
            for _ in range(int(iters)):
                obs, reward, _ = self.env.step_async(self.agent.forward(obs))
                total += reward
            return total.cpu().numpy()
        pop = self.agent.model.population
        if self._loop is None or self._pop_ptr != pop.data_ptr():
            # the graph is bound to the population's device buffer (same-sized populations are written in place).  The
            # capture's two eager warm-up iterations run under this population: part of the env's history, not of a score
            self._loop = GraphedLoop(self.env, self.agent, warmup=0)
            self._pop_ptr = pop.data_ptr()
        self._loop.reward_sum.zero_()
        self._loop.run(int(iters))
        return self._loop.reward_sum.cpu().numpy()


class ClipUp:
    """Toklu et al. 2020: normalised gradient step with momentum, the velocity clipped to ``max_speed``."""

    def __init__(self, n: int, stepsize: float, max_speed: float, momentum: float = 0.9):
        self.stepsize, self.max_speed, self.momentum = float(stepsize), float(max_speed), float(momentum)
        self.velocity = np.zeros(n)

    def ascent(self, grad: np.ndarray) -> np.ndarray:
        norm = float(np.linalg.norm(grad))
        step = grad / norm * self.stepsize if norm > 0 else np.zeros_like(grad)
        self.velocity = self.momentum * self.velocity + step
        vnorm = float(np.linalg.norm(self.velocity))
        if vnorm > self.max_speed:
            self.velocity *= self.max_speed / vnorm
        return self.velocity


def centered_ranks(fitness: np.ndarray) -> np.ndarray:
    """Ranks mapped linearly to [-0.5, 0.5] (the worst candidate -0.5, the best +0.5)."""
    f = np.asarray(fitness, dtype=np.float64)
    ranks = np.empty(f.size)
    ranks[np.argsort(f, kind='stable')] = np.arange(f.size)
    return ranks / max(f.size - 1, 1) - 0.5


class PGPE:
    """Parameter-exploring policy gradients with symmetric sampling (maximisation).

    ask() -> [popsize, n] candidates (pairs centre + eps, centre - eps); tell(fitness) moves the centre along the
    ranked-fitness-weighted mean of eps (through ClipUp) and adapts the per-parameter standard deviation."""

    def __init__(self, num_parameters: int, popsize: int = 10, radius_init: float = 1.5,
                 center_learning_rate: Optional[float] = None, stdev_learning_rate: float = 0.1,
                 max_speed: Optional[float] = None, momentum: float = 0.9, stdev_max_change: float = 0.2,
                 center_init: Optional[np.ndarray] = None, initial_bounds=(-0.5, 0.5), seed: Optional[int] = None):
        if popsize < 2 or popsize % 2:
            raise ValueError("symmetric sampling needs an even population size")
        self.n, self.popsize = int(num_parameters), int(popsize)
        self.rng = np.random.default_rng(seed)
        self.center = (np.asarray(center_init, dtype=np.float64).copy() if center_init is not None
                       else self.rng.uniform(initial_bounds[0], initial_bounds[1], self.n))
        self.stdev = np.full(self.n, radius_init / np.sqrt(self.n))
        max_speed = radius_init / 15. if max_speed is None else max_speed               # the example's rule of thumb
        lr = max_speed / 2. if center_learning_rate is None else center_learning_rate
        self.optimizer = ClipUp(self.n, lr, max_speed, momentum)
        self.stdev_learning_rate, self.stdev_max_change = float(stdev_learning_rate), float(stdev_max_change)
        self._eps = None
        self.generation = 0
        self.best = (-np.inf, None)              # (fitness, candidate) over all generations: "pop_best" of each, kept

    def ask(self) -> np.ndarray:
        self._eps = self.rng.standard_normal((self.popsize // 2, self.n)) * self.stdev
        out = np.empty((self.popsize, self.n))
        out[0::2] = self.center + self._eps
        out[1::2] = self.center - self._eps
        self._asked = out
        return out

    def tell(self, fitness) -> None:
        f = np.asarray(fitness, dtype=np.float64)
        if self._eps is None or f.shape != (self.popsize,):
            raise ValueError("tell() takes the fitness of the candidates of the last ask()")
        k = int(np.argmax(f))
        if f[k] > self.best[0]:
            self.best = (float(f[k]), self._asked[k].copy())
        u = centered_ranks(f)
        up, um = u[0::2], u[1::2]
        grad_center = ((up - um) / 2.)[:, None] * self._eps
        grad_center = grad_center.mean(axis=0)
        baseline = u.mean()
        grad_stdev = (((up + um) / 2. - baseline)[:, None] * (self._eps ** 2 - self.stdev ** 2) / self.stdev).mean(axis=0)
        self.center = self.center + self.optimizer.ascent(grad_center)
        change = np.clip(self.stdev_learning_rate * grad_stdev,
                         -self.stdev_max_change * self.stdev, self.stdev_max_change * self.stdev)
        self.stdev = self.stdev + change
        self._eps = None
        self.generation += 1
