"""The reference's run loop (examples/minimal_run.py:21-25, README.md:30-34) captured as a CUDA graph.

    loop = GraphedLoop(env, agent)
    obs, reward_sum = loop.run(300)          # == 300 x { action = agent.forward(obs); obs, reward, ... = env.step(action) }

ONE small environment is bound by the host, not by the GPU: ~5 kernel launches of a few microseconds each per iteration
cost ~40 us of CPU time through Python + ctypes (and `Env.step` adds a host synchronisation for the reward).  Here two
consecutive iterations -- one period of the medium's ping-pong buffers -- are captured once (torch.cuda.CUDAGraph over the
library's launches) and replayed, so an iteration costs its kernels' time.  What a replay cannot change are kernel
arguments, so everything that varies per iteration lives on the device:
  * the agents' call counter (the Philox key of the in-kernel random draws): a uint64 tensor the kernels read
    (DIE_FWD_STEP_ON_DEVICE, die_brownian_forward_dev) and a captured `counter += 1` advances -- every replayed
    iteration draws the numbers the eager loop would have drawn, so results are bit-identical to the eager loop;
  * rewards: accumulated into `reward_sum` (and the last one kept) on the device; read them back when needed.
`Env.reset()` (in place), `set_state` and edits of the env's tensors between runs are fine: `run` notices stale caches and
takes up to two eager iterations before it replays.  Restrictions (checked): device observations, in-kernel randomness (rng='philox': host draws cannot be captured), the
identity or wave / tabulated food flow is NOT supported (its time index is host state), no speculative move.
"""
from __future__ import annotations

import torch

from .env import Env, identity_food_flow


class GraphedLoop:
    def __init__(self, env: Env, agent, warmup: int = 2):
        if env.dynamics.op_food_flow is not identity_food_flow and env.dynamics.op_food_flow is not None:
            raise NotImplementedError("GraphedLoop: a time-varying food flow advances a host-side time index every step")
        if getattr(agent, '_rng_mode', getattr(agent, '_rng', 'philox')) != 'philox':
            raise NotImplementedError("GraphedLoop: host-side random draws (rng='numpy') cannot be captured; use rng='philox'")
        if not hasattr(agent, '_step_dev'):
            raise NotImplementedError(f"GraphedLoop: {type(agent).__name__} has no device-resident call counter")
        if getattr(agent, 'fuse_move', False) not in (False, 'commit'):
            # ('commit' is fine: this loop IS the contract it needs, and the decision is taken once, at capture time)
            raise NotImplementedError("GraphedLoop: the speculative move depends on host-side version counters")
        self.env, self.agent = env, agent
        dev = env.device
        with torch.cuda.device(dev):
            # eager iterations first: every lazy allocation (gradient cache, action buffer, alive bitmask) happens here,
            # and the medium is back in its first buffer after an even number of them
            obs = env._get_current_obs
            for _ in range(2 * max(1, (warmup + 1) // 2)):
                obs, _, _ = env.step_async(agent.forward(obs))
            self._counter = torch.full((1,), int(agent._step), dtype=torch.int64, device=dev)
            self.reward_sum = torch.zeros(env.batch, dtype=torch.float64, device=dev)
            self.last_reward = torch.zeros(env.batch, dtype=torch.float64, device=dev)
            self.last_alive = torch.zeros(env.batch, dtype=torch.int64, device=dev)
            self.iterations = 0
            agent._step_dev = self._counter
            step0, cur0 = agent._step, env._cur
            torch.cuda.synchronize(dev)
            self._dynamics_key = env._dynamics_key
            self._graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(self._graph):
                    for _ in range(2):
                        obs, reward, alive = env.step_async(agent.forward(obs))
                        self.reward_sum.add_(reward)
                        self._counter.add_(1)
                    self.last_reward.copy_(reward)
                    self.last_alive.copy_(alive)
            finally:
                agent._step_dev = None
            # capturing recorded the launches without running them: the host-side counters go back
            agent._step = step0
            assert env._cur == cur0
            self._cur0 = cur0
            self._generation = env._generation           # Env.reset() re-creates its buffers: the graph is bound to these
            self._obs = obs

    def run(self, iterations: int):
        """`iterations` iterations of forward + step -> (obs, reward_sum[B] over every iteration run so far, on device)."""
        env, agent = self.env, self.agent
        if env._handle is None or env._generation != self._generation:
            raise RuntimeError("the Env was reset (new buffers) after this GraphedLoop was captured: build a new GraphedLoop")
        env._sync_dynamics()
        if env._dynamics_key != self._dynamics_key:
            raise RuntimeError("env.dynamics changed after this GraphedLoop was captured (kernel arguments are baked into "
                               "the graph): build a new GraphedLoop")
        with torch.cuda.device(env.device):
            # eager iterations first (at most two) while the replay's baked-in assumptions do not hold: an odd run left the
            # medium in the other buffer; or the env's caches are stale (an in-place reset(), set_state, an edited tensor)
            # -- the captured forward uses them as the eager one did at capture time, an eager step rebuilds them
            while iterations > 0 and (env._cur != self._cur0 or not env._hints_are_current()):
                self._eager_step()
                iterations -= 1
                self.iterations += 1
            self._counter.fill_(int(agent._step))
            for _ in range(iterations // 2):
                self._graph.replay()
            agent._step += 2 * (iterations // 2)
            if iterations % 2:                               # an odd tail runs eagerly
                self._eager_step()
        self.iterations += iterations
        return env._get_current_obs, self.reward_sum

    def _eager_step(self):
        env, agent = self.env, self.agent
        obs, reward, alive = env.step_async(agent.forward(env._get_current_obs))
        self.reward_sum.add_(reward)
        self.last_reward.copy_(reward)
        self.last_alive.copy_(alive)
