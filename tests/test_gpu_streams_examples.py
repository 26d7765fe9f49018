"""The library enqueues everything on the caller's current stream and keeps no global mutable state: two
environments stepped on two different non-default streams must give the results of the same runs on the
default stream.  Plus: the reference's canonical loop (examples/minimal_run.py) runs end to end."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests._parity import make_pair, lattice_theta

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)


def _run(env, agent, steps, stream=None):
    import torch
    ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
    with ctx:
        obs = env._get_current_obs
        for _ in range(steps):
            obs, *_ = env.step_async(agent.forward(obs))
    return env


def test_two_envs_on_two_streams():
    import torch
    import die_b200 as D
    results = {}
    for mode in ("default", "streams"):
        envs, agents = [], []
        for k in range(2):
            (_,), env = make_pair((160, 128), seed=40 + k)
            ag = D.PhysarumAgent(max_agents=env.max_agents, seed=3 + k, **PHYS)
            ag.set_state(theta=lattice_theta(env.max_agents, 30, 40 + k)[0])
            envs.append(env)
            agents.append(ag)
        torch.cuda.synchronize()
        if mode == "default":
            for env, ag in zip(envs, agents):
                _run(env, ag, 25)
        else:
            streams = [torch.cuda.Stream(), torch.cuda.Stream()]
            for s in streams:
                s.wait_stream(torch.cuda.current_stream())
            for it in range(5):                      # interleave the two loops so their kernels overlap
                for env, ag, s in zip(envs, agents, streams):
                    _run(env, ag, 5, stream=s)
            for s in streams:
                torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        results[mode] = [(e.get_state(), a.get_state()[0]) for e, a in zip(envs, agents)]
    for (s0, t0), (s1, t1) in zip(results["default"], results["streams"]):
        assert np.array_equal(s0[0], s1[0]) and np.array_equal(s0[1], s1[1]) and np.array_equal(t0, t1)


@pytest.mark.parametrize("argv", [["--agent", "brownian", "--field", "64", "--iters", "20"],
                                  ["--agent", "physarum", "--field", "96", "--iters", "20", "--waves"]])
def test_minimal_run_example(argv):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "minimal_run.py"), *argv],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ms per iteration" in out.stdout and "frames:" in out.stdout
