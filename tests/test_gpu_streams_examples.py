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


@pytest.mark.parametrize("agent_kind", ["physarum", "gradient"])
def test_chunked_host_path_equals_device_path(agent_kind):
    """die_env_step_host / die_gradient_forward_host cut a batch into chunks of environments on two internal streams
    (PCIe both ways at once).  With the chunk threshold lowered to zero, a batched run through numpy buffers -- in-kernel
    Philox draws included -- must reproduce the device-tensor run bit for bit."""
    import die_b200 as D
    from die_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.die_set_tuning(b"host_chunk_min_kb", 0))
    _lib.check(lib.die_set_tuning(b"host_chunks", 3))
    try:
        B, field = 8, (48, 40)
        refs, dev_env = make_pair(field, seed=50, batch=B)
        _, host_env = make_pair(field, seed=50, batch=B)
        m = dev_env.max_agents
        th = np.stack([lattice_theta(m, 30, 60 + b)[0] for b in range(B)])
        if agent_kind == "physarum":
            mk = lambda: D.PhysarumAgent(max_agents=m, seed=9, **PHYS)
        else:
            mk = lambda: D.GradientAgent(max_agents=m, seed=9, scale=0.01, inertia=0.9, noise_scale=0.025, sense_offset=0.03)
        a_dev, a_host = mk(), mk()
        prev = np.random.default_rng(5).normal(0, 0.4, (B, 2, m))
        for a in (a_dev, a_host):
            a.use_env_hints = False            # the host path cannot use the env's caches: compare like with like
            a._lazy_init(B, m, dev_env.device)
            a.set_state(theta=th, prev_grad=prev)
        dobs = dev_env._get_current_obs
        hobs = tuple(t.cpu().numpy() for t in host_env._get_current_obs)
        for it in range(8):
            dact = a_dev.forward(dobs)
            hact = a_host.forward(hobs)
            assert isinstance(hact, np.ndarray) and np.array_equal(dact.cpu().numpy(), hact), it
            dobs, dr, _, _, di = dev_env.step(dact)
            hobs, hr, _, _, hi = host_env.step(hact)
            assert np.array_equal(dr, hr) and np.array_equal(di['num_agents'], hi['num_agents'])
            assert np.array_equal(dobs[0].cpu().numpy(), hobs[0]) and np.array_equal(dobs[1].cpu().numpy(), hobs[1])
            # from the second step on the alive channel stays in the pinned buffer (DIE_HOST_KEEP_ALIVE_CHANNEL); the host
            # observation is a read-only copy
            assert host_env.last_step_kept_alive_channel == (it > 0) and not hobs[0].flags.writeable
            assert host_env.host_io_bytes_per_step()[1] == 8 * B * ((3 if it > 0 else 4) * m + 3 * field[0] * field[1]) + 16 * B
        # an edit of the agents tensor (here: half of the agents die) is followed by a full download
        for env in (dev_env, host_env):
            env.agents[:, 2, ::2] = 0.0
        dobs, *_ = dev_env.step(a_dev.forward(dobs))
        hobs, *_ = host_env.step(a_host.forward(tuple(np.array(t.cpu().numpy()) for t in host_env._get_current_obs)))
        assert not host_env.last_step_kept_alive_channel
        assert np.array_equal(dobs[0].cpu().numpy(), hobs[0]) and (hobs[0][:, 2, ::2] == 0).all()
    finally:
        _lib.check(lib.die_set_tuning(b"host_chunk_min_kb", 32 << 10))
        _lib.check(lib.die_set_tuning(b"host_chunks", 4))
