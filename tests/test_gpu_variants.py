"""Variants staged at the end of round 1 without GPU budget, run on a B200 in the first GPU call of round 2
(profiles/r02a_staged_ab_summary.txt): everything here passed there and is an ordinary test now.

  grad_f32          the field pass publishes np.gradient(chem1) as float32 pairs for decision-only consumers (the
                    default: DESIGN.md 3.10); the A-B here is float64 vs float32 pairs
  unnormalised      the signed zeros of PhysarumAgent(normalized_grad=False) on a zero gradient
  tabulated flow    Dynamics.op_food_flow of any FieldSequence through tabulated frames
  simple_agents     examples/simple_agents.py (the reference's four hand-written policies on its two dynamics)
Removed after that call, on the numbers: the feed kernel's register caps (10-20 % slower) and the persistent
cp.async.bulk field pass of round 1 (correct on hardware, 2.3x slower than the tile kernel: one small bulk copy per tile
row); the TMA path of round 2 is the cluster-fused environment step (die_env_fused.cuh, tests/test_gpu_fused_step.py).
"""
import os

import numpy as np
import pytest

from tests._parity import make_pair, lattice_theta

pytestmark = pytest.mark.gpu
PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)


def _run(key, value, shape=(128, 96), steps=40, adversarial=False):
    import die_b200 as D
    from die_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.die_set_tuning(key.encode(), value))
    try:
        (_,), env = make_pair(shape, seed=13)
        if adversarial:
            med, _ag = env.get_state()
            rng = np.random.default_rng(5)
            noise = rng.random(shape)
            scales = [1e-300, 1e-50, 1e-46, 1e-44, 1e-40, 1e-10, 3e-6, 1e-5, 1.0, 1e20, 1e37, 1e39, 1e200, 0.0, 1e-5, 2e-5]
            for k, sc in enumerate(scales):
                med[2, 4 * k:4 * k + 4, shape[1] // 2:] = noise[4 * k:4 * k + 4, shape[1] // 2:] * sc
            env.set_state(medium=med)
        m = env.max_agents
        ag = D.PhysarumAgent(max_agents=m, seed=5, **PHYS)
        ag.set_state(theta=lattice_theta(m, 30, 13)[0])
        obs = env._get_current_obs
        total = 0.0
        for _ in range(steps):
            obs, r, *_ = env.step(ag.forward(obs))
            total += r
        return (*env.get_state(), ag.get_state()[0], total), ag.last_hints
    finally:
        _lib.check(lib.die_set_tuning(key.encode(), 1))          # the default of both switches


@pytest.mark.parametrize("adversarial", [False, True])
def test_float32_gradient_cache_does_not_change_results(adversarial):
    from die_b200 import _lib
    base, hints0 = _run("grad_f32", 0, adversarial=adversarial)
    n0 = _lib.load().die_get_counter(b"forward_lean_f32")
    out, hints1 = _run("grad_f32", 1, adversarial=adversarial)
    assert hints0 == hints1 == (True, True)
    assert _lib.load().die_get_counter(b"forward_lean_f32") > n0, "the float32 LEAN forward must be the one that ran"
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(base[:3], out[:3])) and base[3] == out[3]


def test_unnormalised_physarum_keeps_the_momentum_operands():
    """Found on the CPU emulator: with normalized_grad=False a zero gradient gives g' = (0 * cos d, 0 * sin d), signed
    zeros, and the reference's identity momentum step 1.0 * g' + 0.0 * prev + 0.0 * noise turns (-0.0) + (+0.0) into
    +0.0 -- so the new heading (pi only for (-0., -0.)) depends on the signs of prev_grad and of the noise
    draw.  GradientAgent._needs_prev now keeps those operands for every configuration whose g' can hold a zero."""
    import die_b200 as D
    from oracle import die_ref as R
    (ref,), env = make_pair((32, 48), seed=12)
    m = env.max_agents
    theta0, prev = lattice_theta(m, 30, 12)
    kw = dict(PHYS, normalized_grad=False)
    ra = R.PhysarumAgent(max_agents=m, prev_grad=prev, **kw)
    ga = D.PhysarumAgent(max_agents=m, **kw)
    ga.set_state(theta=theta0, prev_grad=prev)
    rng = np.random.default_rng(0)
    coin, noise = rng.integers(0, 2, m), rng.normal(0., 0.4, size=(2, m))
    ract = ra.forward(ref._get_current_obs, coin=coin.copy(), noise=noise.copy())      # chem1 == 0: every gradient is zero
    gact = ga.forward(env._get_current_obs, coin=coin, noise=noise).cpu().numpy()
    assert set(np.unique(ra._direction_rads)) <= {0.0, np.pi}
    assert np.array_equal(ga.get_state()[0], ra._direction_rads)
    assert np.array_equal(gact, ract)


@pytest.mark.parametrize("field,sigma,batch", [((48, 64), 0.5, None), ((37, 53), 0.8, 3)])
def test_tabulated_food_flow(field, sigma, batch):
    """Dynamics.op_food_flow of any FieldSequence through its tabulated frames (die_env_set_food_frames): bit-exact
    against the oracle, the iterator cycling (T = 4 over 10 steps) and surviving the env's own bookkeeping."""
    import die_b200 as D
    from oracle import die_ref as R
    from tests._parity import assert_state_equal
    rng = np.random.default_rng(9)
    frames = rng.normal(0.0, 0.3, size=(4, *field)).round(3)
    gflow = D.TabulatedSequence(frames).get_flow_operator(scale=0.5, decay=0.25)
    B = batch or 1
    refs = []
    for b in range(B):
        np.random.seed(5 + b)
        rflow = R.FrameSequence(frames).get_flow_operator(scale=0.5, decay=0.25)
        refs.append(R.Env(field, R.Dynamics(init_agent_ratio=0.1, op_food_flow=rflow, diffuse_sigma=sigma), noise_seed=5 + b))
    env = D.Env(field, D.Dynamics(init_agent_ratio=0.1, op_food_flow=gflow, diffuse_sigma=sigma), batch=batch,
                init_state=(np.stack([r.medium for r in refs]), np.stack([r.agents for r in refs])))
    ra, ga = R.BrownianAgent(0.01), D.BrownianAgent(move_scale=0.01)
    m = env.max_agents
    obs = env._get_current_obs
    for it in range(10):
        u = rng.random((B, 3, m))
        gact = ga.forward(obs, u=u if batch else u[0])
        for b in range(B):
            refs[b].step(ra.forward(refs[b]._get_current_obs, u=u[b]))
        obs, *_ = env.step(gact)
        med, ag = env.get_state()
        med, ag = (med, ag) if batch else (med[None], ag[None])
        for b in range(B):
            assert_state_equal(refs[b], med[b], ag[b], float_exact=True)
    assert gflow.calls == 10


@pytest.mark.parametrize("argv", [["--agent", "const"], ["--agent", "rand"], ["--agent", "grad", "--dynamics", "dyn-pred"],
                                  ["--agent", "physarum", "--dynamics", "dyn-pred", "--frames-every", "10"],
                                  ["--agent", "jones", "--frames-every", "10"]])
def test_simple_agents_example(argv):
    """examples/simple_agents.py: the reference's four hand-written policies on its two dynamics (callers)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "simple_agents.py"), "--field", "96",
                          "--iters", "30", *argv], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ms per iteration" in out.stdout and "frames:" in out.stdout


@pytest.mark.parametrize("shape,sigma,batch", [((256, 256), 0.5, 6), ((40, 72), 0.5, None), ((70, 200), 0.8, None),
                                               ((128, 96), 1.0, 5), ((96, 130), 0.3, 2)])
def test_vectorised_field_pass_does_not_change_results(shape, sigma, batch):
    """field_vec = 1 (field_step_vec_kernel, opt-in: no faster than the scalar tile kernel on a B200): 128-bit staging
    loads and output stores."""
    import die_b200 as D
    from die_b200 import _lib
    lib = _lib.load()
    outs = []
    n0 = lib.die_get_counter(b"field_vec")
    try:
        for vec in (0, 1):
            _lib.check(lib.die_set_tuning(b"field_vec", vec))
            _, env = make_pair(shape, seed=13, dynamics_kw=dict(diffuse_sigma=sigma), batch=batch)
            m = env.max_agents
            ag = D.PhysarumAgent(max_agents=m, seed=5, **PHYS)
            obs = env._get_current_obs
            for _ in range(12):
                obs, r, *_ = env.step(ag.forward(obs))
            outs.append((*env.get_state(), ag.get_state()[0], np.asarray(r)))
    finally:
        _lib.check(lib.die_set_tuning(b"field_vec", 0))
    assert lib.die_get_counter(b"field_vec") == n0 + 12, "the 128-bit kernel must be the one that ran"
    assert all(np.array_equal(a, b) for a, b in zip(*outs))


@pytest.mark.parametrize("shape,batch", [((512, 384), None), ((200, 136), 3)])
def test_pair_mode_does_not_change_results(shape, batch):
    """pair_mode: {consumed_field, new food} in one 16-byte pair per cell, the food under every slot handed from the
    feed kernel to the next forward pass -- forced here on fields below the automatic threshold."""
    import die_b200 as D
    from die_b200 import _lib
    lib = _lib.load()
    outs = []
    for mode in (0, 2):
        _lib.check(lib.die_set_tuning(b"pair_mode", mode))
        try:
            n0 = lib.die_get_counter(b"forward_food_here")
            refs, env = make_pair(shape, seed=21, batch=batch)
            m = env.max_agents
            ag = D.PhysarumAgent(max_agents=m, seed=5, **PHYS)
            obs = env._get_current_obs
            rewards = []
            for it in range(30):
                obs, r, *_ = env.step(ag.forward(obs))
                rewards.append(np.asarray(r).copy())
            assert lib.die_get_counter(b"forward_food_here") - n0 == (29 if mode else 0)
            outs.append((*env.get_state(), ag.get_state()[0], np.array(rewards)))
        finally:
            _lib.check(lib.die_set_tuning(b"pair_mode", 1))
    for a, b, what in zip(outs[0], outs[1], ("medium", "agents", "theta", "reward")):
        assert np.array_equal(a, b), f"{what} differs"


def test_pair_mode_is_automatic_on_a_large_field():
    """Automatic mode goes by the number of cells per environment (default 2^23; lowered to 2^22
    here to keep the test small): a 2048 x 2048 field runs in pair mode, and equals the run with the mode switched off."""
    import die_b200 as D
    from die_b200 import _lib
    lib = _lib.load()
    outs = []
    for mode in (1, 0):
        _lib.check(lib.die_set_tuning(b"pair_mode", mode))
        _lib.check(lib.die_set_tuning(b"pair_min_cells_log2", 22))
        try:
            n0 = lib.die_get_counter(b"forward_food_here")
            env = D.Env((2048, 2048), D.Dynamics(), init='device', seed=3)
            ag = D.PhysarumAgent(max_agents=env.max_agents, seed=5, **PHYS)
            obs = env._get_current_obs
            total = 0.0
            for it in range(12):
                obs, r, *_ = env.step(ag.forward(obs))
                total += r
            assert lib.die_get_counter(b"forward_food_here") - n0 == (11 if mode else 0)
            med, agents = env.get_state()
            outs.append((med, agents, ag.get_state()[0], total))
            del env, ag, obs
        finally:
            _lib.check(lib.die_set_tuning(b"pair_mode", 1))
            _lib.check(lib.die_set_tuning(b"pair_min_cells_log2", 23))
    assert all(np.array_equal(a, b) for a, b in zip(outs[0][:3], outs[1][:3])) and outs[0][3] == outs[1][3]
