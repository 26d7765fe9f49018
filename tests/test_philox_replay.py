"""The in-kernel random draws replayed on the host (die_b200/philox.py), and with them the BENCHMARKED forward kernel
-- LEAN + float32 gradient cache + in-kernel Philox coins + alive bitmask + PLAIN field pass -- compared with the
oracle directly, not through the general kernel (round 1's verdict: "parity green, but by a two-hop argument").

CPU part: the kernel sources under the emulator (tests/hostsim).  The GPU part is tests/test_gpu_philox_replay.py."""
import numpy as np
import pytest

from oracle import die_ref as R
from tests._parity import assert_state_equal, lattice_theta, ref_cells_linear

S = pytest.importorskip("tests.hostsim.sim")
from die_b200 import philox as P                       # noqa: E402
from tests.test_hostsim_kernels import make_pair       # noqa: E402

PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)


@pytest.fixture
def portable_math():
    R.set_math_backend('portable')
    yield
    R.set_math_backend('numpy')


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (counter, key -> output)."""
    out = P.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(v[0]) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = P.philox4x32_10([0xffffffff], [0xffffffff], [0xffffffff], [0xffffffff], 0xffffffff, 0xffffffff)
    assert [int(v[0]) for v in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = P.philox4x32_10([0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344], 0xa4093822, 0x299f31d0)
    assert [int(v[0]) for v in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_brownian_uniforms_replica_matches_the_kernel():
    """brownian_forward_kernel with in-kernel draws == the same kernel fed the host replica's uniforms."""
    (ref,), env = make_pair((24, 40), seed=1)
    for step in (0, 7):
        a = S.brownian_forward(env.agents, move_scale=0.02, seed=11, step=step)
        u = P.brownian_uniforms(11, step, 1, env.M)
        b = S.brownian_forward(env.agents, move_scale=0.02, u=u, seed=0, step=0)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("field,batch", [((40, 72), None), ((24, 64), 3), ((70, 90), None), ((48, 50), 7), ((24, 64), 21)])
@pytest.mark.parametrize("variant", ["lean", "lean+pair"])
def test_benchmarked_forward_kernel_against_the_oracle(portable_math, field, batch, variant):
    """Free run with in-kernel coins (the LEAN float32-gradient forward from the second step on), the oracle fed the
    replica's coins: every step bit-exact (actions, headings, cells, fields).  Slot counts beyond one CTA's 2048 and
    batches check the (environment, CTA, thread, item) -> coin map of the replica.  variant "lean+pair": pair mode forced,
    the food under the agent handed over per slot by the feed kernel -- the same bits."""
    S.set_tuning("pair_mode", 2 if variant == "lean+pair" else 1)
    try:
        _benchmarked_forward_against_the_oracle(field, batch, variant)
    finally:
        S.set_tuning("pair_mode", 1)


def _benchmarked_forward_against_the_oracle(field, batch, variant):
    refs, env = make_pair(field, seed=7, batch=batch)
    B, m = env.B, env.M
    seed = 21
    ga = S.SimGradientAgent(m, B=B, seed=seed, **PHYS)
    ras = []
    for b in range(B):
        theta0, prev = lattice_theta(m, 30, 7 + b)
        ga.theta[b] = theta0
        ras.append(R.PhysarumAgent(max_agents=m, prev_grad=prev, **PHYS))
    lean0 = S.lib().die_get_counter(b"forward_lean_f32")
    fh0 = S.lib().die_get_counter(b"forward_food_here")
    iters = 12
    for it in range(iters):
        coin = P.physarum_coins(seed, it, B, m)
        gact = ga.forward(env)
        for b in range(B):
            ract = ras[b].forward(refs[b]._get_current_obs, coin=coin[b].astype(np.int64))
            assert np.array_equal(ga.theta[b], ras[b]._direction_rads), f"theta differs at step {it}"
            assert np.array_equal(gact[b], ract), f"action differs at step {it}"
            refs[b].step(ract)
        env.step(gact)
        for b in range(B):
            assert np.array_equal(ref_cells_linear(refs[b]), env.cells()[b])
            assert_state_equal(refs[b], env.medium[b], env.agents[b], float_exact=True)
    assert S.lib().die_get_counter(b"forward_lean_f32") == lean0 + iters - 1, "the benchmarked instantiation must be the one compared"
    assert S.lib().die_get_counter(b"forward_food_here") == fh0 + ((iters - 1) if variant == "lean+pair" else 0)


def test_call_counter_in_device_memory_draws_the_same_numbers():
    """DIE_FWD_STEP_ON_DEVICE / die_brownian_forward_dev (die_b200/graph.py: a CUDA-graph replay cannot change a kernel
    argument, so the call counter is read from memory): identical draws to the counter passed by value."""
    import ctypes as C
    from die_b200 import _lib as L
    lib = S.lib()
    (ref,), env = make_pair((24, 64), seed=2)
    counter = S.fenced((1,), np.uint64, fill=0)
    for step in (0, 5):
        counter[0] = step
        a = S.brownian_forward(env.agents, move_scale=0.02, seed=11, step=step)
        agents = S.fenced_copy(env.agents)
        action = S.fenced((1, 3, env.M), fill=np.nan)
        S.check(lib.die_brownian_forward_dev(S.ptr(agents), S.ptr(action), env.M, 1, 0.02, 0.5, 11, S.ptr(counter), None))
        assert np.array_equal(a, action)
    outs = []
    for on_device in (False, True):
        (ref,), env = make_pair((24, 64), seed=2)
        ga = S.SimGradientAgent(env.M, seed=3, **PHYS)
        ga.theta[0] = lattice_theta(env.M, 30, 2)[0]
        for it in range(4):
            env.want_gradient()
            flags = (L.FWD_USE_CELLS | L.FWD_USE_GRADIENT) if it > 0 else 0
            counter[0] = it
            step_arg = S.ptr(counter) if on_device else it
            S.check(lib.die_env_forward_gradient(env.handle, C.byref(ga.p), S.ptr(env.agents), S.ptr(env.medium), S.ptr(ga.theta),
                                                 None, S.ptr(ga.action), None, None, None,
                                                 flags | (L.FWD_STEP_ON_DEVICE if on_device else 0), 3, step_arg, None))
            env.step(ga.action)
        outs.append((env.medium.copy(), env.agents.copy(), ga.theta.copy()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
