"""The BENCHMARKED kernels against the oracle, directly, on BASELINE.json's own configurations.

`bench.py` times PhysarumAgent(rng='philox').forward -- the LEAN instantiation of gradient_forward_kernel on the float32
gradient cache, with in-kernel Philox coins -- and Env.step with the alive bitmask and the PLAIN field pass.  The other
parity tests inject coins, which selects the general forward kernel.  Here the in-kernel coins are replayed on the host
(die_b200/philox.py) and handed to the oracle (core/agent/gradient.py:96-124, 168-208 restated in oracle/die_ref.py, with
the 'portable' math backend = die_math.h compiled for the host), so the two free runs must agree bit for bit:
  configs[1]  256 x 256, 300 iterations (README parameters)
  configs[2]  4096 x 4096, three iterations (the oracle needs ~20 s per iteration there)
  configs[3]  a batch of 64 environments (and 8 of 256 x 256)
The launch counter proves the LEAN float32 kernel is the one that ran."""
import numpy as np
import pytest

from tests._parity import assert_state_equal, lattice_theta, make_pair, ref_cells_linear

pytestmark = pytest.mark.gpu
PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)


def _replay(field, iters, batch=None, seed=3, agent_seed=1234, check_every=1):
    import die_b200 as D
    from die_b200 import _lib, philox as P
    from oracle import die_ref as R
    lib = _lib.load()
    R.set_math_backend('portable')
    try:
        refs, env = make_pair(field, seed=seed, batch=batch)
        B, m = env.batch, env.max_agents
        ga = D.PhysarumAgent(max_agents=m, rng='philox', seed=agent_seed, **PHYS)
        ras, thetas = [], []
        for b in range(B):
            theta0, prev = lattice_theta(m, 30, seed + b)
            thetas.append(theta0)
            ras.append(R.PhysarumAgent(max_agents=m, prev_grad=prev, **PHYS))
        ga.set_state(theta=np.stack(thetas))
        lean0 = lib.die_get_counter(b"forward_lean_f32")
        fused0 = lib.die_get_counter(b"step_fused")
        gobs = env._get_current_obs
        for it in range(iters):
            coin = P.physarum_coins(agent_seed, it, B, m)
            gact = ga.forward(gobs)
            for b in range(B):
                ract = ras[b].forward(refs[b]._get_current_obs, coin=coin[b].astype(np.int64))
                refs[b].step(ract)
            gobs, gr, _, _, ginfo = env.step(gact)
            if it % check_every == 0 or it == iters - 1:
                act = gact.cpu().numpy().reshape(B, 3, m)
                theta = ga.get_state()[0].reshape(B, m)
                med, ag = env.get_state()
                med, ag = med.reshape(B, 3, *field), ag.reshape(B, 4, m)
                cells = env.last_cells().cpu().numpy().reshape(B, m)
                for b in range(B):
                    assert np.array_equal(theta[b], ras[b]._direction_rads), f"theta differs at step {it}, env {b}"
                    assert np.array_equal(cells[b], ref_cells_linear(refs[b])), f"cells differ at step {it}, env {b}"
                    assert_state_equal(refs[b], med[b], ag[b], float_exact=True)
        assert lib.die_get_counter(b"forward_lean_f32") - lean0 == iters - 1, \
            "the benchmarked forward instantiation (LEAN, float32 gradient cache) must be the one compared"
        assert lib.die_get_counter(b"step_fused") == fused0, "the default step is the three-kernel path"
    finally:
        R.set_math_backend('numpy')


def test_config1_256x256_300_iterations():
    _replay((256, 256), 300, check_every=10)


def test_config2_4096x4096_three_iterations():
    _replay((4096, 4096), 3)


@pytest.mark.parametrize("field,batch,iters", [((64, 64), 64, 25), ((256, 256), 8, 6)])
def test_config3_batches(field, batch, iters):
    _replay(field, iters, batch=batch)
