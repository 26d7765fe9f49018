"""Dynamics.apply_sense_mask (core/env.py:275-294): the observation's medium is the medium zeroed outside
ceil(round(gaussian(occupancy, sigma=2), 3)).  GPU (die_sense_mask) vs the oracle: the mask is integer-valued
and must be exact, the masked observation bit-identical, and agents acting on it must take the same actions."""
import numpy as np
import pytest

from oracle import die_ref as R
from tests._parity import make_pair, lattice_theta, assert_state_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field,ratio", [((64, 80), 0.01), ((37, 53), 0.03), ((256, 256), 0.002), ((12, 9), 0.05)])
def test_sensed_medium_equals_the_oracle(field, ratio):
    import die_b200 as D
    (ref,), gpu = make_pair(field, seed=14, ratio=ratio, dynamics_kw=dict(apply_sense_mask=True))
    m = ref.agents.shape[-1]
    ra, ga = R.BrownianAgent(0.02, 1.0), D.BrownianAgent(move_scale=0.02, deposit_scale=1.0)
    rng = np.random.default_rng(2)
    robs, gobs = ref._get_current_obs, gpu._get_current_obs
    assert np.array_equal(robs[1], gobs[1].cpu().numpy())
    masked_any = False
    for it in range(12):
        u = rng.random((3, m))
        robs, rr, *_ = ref.step(ra.forward(robs, u=u))
        gobs, gr, *_ = gpu.step(ga.forward(gobs, u=u))
        assert_state_equal(ref, *gpu.get_state(), float_exact=True)          # env.medium itself stays unmasked
        obs_med = gobs[1].cpu().numpy()
        assert np.array_equal(robs[1], obs_med), it
        assert np.array_equal(np.signbit(robs[1]), np.signbit(obs_med))
        masked_any |= bool((obs_med[1] != gpu.get_state()[0][1]).any())
    if field[0] * field[1] > 1000:          # (on the 12 x 9 field the sigma-2 neighbourhoods cover everything)
        assert masked_any, "the mask must actually hide something in this configuration"


def test_physarum_acts_on_the_masked_observation():
    import die_b200 as D
    field = (72, 64)
    R.set_math_backend('portable')
    try:
        (ref,), gpu = make_pair(field, seed=15, ratio=0.01, dynamics_kw=dict(apply_sense_mask=True))
        m = ref.agents.shape[-1]
        kw = dict(scale=0.02, turn_angle=30, sense_offset=0.05)
        theta0, prev = lattice_theta(m, 30, 15)
        ra, ga = R.PhysarumAgent(max_agents=m, prev_grad=prev, **kw), D.PhysarumAgent(max_agents=m, **kw)
        ga.set_state(theta=theta0)
        rng = np.random.default_rng(3)
        robs, gobs = ref._get_current_obs, gpu._get_current_obs
        for it in range(25):
            coin = rng.integers(0, 2, m)
            ract = ra.forward(robs, coin=coin.copy())
            gact = ga.forward(gobs, coin=coin)
            assert np.array_equal(ract, gact.cpu().numpy()), it
            robs, *_ = ref.step(ract)
            gobs, *_ = gpu.step(gact)
            assert np.array_equal(robs[1], gobs[1].cpu().numpy())
            assert_state_equal(ref, *gpu.get_state(), float_exact=True)
        assert ga.last_hints == (False, False)      # the env's caches describe the unmasked medium: not applicable
    finally:
        R.set_math_backend('numpy')


def test_sense_mask_host_path_and_batch():
    import die_b200 as D
    field = (40, 48)
    (r0, r1), gpu = make_pair(field, seed=16, ratio=0.01, batch=2, dynamics_kw=dict(apply_sense_mask=True))
    m = r0.agents.shape[-1]
    ga = D.ConstAgent((0.01, 0.004), 0.5)
    ras = [R.ConstAgent((0.01, 0.004), 0.5) for _ in range(2)]
    hobs = tuple(t.cpu().numpy() for t in gpu._get_current_obs)
    for it in range(6):
        hact = ga.forward(hobs)
        hobs, *_ = gpu.step(hact)
        for b, (ref, ra) in enumerate(zip((r0, r1), ras)):
            robs, *_ = ref.step(ra.forward(ref._get_current_obs))
            assert np.array_equal(robs[1], hobs[1][b]), (it, b)
            assert np.array_equal(ref.agents, hobs[0][b])
