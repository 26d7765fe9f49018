"""The float32 FIELD mode (die_env_set_field_dtype): medium and consumed_field in float32, agents /
actions / headings in float64.  Values are widened on load, every operation is the float64 one, results are rounded once
on store.  CPU part: the kernel sources under the emulator.  What is asserted, per step, against the float64 oracle
started from the SAME (float32-representable) state:
  * integers bit-exact: cells, occupancy, alive, num_agents; positions and actions bit-exact too (they are float64
    functions of float64 inputs and of field values both sides read identically);
  * fields within float32 rounding of the oracle's: |chem, food - oracle| <= 6e-8 relative (2^-24); agent_food and the
    reward within 1e-6 relative (consumed_field is rounded once)."""
import numpy as np
import pytest

from oracle import die_ref as R
from tests._parity import lattice_theta, ref_cells_linear

S = pytest.importorskip("tests.hostsim.sim")
import die_b200 as D                                    # noqa: E402
from die_b200 import _lib as L                          # noqa: E402

PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)
EPS32 = 2.0 ** -24
TINY32 = 2.0 ** -149          # the spacing of float32 subnormals: diffusion spreads chem1 down to 1e-40 and below


@pytest.fixture
def portable_math():
    R.set_math_backend('portable')
    yield
    R.set_math_backend('numpy')


def _shadow_step_checks(ref, env, it):
    med32 = env.medium[0]
    assert med32.dtype == np.float32
    assert np.array_equal(med32[0].astype(np.float64), ref.medium[0]), f"occupancy differs at step {it}"
    for ch, name in ((1, "food"), (2, "chem")):
        err = np.abs(med32[ch].astype(np.float64) - ref.medium[ch])
        assert (err <= EPS32 * np.abs(ref.medium[ch]) + TINY32).all(), f"{name} beyond float32 rounding at step {it}"
    assert np.array_equal(env.agents[0, :3], ref.agents[:3]), f"positions / alive differ at step {it}"
    np.testing.assert_allclose(env.agents[0, 3], ref.agents[3], rtol=1e-6, atol=1e-7)   # (one rounding of consumed_field, <= 6e-8 absolute, also where the stock crosses 0)
    assert np.array_equal(ref_cells_linear(ref), env.cells()[0])


@pytest.mark.parametrize("field,sigma", [((40, 72), 0.5), ((37, 53), 0.8), ((24, 64), 0.3), ((20, 20), 0.1)])
def test_float32_fields_physarum_shadowed_per_step(portable_math, field, sigma):
    """Every step the oracle restarts from the emulated kernels' own state (float32 fields widened), so one step's
    rounding is what is compared; the run itself is the float32 one (turn decisions read the float32 gradient cache)."""
    np.random.seed(3)
    ref = R.Env(field, R.Dynamics(init_agent_ratio=0.1, diffuse_sigma=sigma), noise_seed=3)
    env = S.SimEnv(field, ref.medium[None].astype(np.float32), ref.agents[None],
                   D.Dynamics(init_agent_ratio=0.1, diffuse_sigma=sigma), field_dtype=np.float32)
    m = env.M
    theta0, prev = lattice_theta(m, 30, 3)
    ra = R.PhysarumAgent(max_agents=m, prev_grad=prev, **PHYS)
    ga = S.SimGradientAgent(m, **PHYS)
    ga.theta[0] = theta0
    rng = np.random.default_rng(1)
    for it in range(15):
        ref.medium[...] = env.medium[0].astype(np.float64)
        ref.agents[...] = env.agents[0]
        ra._direction_rads = ga.theta[0].copy()
        coin = rng.integers(0, 2, m)
        ract = ra.forward(ref._get_current_obs, coin=coin.copy())
        gact = ga.forward(env, coin=coin)[0]
        assert np.array_equal(gact, ract), f"action differs at step {it}"
        assert np.array_equal(ga.theta[0], ra._direction_rads)
        _, rr, _, _, rinfo = ref.step(ract)
        r, alive = env.step(gact)
        assert rinfo['num_agents'] == alive[0]
        assert abs(rr - r[0]) <= 1e-6 * max(1.0, abs(rr))
        _shadow_step_checks(ref, env, it)
    assert S.lib().die_env_field_dtype(env.handle) == L.FIELD_F32


def test_float32_fields_brownian_free_run_bound():
    """Free run, nothing re-synchronised: Brownian actions do not depend on the fields, so positions, cells and occupancy
    stay bit-exact for the whole run and the fields drift only by accumulated rounding: <= 1e-5 relative after 60 steps."""
    field = (48, 64)
    np.random.seed(5)
    ref = R.Env(field, R.Dynamics(init_agent_ratio=0.1), noise_seed=5)
    env = S.SimEnv(field, ref.medium[None].astype(np.float32), ref.agents[None], D.Dynamics(init_agent_ratio=0.1),
                   field_dtype=np.float32)
    ref.medium[...] = env.medium[0].astype(np.float64)
    ra = R.BrownianAgent(0.01)
    m = env.M
    rng = np.random.default_rng(7)
    for it in range(60):
        u = rng.random((3, m))
        ract = ra.forward(ref._get_current_obs, u=u)
        gact = S.brownian_forward(env.agents[0], u=u)
        assert np.array_equal(ract, gact)
        ref.step(ract)
        env.step(gact)
        assert np.array_equal(env.medium[0][0].astype(np.float64), ref.medium[0])
        assert np.array_equal(env.agents[0, :3], ref.agents[:3])
    for ch in (1, 2):
        np.testing.assert_allclose(env.medium[0][ch].astype(np.float64), ref.medium[ch], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(env.agents[0, 3], ref.agents[3], rtol=1e-5, atol=1e-6)


def test_float32_mode_refuses_what_it_does_not_implement():
    (h, w) = (24, 32)
    np.random.seed(1)
    ref = R.Env((h, w), R.Dynamics(init_agent_ratio=0.1), noise_seed=1)
    env = S.SimEnv((h, w), ref.medium[None].astype(np.float32), ref.agents[None], D.Dynamics(init_agent_ratio=0.1),
                   field_dtype=np.float32)
    with pytest.raises(RuntimeError, match="float64 fields only"):
        env.step_host(np.zeros((1, 3, env.M)))
