"""JonesAgent on the GPU -- the classic three-sensor Physarum particle (optional: the reference has no
such class, so the specification is oracle/die_ref.py:JonesAgent, built from the reference's own pieces: core/utils.py:39-54
nearest-cell lookups, :154-164 polar2xy, :177-179 renormalize_radians, core/agent/gradient.py:113-124 the unmasked action).
Bar: bit-exact with the oracle in its portable math backend (die_math.h's sin / cos on both sides), free running.
"""
import numpy as np
import pytest

from oracle import die_ref as R
from tests._parity import make_pair, lattice_theta, assert_state_equal, ref_cells_linear

pytestmark = pytest.mark.gpu


@pytest.fixture
def portable_math():
    R.set_math_backend('portable')
    yield
    R.set_math_backend('numpy')


def _theta0(m, turn_angle, seed):
    tr = np.radians(turn_angle)
    return (lattice_theta(m, 30, seed)[0] // tr) * tr


@pytest.mark.parametrize("field,iters,kw", [
    ((256, 256), 120, dict(scale=0.007, sense_offset=0.04)),
    ((37, 53), 60, dict(scale=0.02, sense_offset=0.09, turn_angle=22.5, sense_angle=45)),
    ((96, 64), 60, dict(scale=0.015, sense_offset=0.05, turn_angle=60, sense_angle=30, deposit=2.0))])
def test_jones_free_run_bit_exact(portable_math, field, iters, kw):
    import die_b200 as D
    (ref,), gpu = make_pair(field, seed=4, ratio=0.2)
    m = ref.agents.shape[-1]
    th0 = _theta0(m, kw.get('turn_angle', 45), 4)
    ra = R.JonesAgent(max_agents=m, theta0=th0, **kw)
    ga = D.JonesAgent(max_agents=m, **kw)
    ga.set_state(theta=th0)
    rng = np.random.default_rng(8)
    robs, gobs = ref._get_current_obs, gpu._get_current_obs
    for it in range(iters):
        coin = rng.integers(0, 2, m)
        ract = ra.forward(robs, coin=coin.copy())
        gact = ga.forward(gobs, coin=coin)
        assert np.array_equal(ga.get_state()[0], ra._direction_rads), f"heading differs at step {it}"
        assert np.array_equal(ract, gact.cpu().numpy()), f"action differs at step {it}"
        robs, rr, rterm, _, rinfo = ref.step(ract)
        gobs, gr, gterm, _, ginfo = gpu.step(gact)
        assert np.array_equal(ref_cells_linear(ref), gpu.last_cells().cpu().numpy()), f"cells differ at step {it}"
        assert rinfo['num_agents'] == ginfo['num_agents'] and abs(rr - gr) <= 1e-11 * max(abs(rr), 1.0)
        if it % 20 == 0 or it == iters - 1:
            med, ag = gpu.get_state()
            assert_state_equal(ref, med, ag, float_exact=True)
    # the trail field has structure by now: the particles follow each other's deposits
    assert np.std(gpu.get_state()[0][2]) > 0


def test_jones_batched_philox_graph_and_float32_fields():
    """In-kernel coins: the CUDA-graph replay (device-resident call counter) equals the eager loop; a batch equals its
    environments one by one is NOT expected (coins are keyed on the env index) but is reproducible; float32 field mode runs."""
    import torch
    import die_b200 as D
    kw = dict(scale=0.01, sense_offset=0.05)
    finals = []
    for mode in ("eager", "graph"):
        _, env = make_pair((64, 96), seed=12, ratio=0.2, batch=3)
        m = env.max_agents
        ag = D.JonesAgent(max_agents=m, seed=5, **kw)
        ag.set_state(theta=np.stack([_theta0(m, 45, 20 + b) for b in range(3)]))
        if mode == "eager":
            obs = env._get_current_obs
            for _ in range(22):
                obs, _, _ = env.step_async(ag.forward(obs))
        else:
            loop = D.GraphedLoop(env, ag, warmup=2)
            loop.run(20)
        torch.cuda.synchronize()
        finals.append((env.get_state(), ag.get_state()[0]))
    (ma, aa), ta = finals[0]
    (mb, ab), tb = finals[1]
    assert np.array_equal(ma, mb) and np.array_equal(aa, ab) and np.array_equal(ta, tb)
    assert not np.array_equal(ta[0], ta[1])
    # float32 field mode: the agent reads the float32 medium; positions / cells stay float64-exact for a given action
    _, env32 = make_pair((64, 96), seed=12, ratio=0.2)
    env32 = D.Env((64, 96), D.Dynamics(init_agent_ratio=0.2), init_state=env32.get_state(), field_dtype=torch.float32)
    ag = D.JonesAgent(max_agents=env32.max_agents, seed=5, **kw)
    obs = env32._get_current_obs
    for _ in range(10):
        obs, r, *_ = env32.step(ag.forward(obs))
    assert np.isfinite(r) and obs[1].dtype == torch.float32


def test_jones_host_buffers_and_save_load(tmp_path):
    import die_b200 as D
    _, env = make_pair((48, 40), seed=3, ratio=0.2)
    m = env.max_agents
    ag = D.JonesAgent(max_agents=m, seed=2, scale=0.01, sense_offset=0.05)
    med, agents = env.get_state()
    act = ag.forward((agents, med))                       # numpy in -> numpy out
    assert isinstance(act, np.ndarray) and act.shape == (3, m)
    assert np.allclose(np.hypot(act[0], act[1]), 0.01, rtol=1e-12)
    ag.save(tmp_path / "jones.json")
    ag2 = D.JonesAgent.load(tmp_path / "jones.json")
    assert ag2.init_params() == ag.init_params()
