"""Exhaustive-ish check of the kernels' nearest-cell rule against pandas: every grid
coordinate, every midpoint, their float neighbours, random points, out-of-range points -- through
move_claim (positions) on the GPU."""
import numpy as np
import pytest

from oracle import die_ref as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(256, 256), (5, 7), (4096, 33), (1000, 3)])
def test_cells_match_pandas_on_adversarial_coordinates(shape):
    import die_b200 as D
    h, w = shape
    rng = np.random.default_rng(h * 131 + w)

    def adversarial(n):
        g = R.grid_coords(n)
        mids = (g[:-1] + g[1:]) / 2
        pts = [g, mids]
        for base in (g, mids):
            up, dn = base.copy(), base.copy()
            for _ in range(3):
                up, dn = np.nextafter(up, 2), np.nextafter(dn, -2)
                pts += [up.copy(), dn.copy()]
        pts.append(rng.uniform(0, 1, 20000))
        return np.clip(np.concatenate(pts), 0.0, 1.0)

    xs, ys = adversarial(h), adversarial(w)
    m = max(len(xs), len(ys))
    xs = np.resize(xs, m)
    ys = np.resize(rng.permutation(ys), m)
    agents = np.zeros((4, m))
    agents[0], agents[1] = xs, ys
    agents[2, ::3] = 1.0
    medium = np.zeros((3, h, w))
    env = D.Env(shape, D.Dynamics(boundary=D.BoundaryCondition.limit), init_state=(medium, agents))
    import torch
    action = torch.zeros((3, m), dtype=torch.float64, device='cuda')
    env.step(action)                      # 'limit' boundary + zero action: positions unchanged
    got = env.last_cells().cpu().numpy()
    ref = R.nearest_index_pandas(xs, h) * w + R.nearest_index_pandas(ys, w)
    assert np.array_equal(got, ref.astype(np.int32))


def test_sense_cells_clamp_out_of_range():
    """pos + offset outside [0, 1] clamps to the edge cells (Q4), through the forward kernel."""
    import die_b200 as D
    h, w = 64, 48
    rng = np.random.default_rng(0)
    m = 4096
    agents = np.zeros((4, m))
    agents[0], agents[1] = rng.uniform(0, 1, m), rng.uniform(0, 1, m)
    agents[0, :64] = np.linspace(0, 0.02, 64)
    agents[1, 64:128] = np.linspace(0.98, 1.0, 64)
    medium = np.zeros((3, h, w))
    env = D.Env((h, w), init_state=(medium, agents))
    theta = rng.uniform(-np.pi, np.pi, m)
    ag = D.PhysarumAgent(max_agents=m, sense_offset=0.3, scale=0.01)
    ag.set_state(theta=theta)
    ag.record_sense_cells = True
    ag.forward(env._get_current_obs, coin=np.zeros(m, dtype=np.uint8))
    from oracle import portable_math as P
    s, c = P.sincos(theta)
    px, py = agents[0] + 0.3 * c, agents[1] + 0.3 * s
    ref = R.nearest_index_pandas(px, h) * w + R.nearest_index_pandas(py, w)
    assert np.array_equal(ag.sense_cells.cpu().numpy()[0], ref.astype(np.int32))
    assert (px < 0).any() and (px > 1).any() and (py < 0).any() and (py > 1).any()
