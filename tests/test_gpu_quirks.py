"""Known-answer tests of the CUDA path for the parity-critical quirks Q1-Q12 of SURVEY.md section 8 and for the
edge cases of the domain (no agents, every cell occupied, every agent on ONE cell, M != H*W, the smallest field,
moves longer than the field), on hand-built states small enough to verify by inspection.  Each case states
the expected numbers AND is cross-checked against the oracle."""
import numpy as np
import pytest

from oracle import die_ref as R

pytestmark = pytest.mark.gpu
g = R.grid_coords


def _pair(h=6, w=8, agents=None, food=0.5, chem=None, **dyn):
    import die_b200 as D
    medium = np.zeros((3, h, w))
    medium[1] = food
    if chem is not None:
        medium[2] = chem
    gdyn = dict(dyn)
    if gdyn.get('boundary') == 'limit':
        gdyn['boundary'] = D.BoundaryCondition.limit
    ref = R.Env((h, w), R.Dynamics(**dyn), medium=medium.copy(), agents=agents.copy())
    gpu = D.Env((h, w), D.Dynamics(**gdyn), init_state=(medium, agents))
    return ref, gpu


def _agents(m, rows):
    a = np.zeros((4, m))
    for k, (x, y, alive, f) in enumerate(rows):
        a[:, k] = (x, y, alive, f)
    return a


def _step_both(ref, gpu, act):
    import torch
    robs, rr, rt, _, rinfo = ref.step(act.copy())
    gobs, gr, gt, gtr, ginfo = gpu.step(torch.from_numpy(act).cuda())
    med, ag = gpu.get_state()
    assert np.array_equal(med, ref.medium) and np.array_equal(ag, ref.agents)
    assert abs(rr - gr) <= 1e-12 * max(1.0, abs(rr)) and rt == gt and gtr is False
    assert rinfo['num_agents'] == ginfo['num_agents']
    return gr, ginfo, med, ag


def test_q1_q7_ghost_on_occupied_cell_eats_for_reward():
    ag = _agents(4, [(g(6)[2], g(8)[3], 1, .5), (g(6)[2], g(8)[3], 0, 0), (g(6)[4], g(8)[1], 0, 0)])
    ref, gpu = _pair(agents=ag)
    reward, info, med, agn = _step_both(ref, gpu, np.zeros((3, 4)))
    cf = 0.1 * 0.5
    assert np.isclose(reward, 2 * cf) and info['num_agents'] == 1          # both slots on the occupied cell gain
    assert np.isclose(med[1][2, 3], 0.5 - cf) and med[1][4, 1] == 0.5      # ... the cell loses it once
    assert agn[3].tolist() == [0.5 + cf, cf, 0.0, 0.0]


def test_q2_last_writer_wins_deposit():
    ag = _agents(5, [(g(6)[1], g(8)[1], 1, 0), (g(6)[1], g(8)[1], 1, 0), (g(6)[1], g(8)[1], 0, 0),
                     (g(6)[3], g(8)[5], 1, 0)])
    ref, gpu = _pair(agents=ag, diffuse_sigma=0.1, rate_decay_chem=0.0)     # radius 0: blur = identity
    act = np.zeros((3, 5))
    act[2] = [10., 20., 30., 5., 0.]
    _, _, med, _ = _step_both(ref, gpu, act)
    assert med[2][1, 1] == 20.        # highest ALIVE slot on the cell; the dead slot's 30 is ignored
    assert med[2][3, 5] == 5. and med[2].sum() == 25. and med[0].sum() == 2.


def test_q3_q11_move_wrap_and_cell_resolution():
    ag = _agents(3, [(0.0, 0.0, 1, 0), (0.999, 0.5, 1, 0), (0.5, 0.5, 1, 0)])
    ref, gpu = _pair(h=5, w=5, agents=ag)
    act = np.zeros((3, 3))
    act[0] = [-1e-18, 0.002, 0.125]   # tiny negative wraps to exactly 1.0; 1.001 wraps to ~0.001; tie at .625 goes up
    _, _, _, agn = _step_both(ref, gpu, act)
    assert agn[0, 0] == 1.0 and abs(agn[0, 1] - 0.001) < 1e-12
    assert gpu.last_cells().cpu().numpy().tolist() == [4 * 5 + 0, 0 * 5 + 2, 3 * 5 + 2]


def test_limit_boundary_clips():
    ag = _agents(2, [(0.95, 0.05, 1, 0), (0.5, 0.5, 1, 0)])
    ref, gpu = _pair(agents=ag, boundary='limit')
    act = np.zeros((3, 2))
    act[0, 0], act[1, 0] = 0.2, -0.2
    _, _, _, agn = _step_both(ref, gpu, act)
    assert agn[:2, 0].tolist() == [1.0, 0.0]


def test_moves_longer_than_the_field_take_the_generic_remainder_path():
    ag = _agents(4, [(0.25, 0.75, 1, 0), (0.5, 0.5, 1, 0), (0.1, 0.9, 0, 0), (0.0, 1.0, 1, 0)])
    ref, gpu = _pair(agents=ag)
    act = np.zeros((3, 4))
    act[0] = [3.5, -2.25, 17.125, -1.0]
    act[1] = [-7.75, 4.0, -0.5, 2.0]
    _, _, _, agn = _step_both(ref, gpu, act)
    assert np.array_equal(agn[0], (ag[0] + act[0]) % 1.) and np.array_equal(agn[1], (ag[1] + act[1]) % 1.)


def test_q4_q5_q6_sense_clamped_gradient_nonperiodic_signed_zero():
    import torch
    import die_b200 as D
    h = w = 16
    chem = np.zeros((h, w))
    chem[0, :] = np.linspace(1, 2, w)             # a ridge on the first row only
    medium = np.zeros((3, h, w))
    medium[2] = chem
    ag = _agents(2, [(0.98, 0.5, 1, 0), (0.02, 0.5, 1, 0)])
    prev = np.array([[1., -1.], [0., 0.]])        # theta = 0 and pi
    ra = R.PhysarumAgent(max_agents=2, scale=0.01, sense_offset=0.1, prev_grad=prev)
    ga = D.PhysarumAgent(max_agents=2, scale=0.01, sense_offset=0.1)
    ga.set_state(theta=ra._direction_rads.copy())
    ga.record_sense_cells = True
    coin = np.array([0, 1])
    ract = ra.forward((ag, medium), coin=coin.copy())
    gact = ga.forward((torch.from_numpy(ag).cuda(), torch.from_numpy(medium).cuda()), coin=coin).cpu().numpy()
    sx, sy = ra.last_sense_cells
    assert sx.tolist() == [15, 0]                  # clamped to the edge rows, NOT wrapped to the ridge / far side
    assert ga.sense_cells.cpu().numpy()[0].tolist() == (sx * w + sy).tolist()
    assert np.array_equal(gact[2], ract[2]) and np.allclose(gact[:2], ract[:2], rtol=0, atol=1e-16)
    # slot 0 senses the last row: no ridge there (non-periodic gradient) -> undetermined -> coin 0 -> turn -30 deg
    assert np.isclose(np.arctan2(gact[1, 0], gact[0, 0]), -np.radians(30))


def test_q8_physarum_ghosts_move_and_brownian_ghosts_do_not():
    import torch
    import die_b200 as D
    h = w = 8
    medium = np.zeros((3, h, w))
    medium[1] = 0.3
    ag = _agents(4, [(0.5, 0.5, 1, 0.2)])
    obs = (torch.from_numpy(ag).cuda(), torch.from_numpy(medium).cuda())
    phys = D.PhysarumAgent(max_agents=4, scale=0.01)
    phys.set_state(theta=np.full(4, np.pi / 4))
    act = phys.forward(obs, coin=np.array([0, 1, 0, 1])).cpu().numpy()
    assert (np.hypot(act[0], act[1]) > 0.009).all()           # every slot, alive or not, moves
    assert np.allclose(act[2], 4.0 * 0.3 * 0.1)                # zero gradient: undetermined -> 0.1 deposit mask
    actb = D.BrownianAgent(0.01).forward(obs, u=np.full((3, 4), 0.7)).cpu().numpy()
    assert (actb[:, 1:] == 0).all() and (actb[:, 0] != 0).all()


def test_q9_brownian_quantisation():
    import torch
    import die_b200 as D
    ag = _agents(3, [(0, 0, 1, 0), (0, 0, 1, 0), (0, 0, 1, 0)])
    u = np.array([[0.12345, 0.9996, 0.0004]] * 3)
    obs = (torch.from_numpy(ag).cuda(), torch.zeros((3, 4, 4), dtype=torch.float64, device='cuda'))
    act = D.BrownianAgent(0.01, 0.5).forward(obs, u=u).cpu().numpy()
    assert np.array_equal(act[0], 0.02 * np.array([0.123, 1.0, 0.0]) - 0.01)
    assert np.array_equal(act[2], 0.5 * np.array([0.123, 1.0, 0.0]))


def test_q12_info_rounding_and_termination():
    ag = _agents(2, [(0.5, 0.5, 1, 0)])
    ref, gpu = _pair(agents=ag, food=0.123456789)
    reward, info, _, _ = _step_both(ref, gpu, np.zeros((3, 2)))
    assert info['reward'] == np.round(reward, 3) and info['reward'] != reward
    assert info['mean_reward'] == np.round(reward / 1, 5)


def test_no_alive_agents_terminates():
    import torch
    ag = _agents(6, [])                                         # six ghosts at (0, 0)
    ref, gpu = _pair(agents=ag)
    act = np.zeros((3, 6))
    act[0, :3] = 0.3
    robs, rr, rterm, _, rinfo = ref.step(act.copy())
    gobs, gr, gterm, _, ginfo = gpu.step(torch.from_numpy(act).cuda())
    assert bool(rterm) and gterm is True and ginfo['num_agents'] == 0 and ginfo['mean_reward'] == 0
    med, agn = gpu.get_state()
    assert np.array_equal(med, ref.medium) and np.array_equal(agn, ref.agents)
    assert med[0].sum() == 0 and gr == rr                       # nothing eaten; only the ghosts' action cost


def test_every_cell_occupied_and_everybody_on_one_cell():
    """ratio 1: all H*W slots alive (the alive bitmask is all ones); then a limit boundary and a huge move pile all of
    them onto the corner cell: one winner (the highest slot), binary occupancy, everybody eats the same cell."""
    h, w = 7, 9
    m = h * w
    xs, ys = np.meshgrid(g(h), g(w), indexing='ij')
    ag = np.stack([xs.ravel(), ys.ravel(), np.ones(m), np.full(m, 0.25)])
    ref, gpu = _pair(h=h, w=w, agents=ag, boundary='limit', diffuse_sigma=0.1, rate_decay_chem=0.0)
    act = np.zeros((3, m))
    act[2] = np.arange(m) + 1.0
    _, info, med, _ = _step_both(ref, gpu, act)
    assert info['num_agents'] == m and med[0].sum() == m and np.array_equal(med[2].ravel(), np.arange(m) + 1.0)
    act[0], act[1] = 5.0, 5.0                                  # clip(0, 1): everybody to (1, 1)
    reward, info, med, agn = _step_both(ref, gpu, act)
    assert med[0].sum() == 1 and med[0][h - 1, w - 1] == 1
    assert med[2][h - 1, w - 1] == 2.0 * m                      # its own earlier deposit m + the winner's (slot m-1) m
    assert (agn[0] == 1).all() and (agn[1] == 1).all()


@pytest.mark.parametrize("m", [5, 200])
def test_fewer_or_more_slots_than_cells(m):
    """max_agents != H*W (agents_from_medium(max_agents=...), core/data_init.py:143-144)."""
    import die_b200 as D
    h, w = 8, 10
    rng = np.random.default_rng(m)
    n_alive = min(m, 4)
    ag = np.zeros((4, m))
    ag[0, :n_alive], ag[1, :n_alive] = rng.random(n_alive), rng.random(n_alive)
    ag[2, :n_alive], ag[3, :n_alive] = 1, 0.3
    ref, gpu = _pair(h=h, w=w, agents=ag)
    ra, ga = R.BrownianAgent(0.05), D.BrownianAgent(0.05)
    for it in range(5):
        u = rng.random((3, m))
        ract = ra.forward(ref._get_current_obs, u=u)
        gact = ga.forward(gpu._get_current_obs, u=u)
        assert np.array_equal(ract, gact.cpu().numpy())
        _step_both(ref, gpu, ract)
    kw = dict(scale=0.03, turn_angle=30, sense_offset=0.1)
    rp = R.PhysarumAgent(max_agents=m, prev_grad=np.ones((2, m)), **kw)
    gp = D.PhysarumAgent(max_agents=m, **kw)
    gp.set_state(theta=rp._direction_rads.copy())
    for it in range(5):
        coin = rng.integers(0, 2, m)
        rp._direction_rads = gp.get_state()[0].copy()
        ract = rp.forward(ref._get_current_obs, coin=coin.copy())
        gact = gp.forward(gpu._get_current_obs, coin=coin).cpu().numpy()
        assert np.array_equal(ract[2], gact[2]) and np.allclose(ract[:2], gact[:2], rtol=0, atol=1e-15)
        _step_both(ref, gpu, gact)


def test_smallest_field():
    import die_b200 as D
    ag = _agents(4, [(0.0, 1.0, 1, 0.5), (1.0, 0.0, 1, 0.5)])
    ref, gpu = _pair(h=2, w=2, agents=ag, food=np.array([[0.1, 0.2], [0.3, 0.4]]))
    rng = np.random.default_rng(0)
    ra, ga = R.BrownianAgent(0.4, 1.0), D.BrownianAgent(0.4, 1.0)
    for it in range(10):
        u = rng.random((3, 4))
        ract = ra.forward(ref._get_current_obs, u=u)
        assert np.array_equal(ract, ga.forward(gpu._get_current_obs, u=u).cpu().numpy())
        _step_both(ref, gpu, ract)
    with pytest.raises(Exception):
        D.Env((1, 5), D.Dynamics(), init_state=(np.zeros((3, 1, 5)), np.zeros((4, 5))))      # linspace(0, 1, 1) has no step
