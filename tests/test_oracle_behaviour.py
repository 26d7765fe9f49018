"""Known-answer tests of the oracle for the parity-critical quirks Q1-Q12 of SURVEY.md section 8,
on hand-built states small enough to verify by inspection."""
import numpy as np

from oracle import die_ref as R


def _env(h=6, w=8, agents=None, food=0.5, **dyn):
    medium = np.zeros((3, h, w))
    medium[1] = food
    return R.Env((h, w), R.Dynamics(**dyn), medium=medium, agents=agents)


def _agents(m, rows):
    a = np.zeros((4, m))
    for k, (x, y, alive, f) in enumerate(rows):
        a[:, k] = (x, y, alive, f)
    return a


def test_q1_q7_ghost_on_occupied_cell_eats_for_reward():
    g = R.grid_coords
    # slot 0 alive at cell (2,3); slot 1 a ghost (alive=0) standing on the same cell; slot 2 ghost elsewhere
    ag = _agents(4, [(g(6)[2], g(8)[3], 1, .5), (g(6)[2], g(8)[3], 0, 0), (g(6)[4], g(8)[1], 0, 0)])
    env = _env(agents=ag)
    _, reward, _, _, info = env.step(np.zeros((3, 4)))
    cf = 0.1 * 0.5
    assert np.isclose(reward, 2 * cf)                      # both slots on the occupied cell gain
    assert info['num_agents'] == 1
    assert np.isclose(env.medium[1][2, 3], 0.5 - cf)       # ... the cell loses it once
    assert env.medium[1][4, 1] == 0.5
    assert env.agents[3].tolist() == [0.5 + cf, cf, 0.0, 0.0]


def test_q2_last_writer_wins_deposit():
    g = R.grid_coords
    ag = _agents(5, [(g(6)[1], g(8)[1], 1, 0), (g(6)[1], g(8)[1], 1, 0), (g(6)[1], g(8)[1], 0, 0),
                     (g(6)[3], g(8)[5], 1, 0)])
    env = _env(agents=ag, diffuse_sigma=0.1, rate_decay_chem=0.0)        # radius 0: blur = identity
    act = np.zeros((3, 5))
    act[2] = [10., 20., 30., 5., 0.]
    env.step(act)
    assert env.medium[2][1, 1] == 20.       # highest ALIVE slot on the cell; the dead slot's 30 is ignored
    assert env.medium[2][3, 5] == 5.
    assert env.medium[2].sum() == 25.
    assert env.medium[0].sum() == 2.        # binary occupancy


def test_q3_q11_move_wrap_and_cell_resolution():
    ag = _agents(3, [(0.0, 0.0, 1, 0), (0.999, 0.5, 1, 0), (0.5, 0.5, 1, 0)])
    env = _env(h=5, w=5, agents=ag)
    act = np.zeros((3, 3))
    act[0] = [-1e-18, 0.002, 0.125]         # tiny negative wraps to exactly 1.0; 1.001 wraps to ~0.001; tie at .625
    env.step(act)
    assert env.agents[0, 0] == 1.0
    assert abs(env.agents[0, 1] - 0.001) < 1e-12
    ix, iy = env.last_cells
    assert ix.tolist() == [4, 0, 3] and iy.tolist() == [0, 2, 2]


def test_limit_boundary_clips():
    ag = _agents(2, [(0.95, 0.05, 1, 0), (0.5, 0.5, 1, 0)])
    env = _env(agents=ag, boundary='limit')
    act = np.zeros((3, 2))
    act[0, 0], act[1, 0] = 0.2, -0.2
    env.step(act)
    assert env.agents[:2, 0].tolist() == [1.0, 0.0]


def test_q4_q5_sense_is_clamped_and_gradient_nonperiodic():
    h = w = 16
    chem = np.zeros((h, w))
    chem[0, :] = np.linspace(1, 2, w)            # a ridge on the first row only
    medium = np.zeros((3, h, w))
    medium[2] = chem
    ag = _agents(2, [(0.98, 0.5, 1, 0), (0.02, 0.5, 1, 0)])
    prev = np.array([[1., -1.], [0., 0.]])       # theta = 0 and pi
    agent = R.PhysarumAgent(max_agents=2, scale=0.01, sense_offset=0.1, prev_grad=prev)
    agent.forward((ag, medium), coin=np.array([0, 1]))
    sx, sy = agent.last_sense_cells
    assert sx.tolist() == [15, 0]                 # clamped to the edge rows, NOT wrapped to the ridge / far side
    g = R.gradient_field(chem, normalized=False, grad_clip=None)
    assert g[0, 15, 3] == 0.0                     # last row sees no ridge: non-periodic
    assert g[0, 0, 3] == -chem[0, 3]              # one-sided difference at the edge


def test_q8_physarum_ghosts_move_and_brownian_ghosts_do_not():
    h = w = 8
    medium = np.zeros((3, h, w))
    medium[1] = 0.3
    ag = _agents(4, [(0.5, 0.5, 1, 0.2)])
    phys = R.PhysarumAgent(max_agents=4, scale=0.01, prev_grad=np.ones((2, 4)))
    act = phys.forward((ag, medium), coin=np.array([0, 1, 0, 1]))
    assert (np.hypot(act[0], act[1]) > 0.009).all()          # every slot, alive or not, moves
    assert np.allclose(act[2], 4.0 * 0.3 * 0.1)               # zero gradient: undetermined -> 0.1 deposit mask
    brown = R.BrownianAgent(0.01)
    actb = brown.forward((ag, medium), u=np.full((3, 4), 0.7))
    assert (actb[:, 1:] == 0).all() and (actb[:, 0] != 0).all()


def test_q9_brownian_quantisation():
    ag = _agents(3, [(0, 0, 1, 0), (0, 0, 1, 0), (0, 0, 1, 0)])
    u = np.array([[0.12345, 0.9996, 0.0004]] * 3)
    act = R.BrownianAgent(0.01, 0.5).forward((ag, None), u=u)
    assert np.array_equal(act[0], 0.02 * np.array([0.123, 1.0, 0.0]) - 0.01)
    assert np.array_equal(act[2], 0.5 * np.array([0.123, 1.0, 0.0]))


def test_physarum_turn_rules():
    """_choose_turn, core/agent/gradient.py:168-193: gradient to the left -> +turn, to the right ->
    -turn, behind (> sense_angle) or aligned -> coin; deposit mask only for determined turns."""
    agent = R.PhysarumAgent(max_agents=5, prev_grad=np.stack([np.ones(5), np.zeros(5)]))   # theta = 0
    tr = np.radians(30)
    drads = np.array([np.radians(45), np.radians(-45), np.radians(135), 0.0, np.radians(1)])
    turn = agent._choose_turn(drads, coin=np.array([1, 1, 0, 0, 1]))
    assert np.allclose(turn, [tr, -tr, -tr, -tr, tr])
    assert agent._deposit_mask.tolist() == [True, True, True, False, False]


def test_info_rounding_q12():
    ag = _agents(2, [(0.5, 0.5, 1, 0)])
    env = _env(agents=ag, food=0.123456789)
    _, reward, term, trunc, info = env.step(np.zeros((3, 2)))
    assert info['reward'] == np.round(reward, 3) and info['reward'] != reward
    assert info['mean_reward'] == np.round(reward / 1, 5)
    assert term is False and trunc is False


def test_mass_conservation_of_blur_and_decay():
    np.random.seed(0)
    env = R.Env((32, 48), noise_seed=1)
    env.medium[2] = np.random.random((32, 48))
    before = env.medium[2].sum()
    env._medium_diffuse_decay()
    assert np.isclose(env.medium[2].sum(), 0.9 * before, rtol=1e-13)


def test_jones_three_sensor_rule():
    """oracle JonesAgent (the specification of die_b200.JonesAgent; the reference has no three-sensor mode): one particle at
    the centre of a 41 x 41 field heading along +x; sensors FL / F / FR at +45 / 0 / -45 degrees, 0.25 away, i.e. cells
    (27, 27), (30, 20), (27, 13).  The five branches of the rule, by inspection."""
    n = 41
    g = R.grid_coords(n)
    ag = _agents(1, [(g[20], g[20], 1, 0.5)])
    cells = {'FL': (27, 27), 'F': (30, 20), 'FR': (27, 13)}
    def act(values, coin=0):
        med = np.zeros((3, n, n))
        med[1] = 0.25
        for k, v in values.items():
            med[2][cells[k]] = v
        a = R.JonesAgent(max_agents=1, scale=0.01, deposit=4.0, sense_offset=0.25, turn_angle=45, sense_angle=45,
                         theta0=np.zeros(1))
        out = a.forward((ag, med), coin=np.array([coin]))
        return np.degrees(a._direction_rads[0]), out[:, 0]
    assert act(dict(F=3, FL=1, FR=2))[0] == 0                       # front largest: straight on
    assert np.isclose(act(dict(F=1, FL=2, FR=3), coin=1)[0], 45)    # front smallest: by the coin
    assert np.isclose(act(dict(F=1, FL=2, FR=3), coin=0)[0], -45)
    assert np.isclose(act(dict(F=2, FL=1, FR=3))[0], -45)           # towards the larger side
    assert np.isclose(act(dict(F=2, FL=3, FR=1))[0], 45)
    assert act(dict(F=1, FL=1, FR=1))[0] == 0                       # no information: straight on
    th, a = act(dict(F=2, FL=3, FR=1))
    assert np.allclose(a, [0.01 * np.cos(np.pi / 4), 0.01 * np.sin(np.pi / 4), 4.0 * 0.25])
