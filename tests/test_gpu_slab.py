"""One field split into row slabs over G ranks (die_b200/slab.py) must reproduce the single-GPU Env
bit for bit.  Here the G ranks are emulated in one process on one GPU (the multi-rank kernels never
wait on each other, so running the phases rank after rank is exact); the real multi-process /
NVLink path is exercised by tools/slab_check.py under torchrun."""
import numpy as np
import pytest

from tests._parity import make_pair, lattice_theta

pytestmark = pytest.mark.gpu
PHYS = dict(scale=0.02, turn_angle=30, sense_offset=0.08)      # large moves: agents cross slab seams quickly


@pytest.mark.parametrize("shape,G,band", [((64, 64), 2, 0), ((64, 64), 4, 0), ((96, 80), 4, 0), ((48, 36), 3, 0),
                                          ((128, 32), 8, 0), ((64, 64), 2, 5), ((96, 80), 4, 7), ((128, 32), 8, 64),
                                          ((48, 36), 3, 2)])
def test_slab_world_equals_single_env(shape, G, band):
    """band > 0: the corner mirror (the four band x band corner patches copied to every rank after the field
    pass) serves the gathers that fall into it; everything must stay bit-identical."""
    import die_b200 as D
    from die_b200.slab import EmulatedSlabWorld
    (ref,), env = make_pair(shape, seed=31, ratio=0.15)
    medium0, agents0 = env.get_state()
    m = env.max_agents
    theta0, _ = lattice_theta(m, 30, 31)
    agent = D.PhysarumAgent(max_agents=m, **PHYS)
    agent.set_state(theta=theta0)
    world = EmulatedSlabWorld(medium0, agents0, theta0, G, corner_r=band, **PHYS)
    rng = np.random.default_rng(1)
    obs = env._get_current_obs
    crossed = 0
    for it in range(30):
        coin = rng.integers(0, 2, m)
        act = agent.forward(obs, coin=coin)
        world.forward(coin)
        obs, r, _, _, info = env.step(act)
        wr, walive, _ = world.step()
        med, ag = env.get_state()
        wmed, wag, wth, wact, wcells = world.gather()
        assert np.array_equal(wact, act.cpu().numpy()), f"action differs at step {it}"
        assert np.array_equal(wcells, env.last_cells().cpu().numpy()), f"cells differ at step {it}"
        assert np.array_equal(wmed, med), f"medium differs at step {it}"
        assert np.array_equal(wag, ag), f"agents differ at step {it}"
        assert np.array_equal(wth, agent.get_state()[0]), f"theta differs at step {it}"
        assert walive == info['num_agents']
        assert abs(wr - r) <= 1e-11 * max(1.0, abs(r))
        rows = wcells // shape[1]
        owner_of_cell = rows // (shape[0] // G)
        slot_owner = np.concatenate([np.full(world.layout.local_slots(q), q) for q in range(G)])
        ids = np.concatenate([world.layout.global_ids(q) for q in range(G)])
        crossed = max(crossed, int((owner_of_cell[ids] != slot_owner).sum()))
    assert crossed > 0, "the test must exercise agents standing on another rank's slab"
