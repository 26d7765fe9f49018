"""die_b200/device_init.py (the initial state built by tensor ops instead of the reference's
per-cell Python loop) against the host initialiser die_b200/data_init.py, which follows the reference
(core/data_init.py:132-150, 181-231).  The functions are device-agnostic torch code, so the arithmetic is
checked here on the CPU; tests/test_gpu_device_init.py runs them on the GPU."""
import numpy as np
import pytest
import torch

from die_b200 import data_init, device_init


@pytest.mark.parametrize("shape,periods", [((64, 48), 8), ((37, 53), 5), ((256, 256), 8), ((130, 31), 16)])
def test_gradient_noise_matches_the_host_version(shape, periods):
    ang = device_init.lattice_angles(periods, seed=3, device='cpu')
    dev = device_init.gradient_noise_rows(shape[0], shape[1], 0, shape[0], periods, ang).numpy()
    host = data_init.gradient_noise(shape, periods, ang=ang.numpy())
    # same arithmetic; torch's and numpy's cos/sin may differ in the last bit, which can flip a 3-dp rounding
    assert np.abs(dev - host).max() <= 1e-3 + 1e-12
    assert (dev == host).mean() > 0.999
    assert np.array_equal(dev, np.round(dev, 3))
    # rows can be generated in pieces (the slab environment does, rank by rank)
    lo = device_init.gradient_noise_rows(shape[0], shape[1], 0, shape[0] // 2, periods, ang).numpy()
    hi = device_init.gradient_noise_rows(shape[0], shape[1], shape[0] // 2, shape[0], periods, ang).numpy()
    assert np.array_equal(np.concatenate([lo, hi]), dev)


def test_grid_coords_are_numpy_linspace():
    for n in (2, 3, 7, 256, 1000, 4096):
        assert np.array_equal(device_init.grid_coords(n, 0, n, 'cpu').numpy(), np.linspace(0., 1., n))
        assert np.array_equal(device_init.grid_coords(n, n // 2, n, 'cpu').numpy(), np.linspace(0., 1., n)[n // 2:])


@pytest.mark.parametrize("shape,batch", [((40, 56), 1), ((33, 20), 4)])
def test_agents_compaction_matches_the_host_version(shape, batch):
    medium = device_init.init_medium_device(shape, 0.1, seed=5, device='cpu', batch=batch)
    agents = device_init.agents_from_medium_device(medium, seed=5).numpy()
    med = medium.numpy()
    assert agents.shape == (batch, 4, shape[0] * shape[1])
    for b in range(batch):
        np.random.seed(0)
        host = data_init.agents_from_medium(med[b])
        assert np.array_equal(agents[b, :3], host[:3])                 # x, y, alive: slot order + grid coordinates
        n = int(host[2].sum())
        food = agents[b, 3]
        assert (food[n:] == 0).all() and (food[:n] >= 0.1).all() and (food[:n] <= 1.0).all()
        u = (food[:n] - 0.1) / 0.9
        assert np.abs(u - np.round(u, 3)).max() < 1e-12                # 0.9 * round(u, 3) + 0.1


def test_medium_statistics_and_masks():
    shape, ratio = (128, 96), 0.1
    med = device_init.init_medium_device(shape, ratio, seed=11, device='cpu', batch=6).numpy()
    occ, food, chem = med[:, 0], med[:, 1], med[:, 2]
    assert set(np.unique(occ)) <= {0.0, 1.0} and (chem == 0).all()
    assert abs(occ.mean() - ratio) < 0.01                               # P(agent) = 0.100
    assert (food >= 0).all() and (food <= 1).all() and (food > 0).mean() > 0.3
    assert np.array_equal(food, np.round(food, 3))
    assert not np.array_equal(food[0], food[1])                         # independent textures per environment
    again = device_init.init_medium_device(shape, ratio, seed=11, device='cpu', batch=6).numpy()
    assert np.array_equal(med, again)                                   # seeded => reproducible


def test_too_few_slots_is_an_error():
    medium = device_init.init_medium_device((16, 16), 0.5, seed=1, device='cpu')
    with pytest.raises(ValueError):
        device_init.agents_from_medium_device(medium, seed=1, max_agents=3)
