"""Host-side logic of the slab decomposition (die_b200/slab.py): slot ownership ranges and the split of a global
state into per-rank slabs.  Pure numpy -- runs without a GPU; the kernels are covered by tests/test_gpu_slab.py."""
import numpy as np
import pytest

from die_b200 import data_init
from die_b200.slab import make_layout, split_global_state


def test_layout_ranges_cover_every_slot_once():
    L = make_layout((64, 32), 4, 2048, [50, 60, 40, 55])
    assert L.s0 == [0, 50, 110, 150] and L.n0 == [50, 60, 40, 55]
    assert sum(L.n1) == 2048 - 205 and L.s1[0] == 205
    ids = np.concatenate([L.global_ids(q) for q in range(4)])
    assert sorted(ids.tolist()) == list(range(2048))
    assert max(L.local_slots(q) for q in range(4)) - min(L.local_slots(q) for q in range(4)) <= 21   # alive imbalance only
    assert L.rows_per == 16
    c = L.to_c(2)
    assert (c.G, c.rank, c.H, c.W, c.M) == (4, 2, 64, 32, 2048) and c.s0[3] == 150 and c.n1[0] == L.n1[0]


def test_layout_rejects_bad_geometry():
    with pytest.raises(ValueError):
        make_layout((30, 8), 4, 240, [1, 1, 1, 1])            # H not divisible by G
    with pytest.raises(ValueError):
        make_layout((64, 8), 16, 512, [1] * 16)               # more ranks than DIE_MAX_RANKS


@pytest.mark.parametrize("shape,G", [((32, 24), 2), ((48, 16), 3), ((64, 40), 8)])
def test_split_global_state_round_trip(shape, G):
    np.random.seed(3)
    medium = data_init.init_medium(shape, 0.2, noise_seed=1)
    agents = data_init.agents_from_medium(medium)
    layout, mediums, locals_ = split_global_state(medium, agents, G)
    assert np.array_equal(np.concatenate(mediums, axis=1), medium)
    back = np.zeros_like(agents)
    for q in range(G):
        back[:, layout.global_ids(q)] = locals_[q]
    assert np.array_equal(back, agents)
    rows_per = shape[0] // G
    for q in range(G):          # every rank's alive run starts on its own rows
        n0 = layout.n0[q]
        rows = np.rint(locals_[q][0, :n0] * (shape[0] - 1)).astype(int)
        assert ((rows // rows_per) == q).all() and (locals_[q][2, :n0] == 1).all() and (locals_[q][2, n0:] == 0).all()


def test_split_requires_the_reference_slot_order():
    medium = np.zeros((3, 8, 8))
    agents = np.zeros((4, 64))
    agents[2, 5] = 1            # an alive agent that is not in slots [0, A)
    with pytest.raises(ValueError):
        split_global_state(medium, agents, 2)


def test_slab_dynamics_are_validated_up_front():
    """Everything the slab kernels do not implement is refused by name before anything is allocated (advisor finding of
    round 1: apply_sense_mask was silently ignored, other options failed late inside the first kernel call)."""
    import pytest
    from die_b200.env import Dynamics
    from die_b200.slab import validate_slab_dynamics
    from die_b200.data_init import WaveSequence
    validate_slab_dynamics(Dynamics())
    validate_slab_dynamics(Dynamics(diffuse_sigma=0.8, food_infinite=True))
    flow = WaveSequence((16, 16), dt=0.01).get_flow_operator(scale=0.5, decay=0.5)
    for bad in (dict(apply_sense_mask=True), dict(diffuse_mode='reflect'), dict(diffuse_sigma=0.1), dict(diffuse_sigma=1.5),
                dict(agents_die=True), dict(op_food_flow=flow)):
        with pytest.raises(NotImplementedError):
            validate_slab_dynamics(Dynamics(**bad))
