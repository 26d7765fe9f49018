"""Helpers shared by the GPU parity tests: build an oracle (numpy) environment and a
die_b200 (CUDA) environment on the same seeded state and compare them."""
import numpy as np

from oracle import die_ref as R


def make_pair(field_size, ratio=0.1, seed=0, dynamics_kw=None, ref_dynamics_kw=None, batch=None):
    """-> (ref_envs list, gpu_env).  Same initial state in both."""
    import die_b200 as D
    dynamics_kw = dynamics_kw or {}
    ref_dynamics_kw = ref_dynamics_kw if ref_dynamics_kw is not None else dict(dynamics_kw)
    B = batch or 1
    refs = []
    for b in range(B):
        np.random.seed(seed + b)
        refs.append(R.Env(field_size, R.Dynamics(init_agent_ratio=ratio, **ref_dynamics_kw), noise_seed=seed + b))
    medium = np.stack([r.medium for r in refs])
    agents = np.stack([r.agents for r in refs])
    gpu = D.Env(field_size, D.Dynamics(init_agent_ratio=ratio, **dynamics_kw), batch=batch,
                init_state=(medium, agents))
    return refs, gpu


def lattice_theta(m, turn_angle=30, seed=0):
    """theta_0 as PhysarumAgent.__init__ builds it (core/agent/gradient.py:162)."""
    rng = np.random.default_rng(seed)
    prev = rng.normal(0., 0.4, size=(2, m))
    tr = np.radians(turn_angle)
    return (R.get_radians(prev) // tr) * tr, prev


def assert_state_equal(ref_env, med_gpu, ag_gpu, float_exact=True, rtol=1e-12):
    """Integer-valued channels bit-exact; float channels bit-exact or within rtol."""
    assert np.array_equal(ref_env.medium[0], med_gpu[0]), "occupancy differs"
    assert np.array_equal(ref_env.agents[2], ag_gpu[2]), "alive differs"
    if float_exact:
        assert np.array_equal(ref_env.agents[:2], ag_gpu[:2]), "positions differ"
        assert np.array_equal(ref_env.medium[1], med_gpu[1]), "food differs"
        assert np.array_equal(ref_env.medium[2], med_gpu[2]), "chem differs"
        assert np.array_equal(ref_env.agents[3], ag_gpu[3]), "agent_food differs"
    else:
        np.testing.assert_allclose(ag_gpu[:2], ref_env.agents[:2], rtol=0, atol=1e-14)
        np.testing.assert_allclose(med_gpu[1], ref_env.medium[1], rtol=rtol, atol=1e-300)
        np.testing.assert_allclose(med_gpu[2], ref_env.medium[2], rtol=rtol, atol=1e-300)
        np.testing.assert_allclose(ag_gpu[3], ref_env.agents[3], rtol=rtol, atol=1e-15)


def ref_cells_linear(ref_env):
    ix, iy = ref_env.last_cells
    return (ix * ref_env._field_size[1] + iy).astype(np.int32)
