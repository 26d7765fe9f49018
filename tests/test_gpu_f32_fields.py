"""The float32 FIELD mode on the GPU: Env(field_dtype=torch.float32) -- medium and consumed_field in float32, agents /
actions / headings in float64 (SURVEY section 7 "fp32 vs fp64 fields"; die_env_set_field_dtype).

Stated bounds (asserted below):
  per step, against the float64 oracle restarted from the GPU's own state (float32 fields widened): actions, positions,
    cells, occupancy, alive, num_agents BIT-EXACT; chem1 / food within 2^-24 relative (one rounding); agent_food and
    the reward within 1e-6 relative (+ 1e-7 absolute: one rounding of consumed_field);
  free run, 300 steps at 256^2, against the float64 GPU run with the same in-kernel random numbers: a rounding can flip a
    turn decision that sits on a threshold, after which the two trajectories of that slot differ; asserted: >= 99.9 % of
    the slots in the same cell after 300 steps, the summed reward within 1e-6 relative, the fields' means within 1e-6
    relative.  (Measured on a B200: 100.00 % same cell, reward 2e-8, field means 1e-8.)"""
import numpy as np
import pytest

from tests._parity import lattice_theta, make_pair, ref_cells_linear

pytestmark = pytest.mark.gpu
PHYS = dict(scale=0.007, turn_angle=30, sense_offset=0.04)
EPS32 = 2.0 ** -24
TINY32 = 2.0 ** -149          # the spacing of float32 subnormals: diffusion spreads chem1 down to 1e-40 and below


def _f32_env(D, torch, ref, field, batch=None, **dyn):
    med = np.stack([r.medium for r in ref]).astype(np.float32)
    ag = np.stack([r.agents for r in ref])
    return D.Env(field, D.Dynamics(init_agent_ratio=0.1, **dyn), batch=batch, field_dtype=torch.float32,
                 init_state=(torch.from_numpy(med), torch.from_numpy(ag)))


@pytest.mark.parametrize("field,sigma,iters", [((256, 256), 0.5, 40), ((96, 130), 0.8, 25), ((64, 64), 0.1, 10)])
def test_float32_fields_shadowed_per_step(field, sigma, iters):
    """The benchmarked kernels (LEAN forward on the float32 gradient cache, in-kernel coins replayed on the host)."""
    import torch
    import die_b200 as D
    from die_b200 import _lib, philox as P
    from oracle import die_ref as R
    R.set_math_backend('portable')
    try:
        np.random.seed(3)
        ref = R.Env(field, R.Dynamics(init_agent_ratio=0.1, diffuse_sigma=sigma), noise_seed=3)
        env = _f32_env(D, torch, [ref], field, diffuse_sigma=sigma)
        m = env.max_agents
        theta0, prev = lattice_theta(m, 30, 3)
        ra = R.PhysarumAgent(max_agents=m, prev_grad=prev, **PHYS)
        ga = D.PhysarumAgent(max_agents=m, rng='philox', seed=77, **PHYS)
        ga.set_state(theta=theta0)
        gobs = env._get_current_obs
        assert gobs[1].dtype == torch.float32 and gobs[0].dtype == torch.float64
        lean0 = _lib.load().die_get_counter(b"forward_lean_f32")
        for it in range(iters):
            med, ag = env.get_state()
            ref.medium[...] = med.astype(np.float64)
            ref.agents[...] = ag
            ra._direction_rads = ga.get_state()[0].copy()
            coin = P.physarum_coins(77, it, 1, m)[0].astype(np.int64)
            ract = ra.forward(ref._get_current_obs, coin=coin)
            gact = ga.forward(gobs)
            assert np.array_equal(gact.cpu().numpy(), ract), f"action differs at step {it}"
            _, rr, _, _, rinfo = ref.step(ract)
            gobs, gr, _, _, ginfo = env.step(gact)
            med, ag = env.get_state()
            assert med.dtype == np.float32
            assert np.array_equal(med[0].astype(np.float64), ref.medium[0]) and np.array_equal(ag[:3], ref.agents[:3])
            assert np.array_equal(env.last_cells().cpu().numpy(), ref_cells_linear(ref))
            for ch in (1, 2):
                err = np.abs(med[ch].astype(np.float64) - ref.medium[ch])
                assert (err <= EPS32 * np.abs(ref.medium[ch]) + TINY32).all(), f"channel {ch} beyond one float32 rounding, step {it}"
            np.testing.assert_allclose(ag[3], ref.agents[3], rtol=1e-6, atol=1e-7)   # (one rounding of consumed_field, <= 6e-8 absolute, also where the stock crosses 0)
            assert ginfo['num_agents'] == rinfo['num_agents'] and abs(gr - rr) <= 1e-6 * max(1.0, abs(rr))
        if sigma >= 0.3:
            assert _lib.load().die_get_counter(b"forward_lean_f32") - lean0 == iters - 1
    finally:
        R.set_math_backend('numpy')


def test_float32_fields_free_run_300_steps_bound():
    import torch
    import die_b200 as D
    field = (256, 256)
    outs = []
    for dt in (torch.float64, torch.float32):
        refs, env64 = make_pair(field, seed=9)
        env = env64 if dt == torch.float64 else _f32_env(D, torch, refs, field)
        if dt == torch.float64:                       # same float32-representable initial fields for both runs
            med, ag = env.get_state()
            env.set_state(medium=med.astype(np.float32).astype(np.float64))
        m = env.max_agents
        ga = D.PhysarumAgent(max_agents=m, rng='philox', seed=5, **PHYS)
        ga.set_state(theta=lattice_theta(m, 30, 9)[0])
        obs = env._get_current_obs
        total = 0.0
        for _ in range(300):
            obs, r, *_ = env.step(ga.forward(obs))
            total += r
        med, ag = env.get_state()
        outs.append((med.astype(np.float64), ag, env.last_cells().cpu().numpy(), total))
    (m64, a64, c64, r64), (m32, a32, c32, r32) = outs
    same = float((c64 == c32).mean())
    alive = a64[2] > 0
    same_alive = float((c64[alive] == c32[alive]).mean())
    rel_reward = abs(r64 - r32) / abs(r64)
    rel_chem = abs(m64[2].mean() - m32[2].mean()) / m64[2].mean()
    rel_food = abs(m64[1].mean() - m32[1].mean()) / m64[1].mean()
    print(f"float32 fields after 300 free steps: same cell {same:.4f} (alive {same_alive:.4f}), reward rel {rel_reward:.2e}, "
          f"mean chem rel {rel_chem:.2e}, mean food rel {rel_food:.2e}")
    assert np.array_equal(a64[2], a32[2])
    assert same >= 0.999 and rel_reward <= 1e-6 and rel_chem <= 1e-6 and rel_food <= 1e-6


def test_float32_fields_batch_and_brownian():
    """A batch of float32 environments with BrownianAgent: field-independent actions, so positions / cells / occupancy
    are bit-exact against the oracle for the whole free run; fields within 1e-5 relative after 60 steps."""
    import torch
    import die_b200 as D
    from oracle import die_ref as R
    field, B = (64, 96), 3
    refs = []
    for b in range(B):
        np.random.seed(20 + b)
        refs.append(R.Env(field, R.Dynamics(init_agent_ratio=0.1), noise_seed=20 + b))
    env = _f32_env(D, torch, refs, field, batch=B)
    for b in range(B):
        refs[b].medium[...] = refs[b].medium.astype(np.float32).astype(np.float64)
    m = env.max_agents
    ra, ga = R.BrownianAgent(0.02), D.BrownianAgent(move_scale=0.02)
    rng = np.random.default_rng(1)
    gobs = env._get_current_obs
    for it in range(60):
        u = rng.random((B, 3, m))
        gact = ga.forward(gobs, u=u)
        for b in range(B):
            refs[b].step(ra.forward(refs[b]._get_current_obs, u=u[b]))
        gobs, *_ = env.step(gact)
    med, ag = env.get_state()
    for b in range(B):
        assert np.array_equal(med[b, 0].astype(np.float64), refs[b].medium[0]) and np.array_equal(ag[b, :3], refs[b].agents[:3])
        for ch in (1, 2):
            np.testing.assert_allclose(med[b, ch].astype(np.float64), refs[b].medium[ch], rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(ag[b, 3], refs[b].agents[3], rtol=1e-5, atol=1e-6)


def test_float32_mode_refuses_what_it_does_not_implement():
    import torch
    import die_b200 as D
    with pytest.raises(NotImplementedError):
        D.Env((32, 32), D.Dynamics(apply_sense_mask=True), field_dtype=torch.float32)
    with pytest.raises(NotImplementedError):
        D.Env((32, 32), D.Dynamics(diffuse_sigma=1.5), field_dtype=torch.float32)
    env = D.Env((32, 32), field_dtype=torch.float32)
    with pytest.raises(NotImplementedError):
        env.step(np.zeros((3, env.max_agents)))
    with pytest.raises(NotImplementedError):
        env.render()
