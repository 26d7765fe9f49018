"""die_b200.render.EnvRenderer (die_render_frames) against the oracle's restatement of
core/render.py: medium frame, agent trace and agents frame bit for bit, over several steps."""
import numpy as np
import pytest

from oracle import die_ref as R
from tests._parity import make_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field,colors", [((48, 64), 'rgb'), ((37, 53), 'one'), ((64, 40), 'two')])
def test_frames_equal_the_oracle(field, colors):
    import die_b200 as D
    from die_b200.render import EnvRenderer
    (ref,), gpu = make_pair(field, seed=19, ratio=0.2)
    rr, gr = R.EnvRenderer(field, field_colors_id=colors), EnvRenderer(field, field_colors_id=colors)
    ra, ga = R.BrownianAgent(0.02), D.BrownianAgent(0.02)
    m = ref.agents.shape[-1]
    rng = np.random.default_rng(1)
    for it in range(8):
        rf = rr.render(ref.medium, ref.agents)
        gf = gr.render(gpu.medium, gpu.agents)
        for a, b in zip(rf, gf):
            assert np.array_equal(a, b.cpu().numpy()), it
        hf = gr.render_host(gpu.medium, gpu.agents) if it == 7 else None
        if hf is not None:                       # (a second render call advances the trace once more)
            rf2 = rr.render(ref.medium, ref.agents)
            assert all(np.array_equal(a, b) for a, b in zip(rf2, hf))
        u = rng.random((3, m))
        ref.step(ra.forward(ref._get_current_obs, u=u))
        gpu.step(ga.forward(gpu._get_current_obs, u=u))


def test_env_render_batched_and_host():
    import die_b200 as D
    field = (32, 24)
    (r0, r1), gpu = make_pair(field, seed=20, ratio=0.2, batch=2)
    rrs = [R.EnvRenderer(field), R.EnvRenderer(field)]
    ga = D.ConstAgent((0.01, 0.02), 0.3)
    ras = [R.ConstAgent((0.01, 0.02), 0.3), R.ConstAgent((0.01, 0.02), 0.3)]
    for it in range(4):
        frames = gpu.render()
        for b, (ref, rr) in enumerate(zip((r0, r1), rrs)):
            rf = rr.render(ref.medium, ref.agents)
            for a, f in zip(rf, frames):
                assert np.array_equal(a, f[b].cpu().numpy()), (it, b)
            ref.step(ras[b].forward(ref._get_current_obs))
        gpu.step(ga.forward(gpu._get_current_obs))
    single = D.Env(field, D.Dynamics(), init_state=(r0.medium, r0.agents))
    out = single.render(host=True)
    assert isinstance(out[0], np.ndarray) and out[0].shape == (*field, 3) and out[1].shape == field \
        and out[2].shape == (field[1], field[0], 4)


def test_render_with_fewer_slots_than_cells_skips_the_agents_frame():
    import die_b200 as D
    from die_b200.render import EnvRenderer
    field = (16, 16)
    med = np.zeros((3, *field))
    ag = np.zeros((4, 10))
    env = D.Env(field, D.Dynamics(), init_state=(med, ag))
    frames = EnvRenderer(field).render(env.medium, env.agents)
    assert frames[2] is None and frames[0].shape == (*field, 3)
