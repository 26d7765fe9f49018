"""NeuralAutomataAgent (core/agent/evo.py:117-209): the convolution-stack policy.  CPU part: the CUDA kernel sources
(die_conv_kernels.cuh) under the emulator against the oracle, whose restatement equals the reference's own code executed
live (tests/test_golden_oracle.py).  float32 arithmetic whose accumulation order differs from torch's: the tolerance is
1e-5 relative + 5e-6 absolute (sums of up to 147 float32 products of size ~1, layer after layer) on the model output and
the action; the gather (cells) is exact."""
import numpy as np
import pytest

from oracle import die_ref as R

S = pytest.importorskip("tests.hostsim.sim")
import die_b200 as D                                    # noqa: E402
from tests.test_hostsim_kernels import make_pair       # noqa: E402


def _weights(kernel_sizes, cin, seed):
    rng = np.random.default_rng(seed)
    outs = [cin] * (len(kernel_sizes) - 1) + [3]
    return [rng.uniform(-0.3, 0.3, size=(co, cin, k, k)).astype(np.float32) for k, co in zip(kernel_sizes, outs)]


@pytest.mark.parametrize("field,kernel_sizes,with_agent_channel,batch", [((24, 40), (3,), True, None), ((33, 37), (3, 5), True, 2),
                                                                        ((20, 64), (5, 3, 3), False, None), ((12, 12), (7,), True, 3),
                                                                        ((5, 9), (7, 7), True, None)])
def test_conv_policy_kernels_against_the_oracle(field, kernel_sizes, with_agent_channel, batch):
    refs, env = make_pair(field, seed=4, ratio=0.2, batch=batch)
    rng = np.random.default_rng(1)
    for r in refs:                                       # a chem1 field to look at
        r.medium[2] = rng.random(field) * 3.0
    env.medium[...] = np.stack([r.medium for r in refs])
    cin = 3 if with_agent_channel else 2
    weights = _weights(kernel_sizes, cin, 3)
    coefs = (0.05, 0.05, 0.7)
    for use_cells in (False, True):
        if use_cells:                                    # after a step the env's cell cache is valid for its agents
            act0 = S.brownian_forward(env.agents, move_scale=0.02, seed=4, step=0)
            env.step(act0)
            for b, r in enumerate(refs):
                r.step(np.asarray(act0).reshape(env.B, 3, env.M)[b])
        action, sense = S.conv_policy_forward(env, weights, coefs, with_agent_channel, use_cells=use_cells)
        for b, r in enumerate(refs):
            oa = R.NeuralAutomataAgent(weights, scale=coefs[0], deposit=coefs[2], with_agent_channel=with_agent_channel)
            oact = oa.forward(r._get_current_obs)
            np.testing.assert_allclose(sense[b], oa.sense_output.numpy()[0], rtol=1e-5, atol=5e-6)
            np.testing.assert_allclose(action[b], oact.astype(np.float64), rtol=1e-5, atol=5e-6)
            assert np.array_equal(action[b].astype(np.float32).astype(np.float64), action[b]), "float32 values in the float64 action"


def test_population_of_models_one_per_environment():
    """die_conv_policy_forward_population: environment b of a batch is evaluated with weight set b -- equal, bit for bit,
    to evaluating every environment alone with its model, and within the float32 tolerance of the oracle's agent."""
    field, kernel_sizes = (20, 36), (3, 5)
    refs, env = make_pair(field, seed=6, ratio=0.2, batch=3)
    rng = np.random.default_rng(2)
    for r in refs:
        r.medium[2] = rng.random(field) * 2.0
    env.medium[...] = np.stack([r.medium for r in refs])
    population = [_weights(kernel_sizes, 3, 10 + b) for b in range(3)]
    coefs = (0.05, 0.05, 0.7)
    action, sense = S.conv_policy_forward(env, None, coefs, True, population=population)
    for b, r in enumerate(refs):
        _, single = make_pair(field, seed=6 + b, ratio=0.2)
        single.medium[...] = r.medium[None]
        a1, s1 = S.conv_policy_forward(single, population[b], coefs, True)
        assert np.array_equal(a1[0], action[b]) and np.array_equal(s1[0], sense[b])
        oa = R.NeuralAutomataAgent(population[b], scale=coefs[0], deposit=coefs[2])
        np.testing.assert_allclose(action[b], oa.forward(r._get_current_obs).astype(np.float64), rtol=1e-5, atol=5e-6)
    assert not np.array_equal(sense[0], sense[1])


def test_pgpe_climbs_a_quadratic_and_keeps_its_bookkeeping():
    """die_b200.evolve.PGPE (symmetric sampling, centred ranks, ClipUp): host logic only."""
    from die_b200.evolve import PGPE, ClipUp, centered_ranks
    assert np.array_equal(centered_ranks([3.0, -1.0, 10.0]), [0.0, -0.5, 0.5])
    opt = ClipUp(3, stepsize=0.1, max_speed=0.15)
    v1 = opt.ascent(np.array([10.0, 0.0, 0.0])).copy()
    assert np.allclose(v1, [0.1, 0, 0])
    v2 = opt.ascent(np.array([5.0, 0.0, 0.0]))
    assert np.isclose(np.linalg.norm(v2), 0.15)                      # 0.09 + 0.1 clipped to max_speed
    target = np.linspace(-1, 1, 12)
    es = PGPE(12, popsize=16, radius_init=1.5, seed=0)
    f0 = -np.sum((es.center - target) ** 2)
    for gen in range(150):
        cand = es.ask()
        assert cand.shape == (16, 12) and np.allclose(cand[0::2] + cand[1::2], 2 * es.center)
        es.tell(-np.sum((cand - target) ** 2, axis=1))
    assert es.generation == 150
    assert -np.sum((es.center - target) ** 2) > 0.05 * f0            # 20x closer (in squared distance) than the start
    assert es.best[0] >= -np.sum((es.best[1] - target) ** 2) - 1e-12
    with pytest.raises(ValueError):
        es.tell(np.zeros(16))                                        # no ask() before
    with pytest.raises(ValueError):
        PGPE(4, popsize=5)
